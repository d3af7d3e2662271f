// K1T — tile-binned form of the hemicube rasteriser with ProcessHemicube fused into it (opt-in: RAD_RASTER=tiles; replaces
// the same reference code as raster.cu + process.cu: the GL render Main.cpp:1148-1202 and the OpenCL kernel
// Kernel_ProcessHemicube.h:9-70 with its CPU gather Main.cpp:1257-1269).
//
// raster.cu resolves visibility with RED.MIN.64 on a key buffer in global memory (8 B per pixel written and read back
// through L2 / HBM, 24 N^2 bytes per hemicube) and process.cu streams those keys again.  Here the parked records of the
// set-up kernel (small-quad records and large triangles, same formats) are binned by atlas tile instead, and ONE CTA per
// (hemicube, tile) does the rest on chip:
//   bin_kernel<false>   one lane per record: count it in every tile its bbox overlaps (large triangles: only tiles that
//                       pass the corner test)
//   bin_scan_kernel     exclusive scan of the counters -> first reference of every (slot, tile, list)
//   bin_kernel<true>    same walk again, writing the record index into its place
//   tile_kernel         keys of the tile in shared memory (16 KB): small quads four per warp (quarter-warp walk of
//                       (bbox) n (tile), tile_walk.cuh), large triangles shared by the CTA's warps, 64-bit atomicMin of
//                       (depth24 << 32 | id+1) in shared memory — the same deterministic visibility rule; then the tile's
//                       ids and the dFF entries of its pixels go through the run merging of ProcessHemicube (segadd.cuh)
//                       straight into F_h.  The item buffer is only written when asked for (parity / debugging).
// Traffic per hemicube: the records (64 B each) + 4 B dFF per pixel (L2-resident table) + the F reds; no key traffic.
// The arithmetic of every walk is checked on the CPU against a brute-force statement of the raster rules
// (tests/cpu/tile_walk_check.cpp) and on the GPU against the oracle and the global-key path (tests/test_gpu_tiles.py).
#include "rad_internal.cuh"
#include "segadd.cuh"
#include "tile_walk.cuh"
#include <cstdlib>

namespace {

struct SmemMin {                  // emit target of the walks: shared-memory key array of the tile
	unsigned long long* k;
	__device__ __forceinline__ void operator()(int off, unsigned long long key) const { atomicMin(k + off, key); }
};

__device__ __forceinline__ tw::BigTri load_big(const RadBigTri& r) {
	tw::BigTri t;
	t.X0 = r.X0; t.Y0 = r.Y0; t.X1 = r.X1; t.Y1 = r.Y1; t.X2 = r.X2; t.Y2 = r.Y2;
	t.z0 = r.z0; t.dz1 = r.dz1; t.dz2 = r.dz2; t.inv_area = r.inv_area; t.id1 = r.id1;
	t.px0 = r.px0; t.py0 = r.py0; t.px1 = r.px1; t.py1 = r.py1;
	return t;
}

// one lane per parked record of the launch group; FILL = false counts, FILL = true writes the references
template <bool FILL>
__global__ void __launch_bounds__(256) bin_kernel(RadDev D, RadTiles T) {
	const uint32_t nsm = min(D.qc->q_small, D.q_sm_cap), ntri = min(D.qc->q_tris, D.q_tri_cap);
	const uint32_t total = nsm + ntri;
	const uint4* __restrict__ qsm = reinterpret_cast<const uint4*>(D.q_sm);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
		const bool big = i >= nsm;
		int px0, py0, px1, py1; uint32_t slot;
		tw::BigTri bt;
		if (!big) {
			const uint4 c = __ldg(qsm + 4 * (size_t)i + 2);
			const uint2 d = __ldg(reinterpret_cast<const uint2*>(qsm + 4 * (size_t)i + 3));
			slot = c.w & 0xFFFFu;
			px0 = (int)(d.x & 0xFFFFu); py0 = (int)(d.x >> 16);
			px1 = px0 + (int)(d.y & 0xFFu) - 1; py1 = py0 + (int)((d.y >> 8) & 0xFFu) - 1;
		} else {
			const RadBigTri r = D.q_tri[i - nsm];
			bt = load_big(r);
			slot = r.slot; px0 = r.px0; py0 = r.py0; px1 = r.px1; py1 = r.py1;
		}
		int t0x, t0y, t1x, t1y;
		tw::tile_range(px0, py0, px1, py1, t0x, t0y, t1x, t1y);
		const uint32_t jslot = (slot - D.h0) * T.T;
		for (int ty = t0y; ty <= t1y; ty++)
			for (int tx = t0x; tx <= t1x; tx++) {
				if (big) {
					const int ox = tx * RAD_TILE_W, oy = ty * RAD_TILE_H;
					tw::BigWalk w; w.init(bt, ox, oy, min(RAD_TILE_W, (int)D.W - ox), min(RAD_TILE_H, (int)D.H - oy));
					if (w.rejects()) continue;
				}
				const uint32_t j = (jslot + (uint32_t)ty * T.tx + (uint32_t)tx) * 2u + (big ? 1u : 0u);
				if (!FILL) atomicAdd(&T.cnt[j], 1u);
				else {
					const uint32_t r = T.base[j] + atomicAdd(&T.cnt[j], 1u);
					if (r < T.refs_cap) T.refs[r] = big ? i - nsm : i; else D.ctl->q_overflow = 1;
				}
			}
	}
}

// exclusive scan of cnt[0, n) into base[0, n]; the counters are zeroed on the way (the fill pass counts again, the tile
// CTAs zero them once more for the next launch).  One CTA: n = 2 * tiles * slots of a launch group, a few thousand entries
__global__ void __launch_bounds__(1024) bin_scan_kernel(RadDev D, RadTiles T, uint32_t n) {
	__shared__ uint32_t wsum[32];
	__shared__ uint32_t blocksum;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t carry = 0;
	for (uint32_t i0 = 0; i0 < n; i0 += 4096u) {
		const uint32_t i = i0 + 4u * threadIdx.x;
		uint32_t v[4];
		#pragma unroll
		for (int j = 0; j < 4; j++) { v[j] = i + j < n ? T.cnt[i + j] : 0u; if (i + j < n) T.cnt[i + j] = 0u; }
		const uint32_t s = v[0] + v[1] + v[2] + v[3];
		uint32_t inc = s;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += o; }
		if (lane == 31) wsum[warp] = inc;
		__syncthreads();
		if (warp == 0) {
			const uint32_t w = wsum[lane];
			uint32_t wi = w;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL, wi, d); if (lane >= d) wi += o; }
			wsum[lane] = wi - w;
			if (lane == 31) blocksum = wi;
		}
		__syncthreads();
		uint32_t ex = carry + wsum[warp] + (inc - s);
		#pragma unroll
		for (int j = 0; j < 4; j++) { if (i + j < n) T.base[i + j] = ex; ex += v[j]; }
		carry += blocksum;
		__syncthreads();
	}
	if (threadIdx.x == 0) { T.base[n] = carry; if (carry > T.refs_cap) D.ctl->q_overflow = 1; }
}

// one CTA of THREADS threads per (tile, hemicube slot of the launch group)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) tile_kernel(RadDev D, RadTiles T, int keep_items) {
	__shared__ __align__(16) unsigned long long skeys[RAD_TILE_PIX];
	const uint32_t ls = blockIdx.y, slot = D.h0 + ls, tile = blockIdx.x;
	// last consumer of the lane's work lists: recycle them for the next launch group (as process_kernel does)
	if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && (D.qc->q_tris | D.qc->q_small | D.qc->n_pairs)) { D.qc->parked = D.qc->q_tris + D.qc->q_small; D.qc->q_tris = 0; D.qc->q_entries = 0; D.qc->q_small = 0; D.qc->n_pairs = 0; }
	if (!D.em[slot].valid) return;
	if (D.stop_gate && D.ctl->gate) { if (threadIdx.x < 2) T.cnt[(ls * T.T + tile) * 2u + threadIdx.x] = 0u; return; }   // the stop test fired in an earlier batch of this replay
	const uint32_t j = (ls * T.T + tile) * 2u;
	const uint32_t b0 = T.base[j], b1 = T.base[j + 1], b2 = min(T.base[j + 2], T.refs_cap);
	if (threadIdx.x < 2) T.cnt[j + threadIdx.x] = 0u;
	const int tyi = (int)(tile / T.tx), txi = (int)(tile - (uint32_t)tyi * T.tx);
	const int tx0 = txi * RAD_TILE_W, ty0 = tyi * RAD_TILE_H;
	const int tw_ = min(RAD_TILE_W, (int)D.W - tx0), th_ = min(RAD_TILE_H, (int)D.H - ty0);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint4* __restrict__ items4 = reinterpret_cast<uint4*>(D.items + (size_t)slot * D.RES);
	if (b2 <= b0) {                                       // nothing was binned here: every pixel is empty, F gets nothing
		if (keep_items)
			for (int q = threadIdx.x; q < RAD_TILE_PIX / 4; q += blockDim.x) {
				const int row = q / (RAD_TILE_W / 4), c4 = q % (RAD_TILE_W / 4);
				if (4 * c4 < tw_ && row < th_) items4[((size_t)(ty0 + row) * D.W + tx0) / 4 + c4] = make_uint4(0u, 0u, 0u, 0u);
			}
		return;
	}
	{
		ulonglong2* s2 = reinterpret_cast<ulonglong2*>(skeys);
		for (int q = threadIdx.x; q < RAD_TILE_PIX / 2; q += blockDim.x) s2[q] = make_ulonglong2(~0ull, ~0ull);
	}
	__syncthreads();
	const SmemMin emit{ skeys };
	// small quads: four records per warp step, a quarter warp each
	{
		const uint4* __restrict__ qsm = reinterpret_cast<const uint4*>(D.q_sm);
		const int sub = lane >> 3, l8 = lane & 7;
		for (uint32_t r0 = b0 + 4u * warp; r0 < b1; r0 += THREADS / 8) {
			const uint32_t r = r0 + sub;
			tw::QuadWalk q; q.none();
			if (r < b1) {
				const size_t i = T.refs[r];
				const uint4 w0 = __ldg(qsm + 4 * i), w1 = __ldg(qsm + 4 * i + 1), w2 = __ldg(qsm + 4 * i + 2);
				const uint2 w3 = __ldg(reinterpret_cast<const uint2*>(qsm + 4 * i + 3));
				tw::RecWords rec;
				rec.a[0] = w0.x; rec.a[1] = w0.y; rec.a[2] = w0.z; rec.a[3] = w0.w;
				rec.b[0] = w1.x; rec.b[1] = w1.y; rec.b[2] = w1.z; rec.b[3] = w1.w;
				rec.c[0] = w2.x; rec.c[1] = w2.y; rec.c[2] = w2.z; rec.c[3] = w2.w;
				rec.d[0] = w3.x; rec.d[1] = w3.y;
				q.init(rec, tx0, ty0, tw_, th_, l8);
			}
			int msteps = q.steps();
			msteps = max(msteps, __shfl_xor_sync(FULL, msteps, 8)); msteps = max(msteps, __shfl_xor_sync(FULL, msteps, 16));
			for (int s = 0; s < msteps; s++) q.step(emit);
		}
	}
	// large triangles: the CTA's warps share the 8x4-pixel steps of (bbox) n (tile)
	for (uint32_t r = b1; r < b2; r++) {
		const RadBigTri rt = D.q_tri[T.refs[r]];
		const tw::BigTri t = load_big(rt);
		tw::BigWalk w; w.init(t, tx0, ty0, tw_, th_);
		for (int s = warp; s < w.nsteps; s += THREADS / 32) w.step(t, s, lane, emit);
	}
	__syncthreads();
	// ProcessHemicube on the resolved tile: 4 consecutive pixels per lane and step, 16 lanes per tile row
	float* __restrict__ F = D.F + (size_t)slot * D.P;
	const float4* __restrict__ ff4 = reinterpret_cast<const float4*>(D.ff);
	constexpr int kSteps = RAD_TILE_PIX / 4 / THREADS;
	uint4 id[kSteps]; float4 v[kSteps];
	#pragma unroll
	for (int it = 0; it < kSteps; it++) {
		const int q = it * THREADS + (int)threadIdx.x;
		const int row = q / (RAD_TILE_W / 4), c4 = q % (RAD_TILE_W / 4);
		const bool inb = 4 * c4 < tw_ && row < th_;
		id[it] = make_uint4(0u, 0u, 0u, 0u); v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
		if (inb) {
			const ulonglong2 k0 = *reinterpret_cast<const ulonglong2*>(skeys + row * RAD_TILE_W + 4 * c4);
			const ulonglong2 k1 = *reinterpret_cast<const ulonglong2*>(skeys + row * RAD_TILE_W + 4 * c4 + 2);
			id[it] = make_uint4(k0.x == ~0ull ? 0u : (uint32_t)k0.x, k0.y == ~0ull ? 0u : (uint32_t)k0.y,
			                    k1.x == ~0ull ? 0u : (uint32_t)k1.x, k1.y == ~0ull ? 0u : (uint32_t)k1.y);
			const size_t g = ((size_t)(ty0 + row) * D.W + tx0) / 4 + c4;
			v[it] = __ldg(ff4 + g);
			if (keep_items) items4[g] = id[it];
		}
	}
	#pragma unroll
	for (int it = 0; it < kSteps; it++) process4(id[it], v[it], lane, F, D.P);
}

} // namespace

// slots [V.h0 + s0, +n) of the lane view V: records parked by the set-up kernel -> bins -> tile CTAs.  mark(2) after the
// bins, mark(4) after the tile kernel (rad_profile_batch)
void rad_launch_tiles_view(rad_ctx* c, const RadDev& V, const RadTiles& T, cudaStream_t st, uint32_t s0, uint32_t n, bool keep_items,
                           const std::function<void(int)>& mark) {
	RadDev D = V;
	D.h0 = V.h0 + s0; D.h1 = D.h0 + n;
	const uint32_t nlists = n * T.T * 2u;
	bin_kernel<false><<<148 * 4, 256, 0, st>>>(D, T);
	bin_scan_kernel<<<1, 1024, 0, st>>>(D, T, nlists);
	bin_kernel<true><<<148 * 4, 256, 0, st>>>(D, T);
	if (mark) mark(2);
	static const int threads = [] { const char* e = getenv("RAD_TILE_THREADS"); const int v = e ? atoi(e) : 128; return v == 256 ? 256 : 128; }();   // tuning knob
	if (threads == 256) tile_kernel<256><<<dim3(T.T, n), 256, 0, st>>>(D, T, keep_items ? 1 : 0);
	else tile_kernel<128><<<dim3(T.T, n), 128, 0, st>>>(D, T, keep_items ? 1 : 0);
	if (mark) mark(4);
	c->launches += 4;
	c->keys_dirty = false;
}
