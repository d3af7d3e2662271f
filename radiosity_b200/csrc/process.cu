// K2 — ProcessHemicube: delta-form-factor-weighted scatter-add (replaces the OpenCL kernel
// Kernel_ProcessHemicube.h:9-70, its launch Main.cpp:584-596,1218-1226, the four blocking read-backs
// Main.cpp:1232-1240 and the O(k * records) CPU gather Main.cpp:1257-1269).
//
//   F_h[i] = sum over atlas pixels px of hemicube h with item[px] == i+1 of dFF[px]
//
// The reference expresses this as run-length records appended through an atomic counter and
// finished on the CPU.  Here every warp streams 128 consecutive pixels per step (one 16 B load of ids +
// one 16 B load of dFF per lane), merges runs inside the lane, collapses runs of equal ids across lanes
// with a 5-step segmented shuffle reduction
// (ids are spatially coherent, so a run is almost always a contiguous lane range) and issues ONE
// red.global.add.f32 per run straight into F_h — no record stream, no host round trip.
// Algorithmic traffic: 4 B id + 4 B dFF per pixel (the dFF table of one hemicube is shared by all k
// hemicubes and stays L2-resident); see DESIGN.md.
//
// Two entry points: process_kernel<false> reads the uint32 item buffer (the API seam of
// rad_process_hemicubes / rad_bench_process); process_kernel<true> is the fused steady-state form used
// by rad_shoot: it reads the rasteriser's 64-bit keys directly (a key counts iff its top byte is the group's
// epoch tag, so nothing is cleared) and only materialises the item buffer when asked to.
#include "rad_internal.cuh"
#include "segadd.cuh"

namespace {

constexpr int kUnroll = 2;

__device__ __forceinline__ uint32_t key_id(unsigned long long k, uint32_t tag) { return (uint32_t)(k >> 56) == tag ? (uint32_t)(k & 0xFFFFFFFFull) : 0u; }

template <bool FROM_KEYS>
__global__ void __launch_bounds__(256) process_kernel(RadDev D, int keep_items) {
	const uint32_t slot = D.h0 + blockIdx.y;
	if (FROM_KEYS && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && (D.qc->q_tris | D.qc->q_small | D.qc->n_pairs)) { D.qc->parked = D.qc->q_tris + D.qc->q_small; D.qc->q_tris = 0; D.qc->q_entries = 0; D.qc->q_small = 0; D.qc->n_pairs = 0; }
	if (!D.em[slot].valid) return;
	const int lane = threadIdx.x & 31;
	const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	const uint32_t nsteps = D.RES >> 7;               // 128 pixels per warp step; RES = 3 N^2 is a multiple of 768
	float* __restrict__ F = D.F + (size_t)slot * D.P;
	const float4* __restrict__ ff4 = reinterpret_cast<const float4*>(D.ff);
	const ulonglong2* __restrict__ keys2 = reinterpret_cast<const ulonglong2*>(D.keys + (size_t)(slot - D.kbase) * D.RES);
	uint4* __restrict__ items4 = reinterpret_cast<uint4*>(D.items + (size_t)slot * D.RES);

	for (uint32_t g0 = gw * kUnroll; g0 < nsteps; g0 += nw * kUnroll) {
		uint4 id[kUnroll]; float4 v[kUnroll];
		#pragma unroll
		for (int u = 0; u < kUnroll; u++) {
			const uint32_t g = g0 + u;
			id[u] = make_uint4(0u, 0u, 0u, 0u); v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
			if (g < nsteps) {
				const uint32_t q = (g << 5) + lane;   // index of this lane's 4-pixel group
				if (FROM_KEYS) {
					const ulonglong2 k0 = keys2[2 * (size_t)q], k1 = keys2[2 * (size_t)q + 1];
					id[u] = make_uint4(key_id(k0.x, D.tag), key_id(k0.y, D.tag), key_id(k1.x, D.tag), key_id(k1.y, D.tag));
				} else {
					id[u] = __ldcs(items4 + q);
				}
				v[u] = __ldg(ff4 + q);
			}
		}
		#pragma unroll
		for (int u = 0; u < kUnroll; u++) {
			const uint32_t g = g0 + u;
			if (g < nsteps) {
				const uint32_t q = (g << 5) + lane;
				if (FROM_KEYS) {
					if (keep_items) items4[q] = id[u];   // keys are not cleared: the next render uses a smaller epoch tag
				}
				process4(id[u], v[u], lane, F, D.P);
			}
		}
	}
}

} // namespace

static dim3 process_grid(const RadDev& D) {
	const uint32_t nslots = D.h1 - D.h0;
	const uint32_t groups = D.RES >> 7;
	uint32_t warps = (groups + kUnroll - 1) / kUnroll;
	uint32_t bx = (warps + 7) / 8;
	const uint32_t cap = 148 * 16;                    // a few waves of 256-thread CTAs per hemicube
	if (bx > cap) bx = cap;
	if (bx == 0) bx = 1;
	return dim3(bx, nslots);
}

void rad_launch_process(rad_ctx* c) {
	const RadDev& D = c->d;
	if (D.h1 == D.h0) return;
	process_kernel<false><<<process_grid(D), 256, 0, c->stream>>>(D, 0);
	c->launches++;
}

void rad_launch_process_view(rad_ctx* c, const RadDev& V, cudaStream_t st, uint32_t s0, uint32_t n, uint32_t kbase, bool keep_items) {
	RadDev D = V;
	D.h0 = V.h0 + s0; D.h1 = D.h0 + n; D.kbase = kbase;
	process_kernel<true><<<process_grid(D), 256, 0, st>>>(D, keep_items ? 1 : 0);
	c->launches++;
	c->keys_dirty = false;
}
void rad_launch_process_group(rad_ctx* c, uint32_t s0, uint32_t n, uint32_t kbase, bool keep_items) {
	rad_launch_process_view(c, c->d, c->stream, s0, n, kbase, keep_items);
}

void rad_launch_resolve_process(rad_ctx* c, bool keep_items) {
	if (c->d.h1 == c->d.h0) return;
	rad_launch_process_group(c, 0, c->d.h1 - c->d.h0, 0, keep_items);
}
