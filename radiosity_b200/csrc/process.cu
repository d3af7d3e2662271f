// K2 — ProcessHemicube: delta-form-factor-weighted scatter-add (replaces the OpenCL kernel
// Kernel_ProcessHemicube.h:9-70, its launch Main.cpp:584-596,1218-1226, the four blocking read-backs
// Main.cpp:1232-1240 and the O(k * records) CPU gather Main.cpp:1257-1269).
//
//   F_h[i] = sum over atlas pixels px of hemicube h with item[px] == i+1 of dFF[px]
//
// The reference expresses this as run-length records appended through an atomic counter and
// finished on the CPU (the work-item scans W/4 pixels of a row and emits a record per run,
// Kernel_ProcessHemicube.h:40-58).  Here a CTA takes one chunk of 512 consecutive atlas pixels for a group
// of hemicubes: the chunk's dFF entries come in ONCE (TMA bulk copy, shared by the CTA's warps), every warp
// streams the id chunks of its hemicubes through shared memory (cp.async.bulk + mbarrier, the next chunk in
// flight while the current one is reduced), every lane owns a span of 16 CONSECUTIVE pixels, merges the runs
// of its span in registers and the warp issues ONE red.global.add.f32 per run straight into F_h — no record
// stream, no host round trip, no cross-lane stage except for chunks that are one single run.
// Algorithmic traffic: 4 B id + 4 B dFF per pixel (the dFF table of one hemicube is shared by all k
// hemicubes and stays L2-resident); see DESIGN.md.
//
// Two entry points: process_kernel<false> reads the uint32 item buffer (the API seam of
// rad_process_hemicubes / rad_bench_process); process_kernel<true> is the fused steady-state form used
// by rad_shoot: it reads the rasteriser's 64-bit keys directly  (a key counts iff its top byte is the render's
// epoch tag, so nothing is cleared) and only materialises the item buffer when asked to.
#include "rad_internal.cuh"
#include "segadd.cuh"

#ifndef FULL
#define FULL 0xFFFFFFFFu
#endif

namespace {

constexpr int kSpan = 16;                          // consecutive pixels per lane
constexpr int kChunk = 32 * kSpan;                 // pixels per warp step (one TMA bulk copy of ids / keys)
constexpr int kWarps = 4;                          // warps per CTA: they share the chunk's dFF entries, each streams its own hemicubes
constexpr int kMaxGroup = 16;                      // hemicube slots per task (kMaxGroup / kWarps per warp)

// ---- mbarrier + TMA bulk copy (cp.async.bulk: global -> shared, completion counted in bytes on the mbarrier) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// (id+1, sum) of one run into F: predicated red, no branch
__device__ __forceinline__ void flush_run_pred(uint32_t id, float v, float* __restrict__ F, uint32_t P, bool on = true) {
	asm volatile("{\n.reg .pred q, o;\n.reg .u32 c;\n.reg .u64 a;\n"
	             "sub.u32 c, %0, 1;\n"
	             "setp.ne.u32 o, %4, 0;\n"
	             "setp.lt.and.u32 q, c, %3, o;\n"                            // id != 0 (empty pixel) and id - 1 < P
	             "mad.wide.u32 a, c, 4, %2;\n"
	             "@q red.global.add.f32 [a], %1;\n}"
	             :: "r"(id), "f"(v), "l"(F), "r"(P), "r"((int)on) : "memory");
}

// Shared memory of a CTA: [dFF chunk: kChunk floats][per warp: id chunk | run list | one span of zeros][mbarriers].
// PXB = bytes per pixel of the id stream: 8 (64-bit keys of the rasteriser) or 4 (uint32 item buffer).
template <int PXB> struct Smem {
	static constexpr int kRuns = kChunk * PXB;                         // run list: (id+1, sum) pairs, [run][lane], at most kSpan per lane
	static constexpr int kZero = kRuns + kSpan * 32 * 8;               // ids read by the lanes beyond a partial chunk
	static constexpr int kWarp = kZero + kSpan * PXB;
	static constexpr int kBars = kChunk * 4 + kWarps * kWarp;          // kWarps id barriers, then the dFF barrier
	static constexpr int kBytes = kBars + (kWarps + 1) * 8;
};

// One task = (chunk of kChunk consecutive atlas pixels, group of <= kMaxGroup hemicube slots).  The dFF entries of the
// chunk are the same for every hemicube (the reference replicates the table per hemicube, FormFactors.cpp:317-323, and
// reads it per pixel, Kernel_ProcessHemicube.h:45): one bulk copy per task, every lane keeps the entries of its span
// in registers.  Warp w then walks the non-NULL slots w, w + kWarps, ... of the group: lane 0 issues ONE bulk copy per id
// chunk and re-issues the next one as soon as the ids are in registers, so the copy overlaps the reduction.
// A lane's span is kSpan CONSECUTIVE pixels: ids are spatially coherent, so runs are merged with one compare per
// pixel in registers; a run's end parks (id, sum) in the lane's list in shared memory with a PREDICATED store (run
// ends fall on different pixels in different lanes — flushing in place makes the warp execute a divergent block per
// pixel), and the lists are flushed together, one red.global.add.f32 per run.  A chunk that is one single run (large
// patches, the constant-ID extreme) is reduced with five shuffles and flushed by ONE atomic.  The lane reads its span
// in 16-byte pieces starting at a lane-dependent piece (rotation), so that every LDS.128 wavefront hits eight different
// bank groups; the pieces therefore arrive in rotated order — harmless, F_h[i] is a sum, adjacency only saves atomics.
template <bool FROM_KEYS, bool KEEP>
__global__ void __launch_bounds__(kWarps * 32, FROM_KEYS ? 6 : 8) process_kernel(RadDev D, uint32_t group) {
	constexpr int PXB = FROM_KEYS ? 8 : 4;
	constexpr int kPieces = kSpan * PXB / 16;                          // 16-byte pieces of a lane's id span: 8 (keys) / 4 (items)
	constexpr int kPxPiece = 16 / PXB;                                 // pixels per piece: 2 / 4
	extern __shared__ __align__(128) unsigned char smem[];
	if (FROM_KEYS && blockIdx.x == 0 && threadIdx.x == 0 && (D.qc->q_tris | D.qc->q_small | D.qc->n_pairs)) { D.qc->parked = D.qc->q_tris + D.qc->q_small; D.qc->q_tris = 0; D.qc->q_entries = 0; D.qc->q_small = 0; D.qc->n_pairs = 0; }
	if (FROM_KEYS && D.stop_gate && D.ctl->gate) return;              // the stop test fired in an earlier batch of this replay
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t s_ff = smem_u32(smem), s_id = s_ff + kChunk * 4 + warp * Smem<PXB>::kWarp;
	const uint32_t bar_id = s_ff + Smem<PXB>::kBars + warp * 8, bar_ff = s_ff + Smem<PXB>::kBars + kWarps * 8;
	if (lane == 0) {
		mbar_init(bar_id, 1);
		if (warp == 0) mbar_init(bar_ff, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (lane < kSpan * PXB / 4) reinterpret_cast<uint32_t*>(smem + kChunk * 4 + warp * Smem<PXB>::kWarp + Smem<PXB>::kZero)[lane] = 0u;
	__syncthreads();
	uint32_t par_id = 0, par_ff = 0;
	const uint32_t s_runs = s_id + Smem<PXB>::kRuns + (uint32_t)lane * 8u;
	// epoch tags only ever decrease and a minimum survives: a key belongs to this render iff its high word is below (tag + 1) << 24
	const uint32_t tag_end = (D.tag + 1u) << 24, P = D.P;
	const uint32_t nch = (D.RES + kChunk - 1) / kChunk;
	const uint32_t ngroups = (D.h1 - D.h0 + group - 1) / group;
	const uint32_t ntasks = nch * ngroups;
	for (uint32_t task = blockIdx.x; task < ntasks; task += gridDim.x) {
		const uint32_t grp = task / nch, ch = task - grp * nch;        // consecutive CTAs: consecutive chunks of one group
		const uint32_t s_lo = D.h0 + grp * group, s_n = min(group, D.h1 - s_lo);
		const uint32_t valid = __ballot_sync(FULL, (uint32_t)lane < s_n && D.em[s_lo + (uint32_t)lane].valid != 0);   // NULL emitters render nothing (Main.cpp:1253)
		if (valid == 0) continue;                                      // (CTA-uniform)
		const uint32_t px0 = ch * kChunk, npx = min((uint32_t)kChunk, D.RES - px0);   // RES is a multiple of 256: a span is all in or all out
		const bool live = (uint32_t)lane * kSpan < npx;
		// lanes beyond a partial chunk (RES % 512 == 256) read ids from the warp's span of zeros: nothing to add
		const uint32_t id_base = live ? s_id + (uint32_t)lane * (kSpan * PXB) : s_id + Smem<PXB>::kZero;
		uint32_t mine = 0;                                             // this warp's slots: bits warp, warp + kWarps, ...
		#pragma unroll
		for (int i = 0; i < kMaxGroup / kWarps; i++) mine |= 1u << (warp + i * kWarps);
		mine &= valid;
		auto issue = [&](uint32_t bits) {                              // bulk copy of the id chunk of the first slot in `bits`
			if (bits && lane == 0) {
				const uint32_t s = s_lo + (uint32_t)__ffs(bits) - 1u;
				const void* src = FROM_KEYS ? (const void*)(D.keys + (size_t)(s - D.kbase) * D.RES + px0) : (const void*)(D.items + (size_t)s * D.RES + px0);
				mbar_expect_tx(bar_id, npx * PXB);
				tma_bulk_g2s(s_id, src, npx * PXB, bar_id);
			}
		};
		if (threadIdx.x == 0) { mbar_expect_tx(bar_ff, npx * 4); tma_bulk_g2s(s_ff, D.ff + px0, npx * 4, bar_ff); }
		issue(mine);
		// the dFF entries of this lane's span, in the piece order of the id stream
		float f[kSpan];
		mbar_wait(bar_ff, par_ff); par_ff ^= 1u;
		#pragma unroll
		for (int j = 0; j < kPieces; j++) {
			const uint32_t pc = FROM_KEYS ? ((uint32_t)(j + lane + (lane >> 3)) & 7u) : ((uint32_t)(j + (lane >> 1)) & 3u);
			const uint32_t a = s_ff + (live ? (uint32_t)lane * (kSpan * 4) : 0u) + pc * (kPxPiece * 4);
			if (FROM_KEYS) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(f[2 * j]), "=f"(f[2 * j + 1]) : "r"(a));
			else asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f[4 * j]), "=f"(f[4 * j + 1]), "=f"(f[4 * j + 2]), "=f"(f[4 * j + 3]) : "r"(a));
		}
		while (mine) {
			const uint32_t slot = s_lo + (uint32_t)__ffs(mine) - 1u; mine &= mine - 1u;
			mbar_wait(bar_id, par_id); par_id ^= 1u;
			uint32_t id[kSpan];
			#pragma unroll
			for (int j = 0; j < kPieces; j++) {
				const uint32_t pc = FROM_KEYS ? ((uint32_t)(j + lane + (lane >> 3)) & 7u) : ((uint32_t)(j + (lane >> 1)) & 3u);
				uint4 v;
				asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(id_base + pc * 16u));
				if (FROM_KEYS) {
					id[2 * j] = v.y < tag_end ? v.x : 0u; id[2 * j + 1] = v.w < tag_end ? v.z : 0u;
					if (KEEP && live) *reinterpret_cast<uint2*>(D.items + (size_t)slot * D.RES + px0 + (uint32_t)lane * kSpan + pc * 2u) = make_uint2(id[2 * j], id[2 * j + 1]);
				} else { id[4 * j] = v.x; id[4 * j + 1] = v.y; id[4 * j + 2] = v.z; id[4 * j + 3] = v.w; }
			}
			__syncwarp();                                              // the ids are in registers: the next chunk may land
			issue(mine);
			uint32_t cur = id[0], wr = s_runs; float acc = f[0];
			#pragma unroll
			for (int j = 1; j < kSpan; j++) {
				asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %1, %2;\n@p st.shared.v2.b32 [%0], {%2, %3};\n@p add.u32 %0, %0, 256;\n}"
				             : "+r"(wr) : "r"(id[j]), "r"(cur), "r"(__float_as_uint(acc)) : "memory");
				acc = id[j] != cur ? f[j] : acc + f[j];
				cur = id[j];
			}
			float* __restrict__ F = D.F + (size_t)slot * P;
			const uint32_t first = __shfl_sync(FULL, cur, 0);
			if (__all_sync(FULL, live && wr == s_runs && cur == first)) {   // the whole chunk is one run
				#pragma unroll
				for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(FULL, acc, d);
				flush_run_pred(cur, acc, F, P, lane == 0);
			} else {
				flush_run_pred(cur, acc, F, P);
				const uint32_t wmax = __reduce_max_sync(FULL, wr);
				for (uint32_t a = s_runs; a < wmax; a += 512u) {       // two parked runs per step (the loads overlap)
					uint32_t rid0 = 0u, rv0 = 0u, rid1 = 0u, rv1 = 0u;
					asm volatile("{\n.reg .pred p;\nsetp.lt.u32 p, %2, %3;\n@p ld.shared.v2.b32 {%0, %1}, [%2];\n}" : "+r"(rid0), "+r"(rv0) : "r"(a), "r"(wr));
					asm volatile("{\n.reg .pred p;\nsetp.lt.u32 p, %2, %3;\n@p ld.shared.v2.b32 {%0, %1}, [%2];\n}" : "+r"(rid1), "+r"(rv1) : "r"(a + 256u), "r"(wr));
					flush_run_pred(rid0, __uint_as_float(rv0), F, P);
					flush_run_pred(rid1, __uint_as_float(rv1), F, P);
				}
			}
		}
		__syncthreads();                                               // every warp holds its dFF entries: the next task may overwrite the chunk
	}
}

// Fused form for launches of fewer slots than a CTA has warps (k = 1: ONE hemicube per launch, a latency chain): the task
// kernel above would leave three of four warps idle and pay a bulk-copy round trip per chunk, so every warp takes 128-pixel
// steps of the slot directly — two steps' loads in flight, four pixels per lane, runs merged by the segmented shuffle of
// segadd.cuh.  grid: x = warps over the atlas, y = slot.
template <bool KEEP>
__global__ void __launch_bounds__(256) process_few_kernel(RadDev D) {
	pdl_enter();
	const uint32_t slot = D.h0 + blockIdx.y;
	if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && (D.qc->q_tris | D.qc->q_small | D.qc->n_pairs)) { D.qc->parked = D.qc->q_tris + D.qc->q_small; D.qc->q_tris = 0; D.qc->q_entries = 0; D.qc->q_small = 0; D.qc->n_pairs = 0; }
	if (D.stop_gate && D.ctl->gate) return;
	if (!D.em[slot].valid) return;
	const int lane = threadIdx.x & 31;
	const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	const uint32_t nsteps = D.RES >> 7;               // 128 pixels per warp step (RES = 3 N^2, N a multiple of 8)
	const uint32_t tag_end = (D.tag + 1u) << 24;
	float* __restrict__ F = D.F + (size_t)slot * D.P;
	const float4* __restrict__ ff4 = reinterpret_cast<const float4*>(D.ff);
	const uint4* __restrict__ keys4 = reinterpret_cast<const uint4*>(D.keys + (size_t)(slot - D.kbase) * D.RES);
	uint4* __restrict__ items4 = reinterpret_cast<uint4*>(D.items + (size_t)slot * D.RES);
	for (uint32_t g0 = gw * 2; g0 < nsteps; g0 += nw * 2) {
		uint4 id[2]; float4 v[2];
		#pragma unroll
		for (int u = 0; u < 2; u++) {
			const uint32_t g = g0 + u;
			id[u] = make_uint4(0u, 0u, 0u, 0u); v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
			if (g < nsteps) {
				const uint32_t q = (g << 5) + lane;   // this lane's group of four pixels
				const uint4 k0 = __ldcs(keys4 + 2 * (size_t)q), k1 = __ldcs(keys4 + 2 * (size_t)q + 1);
				id[u] = make_uint4(k0.y < tag_end ? k0.x : 0u, k0.w < tag_end ? k0.z : 0u, k1.y < tag_end ? k1.x : 0u, k1.w < tag_end ? k1.z : 0u);
				v[u] = __ldg(ff4 + q);
			}
		}
		#pragma unroll
		for (int u = 0; u < 2; u++) {
			const uint32_t g = g0 + u;
			if (g < nsteps) {
				if (KEEP) items4[(g << 5) + lane] = id[u];
				process4(id[u], v[u], lane, F, D.P);
			}
		}
	}
}

} // namespace

// slots per task: up to kMaxGroup, fewer when the launch would otherwise have too few tasks to fill the GPU
static uint32_t process_group_size(const RadDev& D) {
	const uint32_t nslots = D.h1 - D.h0, nch = (D.RES + kChunk - 1) / kChunk;
	uint32_t g = kMaxGroup;
	while (g > (uint32_t)kWarps && (uint64_t)nch * ((nslots + g - 1) / g) < 148ull * 6ull) g >>= 1;
	return g;
}
template <bool FROM_KEYS, bool KEEP>
static void launch_process(const RadDev& D, cudaStream_t st) {
	constexpr int kSmem = Smem<FROM_KEYS ? 8 : 4>::kBytes;
	static const bool attr = [] { cudaFuncSetAttribute(process_kernel<FROM_KEYS, KEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem); return true; }();
	(void)attr;
	const uint32_t g = process_group_size(D), nch = (D.RES + kChunk - 1) / kChunk;
	const uint64_t tasks = (uint64_t)nch * ((D.h1 - D.h0 + g - 1) / g);
	const uint64_t cap = 148ull * (FROM_KEYS ? 6 : 8);               // persistent: what is resident
	const uint64_t waves = (tasks + cap - 1) / cap;
	uint64_t bx = waves ? (tasks + waves - 1) / waves : 1;            // every CTA the same number of tasks
	if (bx == 0) bx = 1;
	process_kernel<FROM_KEYS, KEEP><<<(uint32_t)bx, kWarps * 32, kSmem, st>>>(D, g);
}

void rad_launch_process(rad_ctx* c) {
	const RadDev& D = c->d;
	if (D.h1 == D.h0) return;
	launch_process<false, false>(D, c->stream);
	c->launches++;
}

void rad_launch_process_view(rad_ctx* c, const RadDev& V, cudaStream_t st, uint32_t s0, uint32_t n, uint32_t kbase, bool keep_items) {
	RadDev D = V;
	D.h0 = V.h0 + s0; D.h1 = D.h0 + n; D.kbase = kbase;
	if (n < (uint32_t)kWarps && (D.RES & 127u) == 0) {         // k = 1 and other tiny launches
		uint32_t bx = ((D.RES >> 7) + 15) / 16;                    // two steps per warp, eight warps per CTA
		if (bx > 148u * 8u) bx = 148u * 8u;
		if (keep_items) rad_launch_pdl(c->pdl, process_few_kernel<true>, dim3(bx, n), dim3(256), 0, st, D);
		else rad_launch_pdl(c->pdl, process_few_kernel<false>, dim3(bx, n), dim3(256), 0, st, D);
	} else if (keep_items) launch_process<true, true>(D, st); else launch_process<true, false>(D, st);
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
