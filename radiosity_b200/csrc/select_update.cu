// K3 — shooter selection, K4 — fused energy transfer + emitter update (+ argmax of the new state).
//
// Replaces, from the reference:
//   ModelContainer::getHighestRadiosityPatchesId  ModelContainer.cpp:259-299   (CPU list + sort)
//   energy transfer                               Main.cpp:1272-1279           (CPU, O(k*P))
//   emitter update + stop test                    Main.cpp:1286-1303
//
// K4 is ONE pass over the patches: B_i += sum_h ((S_h * F_h[i]) * rho) (.) c_h in hemicube order (the
// reference's float association), F is zeroed for the next batch, emitters get I += S, B -= S, and —
// for k == 1 — the squared length of the updated B goes through a warp-shuffle block reduction and a
// single 64-bit atomicMax, which IS the next selection: key = (bits(|B|^2) << 32 | id) reproduces the
// reference's k == 1 result exactly (largest energy, LAST index among equals, patch 0 if all zero).
//
// For k > 1: RAD_SELECT_REFERENCE runs a one-block emulation of the reference's list (seeded patch 0,
// reject-below-minimum while not full, tie groups reversed by every insertion); RAD_SELECT_TOPK runs a
// tournament of block-wide bitonic sorts over keys (bits << 32 | ~id): energy desc, id asc.
#include <stdlib.h>
#include "rad_internal.cuh"
#include <cooperative_groups.h>
#include "camera.cuh"

namespace {

#define FULL 0xFFFFFFFFu

// acquire/release fence at GPU scope.  NOT __threadfence(): that is a sequentially consistent fence here (MEMBAR.SC.GPU +
// L1 invalidate), and SC fences of different blocks serialise — 258 of them cost 70 us in the update kernel.
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// system scope (peer GPUs over NVLink): flags of the fused dB exchange
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ float* xb_planes(const RadDev& D, uint32_t r, uint32_t parity) {
	return reinterpret_cast<float*>(D.xb[r] + RAD_XB_DATA) + (size_t)parity * 3 * D.xPmax;
}
__device__ __forceinline__ float* xb_reduced(const RadDev& D, uint32_t r) { return reinterpret_cast<float*>(D.xb[r] + RAD_XB_DATA) + (size_t)6 * D.xPmax; }
// spin until every rank's flag in row `row_off` of this rank's exchange buffer has reached seq (threads 0 .. world-1)
__device__ __forceinline__ void xb_wait(const RadDev& D, uint32_t row_off, uint32_t seq) {
	if (threadIdx.x < D.xworld && !D.xnowait) {
		const uint32_t* flag = reinterpret_cast<const uint32_t*>(D.xb[D.xrank] + row_off + 128 * threadIdx.x);
		while ((int32_t)(ld_acquire_sys(flag) - seq) < 0) { }
	}
	__syncthreads();
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
// Sequence protocol of the fused exchange: s = the sequence number in this rank's buffer = batches exchanged so far.  The
// batch in flight writes plane set (s + 1) & 1 (local dB: the raster lanes' lane_delta_kernel, or the whole-rank apply_kernel<1>),
// the FIRST kernel after that (xreduce_kernel, or apply_kernel<2> in the one-shot form) publishes s + 1 to every peer right
// at its start — everything the stream ran before it is in this GPU's L2 — and xb_bump_kernel, right behind the update kernel, bumps s.
__device__ __forceinline__ void xb_publish_now(const RadDev& D, uint32_t row_off, uint32_t seq) {
	if (blockIdx.x == 0 && threadIdx.x < D.xworld) {
		fence_acq_rel_sys();
		st_release_sys(reinterpret_cast<uint32_t*>(D.xb[threadIdx.x] + row_off + 128 * D.xrank), seq);
	}
}
// the block that finishes last publishes `seq` into this rank's slot of flag row `row_off` on every peer (row_off != 0)
// and / or stores it as the new sequence number (bump_seq).  Release / acquire fences: the sequentially consistent
// __threadfence_system() of every block serialises (tens of microseconds for a few hundred blocks).
__device__ __forceinline__ void xb_publish_last(const RadDev& D, uint32_t row_off, uint32_t seq, bool bump_seq) {
	__shared__ bool s_lastblk;
	__syncthreads();
	if (threadIdx.x == 0) { fence_acq_rel(); const uint32_t done = atomicAdd(&D.ctl->ticket, 1u); s_lastblk = done == gridDim.x - 1; if (s_lastblk) D.ctl->ticket = 0; }
	__syncthreads();
	if (!s_lastblk) return;
	if (threadIdx.x == 0) { fence_acq_rel_sys(); if (bump_seq) *reinterpret_cast<volatile uint32_t*>(D.xb[D.xrank]) = seq; }
	__syncthreads();
	if (row_off && threadIdx.x < D.xworld) st_release_sys(reinterpret_cast<uint32_t*>(D.xb[threadIdx.x] + row_off + 128 * D.xrank), seq);
}
__device__ __forceinline__ uint32_t xb_slice(const RadDev& D) { return (D.P + D.xworld - 1) / D.xworld; }   // patches per rank slice

__device__ __forceinline__ float len2(float x, float y, float z) { return x * x + y * y + z * z; }   // Vector.h:356-359

__device__ __forceinline__ unsigned long long warp_max(unsigned long long v) {
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		const unsigned long long o = __shfl_xor_sync(FULL, v, d);
		v = o > v ? o : v;
	}
	return v;
}
// block-wide max, result valid in thread 0
__device__ __forceinline__ unsigned long long block_max(unsigned long long v) {
	__shared__ unsigned long long s_w[32];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	v = warp_max(v);
	if (lane == 0) s_w[w] = v;
	__syncthreads();
	if (w == 0) {
		v = lane < nw ? s_w[lane] : 0ull;
		v = warp_max(v);
	}
	return v;
}

__device__ __forceinline__ unsigned long long energy_key_last(float e, uint32_t i) {   // ties -> larger id
	return e > 0.0f ? ((unsigned long long)__float_as_uint(e) << 32) | i : 0ull;
}

// ---- k == 1: argmax of the current B into selkey[parity] ------------------------------------
__global__ void __launch_bounds__(256) argmax_kernel(RadDev D, int parity) {
	unsigned long long best = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < D.P; i += gridDim.x * blockDim.x)
		best = max(best, energy_key_last(len2(D.rad[i], D.rad[D.P + i], D.rad[2 * (size_t)D.P + i]), i));
	best = block_max(best);
	if (threadIdx.x == 0 && best) atomicMax(&D.ctl->selkey[parity], best);
}

// ---- reference list semantics for k > 1 (single block) --------------------------------------
// ModelContainer::getHighestRadiosityPatchesId (ModelContainer.cpp:259-299) scans the patches in id order and keeps a list:
// a patch is appended if the list is empty, or if its energy is > 0 and >= the energy of the list's LAST entry; after every
// append the list is stable-sorted ascending and reversed (so the newcomer leads its tie group and every other tie group
// flips), then cut to `count`.  The result depends on the whole history (the seeded patch 0, candidates refused below the
// minimum while the list is not full, tie groups flipped by later appends), so the scan is emulated as it is — but each
// append is O(1) for a warp: the list lives in the registers of warp 0 (lane l holds entries l and l + 32), the insert
// position is a ballot + popc, the shift a shuffle, and tie groups (rare) are flipped through 512 bytes of shared memory.
// The other warps only compute energies and pre-filter candidates against the minimum at the start of their 1024-patch chunk
// (the minimum never decreases: a superset), so the serial part sees a few hundred candidates per call, not P.
__global__ void __launch_bounds__(1024) select_reference_kernel(RadDev D) {
	__shared__ float s_e[1024];
	__shared__ unsigned s_flags[32];
	__shared__ float s_le[64]; __shared__ uint32_t s_lid[64];
	__shared__ int s_n; __shared__ float s_min;
	const uint32_t count = D.k;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	__shared__ uint32_t s_fast;
	if (threadIdx.x == 0) s_fast = D.ctl->ref_fast;           // the tie-free fast path has written the list (topk_level_kernel, ref_mode)
	__syncthreads();
	if (s_fast) return;
	if (threadIdx.x == 0) { s_n = 0; s_min = 0.0f; }
	// warp 0: the list.  entry i < 32 in (eA, idA) of lane i, entry i >= 32 in (eB, idB) of lane i - 32
	float eA = 0.0f, eB = 0.0f; uint32_t idA = 0u, idB = 0u; int n = 0;
	__syncthreads();
	for (uint32_t base = 0; base < D.P; base += 1024) {
		const uint32_t i = base + threadIdx.x;
		const float e = i < D.P ? len2(D.rad[i], D.rad[D.P + i], D.rad[2 * (size_t)D.P + i]) : 0.0f;
		s_e[threadIdx.x] = e;
		const bool cand = i < D.P && ((s_n == 0 && i == 0) || (e > 0.0f && e >= s_min));
		const unsigned bal = __ballot_sync(FULL, cand);
		if (lane == 0) s_flags[w] = bal;
		__syncthreads();
		if (w == 0) {
			unsigned word = s_flags[lane];
			for (;;) {
				const unsigned any = __ballot_sync(FULL, word != 0u);
				if (any == 0u) break;
				const int ww = __ffs(any) - 1;                          // first warp-word with a candidate left
				const unsigned m = __shfl_sync(FULL, word, ww);
				const int j = ww * 32 + __ffs(m) - 1;
				if (lane == ww) word &= word - 1u;
				const float x = s_e[j];
				const float last = n == 0 ? 0.0f : (n <= 32 ? __shfl_sync(FULL, eA, (n - 1) & 31) : __shfl_sync(FULL, eB, (n - 33) & 31));
				if (!(n == 0 || (x > 0.0f && last <= x))) continue;     // ModelContainer.cpp:266-269
				// every existing tie group is reversed by the stable sort + reverse (ModelContainer.cpp:273-275)
				const bool vA = lane < n, vB = lane + 32 < n;
				const float pA = __shfl_up_sync(FULL, eA, 1), pB0 = __shfl_sync(FULL, eA, 31), pBs = __shfl_up_sync(FULL, eB, 1);
				const float pB = lane == 0 ? pB0 : pBs;
				const unsigned hA = __ballot_sync(FULL, vA && (lane == 0 || eA != pA)), hB = __ballot_sync(FULL, vB && eB != pB);
				const unsigned long long valid = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
				const unsigned long long heads = ((unsigned long long)hB << 32) | hA;
				if (heads != valid) {                                   // some group has more than one member
					s_le[lane] = eA; s_lid[lane] = idA; s_le[lane + 32] = eB; s_lid[lane + 32] = idB;
					__syncwarp();
					const unsigned long long hs = heads | (n >= 64 ? 0ull : (1ull << n));   // sentinel head behind the list
					#pragma unroll
					for (int half = 0; half < 2; half++) {
						const int idx = lane + 32 * half;
						if (idx < n) {
							const int s0 = 63 - __clzll((long long)(heads & ((idx >= 63 ? ~0ull : ((2ull << idx) - 1ull)))));
							const unsigned long long above = idx >= 63 ? 0ull : (hs >> (idx + 1));
							const int t1 = above ? idx + __ffsll((long long)above) : n;      // first index of the next group
							const int src = s0 + (t1 - 1) - idx;
							if (half == 0) { eA = s_le[src]; idA = s_lid[src]; } else { eB = s_le[src]; idB = s_lid[src]; }
						}
					}
					__syncwarp();
				}
				// the newcomer leads its tie group: position = number of entries with a larger energy
				const int pos = __popc(__ballot_sync(FULL, vA && eA > x)) + __popc(__ballot_sync(FULL, vB && eB > x));
				const float sAe = __shfl_up_sync(FULL, eA, 1), sBe = __shfl_up_sync(FULL, eB, 1), cAe = __shfl_sync(FULL, eA, 31);
				const uint32_t sAi = __shfl_up_sync(FULL, idA, 1), sBi = __shfl_up_sync(FULL, idB, 1), cAi = __shfl_sync(FULL, idA, 31);
				if (lane > pos) { eA = sAe; idA = sAi; } else if (lane == pos) { eA = x; idA = base + (uint32_t)j; }
				if (lane + 32 > pos) { eB = lane == 0 ? cAe : sBe; idB = lane == 0 ? cAi : sBi; } else if (lane + 32 == pos) { eB = x; idB = base + (uint32_t)j; }
				n = min(n + 1, (int)count);                             // tops.erase(it, tops.end())
			}
			if (lane == 0) {
				s_n = n;
			}
			const float mn = n == 0 ? 0.0f : (n <= 32 ? __shfl_sync(FULL, eA, (n - 1) & 31) : __shfl_sync(FULL, eB, (n - 33) & 31));
			if (lane == 0) s_min = mn;
		}
		__syncthreads();
	}
	if (w == 0) {
		#pragma unroll
		for (int half = 0; half < 2; half++) {
			const uint32_t idx = (uint32_t)lane + 32u * half;
			if (idx < count) {
				const bool ok = (int)idx < n;
				D.em[idx].id = ok ? (half ? idB : idA) : 0u;
				D.em[idx].valid = ok ? 1u : 0u;
				D.em[idx].order = idx;
			}
		}
	}
}

// ---- clean top-k for k > 1: tournament of block-wide bitonic networks ---------------------------------------------
// key = (bits(|B|^2) << 32 | ~id): unique, and descending key order == (energy desc, id asc).  Every block takes a
// chunk of 2048 keys (two per thread, elements t and t + 1024), sorts groups of `keep` (= k rounded up to a power of
// two, >= 64) and then halves the number of groups by prune-and-merge rounds (a[i] = max(a[i], b[keep-1-i]) leaves a
// bitonic run holding the best `keep` of both groups) until one group is left: the chunk's best `keep`.  Compare-
// exchanges over a distance below 32 elements are warp shuffles, only the longer ones go through shared memory.  The
// chunks' winners are the next level's input; the block that finishes LAST on a level with <= 2048 / keep chunks
// merges them and writes the emitter list — one launch for 16 k patches, two for 1 M.  Exact and deterministic.
constexpr int kTopChunk = 2048, kTopThreads = 1024;
struct TopkSpec { uint32_t count, slot_base, excl_base, excl_n; };   // speculative path (mode 2): list length and destination, slots whose patches are left out

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
	return ((unsigned long long)__shfl_xor_sync(FULL, (unsigned)(v >> 32), m) << 32) | __shfl_xor_sync(FULL, (unsigned)v, m);
}
// one compare-exchange stage of a bitonic network over the 2048 elements of the block: element e meets e ^ j and the
// pair is ordered descending where (e & k) == 0 (k = 0: descending everywhere).  a0 / a1 are elements t / t + 1024.
__device__ __forceinline__ void cmpx_stage(unsigned long long& a0, unsigned long long& a1, unsigned long long* s, int t, int j, int k) {
	if (j == 1024) {                       // the two elements of one thread
		const bool desc = k == 0 || (t & k) == 0;
		const unsigned long long hi = a0 > a1 ? a0 : a1, lo = a0 > a1 ? a1 : a0;
		a0 = desc ? hi : lo; a1 = desc ? lo : hi;
		return;
	}
	unsigned long long b0, b1;
	if (j < 32) { b0 = shfl_xor_u64(a0, j); b1 = shfl_xor_u64(a1, j); }
	else {
		__syncthreads();
		s[t] = a0; s[t + 1024] = a1;
		__syncthreads();
		b0 = s[t ^ j]; b1 = s[(t ^ j) + 1024];
	}
	const bool first = (t & j) == 0;       // this element is the lower index of its pair
	const bool d0 = k == 0 || (t & k) == 0, d1 = k == 0 || ((t + 1024) & k) == 0;
	a0 = ((a0 > b0) == (first == d0)) ? a0 : b0;
	a1 = ((a1 > b1) == (first == d1)) ? a1 : b1;
}
// best `keep` of the block's 2048 keys, sorted descending, left in elements [0, keep) (a0 of threads t < keep)
// (`span` = number of leading elements that can hold non-zero keys, rounded up to a power of two: groups beyond it are empty)
__device__ __forceinline__ void block_topk(unsigned long long& a0, unsigned long long& a1, unsigned long long* s, int t, int keep, bool groups_sorted, int span = kTopChunk) {
	if (!groups_sorted)
		for (int k = 2; k <= keep; k <<= 1)
			for (int j = k >> 1; j > 0; j >>= 1) cmpx_stage(a0, a1, s, t, j, k == keep ? 0 : k);   // last pass: every group descending
	for (int g = keep; g < span; g <<= 1) {
		// prune: groups (L, L + g/keep) -> element i of the first takes max with element g-1-i ... of the second
		__syncthreads();
		s[t] = a0; s[t + 1024] = a1;
		__syncthreads();
		{
			const int e0 = t, e1 = t + 1024;
			if ((e0 & g) == 0 && (e0 & (g - 1)) < keep) { const unsigned long long o = s[(e0 | g) + (keep - 1) - 2 * (e0 & (keep - 1))]; a0 = a0 > o ? a0 : o; }
			if ((e1 & g) == 0 && (e1 & (g - 1)) < keep) { const unsigned long long o = s[(e1 | g) + (keep - 1) - 2 * (e1 & (keep - 1))]; a1 = a1 > o ? a1 : o; }
		}
		for (int j = keep >> 1; j > 0; j >>= 1) cmpx_stage(a0, a1, s, t, j, 0);   // bitonic merge of the surviving runs
	}
}

// ref_mode (RAD_SELECT_REFERENCE, k > 1): the fast path of the reference's list.  The list of ModelContainer.cpp:259-299 only
// ever rejects a patch whose energy is below the list's last entry, and that entry's energy never decreases (it starts as
// patch 0's): the final SET is the top-`count` of S = {0} + {i : |B_i|^2 > 0 and >= |B_0|^2}, and the final ORDER is by energy
// wherever energies differ.  Ties are what makes the list history-dependent (every append reverses every tie group), so:
// the level kernels select the top-(count + 1) of S (first level filters by |B_0|^2), and if those energies are pairwise
// different — no tie inside the list, none across its end — the list is written here and select_reference_kernel, launched
// behind, returns at once (RadControl::ref_fast); otherwise it runs its exact emulation.  14 of the 16 batches of the bench
// run take the fast path (the fresh scene's 99 equal light patches are the other two).
template <bool FIRST>
__global__ void __launch_bounds__(kTopThreads) topk_level_kernel(RadDev D, const unsigned long long* __restrict__ in, uint32_t n_in,
                                                                  unsigned long long* __restrict__ out, int keep, int finish, int ref_mode, TopkSpec sp) {
	__shared__ unsigned long long s[kTopChunk];
	__shared__ bool s_last;
	__shared__ uint32_t s_excl[RAD_SPEC_SLOTS];
	const int t = threadIdx.x;
	unsigned long long a[2];
	if (FIRST && ref_mode == 2 && sp.excl_n) {      // patches already rendered ahead (the half about to be applied) are not candidates
		if ((uint32_t)t < RAD_SPEC_SLOTS) s_excl[t] = ((uint32_t)t < sp.excl_n && D.em[sp.excl_base + t].valid) ? D.em[sp.excl_base + t].id : 0xFFFFFFFFu;
		__syncthreads();
	}
	const uint32_t e0b = (FIRST && ref_mode == 1) ? __float_as_uint(len2(D.rad[0], D.rad[D.P], D.rad[2 * (size_t)D.P])) : 0u;
	#pragma unroll
	for (int r = 0; r < 2; r++) {
		const uint32_t i = blockIdx.x * kTopChunk + t + r * 1024;
		unsigned long long key = 0ull;
		if (i < n_in) {
			if constexpr (FIRST) {
				const uint32_t eb = __float_as_uint(len2(D.rad[i], D.rad[D.P + i], D.rad[2 * (size_t)D.P + i]));
				bool in_s = ref_mode != 1 || eb >= e0b;                         // (positive floats order like their bits)
				if (ref_mode == 2 && sp.excl_n && eb != 0) {
					#pragma unroll 8
					for (uint32_t j = 0; j < RAD_SPEC_SLOTS; j++) in_s = in_s && s_excl[j] != i;
				}
				if (eb != 0 && eb < 0x7F800000u && in_s) key = ((unsigned long long)eb << 32) | (ref_mode == 2 ? i : 0xFFFFFFFFu - i);   // 2: ties -> higher id first, the argmax's rule
			} else key = in[i];
		}
		a[r] = key;
	}
	block_topk(a[0], a[1], s, t, keep, !FIRST);   // later levels read sorted groups of `keep`
	if (gridDim.x > 1) {
		if (t < keep) out[(size_t)blockIdx.x * keep + t] = a[0];
		if (!finish) return;
		// last block of the level merges the chunks' winners (at most 2048 keys, groups of `keep` already sorted)
		// (one fence per block, after the barrier: it is cumulative over the block's writes; a fence in every thread
		// costs tens of microseconds on this part)
		__syncthreads();
		if (t == 0) { fence_acq_rel(); const uint32_t done = atomicAdd(&D.ctl->ticket, 1u); s_last = done == gridDim.x - 1; if (s_last) { D.ctl->ticket = 0; fence_acq_rel(); } }
		__syncthreads();
		if (!s_last) return;
		const uint32_t nc = gridDim.x * (uint32_t)keep;
		a[0] = (uint32_t)t < nc ? __ldcg(out + t) : 0ull;
		a[1] = (uint32_t)t + 1024 < nc ? __ldcg(out + t + 1024) : 0ull;
		int span = keep;
		while ((uint32_t)span < nc) span <<= 1;
		block_topk(a[0], a[1], s, t, keep, true, span);
	}
	if (ref_mode == 1) {
		// sorted descending in a[0] of threads [0, keep), keep > count: any two neighbours of equal energy among the first count + 1?
		const uint32_t count = D.k;
		__syncthreads();
		if (t < keep) s[t] = a[0];
		__syncthreads();
		const uint32_t hi = (uint32_t)(a[0] >> 32);
		const int tie = (uint32_t)t < count && hi != 0u && hi == (uint32_t)(s[t + 1] >> 32);
		const int npos = __syncthreads_count((uint32_t)t < count && a[0] != 0ull);     // list entries with positive energy
		const int ok = !__syncthreads_or(tie);
		if (t == 0) D.ctl->ref_fast = ok ? 1u : 0u;
		if (!ok) return;
		const bool seed0 = __float_as_uint(len2(D.rad[0], D.rad[D.P], D.rad[2 * (size_t)D.P])) == 0u;   // the seeded patch 0 stays behind the positive ones while there is room
		if ((uint32_t)t < count) {
			const bool isseed = seed0 && t == npos;
			D.em[t].id = a[0] ? 0xFFFFFFFFu - (uint32_t)(a[0] & 0xFFFFFFFFull) : 0u;
			D.em[t].valid = (a[0] || isseed) ? 1u : 0u;
			D.em[t].order = (uint32_t)t;
		}
		return;
	}
	if (ref_mode == 2) {                           // speculative path: `count` entries into em[slot_base ..)
		if ((uint32_t)t < sp.count) {
			RadEmitter* e = D.em + sp.slot_base + t;
			e->id = a[0] ? (uint32_t)(a[0] & 0xFFFFFFFFull) : 0u; e->valid = a[0] ? 1u : 0u; e->order = (uint32_t)t;
		}
		return;
	}
	if ((uint32_t)t < D.k) {
		const uint32_t G = D.deal, slot = G > 1 ? ((uint32_t)t % G) * (D.k / G) + (uint32_t)t / G : (uint32_t)t;
		D.em[slot].id = a[0] ? (ref_mode == 2 ? (uint32_t)(a[0] & 0xFFFFFFFFull) : 0xFFFFFFFFu - (uint32_t)(a[0] & 0xFFFFFFFFull)) : 0u;
		D.em[slot].valid = a[0] ? 1u : 0u;
		D.em[slot].order = (uint32_t)t;
	}
}

__global__ void set_emitters_kernel(RadDev D, const uint32_t* __restrict__ ids, uint32_t n) {
	for (uint32_t h = threadIdx.x; h < D.k; h += blockDim.x) { D.em[h].id = h < n ? ids[h] : 0u; D.em[h].valid = (h < n && ids[h] < D.P) ? 1u : 0u; D.em[h].order = h; }
}

// ---- K4: energy transfer + emitter update (+ fused argmax) -----------------------------------
// what K4 needs of an emitter, 32 B = two 16-byte shared-memory loads per (patch, hemicube)
struct EmLite { float S0, S1, S2; uint32_t valid; float c0, c1, c2; uint32_t id; };

// F_h[i] of kFBatch consecutive slots: all the (independent) loads are issued before anything depends on them — the
// kernel is a latency chain otherwise — then the slots are zeroed for the next batch (fill_n(p_tmp_formfactors, 0),
// Main.cpp:1278).  Invalid (NULL) emitters are not read.
constexpr int kFBatch = 16;
// (s_em[h - h_tab] = emitter of slot h: the table starts at slot h_tab)
__device__ __forceinline__ void take_F(const RadDev& D, const EmLite* s_em, uint32_t h_tab, uint32_t hb, uint32_t hend, uint32_t P, uint32_t i, float* f) {
	#pragma unroll
	for (int j = 0; j < kFBatch; j++) {
		const uint32_t h = hb + j;
		f[j] = (h < hend && s_em[h - h_tab].valid) ? __ldcs(D.F + (size_t)h * P + i) : 0.0f;
	}
	#pragma unroll
	for (int j = 0; j < kFBatch; j++)
		if (f[j] != 0.0f) D.F[(size_t)(hb + j) * P + i] = 0.0f;
}
// B += sum over slots [hb0, hend) of ((S_h * F_h[i]) * rho) (.) c_h, in slot order   (Main.cpp:1274)
__device__ __forceinline__ void gather_transfer(const RadDev& D, const EmLite* s_em, uint32_t hb0, uint32_t hend, uint32_t P, uint32_t i, float rho,
                                                float& bx, float& by, float& bz, uint32_t h_tab = 0) {
	for (uint32_t hb = hb0; hb < hend; hb += kFBatch) {
		float f[kFBatch];
		take_F(D, s_em, h_tab, hb, hend, P, i, f);
		#pragma unroll
		for (int j = 0; j < kFBatch; j++) {
			const uint32_t h = hb + j;
			if (h < hend) {
				const EmLite e = s_em[h - h_tab];
				if (e.valid) {
					bx += ((e.S0 * f[j]) * rho) * e.c0;
					by += ((e.S1 * f[j]) * rho) * e.c1;
					bz += ((e.S2 * f[j]) * rho) * e.c2;
				}
			}
		}
	}
}
// emitter h of the batch: lastEnergy (before the subtraction, Main.cpp:1292), I += S, B -= S   (Main.cpp:1286-1295)
__device__ __forceinline__ void emitter_update(const RadDev& D, const EmLite& e, bool is_last, uint32_t P, float& bx, float& by, float& bz) {
	if (is_last) {
		const float l = sqrtf(len2(bx, by, bz));
		D.ctl->last_energy_len = l;
		if ((double)l < 0.1) D.ctl->stopped = 1;                    // Main.cpp:1298
	}
	D.illum[e.id] += e.S0; D.illum[P + e.id] += e.S1; D.illum[2 * (size_t)P + e.id] += e.S2;
	bx -= e.S0; by -= e.S1; bz -= e.S2;
}

// MODE 0: single GPU — reference association B += d_0, += d_1, ...   (reads F of all k slots)
// MODE 1: multi GPU, local part — dB = sum of this rank's slots      (B untouched)
// MODE 2: multi GPU, final part — B += dB (after the all-reduce), emitter update
// "Is this patch an emitter of the batch" is answered in O(1): for every chunk of blockDim consecutive patches the block
// first drops the emitters that fall into the chunk into a shared-memory table (k / blockDim entries per thread), so
// the patch pass never compares against all k emitters.
template <int MODE>
__global__ void __launch_bounds__(256) apply_kernel(RadDev D, int fuse_select, int parity) {
	extern __shared__ float4 s_raw[];
	EmLite* s_em = reinterpret_cast<EmLite*>(s_raw);
	__shared__ int s_slot[256];            // first emitter slot of each patch of the current chunk (INT_MAX: none)
	__shared__ int s_last_h, s_dups; __shared__ uint32_t s_nvalid;
	const uint32_t P = D.P, k = D.k;
	pdl_enter();
	if (D.stop_gate && D.ctl->gate) return;     // the stop test fired in an earlier batch of this replay: the loop has ended (Main.cpp:1137)
	if (threadIdx.x == 0) { s_last_h = -1; s_dups = 0; s_nvalid = 0; }
	__syncthreads();
	for (uint32_t h = threadIdx.x; h < k; h += blockDim.x) {
		const float4 a = D.emlite[2 * h], b = D.emlite[2 * h + 1];        // packed by the camera kernel
		const uint32_t vo = __float_as_uint(a.w), id = __float_as_uint(b.w);
		EmLite l; l.S0 = a.x; l.S1 = a.y; l.S2 = a.z; l.valid = (vo && id < P) ? 1u : 0u; l.c0 = b.x; l.c1 = b.y; l.c2 = b.z; l.id = id;
		s_em[h] = l;
		if (MODE != 1 && l.valid) { atomicMax(&s_last_h, (int)(((vo >> 1) << 10) | h)); atomicAdd(&s_nvalid, 1u); }   // "last" = end of the list
	}
	__syncthreads();
	const int last_h = s_last_h < 0 ? -1 : (s_last_h & 1023);
	const float rho = D.reflectivity;
	unsigned long long best = 0;
	// fused exchange: the batch's sequence number lives in this rank's exchange buffer (device side, so that CUDA-graph
	// replays advance it); MODE 1 writes plane set (seq + 1) & 1 and its last block publishes seq + 1 to every peer,
	// MODE 2 waits until every rank has published the current seq and then sums the ranks' planes in rank order —
	// the same order on every GPU, so the replicas stay bit-identical.
	const bool fused = (MODE != 0) && D.xworld > 0;
	uint32_t xseq = 0;
	float* dB_out = D.dB;
	if (fused) {
		xseq = *reinterpret_cast<const volatile uint32_t*>(D.xb[D.xrank]) + 1u;     // the batch in flight (see xb_publish_now)
		if (MODE == 1) dB_out = xb_planes(D, D.xrank, xseq & 1u);
		if (MODE == 2) {
			if (!D.xtwo) xb_publish_now(D, 128, xseq);
			xb_wait(D, D.xtwo ? RAD_XB_FLAG2 : 128, xseq);
		}
	}
	for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < P; i0 += gridDim.x * blockDim.x) {
		const uint32_t i = i0 + threadIdx.x;
		if (MODE != 1) {
			__syncthreads();
			s_slot[threadIdx.x] = 0x7FFFFFFF;
			__syncthreads();
			for (uint32_t h = threadIdx.x; h < k; h += blockDim.x)
				if (s_em[h].valid && s_em[h].id - i0 < blockDim.x) { if (atomicMin(&s_slot[s_em[h].id - i0], (int)h) != 0x7FFFFFFF) s_dups = 1; }
			__syncthreads();
		}
		if (i >= P) continue;
		if (MODE == 1) {
			float dx = 0.0f, dy = 0.0f, dz = 0.0f;
			gather_transfer(D, s_em, D.h0, D.h1, P, i, rho, dx, dy, dz);
			const size_t pl = fused ? D.xPmax : P;
			dB_out[i] = dx; dB_out[pl + i] = dy; dB_out[2 * pl + i] = dz;
			continue;
		}
		float bx = D.rad[i], by = D.rad[P + i], bz = D.rad[2 * (size_t)P + i];
		if (MODE == 0) gather_transfer(D, s_em, 0, k, P, i, rho, bx, by, bz);
		else if (!fused) {                                    // (dB is left zeroed: the raster lanes of the next batch add into it)
			bx += D.dB[i]; by += D.dB[P + i]; bz += D.dB[2 * (size_t)P + i];
			D.dB[i] = 0.0f; D.dB[P + i] = 0.0f; D.dB[2 * (size_t)P + i] = 0.0f;
		} else if (D.xtwo) {                                  // two-shot: the slice owner has already summed the ranks' planes
			const float* red = xb_reduced(D, i / xb_slice(D));
			bx += __ldcg(red + i); by += __ldcg(red + D.xPmax + i); bz += __ldcg(red + 2 * (size_t)D.xPmax + i);
		} else {
			// peer loads over NVLink, L1 bypassed: ALL ranks' values are requested before the first one is used (a warp is
			// in-order: summing inside the load loop costs one NVLink round trip per rank, 44 us for 8 ranks at 16 k patches)
			float vx[RAD_MAX_PEERS], vy[RAD_MAX_PEERS], vz[RAD_MAX_PEERS];
			#pragma unroll
			for (uint32_t r = 0; r < RAD_MAX_PEERS; r++) {
				vx[r] = vy[r] = vz[r] = 0.0f;
				if (r < D.xworld) {
					const float* pl = xb_planes(D, r, xseq & 1u);
					vx[r] = __ldcg(pl + i); vy[r] = __ldcg(pl + D.xPmax + i); vz[r] = __ldcg(pl + 2 * (size_t)D.xPmax + i);
				}
			}
			float sx = 0.0f, sy = 0.0f, sz = 0.0f;
			#pragma unroll
			for (uint32_t r = 0; r < RAD_MAX_PEERS; r++) if (r < D.xworld) { sx += vx[r]; sy += vy[r]; sz += vz[r]; }   // rank order
			bx += sx; by += sy; bz += sz;
		}
		if (MODE == 2 && fused) {      // the plane set of the NEXT batch: every peer is done with it (it published this batch after its last update); the raster lanes add into it
			float* nx = xb_planes(D, D.xrank, (xseq + 1u) & 1u);
			nx[i] = 0.0f; nx[D.xPmax + i] = 0.0f; nx[2 * (size_t)D.xPmax + i] = 0.0f;
		}
		const int h = s_slot[threadIdx.x];
		if (h != 0x7FFFFFFF) {
			emitter_update(D, s_em[h], h == last_h, P, bx, by, bz);
			if (s_dups)                     // a patch listed twice (only rad_set_emitters can do that): all of them, in slot order
				for (uint32_t g = (uint32_t)h + 1; g < k; g++) {
					const EmLite o = s_em[g];
					if (o.valid && o.id == i) emitter_update(D, o, (int)g == last_h, P, bx, by, bz);
				}
		}
		D.rad[i] = bx; D.rad[P + i] = by; D.rad[2 * (size_t)P + i] = bz;
		if (fuse_select) best = max(best, energy_key_last(len2(bx, by, bz), i));
	}
	if (MODE == 1) return;                      // (published by the next kernel of the stream, see xb_publish_now)
	if (blockIdx.x == 0 && threadIdx.x == 0) { D.ctl->batches_done += 1; D.ctl->shots_done += s_nvalid; }
	// (the sequence number is bumped by xb_bump_kernel, the next launch: a "last block" ticket here costs every block a fence)
	if (fuse_select) {
		best = block_max(best);
		if (threadIdx.x == 0 && best) atomicMax(&D.ctl->selkey[parity ^ 1], best);
		// k == 1: the block that finishes last knows the next shooter — it takes the snapshot and builds the five face
		// matrices right here, so the next shot starts with its raster kernels (one launch less per shot)
		__shared__ bool s_lastblk; __shared__ RadEmitter s_e;
		__syncthreads();
		if (threadIdx.x == 0) { fence_acq_rel(); const uint32_t done = atomicAdd(&D.ctl->ticket, 1u); s_lastblk = done == gridDim.x - 1; if (s_lastblk) { D.ctl->ticket = 0; fence_acq_rel(); } }
		__syncthreads();
		if (s_lastblk) camera_block(D, 0, parity ^ 1, &s_e);
	}
}

// Two-shot exchange, first shot (large P): this rank sums the G ranks' dB planes over ITS slice of the patches, in rank
// order, into its `red` region, and publishes the second flag; the update kernel then reads every slice from its owner
// (2 (G-1)/G P values cross NVLink per rank instead of (G-1) P).
__global__ void __launch_bounds__(256) xreduce_kernel(RadDev D) {
	if (D.stop_gate && D.ctl->gate) return;
	const uint32_t xseq = *reinterpret_cast<const volatile uint32_t*>(D.xb[D.xrank]) + 1u;
	xb_publish_now(D, 128, xseq);
	xb_wait(D, 128, xseq);
	const uint32_t sz = xb_slice(D), lo = D.xrank * sz, hi = min(D.P, lo + sz);
	float* red = xb_reduced(D, D.xrank);
	for (uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
		float vx[RAD_MAX_PEERS], vy[RAD_MAX_PEERS], vz[RAD_MAX_PEERS];      // all requests first, then the sum in rank order
		#pragma unroll
		for (uint32_t r = 0; r < RAD_MAX_PEERS; r++) {
			vx[r] = vy[r] = vz[r] = 0.0f;
			if (r < D.xworld) {
				const float* pl = xb_planes(D, r, xseq & 1u);
				vx[r] = __ldcg(pl + i); vy[r] = __ldcg(pl + D.xPmax + i); vz[r] = __ldcg(pl + 2 * (size_t)D.xPmax + i);
			}
		}
		float sx = 0.0f, sy = 0.0f, sz3 = 0.0f;
		#pragma unroll
		for (uint32_t r = 0; r < RAD_MAX_PEERS; r++) if (r < D.xworld) { sx += vx[r]; sy += vy[r]; sz3 += vz[r]; }
		red[i] = sx; red[D.xPmax + i] = sy; red[2 * (size_t)D.xPmax + i] = sz3;
	}
	xb_publish_last(D, RAD_XB_FLAG2, xseq, false);
}

// the batch is exchanged: bump the sequence number (a launch of its own: every block of the update kernel reads it)
__global__ void xb_bump_kernel(RadDev D) {
	if (D.stop_gate && D.ctl->gate) return;
	if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(D.xb[D.xrank]) += 1u;
}

// Multi-GPU, raster-lane form of the local dB: the slots [D.h0, D.h1) of ONE raster lane, right behind the lane's
// ProcessHemicube on the lane's stream, so that the transfer of a finished lane overlaps the other lanes' rasterisation.
// The lanes of a rank add into the same planes (red.global.add.f32; the planes start out zeroed, see apply_kernel<2>): the
// sum's association differs from run to run like F's own, every replica reads the same published values.
__global__ void __launch_bounds__(256) lane_delta_kernel(RadDev D) {
	__shared__ EmLite s_loc[RAD_RING_SLOTS];
	if (D.stop_gate && D.ctl->gate) return;
	const uint32_t P = D.P, n = D.h1 - D.h0;
	for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
		const uint32_t h = D.h0 + j;
		const float4 a = D.emlite[2 * h], b = D.emlite[2 * h + 1];
		const uint32_t vo = __float_as_uint(a.w), id = __float_as_uint(b.w);
		EmLite l; l.S0 = a.x; l.S1 = a.y; l.S2 = a.z; l.valid = (vo && id < P) ? 1u : 0u; l.c0 = b.x; l.c1 = b.y; l.c2 = b.z; l.id = id;
		s_loc[j] = l;
	}
	__syncthreads();
	const bool fused = D.xworld > 0;
	float* out = fused ? xb_planes(D, D.xrank, (*reinterpret_cast<const volatile uint32_t*>(D.xb[D.xrank]) + 1u) & 1u) : D.dB;
	const size_t pl = fused ? D.xPmax : P;
	const float rho = D.reflectivity;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
		float dx = 0.0f, dy = 0.0f, dz = 0.0f;
		gather_transfer(D, s_loc, D.h0, D.h1, P, i, rho, dx, dy, dz, D.h0);
		if (dx != 0.0f) atomicAdd(out + i, dx);
		if (dy != 0.0f) atomicAdd(out + pl + i, dy);
		if (dz != 0.0f) atomicAdd(out + 2 * pl + i, dz);
	}
}

// ---- speculative strict progressive refinement (k == 1), the sequential part -------------------------------------------
// The hemicubes of the `nslots` strongest patches (D.em[0 .. nslots), chosen by the top-k kernels with the argmax's own tie
// rule) have been rendered and processed: F[s] is complete for every slot.  This kernel replays the reference's loop over
// them, one shot at a time (Main.cpp:1137-1309 with HEMICUBES_CNT = 1): shooter = argmax of the CURRENT |B|^2 (largest energy,
// last index among equals, nothing to do when everything is dark), S = B of that patch at this moment, B_i += ((S F[i]) rho) c,
// emitter update, stop test — the arithmetic of apply_kernel<0>, shot by shot.  The radiosities stay in registers (PPT patches
// per thread, the grid covers all patches); per shot the blocks agree on the argmax through one 64-bit atomicMax and ONE grid
// barrier (three keys in rotation: the key of shot s + 1 is reset before the barrier of shot s, read last in shot s - 2), and
// every block publishes the B of its own candidate next to its key, so the winner's S needs no second barrier.  A shooter that
// is not among the rendered slots ends the batch (the caller selects the next batch from the state reached); the first shot
// of a batch always hits, its shooter is slot 0 by construction.
template <int PPT>
__global__ void __launch_bounds__(1024, 1) spec_apply_kernel(RadDev D, uint32_t slot_base, uint32_t nslots, int stop_armed) {
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	__shared__ uint32_t s_id[RAD_SPEC_SLOTS], s_used[RAD_SPEC_SLOTS];
	__shared__ unsigned long long s_key;
	__shared__ int s_slot;
	if (D.ctl->gate) return;                    // (grid-uniform: latched by the camera kernel of this batch)
	const uint32_t P = D.P, tid = threadIdx.x, stride = gridDim.x * 1024u, i0 = blockIdx.x * 1024u + tid;
	if (tid < RAD_SPEC_SLOTS) { s_id[tid] = (tid < nslots && D.em[slot_base + tid].valid) ? D.em[slot_base + tid].id : 0xFFFFFFFFu; s_used[tid] = 0u; }
	float bx[PPT], by[PPT], bz[PPT];
	#pragma unroll
	for (int j = 0; j < PPT; j++) {
		const uint32_t i = i0 + (uint32_t)j * stride;
		bx[j] = by[j] = bz[j] = 0.0f;
		if (i < P) { bx[j] = D.rad[i]; by[j] = D.rad[P + i]; bz[j] = D.rad[2 * (size_t)P + i]; }
	}
	const uint32_t target = D.ctl->spec_target, done0 = D.ctl->shots_done;
	const float rho = D.reflectivity;
	uint32_t count = 0; int ended = 0, missed = 0;      // ended: the call is over (target reached, stop test, everything dark)
	__syncthreads();
	for (uint32_t shot = 0; done0 + count < target; shot++) {
		unsigned long long mine = 0ull;
		#pragma unroll
		for (int j = 0; j < PPT; j++) {
			const uint32_t i = i0 + (uint32_t)j * stride;
			if (i < P) { const unsigned long long kk = energy_key_last(len2(bx[j], by[j], bz[j]), i); mine = kk > mine ? kk : mine; }
		}
		const unsigned long long blockbest = block_max(mine);
		const uint32_t rk = shot % 3u;
		if (tid == 0) {
			s_key = blockbest;
			if (blockbest) atomicMax(&D.ctl->spec_key[rk], blockbest);
			if (blockIdx.x == 0) D.ctl->spec_key[(shot + 1u) % 3u] = 0ull;
		}
		__syncthreads();
		if (mine != 0ull && mine == s_key) {        // this thread owns the block's candidate: its B rides along with the key
			#pragma unroll
			for (int j = 0; j < PPT; j++) {
				const uint32_t i = i0 + (uint32_t)j * stride;
				if (i == (uint32_t)(mine & 0xFFFFFFFFull)) D.spec_cand[rk * 256u + blockIdx.x] = make_float4(bx[j], by[j], bz[j], 0.0f);
			}
		}
		grid.sync();                                // (orders the candidates and the key for every thread of the grid; no fence of our own: a fence in every thread serialises)
		const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(&D.ctl->spec_key[rk]);
		if (key == 0ull) {                          // everything is dark: the remaining shots shoot patch 0 with S = 0 (Main.cpp:1161: no-ops that still count)
			count = target - done0; ended = 1;
			if (blockIdx.x == 0 && tid == 0) { D.ctl->last_energy_len = 0.0f; D.ctl->stopped = 1; }
			break;
		}
		const uint32_t shooter = (uint32_t)(key & 0xFFFFFFFFull);
		if (tid == 0) s_slot = -1;
		__syncthreads();
		if (tid < nslots && s_id[tid] == shooter && !s_used[tid]) s_slot = (int)tid;
		__syncthreads();
		const int slot = s_slot;
		if (slot < 0) { missed = 1; break; }        // not rendered ahead: the batch ends here
		const float4 Sv = __ldcg(D.spec_cand + rk * 256u + (shooter >> 10) % gridDim.x);     // the owner block's candidate IS the winner
		const float c0 = __ldg(D.color + shooter), c1 = __ldg(D.color + P + shooter), c2 = __ldg(D.color + 2 * (size_t)P + shooter);
		const float* __restrict__ F = D.F + (size_t)(slot_base + (uint32_t)slot) * P;
		#pragma unroll
		for (int j = 0; j < PPT; j++) {
			const uint32_t i = i0 + (uint32_t)j * stride;
			if (i < P) {
				const float f = __ldcg(F + i);
				bx[j] += ((Sv.x * f) * rho) * c0; by[j] += ((Sv.y * f) * rho) * c1; bz[j] += ((Sv.z * f) * rho) * c2;
				if (i == shooter) {                  // emitter update (Main.cpp:1286-1295): lastEnergy before the subtraction
					D.illum[i] += Sv.x; D.illum[P + i] += Sv.y; D.illum[2 * (size_t)P + i] += Sv.z;
					bx[j] -= Sv.x; by[j] -= Sv.y; bz[j] -= Sv.z;
				}
			}
		}
		// the stop test on the emitter's energy after the transfer, before the subtraction — every block evaluates the same floats
		const float fs = __ldcg(F + shooter);
		const float ex = Sv.x + ((Sv.x * fs) * rho) * c0, ey = Sv.y + ((Sv.y * fs) * rho) * c1, ez = Sv.z + ((Sv.z * fs) * rho) * c2;
		const float l = sqrtf(len2(ex, ey, ez));
		const bool stop_now = (double)l < 0.1;       // Main.cpp:1298
		if (blockIdx.x == 0 && tid == 0) { D.ctl->last_energy_len = l; if (stop_now) D.ctl->stopped = 1; }
		if (tid == 0) s_used[slot] = 1u;
		count++;
		if (stop_armed && stop_now) { ended = 1; break; }
		__syncthreads();
	}
	#pragma unroll
	for (int j = 0; j < PPT; j++) {
		const uint32_t i = i0 + (uint32_t)j * stride;
		if (i < P) { D.rad[i] = bx[j]; D.rad[P + i] = by[j]; D.rad[2 * (size_t)P + i] = bz[j]; }
	}
	if (blockIdx.x == 0 && tid == 0) {
		D.ctl->shots_done = done0 + count; D.ctl->batches_done += count;
		D.ctl->spec_hits += count; D.ctl->spec_misses += (uint32_t)missed;
		if (ended || done0 + count >= target) D.ctl->spec_done = 1;
	}
}

// Small scenes (P <= 148 x 512): one patch per thread, blocks of 512 threads, and everything a shot needs staged ON CHIP before
// the loop starts — the F entry of the thread's patch for every rendered slot (64 x 512 floats of shared memory), the slots'
// emitter colours and self form factors — so that a shot's dependent chain is block reduction -> barrier -> slot look-up ->
// update, with no L2 round trip behind the barrier.  Otherwise spec_apply_kernel<1>.
__global__ void __launch_bounds__(512, 1) spec_apply_smem_kernel(RadDev D, uint32_t slot_base, uint32_t nslots, int stop_armed) {
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	extern __shared__ float sF[];                 // [nslots][512]
	__shared__ uint32_t s_id[RAD_SPEC_SLOTS], s_used[RAD_SPEC_SLOTS];
	__shared__ float s_col[RAD_SPEC_SLOTS][3], s_fself[RAD_SPEC_SLOTS];
	__shared__ unsigned long long s_key;
	__shared__ int s_slot;
	if (D.ctl->gate) return;
	const uint32_t P = D.P, tid = threadIdx.x, i = blockIdx.x * 512u + tid;
	const bool live = i < P;
	if (tid < RAD_SPEC_SLOTS) {
		const bool v = tid < nslots && D.em[slot_base + tid].valid;
		const uint32_t id = v ? D.em[slot_base + tid].id : 0xFFFFFFFFu;
		s_id[tid] = id; s_used[tid] = 0u;
		s_col[tid][0] = v ? __ldg(D.color + id) : 0.0f; s_col[tid][1] = v ? __ldg(D.color + P + id) : 0.0f; s_col[tid][2] = v ? __ldg(D.color + 2 * (size_t)P + id) : 0.0f;
		s_fself[tid] = v ? __ldcg(D.F + (size_t)(slot_base + tid) * P + id) : 0.0f;
	}
	for (uint32_t sl = 0; sl < nslots; sl++) sF[sl * 512u + tid] = live ? __ldcg(D.F + (size_t)(slot_base + sl) * P + i) : 0.0f;
	float bx = 0.0f, by = 0.0f, bz = 0.0f;
	if (live) { bx = D.rad[i]; by = D.rad[P + i]; bz = D.rad[2 * (size_t)P + i]; }
	const uint32_t target = D.ctl->spec_target, done0 = D.ctl->shots_done;
	const float rho = D.reflectivity;
	uint32_t count = 0; int ended = 0, missed = 0;
	__syncthreads();
	for (uint32_t shot = 0; done0 + count < target; shot++) {
		const unsigned long long mine = live ? energy_key_last(len2(bx, by, bz), i) : 0ull;
		const unsigned long long blockbest = block_max(mine);
		const uint32_t rk = shot % 3u;
		if (tid == 0) {
			s_key = blockbest;
			if (blockbest) atomicMax(&D.ctl->spec_key[rk], blockbest);
			if (blockIdx.x == 0) D.ctl->spec_key[(shot + 1u) % 3u] = 0ull;
		}
		__syncthreads();
		if (mine != 0ull && mine == s_key) D.spec_cand[rk * 256u + blockIdx.x] = make_float4(bx, by, bz, 0.0f);
		grid.sync();
		const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(&D.ctl->spec_key[rk]);
		if (key == 0ull) {
			count = target - done0; ended = 1;
			if (blockIdx.x == 0 && tid == 0) { D.ctl->last_energy_len = 0.0f; D.ctl->stopped = 1; }
			break;
		}
		const uint32_t shooter = (uint32_t)(key & 0xFFFFFFFFull);
		if (tid == 0) s_slot = -1;
		__syncthreads();
		if (tid < nslots && s_id[tid] == shooter && !s_used[tid]) s_slot = (int)tid;
		__syncthreads();
		const int slot = s_slot;
		if (slot < 0) { missed = 1; break; }
		const float4 Sv = __ldcg(D.spec_cand + rk * 256u + (shooter >> 9));      // the owner block's candidate IS the winner
		const float c0 = s_col[slot][0], c1 = s_col[slot][1], c2 = s_col[slot][2];
		if (live) {
			const float f = sF[(uint32_t)slot * 512u + tid];
			bx += ((Sv.x * f) * rho) * c0; by += ((Sv.y * f) * rho) * c1; bz += ((Sv.z * f) * rho) * c2;
			if (i == shooter) {
				D.illum[i] += Sv.x; D.illum[P + i] += Sv.y; D.illum[2 * (size_t)P + i] += Sv.z;
				bx -= Sv.x; by -= Sv.y; bz -= Sv.z;
			}
		}
		const float fs = s_fself[slot];
		const float ex = Sv.x + ((Sv.x * fs) * rho) * c0, ey = Sv.y + ((Sv.y * fs) * rho) * c1, ez = Sv.z + ((Sv.z * fs) * rho) * c2;
		const float l = sqrtf(len2(ex, ey, ez));
		const bool stop_now = (double)l < 0.1;
		if (blockIdx.x == 0 && tid == 0) { D.ctl->last_energy_len = l; if (stop_now) D.ctl->stopped = 1; }
		if (tid == 0) s_used[slot] = 1u;
		count++;
		if (stop_armed && stop_now) { ended = 1; break; }
		__syncthreads();
	}
	if (live) { D.rad[i] = bx; D.rad[P + i] = by; D.rad[2 * (size_t)P + i] = bz; }
	if (blockIdx.x == 0 && tid == 0) {
		D.ctl->shots_done = done0 + count; D.ctl->batches_done += count;
		D.ctl->spec_hits += count; D.ctl->spec_misses += (uint32_t)missed;
		if (ended || done0 + count >= target) D.ctl->spec_done = 1;
	}
}

// ---- the sequential part without a barrier per shot: simulate, verify, commit ------------------------------------------
// While the shooters come from the rendered-ahead set, the strict loop's decisions depend on the set's OWN radiosities only:
// member m receives ((S F[s][id_m]) rho) c from the shot of slot s, a 64 x 64 matrix.  spec_sim_kernel replays the loop over
// the members alone — one block, no grid barrier — and records every shot (slot, S, colour, the shooter's argmax key, the stop
// test's lastEnergy); it stops at the call's end, at the stop test, or when the strongest member is one whose hemicube is
// already spent.  spec_verify_kernel then walks every patch through the recorded shots with the loop's own arithmetic and
// finds the first shot before which a patch OUTSIDE the simulation would have been the argmax (key > the recorded key:
// largest energy, last index among equals) — up to there the simulation IS the strict loop — and spec_commit_kernel applies
// exactly those shots.  Same shots, same floats as one shot at a time; 3 launches instead of a grid barrier per shot.
__global__ void __launch_bounds__(64) spec_sim_kernel(RadDev D, uint32_t slot_base, uint32_t nslots, int stop_armed) {
	__shared__ float G[RAD_SPEC_SLOTS][RAD_SPEC_SLOTS + 1];      // G[s][m] = F of slot s at member m's patch
	__shared__ unsigned long long s_k[2];
	__shared__ int s_w, s_stop;
	__shared__ float s_S[3], s_c[3];
	if (D.ctl->gate) return;
	const uint32_t P = D.P, m = threadIdx.x;
	const bool valid = m < nslots && D.em[slot_base + m].valid != 0 && D.em[slot_base + m].id < P;
	const uint32_t id = valid ? D.em[slot_base + m].id : 0u;
	float bx = 0.0f, by = 0.0f, bz = 0.0f, c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
	if (valid) {
		bx = D.rad[id]; by = D.rad[P + id]; bz = D.rad[2 * (size_t)P + id];
		c0 = D.color[id]; c1 = D.color[P + id]; c2 = D.color[2 * (size_t)P + id];
	}
	for (uint32_t sl = 0; sl < RAD_SPEC_SLOTS; sl++) G[sl][m] = (valid && sl < nslots) ? __ldcg(D.F + (size_t)(slot_base + sl) * P + id) : 0.0f;
	const uint32_t target = D.ctl->spec_target, done0 = D.ctl->shots_done;
	const uint32_t R = target > done0 ? target - done0 : 0u;
	const float rho = D.reflectivity;
	if (m == 0) s_stop = 0;
	const int nvalid = __syncthreads_count(valid);
	if (nvalid == 0) {                              // the set was the strongest of ALL patches: nothing carries energy
		if (m == 0) { D.ctl->spec_count_sim = 0u; D.ctl->spec_jstar = 0u; D.ctl->spec_alldark = 1u; }
		return;
	}
	bool used = false;
	uint32_t j = 0;
	while (j < R && j < RAD_SPEC_SLOTS) {
		const unsigned long long key = valid ? energy_key_last(len2(bx, by, bz), id) : 0ull;
		const unsigned long long kw = warp_max(key);
		if ((m & 31u) == 0u) s_k[m >> 5] = kw;
		if (m == 0) s_w = -1;
		__syncthreads();
		const unsigned long long K = s_k[0] > s_k[1] ? s_k[0] : s_k[1];
		if (K == 0ull) break;                       // the members have nothing left: the next batch decides
		if (key == K && !used) { s_w = (int)m; s_S[0] = bx; s_S[1] = by; s_S[2] = bz; s_c[0] = c0; s_c[1] = c1; s_c[2] = c2; }
		__syncthreads();
		const int w = s_w;
		if (w < 0) break;                           // the strongest member's hemicube is spent
		const float S0 = s_S[0], S1 = s_S[1], S2 = s_S[2], w0 = s_c[0], w1 = s_c[1], w2 = s_c[2];
		const float f = G[w][m];
		if (valid) { bx += ((S0 * f) * rho) * w0; by += ((S1 * f) * rho) * w1; bz += ((S2 * f) * rho) * w2; }
		if ((int)m == w) {
			const float l = sqrtf(len2(bx, by, bz));   // lastEnergy: after the transfer, before the subtraction (Main.cpp:1292)
			RadSpecStep st; st.key = K; st.slot = m; st.id = id; st.S[0] = S0; st.S[1] = S1; st.S[2] = S2; st.c[0] = w0; st.c[1] = w1; st.c[2] = w2; st.len = l; st.pad = 0.0f;
			D.spec_steps[j] = st;
			bx -= S0; by -= S1; bz -= S2;
			used = true;
			if ((double)l < 0.1) s_stop = 1;
		}
		j++;
		__syncthreads();
		if (stop_armed && s_stop) break;
	}
	if (m == 0) { D.ctl->spec_count_sim = j; D.ctl->spec_jstar = j; D.ctl->spec_alldark = 0u; }
}

// the patch's radiosity after the first n recorded shots, with the strict loop's own arithmetic; verify: the first shot
// before which this patch would have been the argmax instead (atomicMin into spec_jstar)
template <bool VERIFY>
__global__ void __launch_bounds__(256) spec_walk_kernel(RadDev D, uint32_t slot_base, int stop_armed) {
	__shared__ RadSpecStep s_st[RAD_SPEC_SLOTS];
	if (D.ctl->gate) return;
	const uint32_t P = D.P;
	const uint32_t nsim = D.ctl->spec_count_sim;
	const uint32_t n = VERIFY ? nsim : min(nsim, D.ctl->spec_jstar);
	if (!VERIFY && blockIdx.x == 0 && threadIdx.x == 0) {       // the batch's bookkeeping
		const uint32_t target = D.ctl->spec_target, done0 = D.ctl->shots_done;
		if (D.ctl->spec_alldark) {                  // the remaining shots are no-op shots of patch 0 that still count (Main.cpp:1161 with S = 0)
			D.ctl->batches_done += target - done0; D.ctl->shots_done = target;
			D.ctl->last_energy_len = 0.0f; D.ctl->stopped = 1; D.ctl->spec_done = 1;
		} else {
			int low = 0;
			for (uint32_t j = 0; j < n; j++) low |= (double)D.spec_steps[j].len < 0.1;
			if (n) D.ctl->last_energy_len = D.spec_steps[n - 1].len;
			if (low) D.ctl->stopped = 1;
			D.ctl->shots_done = done0 + n; D.ctl->batches_done += n;
			D.ctl->spec_hits += n; D.ctl->spec_misses += (n < target - done0) ? 1u : 0u;
			if (done0 + n >= target || (stop_armed && low)) D.ctl->spec_done = 1;
		}
	}
	if (n == 0) return;
	for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) s_st[t] = D.spec_steps[t];
	__syncthreads();
	const float rho = D.reflectivity;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
		float bx = D.rad[i], by = D.rad[P + i], bz = D.rad[2 * (size_t)P + i];
		for (uint32_t j = 0; j < n; j++) {
			const RadSpecStep& st = s_st[j];
			if (VERIFY) {
				if (energy_key_last(len2(bx, by, bz), i) > st.key) { atomicMin(&D.ctl->spec_jstar, j); break; }
			}
			const float f = __ldcg(D.F + (size_t)(slot_base + st.slot) * P + i);
			bx += ((st.S[0] * f) * rho) * st.c[0]; by += ((st.S[1] * f) * rho) * st.c[1]; bz += ((st.S[2] * f) * rho) * st.c[2];
			if (i == st.id) {
				if (!VERIFY) { D.illum[i] += st.S[0]; D.illum[P + i] += st.S[1]; D.illum[2 * (size_t)P + i] += st.S[2]; }
				bx -= st.S[0]; by -= st.S[1]; bz -= st.S[2];
			}
		}
		if (!VERIFY) { D.rad[i] = bx; D.rad[P + i] = by; D.rad[2 * (size_t)P + i] = bz; }
	}
}

// Cluster form of the same loop for scenes that fit ONE thread-block cluster (P <= 16 x 1024 x PPT patches): the 16 CTAs
// exchange their argmax candidates (key + B of that patch) through distributed shared memory — every CTA stores its candidate
// into all 16 CTAs' tables, double-buffered by shot parity — and meet at the hardware cluster barrier instead of a grid
// barrier through global memory.  Everything else is spec_apply_kernel.
template <int PPT>
__global__ void __launch_bounds__(1024, 1) spec_apply_cluster_kernel(RadDev D, uint32_t slot_base, uint32_t nslots, int stop_armed) {
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
	__shared__ uint32_t s_id[RAD_SPEC_SLOTS], s_used[RAD_SPEC_SLOTS];
	__shared__ unsigned long long s_key;
	__shared__ float4 s_cand;
	__shared__ int s_slot;
	__shared__ unsigned long long c_key[2][16];
	__shared__ float4 c_S[2][16];
	if (D.ctl->gate) return;                    // (uniform over the cluster)
	const uint32_t P = D.P, tid = threadIdx.x, NB = cluster.num_blocks(), rank = cluster.block_rank(), stride = NB * 1024u, i0 = rank * 1024u + tid;
	if (tid < RAD_SPEC_SLOTS) { s_id[tid] = (tid < nslots && D.em[slot_base + tid].valid) ? D.em[slot_base + tid].id : 0xFFFFFFFFu; s_used[tid] = 0u; }
	float bx[PPT], by[PPT], bz[PPT];
	#pragma unroll
	for (int j = 0; j < PPT; j++) {
		const uint32_t i = i0 + (uint32_t)j * stride;
		bx[j] = by[j] = bz[j] = 0.0f;
		if (i < P) { bx[j] = D.rad[i]; by[j] = D.rad[P + i]; bz[j] = D.rad[2 * (size_t)P + i]; }
	}
	const uint32_t target = D.ctl->spec_target, done0 = D.ctl->shots_done;
	const float rho = D.reflectivity;
	uint32_t count = 0; int ended = 0, missed = 0;
	cluster.sync();                              // every CTA of the cluster is running: its shared memory may be written
	for (uint32_t shot = 0; done0 + count < target; shot++) {
		unsigned long long mine = 0ull;
		#pragma unroll
		for (int j = 0; j < PPT; j++) {
			const uint32_t i = i0 + (uint32_t)j * stride;
			if (i < P) { const unsigned long long kk = energy_key_last(len2(bx[j], by[j], bz[j]), i); mine = kk > mine ? kk : mine; }
		}
		const unsigned long long blockbest = block_max(mine);
		const uint32_t par = shot & 1u;
		if (tid == 0) s_key = blockbest;
		__syncthreads();
		if (mine != 0ull && mine == s_key) {
			#pragma unroll
			for (int j = 0; j < PPT; j++) {
				const uint32_t i = i0 + (uint32_t)j * stride;
				if (i == (uint32_t)(mine & 0xFFFFFFFFull)) s_cand = make_float4(bx[j], by[j], bz[j], 0.0f);
			}
		}
		__syncthreads();
		if (tid < NB) {                              // this CTA's candidate into CTA tid's table
			*cluster.map_shared_rank(&c_key[par][rank], tid) = s_key;
			*cluster.map_shared_rank(&c_S[par][rank], tid) = s_cand;
		}
		cluster.sync();
		unsigned long long key = 0ull; uint32_t owner = 0;
		#pragma unroll
		for (uint32_t r = 0; r < 16u; r++) if (r < NB) { const unsigned long long kk = c_key[par][r]; if (kk > key) { key = kk; owner = r; } }
		if (key == 0ull) {
			count = target - done0; ended = 1;
			if (rank == 0 && tid == 0) { D.ctl->last_energy_len = 0.0f; D.ctl->stopped = 1; }
			break;
		}
		const uint32_t shooter = (uint32_t)(key & 0xFFFFFFFFull);
		if (tid == 0) s_slot = -1;
		__syncthreads();
		if (tid < nslots && s_id[tid] == shooter && !s_used[tid]) s_slot = (int)tid;
		__syncthreads();
		const int slot = s_slot;
		if (slot < 0) { missed = 1; break; }
		const float4 Sv = c_S[par][owner];
		const float c0 = __ldg(D.color + shooter), c1 = __ldg(D.color + P + shooter), c2 = __ldg(D.color + 2 * (size_t)P + shooter);
		const float* __restrict__ F = D.F + (size_t)(slot_base + (uint32_t)slot) * P;
		#pragma unroll
		for (int j = 0; j < PPT; j++) {
			const uint32_t i = i0 + (uint32_t)j * stride;
			if (i < P) {
				const float f = __ldcg(F + i);
				bx[j] += ((Sv.x * f) * rho) * c0; by[j] += ((Sv.y * f) * rho) * c1; bz[j] += ((Sv.z * f) * rho) * c2;
				if (i == shooter) {
					D.illum[i] += Sv.x; D.illum[P + i] += Sv.y; D.illum[2 * (size_t)P + i] += Sv.z;
					bx[j] -= Sv.x; by[j] -= Sv.y; bz[j] -= Sv.z;
				}
			}
		}
		const float fs = __ldcg(F + shooter);
		const float ex = Sv.x + ((Sv.x * fs) * rho) * c0, ey = Sv.y + ((Sv.y * fs) * rho) * c1, ez = Sv.z + ((Sv.z * fs) * rho) * c2;
		const float l = sqrtf(len2(ex, ey, ez));
		const bool stop_now = (double)l < 0.1;
		if (rank == 0 && tid == 0) { D.ctl->last_energy_len = l; if (stop_now) D.ctl->stopped = 1; }
		if (tid == 0) s_used[slot] = 1u;
		count++;
		if (stop_armed && stop_now) { ended = 1; break; }
		__syncthreads();
	}
	#pragma unroll
	for (int j = 0; j < PPT; j++) {
		const uint32_t i = i0 + (uint32_t)j * stride;
		if (i < P) { D.rad[i] = bx[j]; D.rad[P + i] = by[j]; D.rad[2 * (size_t)P + i] = bz[j]; }
	}
	if (rank == 0 && tid == 0) {
		D.ctl->shots_done = done0 + count; D.ctl->batches_done += count;
		D.ctl->spec_hits += count; D.ctl->spec_misses += (uint32_t)missed;
		if (ended || done0 + count >= target) D.ctl->spec_done = 1;
	}
	cluster.sync();                              // nobody leaves while a peer may still store into its tables
}

} // namespace

static uint32_t patch_grid(uint32_t P, uint32_t threads) {
	uint32_t b = (P + threads - 1) / threads;
	const uint32_t cap = 148 * 8;
	return b > cap ? cap : (b ? b : 1);
}

void rad_launch_argmax(rad_ctx* c) {
	cudaMemsetAsync(&c->d.ctl->selkey[c->parity], 0, sizeof(unsigned long long), c->stream);
	argmax_kernel<<<patch_grid(c->d.P, 256), 256, 0, c->stream>>>(c->d, (int)c->parity);
	c->launches++;
	c->selkey_valid = true; c->cam_valid = false;
}

void rad_launch_select(rad_ctx* c) {
	const RadDev& D = c->d;
	if (D.k == 1) {
		if (!c->selkey_valid) { rad_launch_argmax(c); c->cam_valid = false; }
		if (!c->cam_valid) rad_launch_camera(c, (int)c->parity);      // decodes the fused argmax key, then snapshot + MVPs
		return;                                                       // (otherwise the previous update's tail already did)
	} else {
		const bool ref = c->cfg.select_mode == RAD_SELECT_REFERENCE && c->select_override == 0;
		const int mode = c->select_override ? c->select_override : (ref ? 1 : 0);
		TopkSpec sp = { D.k, 0u, 0u, 0u };
		if (mode == 2) { sp.count = c->sel_count ? c->sel_count : D.k; sp.slot_base = c->sel_base; sp.excl_base = c->sel_excl; sp.excl_n = c->sel_excl_n; }
		int keep = 64;
		while ((uint32_t)keep < (mode == 2 ? sp.count : D.k) + (ref ? 1u : 0u)) keep <<= 1;             // reference list: one entry beyond its end (tie check)
		uint32_t n = D.P;
		const unsigned long long* in = nullptr;
		unsigned long long* out = D.cand0;
		bool first = true;
		for (;;) {
			const uint32_t nb = (n + kTopChunk - 1) / kTopChunk;
			const int fin = nb * (uint32_t)keep <= (uint32_t)kTopChunk;       // the level's last block can finish the selection
			if (first) topk_level_kernel<true><<<nb, kTopThreads, 0, c->stream>>>(D, in, n, out, keep, fin, mode, sp);
			else topk_level_kernel<false><<<nb, kTopThreads, 0, c->stream>>>(D, in, n, out, keep, fin, mode, sp);
			c->launches++;
			if (fin) break;
			n = nb * keep; in = out; out = out == D.cand0 ? D.cand1 : D.cand0; first = false;
		}
		if (ref) {                                                            // exact emulation of the list; a no-op after the fast path
			select_reference_kernel<<<1, 1024, 0, c->stream>>>(D);
			c->launches++;
		}
	}
	rad_launch_camera(c);
}

void rad_launch_set_emitters(rad_ctx* c, const uint32_t* d_ids, uint32_t n) {
	set_emitters_kernel<<<1, 64, 0, c->stream>>>(c->d, d_ids, n);
	c->launches++;
	rad_launch_camera(c);
}

// small scenes: 64-thread blocks so that the patches spread over all SMs
static uint32_t apply_threads(uint32_t P) { return P <= 148u * 8u * 64u ? 64u : 256u; }

void rad_launch_apply(rad_ctx* c, bool fuse_select) {
	const RadDev& D = c->d;
	const uint32_t T = apply_threads(D.P);
	rad_launch_pdl(c->pdl, apply_kernel<0>, dim3(patch_grid(D.P, T)), dim3(T), D.k * sizeof(EmLite), c->stream, D, fuse_select ? 1 : 0, (int)c->parity);
	c->launches++;
}
void rad_launch_delta(rad_ctx* c) {
	const RadDev& D = c->d;
	const uint32_t T = apply_threads(D.P);
	apply_kernel<1><<<patch_grid(D.P, T), T, D.k * sizeof(EmLite), c->stream>>>(D, 0, 0);
	c->launches++;
}
void rad_launch_lane_delta(rad_ctx* c, const RadDev& V, cudaStream_t st, uint32_t s0, uint32_t n) {
	RadDev D = V;
	D.h0 = V.h0 + s0; D.h1 = D.h0 + n;
	const uint32_t T = apply_threads(D.P);
	lane_delta_kernel<<<patch_grid(D.P, T), T, 0, st>>>(D);
	c->launches++;
	c->lane_delta_done = true;
}
// the grid covers every patch: blocks of 1024 threads, one per SM at most (co-resident: the kernel has a grid barrier per shot)
int rad_launch_spec_apply(rad_ctx* c, const RadDev& S, uint32_t slot_base, uint32_t nslots, int stop_armed) {
	int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->cfg.device);
	uint32_t nb = (S.P + 1023u) / 1024u;
	if (nb > (uint32_t)nsm) nb = (uint32_t)nsm;
	if (nb > 256u) nb = 256u;
	if (nb < 1u) nb = 1u;
	uint32_t ppt = (S.P + nb * 1024u - 1u) / (nb * 1024u);
	// (fewer, fuller blocks do not pay: the per-thread work of a shot, not the grid barrier, is what a shot waits for)
	static const uint32_t min_ppt = [] { const char* e = getenv("RAD_SPEC_PPT"); const int v = e ? atoi(e) : 1; return (uint32_t)(v < 1 ? 1 : (v > 16 ? 16 : v)); }();   // tuning knob (measured: 1 is best — 86 k shots/s against 76 k at 4 and 54 k at 16 on config 2)
	if (ppt < min_ppt) { ppt = min_ppt; nb = (S.P + ppt * 1024u - 1u) / (ppt * 1024u); if (nb < 1u) nb = 1u; }
	// one thread-block cluster of 16 CTAs when the scene fits it (hardware cluster barrier + distributed shared memory per shot)
	// (opt-in, RAD_SPEC_CLUSTER=1: parity-green and no faster, 88.1 k against 89.0 k shots/s on config 2 — a shot is not waiting for its barrier)
	static const bool use_cluster = [] { const char* e = getenv("RAD_SPEC_CLUSTER"); return e && atoi(e) != 0; }();
	if (use_cluster && S.P <= 16u * 1024u * 4u) {
		const uint32_t cppt = (S.P + 16u * 1024u - 1u) / (16u * 1024u);
		const void* cfn = cppt <= 1 ? (const void*)spec_apply_cluster_kernel<1> : (cppt <= 2 ? (const void*)spec_apply_cluster_kernel<2> : (const void*)spec_apply_cluster_kernel<4>);
		static bool attr_done = false;
		if (!attr_done) {
			cudaFuncSetAttribute(spec_apply_cluster_kernel<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
			cudaFuncSetAttribute(spec_apply_cluster_kernel<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
			cudaFuncSetAttribute(spec_apply_cluster_kernel<4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
			attr_done = true;
		}
		cudaLaunchConfig_t ccfg = {};
		ccfg.gridDim = dim3(16); ccfg.blockDim = dim3(1024); ccfg.dynamicSmemBytes = 0; ccfg.stream = c->stream;
		cudaLaunchAttribute cat[1];
		cat[0].id = cudaLaunchAttributeClusterDimension; cat[0].val.clusterDim.x = 16; cat[0].val.clusterDim.y = 1; cat[0].val.clusterDim.z = 1;
		ccfg.attrs = cat; ccfg.numAttrs = 1;
		RadDev Dc = S;
		void* cargs[4] = { (void*)&Dc, (void*)&slot_base, (void*)&nslots, (void*)&stop_armed };
		const cudaError_t ce = cudaLaunchKernelExC(&ccfg, cfn, cargs);
		if (ce == cudaSuccess) { c->launches++; return RAD_OK; }
		cudaGetLastError();                      // (a device that cannot place the cluster: the grid form below)
	}
	// default: simulate over the set's own 64 x 64 transfers, verify against all patches, commit (no barrier per shot);
	// RAD_SPEC_SIM=0: the cooperative one-barrier-per-shot kernels below
	static const bool use_sim = [] { const char* e = getenv("RAD_SPEC_SIM"); return !e || atoi(e) != 0; }();
	if (use_sim && nslots <= RAD_SPEC_SLOTS) {
		spec_sim_kernel<<<1, 64, 0, c->stream>>>(S, slot_base, nslots, stop_armed);
		spec_walk_kernel<true><<<patch_grid(S.P, 256), 256, 0, c->stream>>>(S, slot_base, stop_armed);
		spec_walk_kernel<false><<<patch_grid(S.P, 256), 256, 0, c->stream>>>(S, slot_base, stop_armed);
		c->launches += 3;
		return RAD_OK;
	}
	// small scenes: everything a shot reads staged in shared memory (RAD_SPEC_SMEM=0: the register form below)
	static const bool use_smem = [] { const char* e = getenv("RAD_SPEC_SMEM"); return !e || atoi(e) != 0; }();
	if (use_smem && S.P <= (uint32_t)nsm * 512u && S.P <= 256u * 512u && nslots <= RAD_SPEC_SLOTS) {
		static bool smem_attr = false;
		if (!smem_attr) { cudaFuncSetAttribute(spec_apply_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RAD_SPEC_SLOTS * 512 * 4); smem_attr = true; }
		cudaLaunchConfig_t scfg = {};
		scfg.gridDim = dim3((S.P + 511u) / 512u); scfg.blockDim = dim3(512); scfg.dynamicSmemBytes = (size_t)nslots * 512 * 4; scfg.stream = c->stream;
		cudaLaunchAttribute sat[1];
		sat[0].id = cudaLaunchAttributeCooperative; sat[0].val.cooperative = 1;
		scfg.attrs = sat; scfg.numAttrs = 1;
		RadDev Ds = S;
		void* sargs[4] = { (void*)&Ds, (void*)&slot_base, (void*)&nslots, (void*)&stop_armed };
		const cudaError_t se = cudaLaunchKernelExC(&scfg, (const void*)spec_apply_smem_kernel, sargs);
		if (se == cudaSuccess) { c->launches++; return RAD_OK; }
		cudaGetLastError();
	}
	const void* fn = nullptr;
	if (ppt <= 1) fn = (const void*)spec_apply_kernel<1>; else if (ppt <= 2) fn = (const void*)spec_apply_kernel<2>; else if (ppt <= 4) fn = (const void*)spec_apply_kernel<4>;
	else if (ppt <= 8) fn = (const void*)spec_apply_kernel<8>; else if (ppt <= 16) fn = (const void*)spec_apply_kernel<16>;
	else { c->err = "speculative k = 1 path: too many patches per SM (set RAD_SPEC=0)"; return RAD_E_ARG; }
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(nb); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
	cfg.attrs = at; cfg.numAttrs = 1;
	RadDev D = S;
	void* args[4] = { (void*)&D, (void*)&slot_base, (void*)&nslots, (void*)&stop_armed };
	const cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
	c->launches++;
	if (e != cudaSuccess) { c->err = std::string("spec_apply_kernel launch: ") + cudaGetErrorString(e); return RAD_E_CUDA; }
	return RAD_OK;
}
void rad_launch_xreduce(rad_ctx* c) {
	const RadDev& D = c->d;
	const uint32_t slice = (D.P + D.xworld - 1) / D.xworld;
	xreduce_kernel<<<patch_grid(slice, 256), 256, 0, c->stream>>>(D);
	c->launches++;
}
void rad_launch_finish(rad_ctx* c, bool fuse_select) {
	const RadDev& D = c->d;
	const uint32_t T = apply_threads(D.P);
	apply_kernel<2><<<patch_grid(D.P, T), T, D.k * sizeof(EmLite), c->stream>>>(D, fuse_select ? 1 : 0, (int)c->parity);
	c->launches++;
	if (D.xworld > 0) { xb_bump_kernel<<<1, 32, 0, c->stream>>>(D); c->launches++; }
}
