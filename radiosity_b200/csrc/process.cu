// K2 — ProcessHemicube: delta-form-factor-weighted scatter-add (replaces the OpenCL kernel
// Kernel_ProcessHemicube.h:9-70, its launch Main.cpp:584-596,1218-1226, the four blocking read-backs
// Main.cpp:1232-1240 and the O(k * records) CPU gather Main.cpp:1257-1269).
//
//   F_h[i] = sum over atlas pixels px of hemicube h with item[px] == i+1 of dFF[px]
//
// The reference expresses this as run-length records appended through an atomic counter and
// finished on the CPU.  Here every warp streams 32 consecutive pixels per step (coalesced 128 B of
// ids + 128 B of dFF), collapses runs of equal ids with a 5-step segmented shuffle reduction
// (ids are spatially coherent, so a run is almost always a contiguous lane range) and issues ONE
// red.global.add.f32 per run straight into F_h — no record stream, no host round trip.
// Algorithmic traffic: 4 B id + 4 B dFF per pixel (the dFF table of one hemicube is shared by all k
// hemicubes and stays L2-resident); see DESIGN.md.
//
// Two entry points: process_kernel<false> reads the uint32 item buffer (the API seam of
// rad_process_hemicubes / rad_bench_process); process_kernel<true> is the fused steady-state form used
// by rad_shoot: it reads the rasteriser's 64-bit keys directly, resets them for the next batch and
// only materialises the item buffer when asked to.
#include "rad_internal.cuh"

namespace {

#define FULL 0xFFFFFFFFu
constexpr int kUnroll = 4;

__device__ __forceinline__ void segmented_add(uint32_t id, float v, int lane, float* __restrict__ F, uint32_t P) {
	const uint32_t prev = __shfl_up_sync(FULL, id, 1);
	const bool head = lane == 0 || id != prev;
	const unsigned heads = __ballot_sync(FULL, head);
	const unsigned right = (heads >> lane) >> 1;      // head flags of the lanes to my right
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const float o = __shfl_down_sync(FULL, v, d);
		if (lane + d < 32 && (right & ((1u << d) - 1u)) == 0) v += o;
	}
	if (head && id != 0 && id - 1 < P) atomicAdd(F + (id - 1), v);   // RED.E.ADD.F32 (result unused)
}

template <bool FROM_KEYS>
__global__ void __launch_bounds__(256) process_kernel(RadDev D, int keep_items) {
	const uint32_t slot = D.h0 + blockIdx.y;
	if (FROM_KEYS && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { D.ctl->pad = D.ctl->q_tris; D.ctl->q_tris = 0; D.ctl->q_entries = 0; }
	if (!D.em[slot].valid) return;
	const int lane = threadIdx.x & 31;
	const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	const uint32_t ngroups = D.RES >> 5;              // RES is a multiple of 32 (N % 16 == 0)
	float* __restrict__ F = D.F + (size_t)slot * D.P;
	const float* __restrict__ ff = D.ff;
	unsigned long long* __restrict__ keys = D.keys + (size_t)slot * D.RES;
	uint32_t* __restrict__ items = D.items + (size_t)slot * D.RES;

	for (uint32_t g0 = gw * kUnroll; g0 < ngroups; g0 += nw * kUnroll) {
		uint32_t id[kUnroll]; float v[kUnroll];
		#pragma unroll
		for (int u = 0; u < kUnroll; u++) {
			const uint32_t g = g0 + u;
			id[u] = 0; v[u] = 0.0f;
			if (g < ngroups) {
				const uint32_t i = (g << 5) + lane;
				if (FROM_KEYS) {
					const unsigned long long k = keys[i];
					id[u] = k == RAD_CLEAR_KEY ? 0u : (uint32_t)(k & 0xFFFFFFFFull);
				} else {
					id[u] = __ldcs(items + i);
				}
				v[u] = __ldg(ff + i);
			}
		}
		#pragma unroll
		for (int u = 0; u < kUnroll; u++) {
			const uint32_t g = g0 + u;
			if (g < ngroups) {
				const uint32_t i = (g << 5) + lane;
				if (FROM_KEYS) {
					keys[i] = RAD_CLEAR_KEY;
					if (keep_items) items[i] = id[u];
				}
				segmented_add(id[u], v[u], lane, F, D.P);
			}
		}
	}
}

} // namespace

static dim3 process_grid(const RadDev& D) {
	const uint32_t nslots = D.h1 - D.h0;
	const uint32_t groups = D.RES >> 5;
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

void rad_launch_resolve_process(rad_ctx* c, bool keep_items) {
	const RadDev& D = c->d;
	if (D.h1 == D.h0) return;
	process_kernel<true><<<process_grid(D), 256, 0, c->stream>>>(D, keep_items ? 1 : 0);
	c->launches++;
	c->keys_dirty = false;
}
