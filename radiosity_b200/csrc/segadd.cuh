// Warp-level run merging of ProcessHemicube (K2): shared by process.cu and the tile-binned rasteriser's fused
// process stage (raster_tiles.cu).  Equal ids in neighbouring pixels / lanes collapse into ONE red.global.add.f32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

#ifndef FULL
#define FULL 0xFFFFFFFFu
#endif

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

__device__ __forceinline__ void flush_run(uint32_t id, float v, float* __restrict__ F, uint32_t P) {
	if (id != 0 && id - 1 < P) atomicAdd(F + (id - 1), v);
}

// four consecutive pixels per lane: runs that end inside the lane are flushed by the lane itself, the lane's last run
// joins the warp-wide segmented reduction (equal ids in neighbouring lanes collapse into one atomic)
__device__ __forceinline__ void process4(const uint4 id, const float4 f, int lane, float* __restrict__ F, uint32_t P) {
	uint32_t cur = id.x; float acc = f.x;
	if (id.y == cur) acc += f.y; else { flush_run(cur, acc, F, P); cur = id.y; acc = f.y; }
	if (id.z == cur) acc += f.z; else { flush_run(cur, acc, F, P); cur = id.z; acc = f.z; }
	if (id.w == cur) acc += f.w; else { flush_run(cur, acc, F, P); cur = id.w; acc = f.w; }
	segmented_add(cur, acc, lane, F, P);
}

} // namespace
