// Micro-benchmark behind the rasteriser's visibility-update design: what does one per-fragment
// "min" cost on a B200 in each of the forms the kernel could use?  Access pattern = the small-quad walk: every quarter
// warp walks 8 rows of an 8-pixel-wide box at a random position of its tile (shared memory: a 64 x 32 tile per CTA;
// global: 1024 x 768 key images, `nslots` of them), all lanes or about half of them active.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics_ubench atomics_ubench.cu && ./atomics_ubench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define TW 64
#define TH 32
enum { S_MIN32, S_MIN32_RET, S_MIN64, S_LDST32, S_ADDF32, G_RED64, G_RED32, G_RED64_ROWS, NMODES };
static const char* kNames[NMODES] = { "smem atomicMin u32 (ATOMS.MIN)", "smem atomicMin u32, result used", "smem atomicMin u64 (CAS loop)",
	"smem LDS + compare + STS u32 (no atomic)", "smem atomicAdd f32", "global RED.MIN.64", "global RED.MIN.32", "global RED.MIN.64, full-warp rows" };

template <int MODE>
__global__ void __launch_bounds__(128) bench(unsigned long long* __restrict__ g64, uint32_t* __restrict__ g32, uint32_t nslots, int iters, int half, uint32_t* sink) {
	__shared__ __align__(16) unsigned long long s64[TW * TH];
	uint32_t* s32 = reinterpret_cast<uint32_t*>(s64);
	float* sf = reinterpret_cast<float*>(s64);
	for (int i = threadIdx.x; i < TW * TH; i += blockDim.x) s64[i] = ~0ull;
	__syncthreads();
	const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, l8 = lane & 7;
	uint32_t h = (MODE == G_RED64_ROWS ? gt >> 5 : gt >> 3) * 2654435761u + 12345u;
	uint32_t acc = 0;
	const size_t RES = 1024u * 768u;
	for (int it = 0; it < iters; it++) {
		h = h * 1664525u + 1013904223u;
		const uint32_t key = (h & 0xFFFFFF00u) | (gt & 0xFFu);
		const bool act = !half || (((h >> 5) + l8 * 0x9E3779B9u) >> 31) != 0;     // about half of the lanes
		if (MODE < G_RED64) {
			const uint32_t x = (h >> 3) % (TW - 8 + 1) + l8, y = (h >> 17) % (TH - 8 + 1);
			#pragma unroll
			for (int r = 0; r < 8; r++) {
				const uint32_t a = (y + r) * TW + x;
				if (act) {
					if (MODE == S_MIN32) atomicMin(&s32[a], key + r);
					if (MODE == S_MIN32_RET) acc += atomicMin(&s32[a], key + r) >> 8 == (key + r) >> 8;
					if (MODE == S_MIN64) atomicMin(&s64[a], ((unsigned long long)(key + r) << 32) | gt);
					if (MODE == S_LDST32) { if (s32[a] > key + r) s32[a] = key + r; }
					if (MODE == S_ADDF32) atomicAdd(&sf[a], 1.0f);
				}
			}
		} else {
			const uint32_t slot = (h >> 8) % nslots;
			uint32_t x = (h >> 3) % (1024 - 40), y = (h >> 17) % (768 - 8);
			x += MODE == G_RED64_ROWS ? lane : l8;
			const size_t a = slot * RES + (size_t)y * 1024 + x;
			#pragma unroll
			for (int r = 0; r < 8; r++) {
				if (act) {
					if (MODE == G_RED32) atomicMin(&g32[a + (size_t)r * 1024], key + r);
					else atomicMin(&g64[a + (size_t)r * 1024], ((unsigned long long)(key + r) << 32) | gt);
				}
			}
		}
	}
	__syncthreads();
	if (MODE < G_RED64) acc += s32[threadIdx.x];
	if (acc == 0x12345678u) *sink = acc;
}

template <int MODE>
static void run(unsigned long long* g64, uint32_t* g32, uint32_t nslots, int half, uint32_t* sink, int ctas_per_sm) {
	const int blocks = 148 * ctas_per_sm, iters = MODE < G_RED64 ? 4096 : 512;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	bench<MODE><<<blocks, 128>>>(g64, g32, nslots, 16, half, sink);
	float best = 1e30f;
	for (int rep = 0; rep < 3; rep++) {
		if (MODE >= G_RED64) { cudaMemset(g64, 0xFF, (size_t)nslots * 1024 * 768 * 8); }
		cudaEventRecord(e0);
		bench<MODE><<<blocks, 128>>>(g64, g32, nslots, iters, half, sink);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
	}
	const double lanes = (double)blocks * 128 * iters * 8 * (half ? 0.5 : 1.0);
	const double gops = lanes / (best * 1e-3) / 1e9;
	printf("{\"mode\": \"%s\", \"half_lanes\": %d, \"ctas_per_sm\": %d, \"footprint_mb\": %.0f, \"gops\": %.1f, \"cyc_per_lane_per_sm_at_1965\": %.3f, \"ms\": %.3f}\n",
	       kNames[MODE], half, ctas_per_sm, MODE >= G_RED64 ? nslots * 1024.0 * 768 * (MODE == G_RED32 ? 4 : 8) / 1e6 : 0.0, gops, 148 * 1.965 / gops, best);
	cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
}

int main() {
	const uint32_t maxslots = 64;
	unsigned long long* g64; uint32_t* sink;
	cudaMalloc(&g64, (size_t)maxslots * 1024 * 768 * 8); cudaMalloc(&sink, 4);
	uint32_t* g32 = reinterpret_cast<uint32_t*>(g64);
	for (int half = 0; half < 2; half++) {
		for (int c : { 8, 16 }) {
			run<S_MIN32>(g64, g32, 1, half, sink, c);
			run<S_MIN32_RET>(g64, g32, 1, half, sink, c);
			run<S_MIN64>(g64, g32, 1, half, sink, c);
			run<S_LDST32>(g64, g32, 1, half, sink, c);
			run<S_ADDF32>(g64, g32, 1, half, sink, c);
		}
		for (uint32_t ns : { 8u, 16u, 64u }) {
			run<G_RED64>(g64, g32, ns, half, sink, 8);
			run<G_RED32>(g64, g32, ns, half, sink, 8);
		}
		run<G_RED64_ROWS>(g64, g32, 8, half, sink, 8);
		run<G_RED64_ROWS>(g64, g32, 64, half, sink, 8);
	}
	return 0;
}
