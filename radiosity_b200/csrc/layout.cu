// Layout conversion between the reference's host layouts and the device layouts, on the GPU:
//   float[P*3] AoS  <->  three planes [3][P]           (colour, radiosity, illumination, dB)
//   float[P*12] quads (ModelContainer.cpp:100-107)  ->  three float4 streams
#include "rad_internal.cuh"

namespace {
__global__ void aos3_to_planes_kernel(const float* __restrict__ aos, float* __restrict__ planes, uint32_t P) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
		planes[i] = aos[3 * (size_t)i]; planes[P + i] = aos[3 * (size_t)i + 1]; planes[2 * (size_t)P + i] = aos[3 * (size_t)i + 2];
	}
}
__global__ void planes_to_aos3_kernel(const float* __restrict__ planes, float* __restrict__ aos, uint32_t P) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
		aos[3 * (size_t)i] = planes[i]; aos[3 * (size_t)i + 1] = planes[P + i]; aos[3 * (size_t)i + 2] = planes[2 * (size_t)P + i];
	}
}
__global__ void split_quads_kernel(const float4* __restrict__ q, float4* __restrict__ v0, float4* __restrict__ v1, float4* __restrict__ v2, uint32_t P) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
		v0[i] = q[3 * (size_t)i]; v1[i] = q[3 * (size_t)i + 1]; v2[i] = q[3 * (size_t)i + 2];
	}
}
__global__ void nb_to_planes_kernel(const int32_t* __restrict__ nb8, int32_t* __restrict__ planes, uint32_t P) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
		for (int j = 0; j < 8; j++) planes[(size_t)j * P + i] = nb8[8 * (size_t)i + j];
}
uint32_t grid_for(uint32_t P) { uint32_t b = (P + 255) / 256; return b > 148 * 8 ? 148 * 8 : (b ? b : 1); }
} // namespace

void rad_launch_aos3_to_planes(rad_ctx* c, const float* aos, float* planes, uint32_t P) {
	aos3_to_planes_kernel<<<grid_for(P), 256, 0, c->stream>>>(aos, planes, P);
}
void rad_launch_planes_to_aos3(rad_ctx* c, const float* planes, float* aos, uint32_t P) {
	planes_to_aos3_kernel<<<grid_for(P), 256, 0, c->stream>>>(planes, aos, P);
}
void rad_launch_split_quads(rad_ctx* c, const float* verts12, uint32_t P) {
	split_quads_kernel<<<grid_for(P), 256, 0, c->stream>>>(reinterpret_cast<const float4*>(verts12), const_cast<float4*>(c->d.v0),
	                                                        const_cast<float4*>(c->d.v1), const_cast<float4*>(c->d.v2), P);
}
void rad_launch_nb_to_planes(rad_ctx* c, const int32_t* nb8, uint32_t P) {
	nb_to_planes_kernel<<<grid_for(P), 256, 0, c->stream>>>(nb8, c->d.nb, P);
}
