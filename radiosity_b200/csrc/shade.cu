// K5 — display stage: Colors::smoothShadePatch for every patch (Colors.cpp:198-261; CPU loop over all
// patches inside OnIdle, Main.cpp:1323-1341).  Two passes: E_i = colour_i (.) (I_i + B_i), then per patch a gather of the 8
// neighbours' E and four 4-term means in the reference's summation order (bit-identical to the reference's CPU code).
// Algorithmic bytes per patch: 36 (colour, I, B) + 32 (neighbour ids) + 48 (12 output floats); the E gather hits L2.
#include "rad_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) shade_energy_kernel(RadDev D) {
	const size_t P = D.P;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < D.P; i += gridDim.x * blockDim.x)
		for (int c = 0; c < 3; c++)
			D.shade_e[c * P + i] = D.color[c * P + i] * (D.illum[c * P + i] + D.rad[c * P + i]);
}

__global__ void __launch_bounds__(256) shade_gather_kernel(RadDev D, float* __restrict__ out12) {
	const size_t P = D.P;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < D.P; i += gridDim.x * blockDim.x) {
		int nb[8];
		#pragma unroll
		for (int j = 0; j < 8; j++) nb[j] = D.nb[(size_t)j * P + i];
		float o[12];
		#pragma unroll
		for (int c = 0; c < 3; c++) {
			const float* __restrict__ E = D.shade_e + c * P;
			const float self = E[i];
			float e[8];
			#pragma unroll
			for (int j = 0; j < 8; j++) e[j] = E[nb[j]];
			// output order lb, rb, rt, lt; each sum in the reference's order, then * (1/4)
			o[0 + c] = (((self + e[5]) + e[6]) + e[7]) * 0.25f;
			o[3 + c] = (((self + e[3]) + e[4]) + e[5]) * 0.25f;
			o[6 + c] = (((self + e[1]) + e[2]) + e[3]) * 0.25f;
			o[9 + c] = (((self + e[7]) + e[0]) + e[1]) * 0.25f;
		}
		float4* dst = reinterpret_cast<float4*>(out12 + 12 * (size_t)i);
		dst[0] = make_float4(o[0], o[1], o[2], o[3]);
		dst[1] = make_float4(o[4], o[5], o[6], o[7]);
		dst[2] = make_float4(o[8], o[9], o[10], o[11]);
	}
}

} // namespace

void rad_launch_shade(rad_ctx* c, float* out12) {
	const RadDev& D = c->d;
	uint32_t b = (D.P + 255) / 256;
	if (b > 148 * 8) b = 148 * 8;
	shade_energy_kernel<<<b, 256, 0, c->stream>>>(D);
	shade_gather_kernel<<<b, 256, 0, c->stream>>>(D, out12);
	c->launches += 2;
}
