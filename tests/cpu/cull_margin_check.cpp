// CPU check of the cull margin of the CUDA path: radiosity_b200/csrc/camera.cuh is compiled for the host (g++, the few
// CUDA built-ins it uses shimmed below) and cull_margin2() is evaluated for quads read from stdin — tests/
// test_cull_rule_cpu.py compares the values with the numpy restatement (tests/cull_rule.py) that is checked against the
// oracle's item buffers.  Build: g++ -O1 -ffp-contract=off -std=c++17 -I/usr/local/cuda/include cull_margin_check.cpp
// (test infrastructure only)
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static struct { unsigned x, y, z; } threadIdx;
static inline void __syncthreads() {}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
#include "../../radiosity_b200/csrc/camera.cuh"

int main() {
	float v[12];
	while (fread(v, 4, 12, stdin) == 12) {
		Quad q;
		q.a = mk(v[0], v[1], v[2]); q.b = mk(v[3], v[4], v[5]); q.c = mk(v[6], v[7], v[8]); q.d = mk(v[9], v[10], v[11]);
		const float c = cull_margin2(q);
		fwrite(&c, 4, 1, stdout);
	}
	return 0;
}
