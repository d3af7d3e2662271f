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
#include "rad_internal.cuh"

namespace {

#define FULL 0xFFFFFFFFu

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
__global__ void __launch_bounds__(1024) select_reference_kernel(RadDev D) {
	__shared__ float s_e[1024];
	__shared__ unsigned s_flags[32];
	__shared__ uint32_t l_id[65]; __shared__ float l_e[65];
	__shared__ int s_n; __shared__ float s_min;
	const uint32_t count = D.k;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (threadIdx.x == 0) { s_n = 0; s_min = 0.0f; }
	__syncthreads();
	for (uint32_t base = 0; base < D.P; base += 1024) {
		const uint32_t i = base + threadIdx.x;
		const float e = i < D.P ? len2(D.rad[i], D.rad[D.P + i], D.rad[2 * (size_t)D.P + i]) : 0.0f;
		s_e[threadIdx.x] = e;
		// superset of the accepted candidates: the list minimum never decreases
		const bool cand = i < D.P && ((s_n == 0 && i == 0) || (e > 0.0f && e >= s_min));
		const unsigned b = __ballot_sync(FULL, cand);
		if (lane == 0) s_flags[w] = b;
		__syncthreads();
		if (threadIdx.x == 0) {
			int n = s_n;
			for (int ww = 0; ww < 32; ww++) {
				unsigned m = s_flags[ww];
				while (m) {
					const int j = ww * 32 + __ffs(m) - 1;
					m &= m - 1;
					const float x = s_e[j];
					if (!(n == 0 || (x > 0.0f && l_e[n - 1] <= x))) continue;
					// every tie group is reversed by the stable-sort + reverse of the reference (ModelContainer.cpp:273-275)
					for (int a = 0; a < n;) {
						int z = a;
						while (z + 1 < n && l_e[z + 1] == l_e[a]) z++;
						for (int lo = a, hi = z; lo < hi; lo++, hi--) { const uint32_t t = l_id[lo]; l_id[lo] = l_id[hi]; l_id[hi] = t; }
						a = z + 1;
					}
					int pos = 0;
					while (pos < n && l_e[pos] > x) pos++;      // the newcomer leads its tie group
					for (int q = n; q > pos; q--) { l_id[q] = l_id[q - 1]; l_e[q] = l_e[q - 1]; }
					l_id[pos] = base + j; l_e[pos] = x;
					n++;
					if (n > (int)count) n = (int)count;
				}
			}
			s_n = n;
			s_min = n > 0 ? l_e[n - 1] : 0.0f;
		}
		__syncthreads();
	}
	if (threadIdx.x < count) {
		const bool ok = (int)threadIdx.x < s_n;
		D.em[threadIdx.x].id = ok ? l_id[threadIdx.x] : 0u;
		D.em[threadIdx.x].valid = ok ? 1u : 0u;
	}
}

// ---- clean top-k for k > 1: tournament of block-wide bitonic sorts -----------------------------
// key = (bits(|B|^2) << 32 | ~id): unique, and descending key order == (energy desc, id asc).  Every block sorts a
// chunk of 2048 keys in shared memory and keeps its best 64; levels repeat (P -> P/32 -> ...) until one block is
// left, which writes the emitter list.  Exact and deterministic; 2 launches for 16 k patches, 3 for 1 M.
constexpr int kTopChunk = 2048, kTopKeep = 64, kTopThreads = 1024;   // one compare-exchange per thread per bitonic stage

template <bool FIRST>
__global__ void __launch_bounds__(kTopThreads) topk_level_kernel(RadDev D, const unsigned long long* __restrict__ in, uint32_t n_in,
                                                                  unsigned long long* __restrict__ out, int final_level) {
	__shared__ unsigned long long s[kTopChunk];
	const uint32_t base = blockIdx.x * kTopChunk;
	for (int t = threadIdx.x; t < kTopChunk; t += kTopThreads) {
		const uint32_t i = base + t;
		unsigned long long key = 0ull;
		if (i < n_in) {
			if (FIRST) {
				const uint32_t eb = __float_as_uint(len2(D.rad[i], D.rad[D.P + i], D.rad[2 * (size_t)D.P + i]));
				if (eb != 0 && eb < 0x7F800000u) key = ((unsigned long long)eb << 32) | (0xFFFFFFFFu - i);
			} else key = in[i];
		}
		s[t] = key;
	}
	__syncthreads();
	for (int k = 2; k <= kTopChunk; k <<= 1)
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int t = threadIdx.x; t < kTopChunk / 2; t += kTopThreads) {
				const int i = ((t / j) * 2 * j) + (t % j), l = i + j;
				const unsigned long long a = s[i], b = s[l];
				const bool desc = (i & k) == 0;
				if ((a < b) == desc) { s[i] = b; s[l] = a; }
			}
			__syncthreads();
		}
	if (final_level) {
		if (threadIdx.x < D.k) {
			const unsigned long long key = s[threadIdx.x];
			D.em[threadIdx.x].id = key ? 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull) : 0u;
			D.em[threadIdx.x].valid = key ? 1u : 0u;
		}
	} else if (threadIdx.x < kTopKeep) out[(size_t)blockIdx.x * kTopKeep + threadIdx.x] = s[threadIdx.x];
}

__global__ void set_emitters_kernel(RadDev D, const uint32_t* __restrict__ ids, uint32_t n) {
	const uint32_t h = threadIdx.x;
	if (h < D.k) { D.em[h].id = h < n ? ids[h] : 0u; D.em[h].valid = (h < n && ids[h] < D.P) ? 1u : 0u; }
}

// ---- K4: energy transfer + emitter update (+ fused argmax) -----------------------------------
// MODE 0: single GPU — reference association B += d_0, += d_1, ...   (reads F of all k slots)
// MODE 1: multi GPU, local part — dB = sum of this rank's slots      (B untouched)
// MODE 2: multi GPU, final part — B += dB (after the all-reduce), emitter update
// F_h[i] of kFBatch consecutive slots: all the (independent) loads are issued before anything depends on them — the
// kernel is a latency chain otherwise — then the slots are zeroed for the next batch (fill_n(p_tmp_formfactors, 0),
// Main.cpp:1278).  Invalid (NULL) emitters are not read.
constexpr int kFBatch = 16;
__device__ __forceinline__ void take_F(const RadDev& D, const RadEmitter* s_em, uint32_t hb, uint32_t hend, uint32_t P, uint32_t i, float* f) {
	#pragma unroll
	for (int j = 0; j < kFBatch; j++) {
		const uint32_t h = hb + j;
		f[j] = (h < hend && s_em[h].valid) ? __ldcs(D.F + (size_t)h * P + i) : 0.0f;
	}
	#pragma unroll
	for (int j = 0; j < kFBatch; j++)
		if (f[j] != 0.0f) D.F[(size_t)(hb + j) * P + i] = 0.0f;
}

template <int MODE>
__global__ void __launch_bounds__(256) apply_kernel(RadDev D, int fuse_select, int parity) {
	extern __shared__ RadEmitter s_em[];
	for (uint32_t h = threadIdx.x; h < D.k; h += blockDim.x) s_em[h] = D.em[h];
	__syncthreads();
	const uint32_t P = D.P, k = D.k;
	const float rho = D.reflectivity;
	int last_h = -1; uint32_t nvalid = 0;
	for (uint32_t h = 0; h < k; h++) if (s_em[h].valid) { last_h = (int)h; nvalid++; }
	unsigned long long best = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
		if (MODE == 1) {
			float dx = 0.0f, dy = 0.0f, dz = 0.0f;
			for (uint32_t hb = D.h0; hb < D.h1; hb += kFBatch) {
				float f[kFBatch];
				take_F(D, s_em, hb, D.h1, P, i, f);
				#pragma unroll
				for (int j = 0; j < kFBatch; j++) {
					const uint32_t h = hb + j;
					if (h >= D.h1 || !s_em[h].valid) continue;
					dx += ((s_em[h].S[0] * f[j]) * rho) * s_em[h].color[0];
					dy += ((s_em[h].S[1] * f[j]) * rho) * s_em[h].color[1];
					dz += ((s_em[h].S[2] * f[j]) * rho) * s_em[h].color[2];
				}
			}
			D.dB[i] = dx; D.dB[P + i] = dy; D.dB[2 * (size_t)P + i] = dz;
			continue;
		}
		float bx = D.rad[i], by = D.rad[P + i], bz = D.rad[2 * (size_t)P + i];
		if (MODE == 0) {
			for (uint32_t hb = 0; hb < k; hb += kFBatch) {
				float f[kFBatch];
				take_F(D, s_em, hb, k, P, i, f);
				#pragma unroll
				for (int j = 0; j < kFBatch; j++) {
					const uint32_t h = hb + j;
					if (h >= k || !s_em[h].valid) continue;
					// p->radiosity += S_h * F[i] * reflectivity * colour(emitter_h)   (Main.cpp:1274)
					bx += ((s_em[h].S[0] * f[j]) * rho) * s_em[h].color[0];
					by += ((s_em[h].S[1] * f[j]) * rho) * s_em[h].color[1];
					bz += ((s_em[h].S[2] * f[j]) * rho) * s_em[h].color[2];
				}
			}
		} else {
			bx += D.dB[i]; by += D.dB[P + i]; bz += D.dB[2 * (size_t)P + i];
		}
		// emitters: lastEnergy, I += S, B -= S   (Main.cpp:1286-1295)
		for (uint32_t h = 0; h < k; h++) {
			if (!s_em[h].valid || s_em[h].id != i) continue;
			if ((int)h == last_h) {
				const float l = sqrtf(len2(bx, by, bz));
				D.ctl->last_energy_len = l;
				if ((double)l < 0.1) D.ctl->stopped = 1;                    // Main.cpp:1298
			}
			D.illum[i] += s_em[h].S[0]; D.illum[P + i] += s_em[h].S[1]; D.illum[2 * (size_t)P + i] += s_em[h].S[2];
			bx -= s_em[h].S[0]; by -= s_em[h].S[1]; bz -= s_em[h].S[2];
		}
		D.rad[i] = bx; D.rad[P + i] = by; D.rad[2 * (size_t)P + i] = bz;
		if (fuse_select) best = max(best, energy_key_last(len2(bx, by, bz), i));
	}
	if (MODE != 1) {
		if (blockIdx.x == 0 && threadIdx.x == 0) { D.ctl->batches_done += 1; D.ctl->shots_done += nvalid; }
		if (fuse_select) {
			best = block_max(best);
			if (threadIdx.x == 0 && best) atomicMax(&D.ctl->selkey[parity ^ 1], best);
		}
	}
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
	c->selkey_valid = true;
}

void rad_launch_select(rad_ctx* c) {
	const RadDev& D = c->d;
	if (D.k == 1) {
		if (!c->selkey_valid) rad_launch_argmax(c);
		rad_launch_camera(c, (int)c->parity);      // decodes the fused argmax key, then snapshot + MVPs
		return;
	} else if (c->cfg.select_mode == RAD_SELECT_REFERENCE) {
		select_reference_kernel<<<1, 1024, 0, c->stream>>>(D);
		c->launches++;
	} else {
		uint32_t n = D.P;
		const unsigned long long* in = nullptr;
		unsigned long long* out = D.cand0;
		bool first = true;
		for (;;) {
			const uint32_t nb = (n + kTopChunk - 1) / kTopChunk;
			const int fin = nb == 1;
			if (first) topk_level_kernel<true><<<nb, kTopThreads, 0, c->stream>>>(D, in, n, out, fin);
			else topk_level_kernel<false><<<nb, kTopThreads, 0, c->stream>>>(D, in, n, out, fin);
			c->launches++;
			if (fin) break;
			n = nb * kTopKeep; in = out; out = out == D.cand0 ? D.cand1 : D.cand0; first = false;
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
	apply_kernel<0><<<patch_grid(D.P, T), T, D.k * sizeof(RadEmitter), c->stream>>>(D, fuse_select ? 1 : 0, (int)c->parity);
	c->launches++;
}
void rad_launch_delta(rad_ctx* c) {
	const RadDev& D = c->d;
	const uint32_t T = apply_threads(D.P);
	apply_kernel<1><<<patch_grid(D.P, T), T, D.k * sizeof(RadEmitter), c->stream>>>(D, 0, 0);
	c->launches++;
}
void rad_launch_finish(rad_ctx* c, bool fuse_select) {
	const RadDev& D = c->d;
	const uint32_t T = apply_threads(D.P);
	apply_kernel<2><<<patch_grid(D.P, T), T, D.k * sizeof(RadEmitter), c->stream>>>(D, fuse_select ? 1 : 0, (int)c->parity);
	c->launches++;
}
