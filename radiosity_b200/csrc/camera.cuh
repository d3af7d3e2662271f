// Shooter camera: radiosity snapshot, emitter record and the five face MVPs (replaces the per-emitter CPU work of OnIdle,
// Main.cpp:1161,1172-1183 -> Camera.cpp:19-52,97-103, Transform.cpp:26-46,70-80,127-156).  Shared by camera_kernel
// (raster.cu) and by the tail of the fused k == 1 update kernel (select_update.cu), which prepares the next shot's camera
// as soon as its argmax is known.
#pragma once
#include "rad_internal.cuh"

namespace {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 vneg(V3 a) { return mk(-a.x, -a.y, -a.z); }
// the reference's v_Cross: a.v_Cross(b) == b x a   (Vector.h:534-537)
__device__ __forceinline__ V3 rcross(V3 a, V3 b) {
	return mk(b.y * a.z - b.z * a.y, b.z * a.x - b.x * a.z, b.x * a.y - b.y * a.x);
}
__device__ __forceinline__ V3 vnormalize(V3 a) {   // Vector.h:390-400: t = 1/len, then scale
	float t = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
	if (t != 0) { t = 1 / t; a.x *= t; a.y *= t; a.z *= t; }
	return a;
}

// out = a * b, column-major m[c*4+r], term order of Matrix4f::ProductOf (Vector.cpp:445-458)
__device__ __forceinline__ void mat_product(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b) {
	for (int c = 0; c < 4; c++)
		for (int r = 0; r < 4; r++)
			out[c * 4 + r] = a[r] * b[c * 4] + a[4 + r] * b[c * 4 + 1] + a[8 + r] * b[c * 4 + 2] + a[12 + r] * b[c * 4 + 3];
}

struct Quad { V3 a, b, c, d; };
__device__ __forceinline__ Quad load_quad(const float4* __restrict__ v0, const float4* __restrict__ v1, const float4* __restrict__ v2, uint32_t p) {
	float4 q0 = __ldg(v0 + p), q1 = __ldg(v1 + p), q2 = __ldg(v2 + p);
	Quad q;
	q.a = mk(q0.x, q0.y, q0.z); q.b = mk(q0.w, q1.x, q1.y); q.c = mk(q1.z, q1.w, q2.x); q.d = mk(q2.y, q2.z, q2.w);
	return q;
}
__device__ __forceinline__ Quad load_quad(const RadDev& D, uint32_t p) { return load_quad(D.v0, D.v1, D.v2, p); }

// face order in the atlas = p_patchlook_perm (Main.h:210-211): UP, DOWN, LEFT, RIGHT, FRONT
// MVP = Perspective * LookAt(eye, target + eye, up), column-major m[c*4+r]
__device__ void build_mvp(const Quad& q, int face, const float* __restrict__ proj, float* __restrict__ out) {
	V3 eye = mk((q.a.x + q.b.x + q.c.x + q.d.x) / 4.0f, (q.a.y + q.b.y + q.c.y + q.d.y) / 4.0f, (q.a.z + q.b.z + q.c.z + q.d.z) / 4.0f);
	V3 normal = rcross(vsub(q.b, q.a), vsub(q.d, q.a));   // Patch::getNormal, Patch.cpp:272-276
	V3 pup = vsub(q.d, q.a);                              // Patch::getUp, Patch.cpp:253-255
	V3 target, up;
	switch (face) {                                       // Camera::lookFromPatch, Camera.cpp:19-52
	case 0: target = pup; up = vneg(normal); break;                       // UP
	case 1: target = vneg(pup); up = normal; break;                       // DOWN
	case 2: target = vneg(rcross(normal, pup)); up = pup; break;          // LEFT
	case 3: target = rcross(normal, pup); up = pup; break;                // RIGHT
	default: target = normal; up = pup; break;                            // FRONT
	}
	// CGLTransform::LookAt(eye, target + eye, up)
	V3 dir = vnormalize(vsub(vadd(target, eye), eye));
	V3 right = vnormalize(rcross(dir, up));
	up = rcross(right, dir);
	float la[16], tr[16], mv[16];
	la[0] = right.x; la[4] = right.y; la[8] = right.z;
	la[1] = up.x; la[5] = up.y; la[9] = up.z;
	la[2] = -dir.x; la[6] = -dir.y; la[10] = -dir.z;
	la[3] = 0; la[7] = 0; la[11] = 0; la[12] = 0; la[13] = 0; la[14] = 0; la[15] = 1;
	// Translate(-eye): (*this) *= Translation  (Vector.cpp:325-330,478-527) — full products so that signed
	// zeros come out exactly as in the reference
	for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) tr[c * 4 + r] = (c < 3) ? (float)(c == r) : 0.0f;
	tr[12] = -eye.x; tr[13] = -eye.y; tr[14] = -eye.z; tr[15] = 1;
	mat_product(mv, la, tr);
	mat_product(out, proj, mv);     // t_projection * t_modelview (Main.cpp:1183)
}

// The view basis (rows right, up, -dir of the LookAt matrix) of a face.  ideal == false: what the face really gets — the same
// float32 operations as build_mvp above (Camera.cpp:19-52, Transform.cpp:26-46; kept separate so that the MVP code stays
// as verified).  ideal == true: the same without the round trip of the target through eye; these vectors are +-axes of the
// orthonormal shooter frame the conservative culls work in.
__device__ void look_basis(const Quad& q, int face, bool ideal, V3& right, V3& up, V3& dir) {
	V3 eye = mk((q.a.x + q.b.x + q.c.x + q.d.x) / 4.0f, (q.a.y + q.b.y + q.c.y + q.d.y) / 4.0f, (q.a.z + q.b.z + q.c.z + q.d.z) / 4.0f);
	V3 normal = rcross(vsub(q.b, q.a), vsub(q.d, q.a));
	V3 pup = vsub(q.d, q.a);
	V3 target;
	switch (face) {
	case 0: target = pup; up = vneg(normal); break;
	case 1: target = vneg(pup); up = normal; break;
	case 2: target = vneg(rcross(normal, pup)); up = pup; break;
	case 3: target = rcross(normal, pup); up = pup; break;
	default: target = normal; up = pup; break;
	}
	dir = vnormalize(ideal ? target : vsub(vadd(target, eye), eye));
	right = vnormalize(rcross(dir, up));
	up = rcross(right, dir);
}
__device__ __forceinline__ float vdist(V3 a, V3 b) { const V3 d = vsub(a, b); return sqrtf(d.x * d.x + d.y * d.y + d.z * d.z); }

// Squared relative margin of the conservative culls for shooter q (RadEmitter::ctol).
// The reference builds a face's view matrix from LookAt(eye, target + eye, up) in float32: for a small patch |target| is
// tiny against |eye| (a side face's target is n x u, ~edge^3) and "target + eye - eye" turns the face — 1e-4 rad at 16 k
// patches, 0.13 degrees at 250 k, 0.9 degrees at 1 M, anything (axes swapped) for patches a thousand times smaller than
// the scene — faithfully reproduced in the MVPs.  The culls work in the ideal frame, and every plane they test is spanned
// by at most two basis vectors of a face, so a vertex that is outside an ideal plane by more than 2 * dev * |r| is outside
// the real one: margin = 2e-3 + 3 * dev, dev = the largest |real - ideal| over the 15 basis vectors.  A margin^2 >= 2 switches
// the frustum and horizon culls off (|h| <= sqrt(2) |r|); the facing cull does not depend on the frame.
__device__ float face_deviation(const Quad& q, int face) {        // largest |real - ideal| over one face's three basis vectors
	V3 rr, ru, rd, ir, iu, id;
	look_basis(q, face, false, rr, ru, rd);
	look_basis(q, face, true, ir, iu, id);
	return fmaxf(vdist(rr, ir), fmaxf(vdist(ru, iu), vdist(rd, id)));
}
__device__ __forceinline__ float margin2_of(float dev) { const float margin = 2e-3f + 3.0f * dev; return margin * margin; }
// serial form (camera_block below spreads the faces over five lanes); also what the CPU check of the margin evaluates
__device__ float cull_margin2(const Quad& q) {
	float dev = 0.0f;
	for (int face = 0; face < RAD_NFACES; face++) dev = fmaxf(dev, face_deviation(q, face));
	return margin2_of(dev);
}

// thread 0 of a block: emitter record of slot h.  sel_parity >= 0 (k == 1 only): the emitter is first decoded from the
// fused argmax key selkey[sel_parity] (all-zero energies leave key 0 == patch 0, the reference's seeded entry) and the
// other key is recycled.  Loads bypass L1: in the update kernel's tail the state was just written by other blocks.
__device__ __forceinline__ RadEmitter camera_emitter(const RadDev& D, uint32_t h, int sel_parity) {
	RadEmitter e = D.em[h];
	if (sel_parity >= 0) {
		const unsigned long long key = __ldcg(&D.ctl->selkey[sel_parity]);
		e.id = (uint32_t)(key & 0xFFFFFFFFull); e.valid = 1;
		D.ctl->selkey[sel_parity ^ 1] = 0ull;
	}
	if (e.valid && e.id < D.P) {
		for (int c = 0; c < 3; c++) {
			e.S[c] = __ldcg(D.rad + (size_t)c * D.P + e.id);            // p_tmp_radiosities[hi] (Main.cpp:1161)
			e.color[c] = D.color[(size_t)c * D.P + e.id];
		}
		const Quad q = load_quad(D, e.id);
		e.eye[0] = (q.a.x + q.b.x + q.c.x + q.d.x) / 4.0f; e.eye[1] = (q.a.y + q.b.y + q.c.y + q.d.y) / 4.0f; e.eye[2] = (q.a.z + q.b.z + q.c.z + q.d.z) / 4.0f;
		const V3 n = rcross(vsub(q.b, q.a), vsub(q.d, q.a));
		e.nrm[0] = n.x; e.nrm[1] = n.y; e.nrm[2] = n.z;
		// orthonormal shooter frame for the conservative culls: s = n x u, t = u, f = n (u = v4 - v1 lies in the patch plane)
		const V3 u = vsub(q.d, q.a);
		const V3 sx = rcross(u, n);               // rcross(a, b) = b x a  ->  n x u
		const float ls = rsqrtf(fmaxf(sx.x * sx.x + sx.y * sx.y + sx.z * sx.z, 1e-30f)), lt = rsqrtf(fmaxf(u.x * u.x + u.y * u.y + u.z * u.z, 1e-30f)), lf = rsqrtf(fmaxf(n.x * n.x + n.y * n.y + n.z * n.z, 1e-30f));
		e.ax[0] = sx.x * ls; e.ax[1] = sx.y * ls; e.ax[2] = sx.z * ls;
		e.ax[3] = u.x * lt; e.ax[4] = u.y * lt; e.ax[5] = u.z * lt;
		e.ax[6] = n.x * lf; e.ax[7] = n.y * lf; e.ax[8] = n.z * lf;
	} else e.valid = 0;
	D.em[h] = e;
	D.emlite[2 * h] = make_float4(e.S[0], e.S[1], e.S[2], __uint_as_float(e.valid ? (1u | (e.order << 1)) : 0u));
	D.emlite[2 * h + 1] = make_float4(e.color[0], e.color[1], e.color[2], __uint_as_float(e.id));
	return e;
}
// whole block: slot h's emitter record (thread 0), its five MVPs and the cull margin (threads 0..4: one face each; the
// margin is written into the record afterwards — the culls that read it run in a later kernel)
__device__ __forceinline__ void camera_block(const RadDev& D, uint32_t h, int sel_parity, RadEmitter* s_e) {
	if (threadIdx.x == 0) *s_e = camera_emitter(D, h, sel_parity);
	__syncthreads();
	float dev = 0.0f;
	if (s_e->valid && threadIdx.x < RAD_NFACES) {
		const Quad q = load_quad(D, s_e->id);
		build_mvp(q, threadIdx.x, D.proj, D.mvp + ((size_t)h * RAD_NFACES + threadIdx.x) * 16);
		dev = face_deviation(q, threadIdx.x);
	}
	if (threadIdx.x < 32) {                       // first warp, all of its lanes: max over lanes 0..7 ends up in lane 0
		#pragma unroll
		for (int o = 4; o >= 1; o >>= 1) dev = fmaxf(dev, __shfl_xor_sync(0xFFFFFFFFu, dev, o));
		if (threadIdx.x == 0 && s_e->valid) D.em[h].ctol = margin2_of(dev);
	}
}

} // namespace
