// K1 — five-face hemicube item-buffer rasteriser (replaces the OpenGL render of the reference:
// Shaders.cpp:235-260 patch-view program, DrawPatchLook Main.cpp:689-725, GL state Main.cpp:12,1104,
// 1148-1200, atlas/FBO Main.cpp:271-294) and the per-emitter camera set-up done on the CPU in
// OnIdle (Main.cpp:1161,1172-1183 -> Camera.cpp:19-52,97-103, Transform.cpp:26-46,70-80,127-156).
//
// Pipeline per batch (all on the context's stream):
//   camera_kernel        k x 5 lanes: radiosity snapshot, emitter colour, the five MVPs
//   raster_cull_kernel   one lane per (patch, hemicube): conservative back-face / horizon / per-face frustum culls in
//                        the shooter's frame; surviving (patch, face) pairs are compacted (one atomic per warp)
//   raster_setup_kernel  one lane per surviving pair: vertex transform, near-plane clip, viewport, 8-bit sub-pixel
//                        snap, exact back-face cull, scissored bbox; then
//                          tiny bbox       -> the owning lane walks it alone
//                          anything larger -> parked in the small / chunk queues (slots claimed with one atomic
//                                             per warp and queue)
//   raster_queue_kernel  persistent warps drain the two queues: small triangles four per warp (a quarter warp each,
//                        8x1 pixel rows, int32 walk), chunks one warp each, 8x4 pixels per step
//   resolve_kernel       64-bit keys -> uint32 item buffer (id+1), keys reset for the next batch
// Visibility is a deterministic 64-bit atomicMin of (depth24 << 32 | id+1) per pixel: equal to GL_LESS
// with patches drawn in id order (Main.cpp:715-720), independent of thread scheduling.
//
// Raster rules (identical, operation for operation, to oracle/oracle.cpp): clip = MVP*(p,1) with
// row r = ((m0r*x + m1r*y) + m2r*z) + m3r; near clip z+w>=0 with new vertices interpolated from the
// inside vertex; iw = 1/w, xw = (x*iw)*(N/2) + (vx+N/2); RNE snap to 1/256 px; keep signed area > 0 (CCW front,
// GL_CULL_FACE back); pixel centres, exact int64 edge functions, top-left tie rule; depth =
// barycentric interpolation of (z*iw)*0.5+0.5, RNE-quantised to 24 bit, d >= 0xFFFFFF fails (LESS vs 1.0).
#include <stdio.h>
#include <stdlib.h>
#include "rad_internal.cuh"
#include "camera.cuh"
#include "segadd.cuh"

namespace {

// one block per hemicube slot: lanes 0..4 build the face matrices, lane 0 takes the snapshot (camera.cuh)
__global__ void camera_kernel(RadDev D, int sel_parity, uint32_t slot_base) {
	__shared__ RadEmitter s_e;
	pdl_enter();
	if (D.stop_gate && blockIdx.x == 0 && threadIdx.x == 0 && (D.spec ? D.ctl->spec_done : D.ctl->stopped)) D.ctl->gate = 1;   // the previous batch was the last one (Main.cpp:1137,1298)
	camera_block(D, slot_base + blockIdx.x, sel_parity, &s_e);
}

struct CV { float x, y, z, w; };
__device__ __forceinline__ CV xform(const float* __restrict__ m, V3 p) {
	CV c;
	c.x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12];
	c.y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13];
	c.z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14];
	c.w = ((m[3] * p.x + m[7] * p.y) + m[11] * p.z) + m[15];
	return c;
}
__device__ __noinline__ CV clip_lerp(const CV& in, const CV& out, float din, float dout) {
	float t = din / (din - dout);
	CV r;
	r.x = in.x + t * (out.x - in.x);
	r.y = in.y + t * (out.y - in.y);
	r.z = in.z + t * (out.z - in.z);
	r.w = in.w + t * (out.w - in.w);
	return r;
}
__device__ __forceinline__ int snap256(float v) {
	float s = v * 256.0f;
	s = fminf(fmaxf(s, -536870912.0f), 536870912.0f);
	return __float2int_rn(s);
}
__device__ __forceinline__ long long edge_fn(int ax, int ay, int bx, int by, int cx, int cy) {
	return (long long)(bx - ax) * (long long)(cy - ay) - (long long)(by - ay) * (long long)(cx - ax);
}
__device__ __forceinline__ int edge_bias(int ax, int ay, int bx, int by) {
	int dx = bx - ax, dy = by - ay;
	return (dy < 0 || (dy == 0 && dx < 0)) ? 0 : -1;
}

struct Tri {                 // screen-space triangle ready for coverage
	int X0, Y0, X1, Y1, X2, Y2;
	float z0, dz1, dz2, inv_area;
	float z1, z2;            // raw depths of vertices 1 and 2 (set-up kernel only: small-quad records keep the raw values)
	int bx;                  // px0 | px1 << 16
	int by;                  // py0 | py1 << 16
};
struct PV { int X, Y; float Z; };   // projected, snapped vertex

// Edge functions are exact int64 and stepped incrementally; E[i] carries the top-left bias (0 / -1) so that
// "inside" is (E0 | E1 | E2) >= 0.  Depth uses the unbiased values.  The atomicMin is fire-and-forget (RED.MIN.64):
// nothing in a pixel loop waits on memory.
struct EdgeSet {
	long long e0, e1, e2;        // biased edge values at the current pixel centre
	long long sx0, sx1, sx2;     // step for +1 pixel in x
	long long sy0, sy1, sy2;     // step for +1 pixel in y
	int b1, b2;
};
__device__ __forceinline__ void edges_at(const Tri& t, int px, int py, EdgeSet& E) {
	const int cx = px * 256 + 128, cy = py * 256 + 128;
	const int b0 = edge_bias(t.X1, t.Y1, t.X2, t.Y2);
	E.b1 = edge_bias(t.X2, t.Y2, t.X0, t.Y0); E.b2 = edge_bias(t.X0, t.Y0, t.X1, t.Y1);
	E.e0 = edge_fn(t.X1, t.Y1, t.X2, t.Y2, cx, cy) + b0;
	E.e1 = edge_fn(t.X2, t.Y2, t.X0, t.Y0, cx, cy) + E.b1;
	E.e2 = edge_fn(t.X0, t.Y0, t.X1, t.Y1, cx, cy) + E.b2;
	E.sx0 = -(long long)(t.Y2 - t.Y1) * 256; E.sy0 = (long long)(t.X2 - t.X1) * 256;
	E.sx1 = -(long long)(t.Y0 - t.Y2) * 256; E.sy1 = (long long)(t.X0 - t.X2) * 256;
	E.sx2 = -(long long)(t.Y1 - t.Y0) * 256; E.sy2 = (long long)(t.X1 - t.X0) * 256;
}
__device__ __forceinline__ void shade_covered(const Tri& t, long long e1, long long e2, uint32_t id1, unsigned long long* __restrict__ a, uint32_t tagsh) {
	const float l1 = (float)e1 * t.inv_area, l2 = (float)e2 * t.inv_area;
	float z = (t.z0 + l1 * t.dz1) + l2 * t.dz2;
	z = fminf(fmaxf(z, 0.0f), 1.0f);
	const uint32_t dq = __float2uint_rn(z * 16777215.0f);
	if (dq < 0xFFFFFFu) atomicMin(a, ((unsigned long long)(tagsh | dq) << 32) | id1);
}

// 32-bit form of the same edge functions for triangles that are small around the walk origin: with R = largest
// |vertex - origin| and S = largest pixel-centre offset (sub-pixel units), |E| <= 4 R (R + S), so R (R + S) < 2^29 keeps
// every product and sum inside int32.  Same integers, hence the same coverage and (float)E as the 64-bit walk.
struct EdgeSet32 { int e0, e1, e2, sx0, sx1, sx2, sy0, sy1, sy2, b1, b2; };
__device__ __forceinline__ bool fits32(const Tri& t, int px, int py, int span_px) {
	const int cx = px * 256 + 128, cy = py * 256 + 128;
	const int r = max(max(max(abs(t.X0 - cx), abs(t.Y0 - cy)), max(abs(t.X1 - cx), abs(t.Y1 - cy))), max(abs(t.X2 - cx), abs(t.Y2 - cy)));
	return (long long)r * (long long)(r + span_px * 256) < (1ll << 29);
}
__device__ __forceinline__ void edges_at32(const Tri& t, int px, int py, EdgeSet32& E) {
	const int cx = px * 256 + 128, cy = py * 256 + 128;
	const int x0 = t.X0 - cx, y0 = t.Y0 - cy, x1 = t.X1 - cx, y1 = t.Y1 - cy, x2 = t.X2 - cx, y2 = t.Y2 - cy;
	const int b0 = edge_bias(t.X1, t.Y1, t.X2, t.Y2);
	E.b1 = edge_bias(t.X2, t.Y2, t.X0, t.Y0); E.b2 = edge_bias(t.X0, t.Y0, t.X1, t.Y1);
	// edge_fn(a, b, c) with c at the origin: (bx-ax)(0-ay) - (by-ay)(0-ax)
	E.e0 = (x2 - x1) * (-y1) - (y2 - y1) * (-x1) + b0;
	E.e1 = (x0 - x2) * (-y2) - (y0 - y2) * (-x2) + E.b1;
	E.e2 = (x1 - x0) * (-y0) - (y1 - y0) * (-x0) + E.b2;
	E.sx0 = -(y2 - y1) * 256; E.sy0 = (x2 - x1) * 256;
	E.sx1 = -(y0 - y2) * 256; E.sy1 = (x0 - x2) * 256;
	E.sx2 = -(y1 - y0) * 256; E.sy2 = (x1 - x0) * 256;
}
__device__ __forceinline__ void shade_covered32(const Tri& t, int e1, int e2, uint32_t id1, unsigned long long* __restrict__ a, uint32_t tagsh) {
	const float l1 = (float)e1 * t.inv_area, l2 = (float)e2 * t.inv_area;
	float z = (t.z0 + l1 * t.dz1) + l2 * t.dz2;
	z = fminf(fmaxf(z, 0.0f), 1.0f);
	const uint32_t dq = __float2uint_rn(z * 16777215.0f);
	if (dq < 0xFFFFFFu) atomicMin(a, ((unsigned long long)(tagsh | dq) << 32) | id1);
}

// perspective divide (reciprocal, then multiply), viewport, 8-bit sub-pixel snap
__device__ __forceinline__ PV project(const CV& c, float hw, float ox, float oy) {
	const float iw = 1.0f / c.w;
	PV p;
	p.X = snap256((c.x * iw) * hw + ox);
	p.Y = snap256((c.y * iw) * hw + oy);
	p.Z = (c.z * iw) * 0.5f + 0.5f;
	return p;
}

// cull + scissored bbox + depth plane.  Returns the bbox area in pixels (0 = nothing to draw).
__device__ __forceinline__ int setup_tri(const PV& a, const PV& b, const PV& c, int scx, int scy, int scw, int sch, Tri& t) {
	const long long area2 = edge_fn(a.X, a.Y, b.X, b.Y, c.X, c.Y);
	if (area2 <= 0) return 0;                    // back-facing (CW in window space) or degenerate
	const int minx = min(a.X, min(b.X, c.X)), maxx = max(a.X, max(b.X, c.X));
	const int miny = min(a.Y, min(b.Y, c.Y)), maxy = max(a.Y, max(b.Y, c.Y));
	const int px0 = max((minx - 128 + 255) >> 8, scx), px1 = min((maxx - 128) >> 8, scx + scw - 1);
	const int py0 = max((miny - 128 + 255) >> 8, scy), py1 = min((maxy - 128) >> 8, scy + sch - 1);
	if (px0 > px1 || py0 > py1) return 0;
	t.X0 = a.X; t.Y0 = a.Y; t.X1 = b.X; t.Y1 = b.Y; t.X2 = c.X; t.Y2 = c.Y;
	t.z0 = a.Z; t.dz1 = b.Z - a.Z; t.dz2 = c.Z - a.Z; t.z1 = b.Z; t.z2 = c.Z;
	t.inv_area = 1.0f / (float)area2;
	t.bx = px0 | (px1 << 16); t.by = py0 | (py1 << 16);
	return (px1 - px0 + 1) * (py1 - py0 + 1);
}

#define FULL 0xFFFFFFFFu

// Small-quad record (RadSmallQuad, packed by hand into four 16-byte words so that it never touches local memory):
// vertices relative to the centre of the bbox origin pixel (int16 by the fits32 bound).
struct SmallRec { uint4 a, b, c, d; };
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ SmallRec make_small(int X0, int Y0, int X1, int Y1, int X2, int Y2, int X3, int Y3,
                                                float Z0, float Z1, float Z2, float Z3, float invA, float invB,
                                                uint32_t id1, uint32_t slot, int px0, int py0, int bw, int bh) {
	const int cx = px0 * 256 + 128, cy = py0 * 256 + 128;
	SmallRec r;
	r.a = make_uint4(pack16(X0 - cx, Y0 - cy), pack16(X1 - cx, Y1 - cy), pack16(X2 - cx, Y2 - cy), pack16(X3 - cx, Y3 - cy));
	r.b = make_uint4(__float_as_uint(Z0), __float_as_uint(Z1), __float_as_uint(Z2), __float_as_uint(Z3));
	r.c = make_uint4(__float_as_uint(invA), __float_as_uint(invB), id1, pack16((int)slot, (32768 + bw - 1) / bw));
	r.d = make_uint4(pack16(px0, py0), (uint32_t)bw | ((uint32_t)bh << 8), 0u, 0u);
	return r;
}
__device__ __forceinline__ int radius_about(int cx, int cy, int X, int Y) { return max(abs(X - cx), abs(Y - cy)); }

// The work lists a set-up warp appends to: the lane's share of the context's lists (RadDev::qc ...) or, on the ring path,
// the lists of ONE hemicube slot (all pairs of a warp step belong to the same slot there).
struct QV { RadQueueCtl* qc; RadBigTri* q_tri; RadQueueEntry* q_ent; RadSmallQuad* q_sm; uint32_t q_tri_cap, q_ent_cap, q_sm_cap; };
__device__ __forceinline__ QV qv_lane(const RadDev& D) {
	QV q; q.qc = D.qc; q.q_tri = D.q_tri; q.q_ent = D.q_ent; q.q_sm = D.q_sm; q.q_tri_cap = D.q_tri_cap; q.q_ent_cap = D.q_ent_cap; q.q_sm_cap = D.q_sm_cap;
	return q;
}
__device__ __forceinline__ QV qv_slot(const RadDev& D, const RadRing& R, uint32_t sl) {
	QV q; q.qc = &D.rc->slot[sl];
	q.q_tri = D.q_tri + (size_t)sl * R.cap_tri; q.q_ent = D.q_ent + (size_t)sl * R.cap_ent; q.q_sm = D.q_sm + (size_t)sl * R.cap_sm;
	q.q_tri_cap = R.cap_tri; q.q_ent_cap = R.cap_ent; q.q_sm_cap = R.cap_sm;
	return q;
}

// Warp-collective append to the small-quad queue: ONE atomic per warp.
__device__ __forceinline__ void push_small(const RadDev& D, const QV& Q, bool take, const SmallRec& r, int lane) {
	const unsigned ms = __ballot_sync(FULL, take);
	if (ms == 0) return;
	uint32_t sbase = 0;
	if (lane == 0) sbase = atomicAdd(&Q.qc->q_small, (uint32_t)__popc(ms));
	sbase = __shfl_sync(FULL, sbase, 0);
	if (take) {
		const uint32_t si = sbase + __popc(ms & ((1u << lane) - 1u));
		if (si < Q.q_sm_cap) {
			uint4* dst = reinterpret_cast<uint4*>(Q.q_sm + si);
			dst[0] = r.a; dst[1] = r.b; dst[2] = r.c; dst[3] = r.d;
		} else D.ctl->q_overflow = 1;
	}
}

// Both triangles of an unclipped, front-facing patch whose common bbox is small: parked as ONE record, walked once.
// Returns true when the quad was taken (the caller then skips its two triangles).
__device__ __forceinline__ bool emit_quad(const RadDev& D, const QV& Q, const Tri& ta, const Tri& tb, int areaA, int areaB, int X3, int Y3, float Z3,
                                          uint32_t id1, uint32_t slot, int lane) {
	bool take = false;
	int px0 = 0, py0 = 0, bw = 1, bh = 1;
	if (areaA > 0 && areaB > 0) {
		px0 = min(ta.bx & 0xFFFF, tb.bx & 0xFFFF); py0 = min(ta.by & 0xFFFF, tb.by & 0xFFFF);
		bw = max(ta.bx >> 16, tb.bx >> 16) - px0 + 1; bh = max(ta.by >> 16, tb.by >> 16) - py0 + 1;
		if (bw * bh > (int)D.inline_area && bw * bh <= 8 * (int)D.small_steps && bw <= 255 && bh <= 255) {   // w, h are bytes in the record
			const int cx = px0 * 256 + 128, cy = py0 * 256 + 128;
			const int r = max(max(radius_about(cx, cy, ta.X0, ta.Y0), radius_about(cx, cy, ta.X1, ta.Y1)), max(radius_about(cx, cy, ta.X2, ta.Y2), radius_about(cx, cy, X3, Y3)));
			take = (long long)r * (long long)(r + (max(bw, bh) + 8) * 256) < (1ll << 29);
		}
	}
	if (!__any_sync(FULL, take)) return false;
	SmallRec r;
	if (take) r = make_small(ta.X0, ta.Y0, ta.X1, ta.Y1, ta.X2, ta.Y2, X3, Y3, ta.z0, ta.z1, ta.z2, Z3, ta.inv_area, tb.inv_area, id1, slot, px0, py0, bw, bh);
	push_small(D, Q, take, r, lane);
	return take;
}

// One triangle per lane (area == 0: none).  Small bboxes are walked by the owning lane; everything else is parked:
// short int32-safe walks in the small-quad queue (as a quad whose second triangle is degenerate), the rest in the chunk
// queue — bbox-relative chunks of tile x tile pixels (RadDev::tile), one warp each in raster_queue_kernel — so that no warp of
// the set-up kernel ever carries a long pixel loop (load balance).  Queue slots are claimed with one atomic per warp.
__device__ __forceinline__ void emit_tri(const RadDev& D, const QV& Q, const Tri& tr, int area, uint32_t id1, uint32_t slot, int lane,
                                         unsigned long long* __restrict__ keys) {
	const uint32_t tagsh = D.tag << 24;
	const int bw = (tr.bx >> 16) - (tr.bx & 0xFFFF) + 1, bh = (tr.by >> 16) - (tr.by & 0xFFFF) + 1;
	// tiny bbox: walked here by the owning lane (int32 edge functions; the rare tiny bbox of a triangle whose vertices lie
	// far away goes to the chunk queue, which has the int64 walk)
	const bool tiny = area > 0 && area <= (int)D.inline_area && fits32(tr, tr.bx & 0xFFFF, tr.by & 0xFFFF, max(bw, bh));
	if (tiny) {
		const int px0 = tr.bx & 0xFFFF, px1 = tr.bx >> 16, py0 = tr.by & 0xFFFF, py1 = tr.by >> 16;
		EdgeSet32 E; edges_at32(tr, px0, py0, E);
		for (int py = py0; py <= py1; py++, E.e0 += E.sy0, E.e1 += E.sy1, E.e2 += E.sy2) {
			int e0 = E.e0, e1 = E.e1, e2 = E.e2;
			unsigned long long* row = keys + (size_t)py * D.W;
			for (int px = px0; px <= px1; px++, e0 += E.sx0, e1 += E.sx1, e2 += E.sx2)
				if ((e0 | e1 | e2) >= 0) shade_covered32(tr, e1 - E.b1, e2 - E.b2, id1, row + px, tagsh);
		}
	}
	const bool parked = area > 0 && !tiny;
	const bool small = parked && bw * bh <= 8 * (int)D.small_steps && bw <= 255 && bh <= 255 && fits32(tr, tr.bx & 0xFFFF, tr.by & 0xFFFF, max(bw, bh) + 8);   // +8: lanes start up to 7 px right of the origin
	const bool big = parked && !small;
	const unsigned ms = __ballot_sync(FULL, small), mb = __ballot_sync(FULL, big);
	if ((ms | mb) == 0) return;
	if (ms) {
		SmallRec q;
		if (small) q = make_small(tr.X0, tr.Y0, tr.X1, tr.Y1, tr.X2, tr.Y2, tr.X0, tr.Y0, tr.z0, tr.z1, tr.z2, tr.z0, tr.inv_area, 0.0f,
		                          id1, slot, tr.bx & 0xFFFF, tr.by & 0xFFFF, bw, bh);
		push_small(D, Q, small, q, lane);
	}
	if (mb == 0) return;
	RadBigTri r;
	if (big) {
		r.X0 = tr.X0; r.Y0 = tr.Y0; r.X1 = tr.X1; r.Y1 = tr.Y1; r.X2 = tr.X2; r.Y2 = tr.Y2;
		r.z0 = tr.z0; r.dz1 = tr.dz1; r.dz2 = tr.dz2; r.inv_area = tr.inv_area;
		r.id1 = id1; r.slot = slot;
		r.px0 = tr.bx & 0xFFFF; r.px1 = tr.bx >> 16; r.py0 = tr.by & 0xFFFF; r.py1 = tr.by >> 16;
	}
	int ncx = 0, ncy = 0, nent = 0;
	if (big) { ncx = (bw - 1) / (int)D.tile + 1; ncy = (bh - 1) / (int)D.tile + 1; nent = ncx * ncy; }
	int pre = nent;                               // inclusive warp scan of the entry counts
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, pre, d); if (lane >= d) pre += o; }
	const int total = __shfl_sync(FULL, pre, 31);
	uint32_t tbase = 0, ebase = 0;
	if (lane == 0) { tbase = atomicAdd(&Q.qc->q_tris, (uint32_t)__popc(mb)); ebase = atomicAdd(&Q.qc->q_entries, (uint32_t)total); }
	tbase = __shfl_sync(FULL, tbase, 0); ebase = __shfl_sync(FULL, ebase, 0);
	if (!big) return;
	const uint32_t ti = tbase + __popc(mb & ((1u << lane) - 1u));
	uint32_t e = ebase + (uint32_t)(pre - nent);
	if (ti >= Q.q_tri_cap || e + nent > Q.q_ent_cap) { D.ctl->q_overflow = 1; return; }
	Q.q_tri[ti] = r;
	for (int cy = 0; cy < ncy; cy++)
		for (int cx = 0; cx < ncx; cx++) {
			RadQueueEntry q; q.tri = ti; q.tx = (uint16_t)cx; q.ty = (uint16_t)cy;
			Q.q_ent[e++] = q;
		}
}

// ---- stage 1: conservative culling, one lane per (patch, hemicube) ---------------------------------------------------
// Emits the (patch, face) pairs that MAY produce pixels.  Every test has a wide margin (2e-3 of the distance to the eye
// plus the deviation of the faces' real float32 view bases from the ideal frame, RadEmitter::ctol;
// ~6 degrees for facing) and only ever drops work that the exact stage would drop too:
//  - both triangles clearly face away from the eye: the exact path would get a negative window-space area on every face;
//  - per face, all four vertices clearly outside one side plane of the face's 90-degree frustum, or clearly below the
//    shooter's horizon (clipped by the near plane on FRONT, projected below the scissor on the four side faces).
// The frusta are evaluated in the shooter's orthonormal frame (s = n x u, t = u, f = n; Camera.cpp:19-52): FRONT looks
// along +f, UP/DOWN along +-t, LEFT/RIGHT along +-s, and the scissored half of every side face is the f >= 0 half.
// grid: x = patch chunk, z = local hemicube slot
// SL (ring path): the pairs of every slot go to the slot's own list (RadControl::slot[z].n_pairs, z-th share of D.pairs)
template <bool SL>
__global__ void __launch_bounds__(256) raster_cull_kernel(RadDev D, RadRing R) {
	pdl_enter();
	const uint32_t slot = D.h0 + blockIdx.z;
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (D.stop_gate && p == 0 && blockIdx.z == 0 && (D.spec ? D.ctl->spec_done : D.ctl->stopped)) D.ctl->gate = 1;   // first raster kernel of a batch: latch the stop test (k == 1 has no camera kernel)
	if (D.spec && D.ctl->gate) return;            // speculative path: batches behind the end of the call render nothing (the camera kernel has latched the gate)
	const bool live = p < D.P;
	Quad q;
	if (live) q = load_quad(D, p);
	const RadEmitter em = D.em[slot];
	if (!em.valid) return;
	const int lane = threadIdx.x & 31;
	unsigned faces = 0;                           // bit f set: face f needs the exact stage
	if (live) {
		const V3 eye = mk(em.eye[0], em.eye[1], em.eye[2]);
		const V3 da = vsub(eye, q.a);
		const V3 n1 = rcross(vsub(q.b, q.a), vsub(q.c, q.a)), n2 = rcross(vsub(q.c, q.a), vsub(q.d, q.a));
		const float d2 = da.x * da.x + da.y * da.y + da.z * da.z;
		const float s1 = n1.x * da.x + n1.y * da.y + n1.z * da.z, s2 = n2.x * da.x + n2.y * da.y + n2.z * da.z;
		const float m1 = 0.01f * (n1.x * n1.x + n1.y * n1.y + n1.z * n1.z) * d2, m2 = 0.01f * (n2.x * n2.x + n2.y * n2.y + n2.z * n2.z) * d2;
		const bool back = (s1 < 0.0f && s1 * s1 > m1 || m1 == 0.0f) && (s2 < 0.0f && s2 * s2 > m2 || m2 == 0.0f) && (m1 > 0.0f || m2 > 0.0f);
		if (!back) {
			// g[k] = largest (h_k(v) + margin(v)) over the four vertices: negative <=> all of them are clearly outside plane k
			// (h < -margin; margin(v) = sqrt(ctol) |r_v|, a hair wider than the squared form h^2 > ctol |r|^2 it replaces).
			// planes: 0 c>=0 | 1..4 FRONT c-a,c+a,c-b,c+b | 5..7 UP b-a,b+a,b-c | 8..10 DOWN -b-a,-b+a,-b-c |
			// 11..13 LEFT a-b,a+b,a-c | 14..16 RIGHT -a-b,-a+b,-a-c  (a+b = b+a, -a-b = -b-a, -a+b = b-a: 14 distinct values)
			float g[14];
			#pragma unroll
			for (int k = 0; k < 14; k++) g[k] = -3.0e38f;
			const V3* vv[4] = { &q.a, &q.b, &q.c, &q.d };
			#pragma unroll
			for (int v = 0; v < 4; v++) {
				const V3 r = vsub(*vv[v], eye);
				const float a = r.x * em.ax[0] + r.y * em.ax[1] + r.z * em.ax[2];
				const float b = r.x * em.ax[3] + r.y * em.ax[4] + r.z * em.ax[5];
				const float c = r.x * em.ax[6] + r.y * em.ax[7] + r.z * em.ax[8];
				const float mr = sqrtf(em.ctol * (r.x * r.x + r.y * r.y + r.z * r.z)) * 1.0001f;     // ((2e-3 + 3 dev) |r|), see camera_emitter
				const float am = a + mr, bm = b + mr, cm = c + mr, na = mr - a, nb = mr - b;
				const float h[14] = { cm, cm - a, cm + a, cm - b, cm + b, bm - a, bm + a, bm - c, nb - a, nb + a, nb - c, am - b, am - c, na - c };
				#pragma unroll
				for (int k = 0; k < 14; k++) g[k] = fmaxf(g[k], h[k]);
			}
			const bool below = g[0] < 0.0f;
			if (!below) {
				if (!(g[5] < 0.0f || g[6] < 0.0f || g[7] < 0.0f)) faces |= 1u;                     // UP: b-a, b+a, b-c
				if (!(g[8] < 0.0f || g[9] < 0.0f || g[10] < 0.0f)) faces |= 2u;                    // DOWN: -b-a, -b+a, -b-c
				if (!(g[11] < 0.0f || g[6] < 0.0f || g[12] < 0.0f)) faces |= 4u;                   // LEFT: a-b, a+b, a-c
				if (!(g[8] < 0.0f || g[5] < 0.0f || g[13] < 0.0f)) faces |= 8u;                    // RIGHT: -a-b, -a+b, -a-c
				if (!(g[1] < 0.0f || g[2] < 0.0f || g[3] < 0.0f || g[4] < 0.0f)) faces |= 16u;     // FRONT: c-a, c+a, c-b, c+b
			}
		}
	}
	// pairs of one face are written as one contiguous run per warp; ONE atomic per warp claims the space
	unsigned mf[RAD_NFACES]; int total = 0;
	#pragma unroll
	for (int f = 0; f < RAD_NFACES; f++) { mf[f] = __ballot_sync(FULL, (faces >> f) & 1u); total += __popc(mf[f]); }
	if (total == 0) return;
	uint32_t base = 0;
	if (lane == 0) base = atomicAdd(SL ? &D.rc->slot[blockIdx.z].n_pairs : &D.qc->n_pairs, (uint32_t)total);
	base = __shfl_sync(FULL, base, 0);
	if (base + total > (SL ? R.cap_pairs : D.pairs_cap)) { if (lane == 0) D.ctl->q_overflow = 1; return; }
	uint32_t* __restrict__ list = SL ? D.pairs + (size_t)blockIdx.z * R.cap_pairs : D.pairs;
	#pragma unroll
	for (int f = 0; f < RAD_NFACES; f++) {
		if ((faces >> f) & 1u) list[base + __popc(mf[f] & ((1u << lane) - 1u))] = p | ((uint32_t)f << 23) | ((slot - D.h0) << 26);
		base += __popc(mf[f]);
	}
}

// ---- stage 2: exact set-up, one lane per surviving (patch, face) pair ----------------------------------------------------
struct FaceWin { int scx, scy, scw, sch; float ox, oy; };
// viewport origin and scissor of a face (Main.cpp:314-389)
__device__ __forceinline__ FaceWin face_window(int f, int N) {
	int vpx, vpy; FaceWin w;
	switch (f) {
	case 0: vpx = 0; vpy = N; w.scx = 0; w.scy = N; w.scw = N; w.sch = N / 2; break;
	case 1: vpx = N; vpy = N / 2; w.scx = N; w.scy = N; w.scw = N; w.sch = N / 2; break;
	case 2: vpx = -(N / 2); vpy = 0; w.scx = 0; w.scy = 0; w.scw = N / 2; w.sch = N; break;
	case 3: vpx = N + N / 2; vpy = 0; w.scx = N + N / 2; w.scy = 0; w.scw = N / 2; w.sch = N; break;
	default: vpx = N / 2; vpy = 0; w.scx = N / 2; w.scy = 0; w.scw = N; w.sch = N; break;
	}
	const float hw = (float)N * 0.5f;
	w.ox = (float)vpx + hw; w.oy = (float)vpy + hw;
	return w;
}
__device__ __forceinline__ void load_mvp(const float* __restrict__ mvp, uint32_t slot, int f, float* __restrict__ m) {
	const float4* mp = reinterpret_cast<const float4*>(mvp + ((size_t)slot * RAD_NFACES + f) * 16);
	const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2), m3 = __ldg(mp + 3);
	m[0] = m0.x; m[1] = m0.y; m[2] = m0.z; m[3] = m0.w; m[4] = m1.x; m[5] = m1.y; m[6] = m1.z; m[7] = m1.w;
	m[8] = m2.x; m[9] = m2.y; m[10] = m2.z; m[11] = m2.w; m[12] = m3.x; m[13] = m3.y; m[14] = m3.z; m[15] = m3.w;
}

// Rare path, out of line: triangle t (0: (0,1,2), 1: (0,2,3)) of a patch that crosses the near plane (z + w >= 0).  The
// patch is transformed again here so that the common path does not keep its clip-space vertices alive.  New vertices
// are interpolated from the inside vertex; vertex order as produced by walking the edges 0-1, 1-2, 2-0; the clipped
// polygon (3 or 4 vertices) is drawn as the fan (0,1,2), (0,2,3): sub selects the fan triangle.
__device__ __noinline__ int clip_tri(const float4* __restrict__ v0, const float4* __restrict__ v1, const float4* __restrict__ v2,
                                      const float* __restrict__ mvp, uint32_t h0, int N, uint32_t pair, int t, int sub, Tri* out) {
	const uint32_t p = pair & 0x7FFFFFu, slot = h0 + (pair >> 26);
	const int f = (int)((pair >> 23) & 7u);
	const float hw = (float)N * 0.5f;
	const FaceWin fw = face_window(f, N);
	float m[16]; load_mvp(mvp, slot, f, m);
	const Quad q = load_quad(v0, v1, v2, p);
	const CV in0 = xform(m, q.a), in1 = xform(m, t ? q.c : q.b), in2 = xform(m, t ? q.d : q.c);
	const float d0 = in0.z + in0.w, d1 = in1.z + in1.w, d2 = in2.z + in2.w;
	CV p0, p1, p2, p3; int n = 0;
	switch ((d0 >= 0.0f ? 1 : 0) | (d1 >= 0.0f ? 2 : 0) | (d2 >= 0.0f ? 4 : 0)) {
	case 7: p0 = in0; p1 = in1; p2 = in2; n = 3; break;
	case 1: p0 = in0; p1 = clip_lerp(in0, in1, d0, d1); p2 = clip_lerp(in0, in2, d0, d2); n = 3; break;
	case 2: p0 = clip_lerp(in1, in0, d1, d0); p1 = in1; p2 = clip_lerp(in1, in2, d1, d2); n = 3; break;
	case 4: p0 = clip_lerp(in2, in1, d2, d1); p1 = in2; p2 = clip_lerp(in2, in0, d2, d0); n = 3; break;
	case 3: p0 = in0; p1 = in1; p2 = clip_lerp(in1, in2, d1, d2); p3 = clip_lerp(in0, in2, d0, d2); n = 4; break;
	case 6: p0 = clip_lerp(in1, in0, d1, d0); p1 = in1; p2 = in2; p3 = clip_lerp(in2, in0, d2, d0); n = 4; break;
	case 5: p0 = in0; p1 = clip_lerp(in0, in1, d0, d1); p2 = clip_lerp(in2, in1, d2, d1); p3 = in2; n = 4; break;
	default: break;
	}
	if (n < 3 + sub) return 0;
	const PV a = project(p0, hw, fw.ox, fw.oy), b = project(sub ? p2 : p1, hw, fw.ox, fw.oy), cc = project(sub ? p3 : p2, hw, fw.ox, fw.oy);
	return setup_tri(a, b, cc, fw.scx, fw.scy, fw.scw, fw.sch, *out);
}

// one warp step of the exact stage: pair e per lane (live: the lane has one), records appended to the lists Q
__device__ __forceinline__ void setup_pairs(const RadDev& D, const QV& Q, uint32_t e, bool live, int lane) {
	const int N = (int)D.N;
	const float hw = (float)N * 0.5f;
	{
		const uint32_t p = e & 0x7FFFFFu, slot = D.h0 + (e >> 26);
		const int f = (int)((e >> 23) & 7u);
		const FaceWin fw = face_window(f, N);
		// transform; nin = vertices inside the near plane (4: the common, unclipped case; 1..3: clip path; 0: nothing)
		PV pv[4]; int nin = 0;
		if (live) {
			float m[16]; load_mvp(D.mvp, slot, f, m);
			const Quad q = load_quad(D, p);
			CV c[4];
			c[0] = xform(m, q.a); c[1] = xform(m, q.b); c[2] = xform(m, q.c); c[3] = xform(m, q.d);
			#pragma unroll
			for (int k = 0; k < 4; k++) nin += (c[k].z + c[k].w) >= 0.0f;
			// exact trivial reject of the whole quad (all four inside the near plane, so w > 0): wholly beyond one
			// viewport edge means no snapped vertex can bring a pixel centre inside the scissor
			if (nin == 4 && ((c[0].x > c[0].w && c[1].x > c[1].w && c[2].x > c[2].w && c[3].x > c[3].w) ||
			                 (c[0].x < -c[0].w && c[1].x < -c[1].w && c[2].x < -c[2].w && c[3].x < -c[3].w) ||
			                 (c[0].y > c[0].w && c[1].y > c[1].w && c[2].y > c[2].w && c[3].y > c[3].w) ||
			                 (c[0].y < -c[0].w && c[1].y < -c[1].w && c[2].y < -c[2].w && c[3].y < -c[3].w))) nin = 0;
			if (nin == 4) {
				#pragma unroll
				for (int k = 0; k < 4; k++) pv[k] = project(c[k], hw, fw.ox, fw.oy);
			}
		}
		if (!__any_sync(FULL, nin > 0)) return;
		unsigned long long* __restrict__ keys = D.keys + (size_t)(slot - D.kbase) * D.RES;
		const uint32_t id1 = p + 1;
		const bool clipped = nin > 0 && nin < 4;
		// unclipped patch: both triangles are set up together; a small common bbox is parked as ONE quad record
		Tri trA, trB; int areaA = 0, areaB = 0;
		if (nin == 4) {
			areaA = setup_tri(pv[0], pv[1], pv[2], fw.scx, fw.scy, fw.scw, fw.sch, trA);
			areaB = setup_tri(pv[0], pv[2], pv[3], fw.scx, fw.scy, fw.scw, fw.sch, trB);
		}
		if (emit_quad(D, Q, trA, trB, areaA, areaB, pv[3].X, pv[3].Y, pv[3].Z, id1, slot, lane)) { areaA = 0; areaB = 0; }
		// triangles (0,1,2) and (0,2,3) (ModelContainer.cpp:112-117) on their own, then — rare — the clipped fans
		const int nt = __any_sync(FULL, clipped) ? 6 : 2;
		#pragma unroll 1
		for (int t = 0; t < nt; t++) {
			Tri tr = t ? trB : trA; int area = t == 0 ? areaA : (t == 1 ? areaB : 0);
			if (t >= 2 && clipped) { Tri ct; area = clip_tri(D.v0, D.v1, D.v2, D.mvp, D.h0, N, e, (t - 2) >> 1, (t - 2) & 1, &ct); if (area > 0) tr = ct; }
			if (__any_sync(FULL, area > 0)) emit_tri(D, Q, tr, area, id1, slot, lane, keys);
		}
	}
}

// SL (ring path): the pair lists are per slot; a warp step takes 32 consecutive pairs of ONE slot and appends to that slot's
// work lists.  s_wbase = first warp step of every slot (running sum of ceil(pairs / 32)).
template <int MINB, bool SL>
__global__ void __launch_bounds__(128, MINB) raster_setup_kernel(RadDev D, RadRing R) {
	pdl_enter();
	const int lane = threadIdx.x & 31;
	if (!SL) {
		const uint32_t npairs = min(D.qc->n_pairs, D.pairs_cap);
		const uint32_t stride = gridDim.x * blockDim.x;
		const QV Q = qv_lane(D);
		for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; i0 < npairs; i0 += stride) {
			const uint32_t i = i0 + lane;
			const bool live = i < npairs;
			setup_pairs(D, Q, live ? D.pairs[i] : 0u, live, lane);
		}
		return;
	}
	__shared__ uint32_t s_wbase[RAD_RING_SLOTS + 1];
	if (threadIdx.x < 32) {                        // warp 0: scan of the slots' warp-step counts (two slots per lane)
		const uint32_t a = 2u * lane, b = a + 1u;
		const uint32_t na = a < R.nslots ? (min(D.rc->slot[a].n_pairs, R.cap_pairs) + 31u) >> 5 : 0u;
		const uint32_t nb = b < R.nslots ? (min(D.rc->slot[b].n_pairs, R.cap_pairs) + 31u) >> 5 : 0u;
		uint32_t incl = na + nb;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
		s_wbase[a] = incl - na - nb; s_wbase[b] = incl - nb;
		if (lane == 31) s_wbase[RAD_RING_SLOTS] = incl;
	}
	__syncthreads();
	const uint32_t total = s_wbase[RAD_RING_SLOTS];
	const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t w = gw; w < total; w += nw) {
		uint32_t sl = 0;                           // last slot whose first warp step is <= w
		#pragma unroll
		for (uint32_t d = RAD_RING_SLOTS / 2; d > 0; d >>= 1) if (sl + d < R.nslots && s_wbase[sl + d] <= w) sl += d;
		const uint32_t npairs = min(D.rc->slot[sl].n_pairs, R.cap_pairs);
		const uint32_t i = ((w - s_wbase[sl]) << 5) + lane;
		const bool live = i < npairs;
		const QV Q = qv_slot(D, R, sl);
		setup_pairs(D, Q, live ? D.pairs[(size_t)sl * R.cap_pairs + i] : 0u, live, lane);
	}
}

// biased edge function of (a -> b) at the origin (vertices are origin-relative) and its per-pixel steps; same integers
// as edges_at32
__device__ __forceinline__ void edge_origin(int ax, int ay, int bx, int by, int& e, int& sx, int& sy, int& bias) {
	const int dx = bx - ax, dy = by - ay;
	bias = (dy < 0 || (dy == 0 && dx < 0)) ? 0 : -1;
	e = dx * (-ay) - dy * (-ax) + bias;
	sx = -dy * 256; sy = dx * 256;
}

// The two walks of the parked records, shared by raster_queue_kernel (whole-batch key buffers) and raster_ring_kernel (ring).
//  1. small quads: FOUR records per warp step, one quarter warp (8 lanes) each.  A record holds both triangles of a
//     patch, A = (0,1,2) and B = (0,2,3); the common bbox is walked ONCE as one linear run of w * ceil(h/2) two-row
//     columns, 8 per step (16 pixels per quarter warp and step),
//     with the five distinct int32 edge functions (the diagonal is shared: B's edge (0->2) is minus A's edge (2->0)).
//     A pixel belongs to A or to B (never both: they lie on opposite sides of the diagonal and the top-left rule gives
//     the diagonal itself to exactly one) and takes its depth from that triangle's plane — the same integers and the
//     same float operations as two separate triangle walks, in about half the pixel visits;
//  2. chunks: one warp per (triangle, chunk) of at most tile x tile pixels, 8x4 pixels per step.
// RING: every record of the list belongs to the slot whose key buffer is `ring_keys`; otherwise the record names its slot.
// (Measured and dropped: the next step's records fetched into shared memory with cp.async during the walk — 0.371 against
// 0.331 ms per batch; both rows' keys computed first and the two REDs issued back to back from one asm block, against the
// write-after-read waits on the REDs' operand registers — 0.342 against 0.333 ms; prefetch.global.L2 of the next step's records
// — 0.338 against 0.334 ms; every append of the set-up kernel ordered by walk length — no change in the walk, set-up slower.)
template <bool RING>
__device__ __forceinline__ void walk_small4(const RadDev& D, const uint4* __restrict__ qsm, uint32_t base, uint32_t nsm,
                                            unsigned long long* __restrict__ ring_keys, uint32_t tagsh, int lane) {
	const int sub = lane >> 3, l8 = lane & 7;
	const int W = (int)D.W;
	const uint32_t i = base + sub;
	int npx = 0, w8 = 1, hh = 0, q8 = 0, r8 = 0, x = 0, y = 0;
	int a0y2 = 0, a1y2 = 0, a2y2 = 0, b0y2 = 0, b1y2 = 0;
	int a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0;                         // biased edge values at the bbox origin
	int a0x = 0, a1x = 0, a2x = 0, b0x = 0, b1x = 0, a0y = 0, a1y = 0, a2y = 0, b0y = 0, b1y = 0;
	int bA1 = 0, bA2 = 0, bB1 = 0, bB2 = 0;
	float Z0 = 0, dA1 = 0, dA2 = 0, dB2 = 0, invA = 0, invB = 0;
	uint32_t id1 = 0; unsigned long long* kp = nullptr;
	if (i < nsm) {
		const uint4 r0 = __ldg(qsm + 4 * (size_t)i), r1 = __ldg(qsm + 4 * (size_t)i + 1), r2 = __ldg(qsm + 4 * (size_t)i + 2);
		const uint2 r3 = __ldg(reinterpret_cast<const uint2*>(qsm + 4 * (size_t)i + 3));
		const int x0 = (int)(r0.x << 16) >> 16, y0 = (int)r0.x >> 16, x1 = (int)(r0.y << 16) >> 16, y1 = (int)r0.y >> 16;
		const int x2 = (int)(r0.z << 16) >> 16, y2 = (int)r0.z >> 16, x3 = (int)(r0.w << 16) >> 16, y3 = (int)r0.w >> 16;
		int bA0, bB0;
		edge_origin(x1, y1, x2, y2, a0, a0x, a0y, bA0);      // A: edges (1->2), (2->0), (0->1)
		edge_origin(x2, y2, x0, y0, a1, a1x, a1y, bA1);
		edge_origin(x0, y0, x1, y1, a2, a2x, a2y, bA2);
		edge_origin(x2, y2, x3, y3, b0, b0x, b0y, bB0);      // B: edges (2->3), (3->0), (0->2)
		edge_origin(x3, y3, x0, y0, b1, b1x, b1y, bB1);
		{ const int dx = x2 - x0, dy = y2 - y0; bB2 = (dy < 0 || (dy == 0 && dx < 0)) ? 0 : -1; }
		Z0 = __uint_as_float(r1.x);
		dA1 = __uint_as_float(r1.y) - Z0; dA2 = __uint_as_float(r1.z) - Z0; dB2 = __uint_as_float(r1.w) - Z0;
		invA = __uint_as_float(r2.x); invB = __uint_as_float(r2.y);
		id1 = r2.z;
		const uint32_t slot = r2.w & 0xFFFFu, rcpw = r2.w >> 16;
		const int px0 = (int)(r3.x & 0xFFFFu), py0 = (int)(r3.x >> 16);
		w8 = (int)(r3.y & 0xFFu);
		hh = (int)((r3.y >> 8) & 0xFFu);
		npx = w8 * ((hh + 1) >> 1);                               // positions of the walk: every one is a column of TWO rows
		q8 = (int)((8u * rcpw) >> 15); r8 = 8 - q8 * w8;          // 8 / w, 8 % w
		y = (int)(((uint32_t)l8 * rcpw) >> 15); x = l8 - y * w8;   // this lane's first position (y counts row pairs)
		kp = (RING ? ring_keys : D.keys + (size_t)(slot - D.kbase) * D.RES) + (size_t)py0 * D.W + px0;
		a0y2 = 2 * a0y; a1y2 = 2 * a1y; a2y2 = 2 * a2y; b0y2 = 2 * b0y; b1y2 = 2 * b1y;
	}
	int msteps = (npx + 7) >> 3;
	msteps = max(msteps, __shfl_xor_sync(FULL, msteps, 8)); msteps = max(msteps, __shfl_xor_sync(FULL, msteps, 16));
	int idx = l8;
	// one covered-pixel test + depth + RED; the unbiased A edge (2->0) is minus B's edge (0->2)
	auto pixel = [&](int e0, int e1, int e2, int f0, int f1, int off) {
		const int e1u = e1 - bA1;
		const int f2 = bB2 - e1u;
		const bool inA = (e0 | e1 | e2) >= 0, inB = (f0 | f1 | f2) >= 0;
		if (inA || inB) {
			const float inv = inA ? invA : invB;
			const float l1 = (float)(inA ? e1u : f1 - bB1) * inv, l2 = (float)(inA ? e2 - bA2 : -e1u) * inv;
			float z = (Z0 + l1 * (inA ? dA1 : dA2)) + l2 * (inA ? dA2 : dB2);
			z = fminf(fmaxf(z, 0.0f), 1.0f);
			const uint32_t dq = __float2uint_rn(z * 16777215.0f);
			if (dq < 0xFFFFFFu) atomicMin(kp + off, ((unsigned long long)(tagsh | dq) << 32) | id1);
		}
	};
	for (int s = 0; s < msteps; s++) {
		if (idx < npx) {
			// rows 2y and 2y + 1 of column x: the second row's edge values are one add away from the first's
			const int e0 = a0 + x * a0x + y * a0y2, e1 = a1 + x * a1x + y * a1y2, e2 = a2 + x * a2x + y * a2y2;
			const int f0 = b0 + x * b0x + y * b0y2, f1 = b1 + x * b1x + y * b1y2;
			const int off = 2 * y * W + x;
			pixel(e0, e1, e2, f0, f1, off);
			if (2 * y + 1 < hh) pixel(e0 + a0y, e1 + a1y, e2 + a2y, f0 + b0y, f1 + b1y, off + W);
		}
		idx += 8; x += r8; y += q8;
		if (x >= w8) { x -= w8; y++; }
	}
}

template <bool RING>
__device__ __forceinline__ void walk_chunk(const RadDev& D, const RadQueueEntry e, const RadBigTri* __restrict__ q_tri, uint32_t q_tri_cap,
                                           unsigned long long* __restrict__ ring_keys, uint32_t tagsh, int lane) {
	if (e.tri >= q_tri_cap) return;
	const RadBigTri r = q_tri[e.tri];
	Tri w;
	w.X0 = r.X0; w.Y0 = r.Y0; w.X1 = r.X1; w.Y1 = r.Y1; w.X2 = r.X2; w.Y2 = r.Y2;
	w.z0 = r.z0; w.dz1 = r.dz1; w.dz2 = r.dz2; w.inv_area = r.inv_area;
	const int T = (int)D.tile;
	const int px0 = r.px0 + (int)e.tx * T, px1 = min(r.px1, px0 + T - 1);
	const int py0 = r.py0 + (int)e.ty * T, py1 = min(r.py1, py0 + T - 1);
	unsigned long long* __restrict__ keys = RING ? ring_keys : D.keys + (size_t)(r.slot - D.kbase) * D.RES;
	// this lane's first pixel; steps of 8 pixels in x and 4 in y
	const int lx = px0 + (lane & 7), ly = py0 + (lane >> 3);
	if (fits32(w, px0, py0, T)) {          // warp-uniform: small triangle around this chunk -> int32 walk
		EdgeSet32 E; edges_at32(w, lx, ly, E);
		for (int py = ly; py <= py1; py += 4, E.e0 += 4 * E.sy0, E.e1 += 4 * E.sy1, E.e2 += 4 * E.sy2) {
			int e0 = E.e0, e1 = E.e1, e2 = E.e2;
			unsigned long long* row = keys + (size_t)py * D.W;
			for (int px = lx; px <= px1; px += 8, e0 += 8 * E.sx0, e1 += 8 * E.sx1, e2 += 8 * E.sx2)
				if ((e0 | e1 | e2) >= 0) shade_covered32(w, e1 - E.b1, e2 - E.b2, r.id1, row + px, tagsh);
		}
		return;
	}
	EdgeSet E; edges_at(w, lx, ly, E);
	// a triangle spread over several chunks: skip the chunk when one edge has all four corner pixels outside
	if (r.px1 - r.px0 >= T || r.py1 - r.py0 >= T) {
		// corner values from lane 0's origin value: e(px0,py0) + dx*sx + dy*sy
		const long long dx = px1 - px0, dy = py1 - py0;
		const long long c0 = __shfl_sync(FULL, E.e0, 0), c1 = __shfl_sync(FULL, E.e1, 0), c2 = __shfl_sync(FULL, E.e2, 0);
		const bool out = (c0 < 0 && c0 + dx * E.sx0 < 0 && c0 + dy * E.sy0 < 0 && c0 + dx * E.sx0 + dy * E.sy0 < 0) ||
		                 (c1 < 0 && c1 + dx * E.sx1 < 0 && c1 + dy * E.sy1 < 0 && c1 + dx * E.sx1 + dy * E.sy1 < 0) ||
		                 (c2 < 0 && c2 + dx * E.sx2 < 0 && c2 + dy * E.sy2 < 0 && c2 + dx * E.sx2 + dy * E.sy2 < 0);
		if (out) return;
	}
	for (int py = ly; py <= py1; py += 4, E.e0 += 4 * E.sy0, E.e1 += 4 * E.sy1, E.e2 += 4 * E.sy2) {
		long long e0 = E.e0, e1 = E.e1, e2 = E.e2;
		unsigned long long* row = keys + (size_t)py * D.W;
		for (int px = lx; px <= px1; px += 8, e0 += 8 * E.sx0, e1 += 8 * E.sx1, e2 += 8 * E.sx2)
			if ((e0 | e1 | e2) >= 0) shade_covered(w, e1 - E.b1, e2 - E.b2, r.id1, row + px, tagsh);
	}
}

// persistent warps drain both queues of the launch (whole-batch key buffers)
__global__ void __launch_bounds__(128) raster_queue_kernel(RadDev D) {
	pdl_enter();
	const int lane = threadIdx.x & 31;
	const uint32_t tagsh = D.tag << 24;
	const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	const uint32_t nsm = min(D.qc->q_small, D.q_sm_cap);
	const uint4* __restrict__ qsm = reinterpret_cast<const uint4*>(D.q_sm);
	for (uint32_t base = gw * 4; base < nsm; base += nw * 4) walk_small4<false>(D, qsm, base, nsm, nullptr, tagsh, lane);
	const uint32_t nent = min(D.qc->q_entries, D.q_ent_cap);
	for (uint32_t i = gw; i < nent; i += nw) walk_chunk<false>(D, D.q_ent[i], D.q_tri, D.q_tri_cap, nullptr, tagsh, lane);
}

// ---- ring path: walk + ProcessHemicube in one persistent kernel, stage by stage through L2-resident key buffers ---------
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
	uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// The warp waits until stage `st` is ready (flag set by the last CTA that finished it).  s_seen caches, per CTA, how many
// stages are known to be ready.  false: the watchdog fired — somebody waited for two seconds — give up.
__device__ __forceinline__ bool ring_wait(const uint32_t* flag, uint32_t st, uint32_t* s_seen, uint32_t* abort_flag, int lane) {
	int ok = 1;
	if (lane == 0 && *reinterpret_cast<volatile uint32_t*>(s_seen) <= st) {
		uint32_t spins = 0, ns = 200; unsigned long long t0 = 0;
		while (ld_acquire_u32(flag) == 0u) {
			if (*reinterpret_cast<volatile uint32_t*>(s_seen) > st) break;      // a warp of this CTA has seen it
			__nanosleep(ns); if (ns < 1600) ns += ns;
			if ((++spins & 63u) == 0u) {
				if (*reinterpret_cast<volatile uint32_t*>(abort_flag)) { ok = 0; break; }
				const unsigned long long t = global_ns();
				if (t0 == 0) t0 = t; else if (t - t0 > 2000000000ull) { *reinterpret_cast<volatile uint32_t*>(abort_flag) = 1u; ok = 0; break; }
			}
		}
		if (ok) { __threadfence_block(); atomicMax(s_seen, st + 1u); }
	}
	__threadfence_block();
	return __shfl_sync(FULL, ok, 0) != 0;
}
// This warp has finished the stage (all its REDs / key reads are done).  The last warp of the CTA counts the CTA in; the
// last CTA raises the stage's ready flag.  fence.acq_rel, not the sequentially consistent __threadfence(): every hop is a
// release (writes before it) / acquire (the counter it has just read) pair.
__device__ __forceinline__ void fence_ar_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void ring_signal(uint32_t* s_cnt, uint32_t* cnt, uint32_t* flag, uint32_t nctas, int lane) {
	__syncwarp();
	if (lane == 0) {
		fence_ar_gpu();
		if (atomicAdd(s_cnt, 1u) == (blockDim.x >> 5) - 1u) {
			fence_ar_gpu();
			if (atomicAdd(cnt, 1u) == nctas - 1u) { fence_ar_gpu(); *reinterpret_cast<volatile uint32_t*>(flag) = 1u; }
		}
	}
}

// CTAs [0, R.walk_ctas) walk, the others process.  Stage st = slots [st * sg, (st + 1) * sg) of the launch, key buffers
// (st % rs) * sg ... of the ring, epoch tag tag0 - st / rs.  Work inside a stage is dealt out statically: item g (counted
// over the whole launch, so that the remainders do not always hit the same warps) belongs to warp g % nw.
//   walk:    wait until the process warps have finished stage st - rs (the buffers' previous user), walk, signal walk_done[st]
//   process: wait until all walk warps have signalled walk_done[st], resolve + ProcessHemicube, signal proc_done[st]
// Every warp of the grid must be resident for this to make progress: the launch is cooperative (co-residency checked by
// the driver) and every wait has a watchdog.  Keys are read with ld.global.cg: the ring re-uses addresses inside one
// launch, L1 must not serve a line of the previous round.
template <bool KEEP>
__global__ void __launch_bounds__(128, 8) raster_ring_kernel(RadDev D, RadRing R) {
	__shared__ uint32_t s_nsm[RAD_RING_SLOTS], s_nent[RAD_RING_SLOTS], s_valid[RAD_RING_SLOTS], s_cnt[RAD_RING_SLOTS], s_seen;
	if (D.stop_gate && D.ctl->gate) return;                           // the stop test fired in an earlier batch of this replay
	if (threadIdx.x == 0) s_seen = 0;
	for (uint32_t t = threadIdx.x; t < RAD_RING_SLOTS; t += blockDim.x) s_cnt[t] = 0;
	for (uint32_t t = threadIdx.x; t < R.nslots; t += blockDim.x) {
		s_nsm[t] = min(D.rc->slot[t].q_small, R.cap_sm); s_nent[t] = min(D.rc->slot[t].q_entries, R.cap_ent);
		s_valid[t] = D.em[D.h0 + t].valid;
	}
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const bool walker = blockIdx.x < R.walk_ctas;
	const uint32_t nw = walker ? R.nw_walk : R.nw_proc;
	const uint32_t gw = (walker ? blockIdx.x : blockIdx.x - R.walk_ctas) * (blockDim.x >> 5) + (threadIdx.x >> 5);
	uint32_t rot = 0;                                                   // items dealt so far, mod nw
	const bool stamp = (R.debug & 16u) && threadIdx.x == 0;
	if (stamp) atomicMin(&D.rc->t_start, global_ns());
	unsigned long long waited = 0;
	if (walker) {
		for (uint32_t st = 0; st < R.nst; st++) {
			const unsigned long long tw0 = stamp ? global_ns() : 0ull;
			if (st >= R.rs && !(R.debug & 1u) && !ring_wait(&D.rc->proc_ready[st - R.rs][0], st - R.rs, &s_seen, &D.ctl->ring_abort, lane)) return;
			if (stamp) waited += global_ns() - tw0;
			const uint32_t tagsh = (R.tag0 - st / R.rs) << 24;
			const uint32_t s1 = min((st + 1) * R.sg, R.nslots);
			for (uint32_t sl = st * R.sg; sl < s1; sl++) {
				unsigned long long* __restrict__ keys = D.keys + (size_t)((st % R.rs) * R.sg + (sl - st * R.sg)) * D.RES;
				const uint32_t nsm = s_nsm[sl], ws = (nsm + 3u) >> 2;
				const uint4* __restrict__ qsm = reinterpret_cast<const uint4*>(D.q_sm + (size_t)sl * R.cap_sm);
				for (uint32_t j = gw >= rot ? gw - rot : gw + nw - rot; j < ws; j += nw) walk_small4<true>(D, qsm, 4u * j, nsm, keys, tagsh, lane);
				rot = (rot + ws) % nw;
				const uint32_t nent = s_nent[sl];
				const RadQueueEntry* __restrict__ qe = D.q_ent + (size_t)sl * R.cap_ent;
				for (uint32_t j = gw >= rot ? gw - rot : gw + nw - rot; j < nent; j += nw)
					walk_chunk<true>(D, qe[j], D.q_tri + (size_t)sl * R.cap_tri, R.cap_tri, keys, tagsh, lane);
				rot = (rot + nent) % nw;
			}
			ring_signal(&s_cnt[st], &D.rc->walk_cnt[st][0], &D.rc->walk_ready[st][0], R.walk_ctas, lane);
		}
		if (stamp) { atomicMax(&D.rc->t_walk_end, global_ns()); atomicAdd(&D.rc->t_walk_wait, waited); }
		return;
	}
	const uint32_t nsteps = D.RES >> 7;                                 // 128 pixels per warp step
	const float4* __restrict__ ff4 = reinterpret_cast<const float4*>(D.ff);
	for (uint32_t st = 0; st < R.nst; st++) {
		const unsigned long long tw0 = stamp ? global_ns() : 0ull;
		if (!ring_wait(&D.rc->walk_ready[st][0], st, &s_seen, &D.ctl->ring_abort, lane)) return;
		if (stamp) waited += global_ns() - tw0;
		const uint32_t tag_end = (R.tag0 - st / R.rs + 1u) << 24;      // tags only decrease and a minimum survives: mine iff high word < (tag + 1) << 24
		const uint32_t s1 = min((st + 1) * R.sg, R.nslots);
		for (uint32_t sl = st * R.sg; sl < s1; sl++) {
			if (!s_valid[sl]) continue;                                 // NULL emitters render nothing (Main.cpp:1253)
			const uint32_t slot = D.h0 + sl;
			const uint4* __restrict__ keys4 = reinterpret_cast<const uint4*>(D.keys + (size_t)((st % R.rs) * R.sg + (sl - st * R.sg)) * D.RES);
			float* __restrict__ F = D.F + (size_t)slot * D.P;
			uint4* __restrict__ items4 = reinterpret_cast<uint4*>(D.items + (size_t)slot * D.RES);
			const uint32_t pairs = (nsteps + 1u) >> 1;                   // two warp steps per item: their loads are in flight together
			for (uint32_t j = gw >= rot ? gw - rot : gw + nw - rot; j < pairs; j += nw) {
				uint4 id[2]; float4 v[2];
				#pragma unroll
				for (int u = 0; u < 2; u++) {
					const uint32_t g = 2u * j + u;
					id[u] = make_uint4(0u, 0u, 0u, 0u); v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
					if (g < nsteps) {
						const uint32_t q = (g << 5) + lane;               // this lane's group of four pixels
						const uint4 k0 = __ldcg(keys4 + 2 * (size_t)q), k1 = __ldcg(keys4 + 2 * (size_t)q + 1);
						id[u] = make_uint4(k0.y < tag_end ? k0.x : 0u, k0.w < tag_end ? k0.z : 0u, k1.y < tag_end ? k1.x : 0u, k1.w < tag_end ? k1.z : 0u);
						v[u] = __ldg(ff4 + q);
					}
				}
				#pragma unroll
				for (int u = 0; u < 2; u++) {
					const uint32_t g = 2u * j + u;
					if (g < nsteps) {
						if (KEEP) items4[(g << 5) + lane] = id[u];
						process4(id[u], v[u], lane, F, D.P);
					}
				}
			}
			rot = (rot + pairs) % nw;
		}
		ring_signal(&s_cnt[st], &D.rc->proc_cnt[st][0], &D.rc->proc_ready[st][0], gridDim.x - R.walk_ctas, lane);
	}
	if (stamp) { atomicMax(&D.rc->t_proc_end, global_ns()); atomicAdd(&D.rc->t_proc_wait, waited); }
}

// recycles the chunk queue between hemicube groups of one batch (see rad_launch_raster)
__global__ void queue_reset_kernel(RadDev D, int first_group) {
	if (threadIdx.x == 0) {
		D.qc->parked = (first_group ? 0u : D.qc->parked) + D.qc->q_tris + D.qc->q_small;
		D.qc->q_tris = 0; D.qc->q_entries = 0; D.qc->q_small = 0; D.qc->n_pairs = 0;
	}
}

// keys -> item buffer (id+1, 0 = empty).  A key belongs to this render iff its top byte equals the launch's epoch tag
// (see RadDev::tag), so nothing is ever cleared in the steady state.  Also recycles the queues.
__global__ void __launch_bounds__(256) resolve_kernel(RadDev D) {
	const uint32_t slot = D.h0 + blockIdx.y;
	if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && (D.qc->q_tris | D.qc->q_small | D.qc->n_pairs)) { D.qc->parked = D.qc->q_tris + D.qc->q_small; D.qc->q_tris = 0; D.qc->q_entries = 0; D.qc->q_small = 0; D.qc->n_pairs = 0; }
	const unsigned long long* __restrict__ keys = D.keys + (size_t)(slot - D.kbase) * D.RES;
	uint32_t* __restrict__ items = D.items + (size_t)slot * D.RES;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < D.RES; i += gridDim.x * blockDim.x) {
		const unsigned long long k = keys[i];
		items[i] = (uint32_t)(k >> 56) == D.tag ? (uint32_t)(k & 0xFFFFFFFFull) : 0u;
	}
}

__global__ void __launch_bounds__(256) clear_keys_kernel(unsigned long long* __restrict__ keys, size_t n) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) keys[i] = RAD_CLEAR_KEY;
}

__global__ void read_depth_kernel(const unsigned long long* __restrict__ keys, uint32_t* __restrict__ out, uint32_t n, uint32_t tag) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { unsigned long long k = keys[i]; out[i] = (uint32_t)(k >> 56) == tag ? (uint32_t)(k >> 32) & 0xFFFFFFu : 0xFFFFFFu; }
}

// L2 atomic roofline micro-benchmark (rad_bench_atomics): the rasteriser's RED.MIN.64 with nothing around it
__global__ void __launch_bounds__(128) atomic_bench_kernel(unsigned long long* __restrict__ keys, uint32_t W, uint32_t H, uint32_t nslots,
                                                            uint32_t pattern, uint32_t steps) {
	const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
	const size_t RES = (size_t)W * H;
	uint32_t h = (pattern == 2 ? gt : (pattern == 1 ? gt >> 5 : gt >> 3)) * 2654435761u + 12345u;
	for (uint32_t s = 0; s < steps; s += 8) {
		h = h * 1664525u + 1013904223u;
		const uint32_t slot = (h >> 8) % nslots;
		uint32_t x = (h >> 3) % (W - 40), y = (h >> 17) % (H - 8);
		x += pattern == 1 ? lane : (pattern == 0 ? (lane & 7) : 0);
		unsigned long long* a = keys + slot * RES + (size_t)y * W + x;
		#pragma unroll
		for (int r = 0; r < 8; r++) atomicMin(a + (size_t)r * W, ((unsigned long long)(0xFE000000u | (h & 0xFFFFFFu)) << 32) | gt);
	}
}

} // namespace

void rad_launch_atomic_bench(rad_ctx* c, uint32_t pattern, uint32_t steps, uint32_t blocks) {
	const RadDev& D = c->d;
	atomic_bench_kernel<<<blocks, 128, 0, c->stream>>>(D.keys, D.W, D.H, D.k, pattern, steps);
	c->launches++;
}

void rad_launch_camera(rad_ctx* c, int sel_parity) {
	rad_launch_pdl(c->pdl, camera_kernel, dim3(c->cam_count ? c->cam_count : c->d.k), dim3(32), 0, c->stream, c->d, sel_parity, c->cam_count ? c->cam_base : 0u);
	c->launches++;
}

// hemicube slots rendered per set-up launch: bounded so that the chunk queue cannot overflow even if every patch
// parked both of its triangles (in practice well under half of them do)
static uint32_t queue_group(const RadDev& D) {
	const uint32_t nslots = D.h1 - D.h0;
	uint64_t g = (uint64_t)min(min(D.q_tri_cap, D.q_sm_cap) / 2u, D.pairs_cap / 5u) / (D.P ? D.P : 1);
	if (g < 1) g = 1;
	if (g > 64) g = 64;                       // the pair list carries the group-local slot in 6 bits
	return g > nslots ? nslots : (uint32_t)g;
}
// View of the device block for raster lane `lane` of `L`: its own counters and its own share of the work lists.  The
// lists are sized for the worst case of the whole batch, so a lane's share covers its share of the slots.
static RadDev lane_view(const rad_ctx* c, uint32_t lane, uint32_t L) {
	RadDev D = c->d;
	D.qc = &c->d.ctl->lane[lane];
	if (L > 1) {
		const uint32_t tc = D.q_tri_cap / L, ec = D.q_ent_cap / L, sc = D.q_sm_cap / L, pc = D.pairs_cap / L;
		D.q_tri += (size_t)lane * tc; D.q_tri_cap = tc;
		D.q_ent += (size_t)lane * ec; D.q_ent_cap = ec;
		D.q_sm += (size_t)lane * sc; D.q_sm_cap = sc;
		D.pairs += (size_t)lane * pc; D.pairs_cap = pc;
	}
	return D;
}

// exact stage: persistent grid over the surviving pairs (their number is only known on the device)
template <bool SL>
static void launch_setup_kernel(const RadDev& D, const RadRing& R, uint32_t n, cudaStream_t st, bool pdl = false) {
	uint64_t want = ((uint64_t)D.P * n * 2 + 127) / 128;      // typically ~1 of 5 (patch, face) pairs survives
	static const int sctas = [] { const char* e = getenv("RAD_SETUP_CTAS"); const int v = e ? atoi(e) : 3; return v < 1 ? 1 : (v > 16 ? 16 : v); }();   // tuning knob: CTAs per SM of the set-up grid (3 = what fits; more only queue up behind the other lanes)
	const uint32_t blocks = (uint32_t)(want < 148 ? 148 : (want > 148u * sctas ? 148u * sctas : want));
	static const int minb = [] { const char* e = getenv("RAD_SETUP_MINB"); const int v = e ? atoi(e) : 3; return v < 3 ? 3 : (v > 5 ? 5 : v); }();   // tuning knob: resident CTAs per SM the set-up kernel is compiled for
	if (minb == 3) rad_launch_pdl(pdl, raster_setup_kernel<3, SL>, dim3(blocks), dim3(128), 0, st, D, R);
	else if (minb == 4) rad_launch_pdl(pdl, raster_setup_kernel<4, SL>, dim3(blocks), dim3(128), 0, st, D, R);
	else rad_launch_pdl(pdl, raster_setup_kernel<5, SL>, dim3(blocks), dim3(128), 0, st, D, R);
}

// slots [D.h0 + s0, +n) of the view V on stream st
static void launch_setup(rad_ctx* c, const RadDev& V, cudaStream_t st, uint32_t s0, uint32_t n, uint32_t kbase, bool tiles = false) {
	RadDev D = V;
	D.h0 = V.h0 + s0; D.h1 = D.h0 + n; D.kbase = kbase;
	// only bboxes of a few pixels are walked by the set-up lane itself; everything else goes through the balanced queues
	// measured: with a few pixels per patch (1 M patches at hemicube 1024) the queues' per-record overhead is not worth it
	// and the owning lane walks bboxes up to 64 px itself; from ~8 pixels per patch on, everything above 2 px is parked
	if (!c->inline_area_forced) D.inline_area = D.RES / (D.P ? D.P : 1u) >= 8u ? 2u : 64u;
	// tile-binned form (raster_tiles.cu): every triangle is parked (nothing touches the global key buffer), a large
	// triangle gets one chunk entry (unused there: the bins cut it by atlas tile)
	if (tiles) { D.inline_area = 0u; D.tile = 1u << 15; }
	const RadRing R0 = {};
	rad_launch_pdl(c->pdl, raster_cull_kernel<false>, dim3((D.P + 255) / 256, 1, n), dim3(256), 0, st, D, R0);
	// exact stage: persistent grid over the surviving pairs (their number is only known on the device)
	launch_setup_kernel<false>(D, R0, n, st, c->pdl);
	c->launches += 2;
}
static void launch_chunks(rad_ctx* c, const RadDev& V, cudaStream_t st, uint32_t kbase) {
	RadDev D = V;
	D.kbase = kbase;
	static const int ctas = [] { const char* e = getenv("RAD_QUEUE_CTAS"); const int v = e ? atoi(e) : 8; return v < 1 ? 1 : (v > 8 ? 8 : v); }();   // tuning knob: persistent CTAs per SM
	rad_launch_pdl(c->pdl, raster_queue_kernel, dim3(148 * ctas), dim3(128), 0, st, D);
	c->launches++;
}

void rad_launch_raster_setup_only(rad_ctx* c) {
	if (c->d.h1 == c->d.h0) return;
	c->d.tag = rad_next_tag(c);               // staged path: every slot has its own key buffer, one tag for the render
	launch_setup(c, c->d, c->stream, 0, queue_group(c->d), 0);
}

void rad_launch_raster_tiles_only(rad_ctx* c) {
	if (c->d.h1 == c->d.h0) return;
	launch_chunks(c, c->d, c->stream, 0);
	// remaining hemicube groups of the batch (only when the batch does not fit the chunk queue at once)
	const uint32_t nslots = c->d.h1 - c->d.h0, g = queue_group(c->d);
	for (uint32_t s0 = g; s0 < nslots; s0 += g) {
		queue_reset_kernel<<<1, 32, 0, c->stream>>>(c->d, s0 == g ? 1 : 0);
		c->launches++;
		launch_setup(c, c->d, c->stream, s0, nslots - s0 < g ? nslots - s0 : g, 0);
		launch_chunks(c, c->d, c->stream, 0);
	}
}

// staged API: all slots rendered into their own key buffers (kbase = 0: slot s uses key buffer s)
void rad_launch_raster(rad_ctx* c) {
	rad_launch_raster_setup_only(c);
	rad_launch_raster_tiles_only(c);
}

void rad_launch_process_view(rad_ctx* c, const RadDev& V, cudaStream_t st, uint32_t s0, uint32_t n, uint32_t kbase, bool keep_items);   // process.cu

// steady state (rad_shoot): set-up -> queues -> fused resolve+ProcessHemicube.  The batch's slots are split over
// `lanes` concurrent streams (forked from and joined to the context's stream, inside the CUDA graph too): the lanes'
// kernels are bound by different things — set-up by latency, the queues by integer issue and REDs, ProcessHemicube by
// HBM — and overlap when they run side by side (+28 % shots/s measured at 4 lanes on the 16 k-patch scene).
// Within a lane the slots are rendered in groups only if its share of the work lists cannot hold the worst case (or a
// key-buffer cap is set, RAD_L2_GROUP_MB); a group's key buffers are then recycled by the next one under a new tag.
// ---- ring path (RadRing, raster_ring_kernel) -------------------------------------------------------------------------------
// Worth it when the walks dominate (a handful of pixels per patch or more; micro-triangle scenes keep the inline tier of the
// lane path, which needs every slot's key buffer at set-up time) and there are enough slots to pipeline.
static bool ring_eligible(const rad_ctx* c, uint32_t nslots) {
	const RadDev& D = c->d;
	return c->ring_mode && !c->tile_mode && c->world == 1 && nslots >= 4 && D.k >= 4 && D.P > 0 && D.RES / D.P >= 8u && (D.RES & 127u) == 0;
}
template <bool KEEP>
static cudaError_t launch_ring_kernel(rad_ctx* c, const RadDev& D, RadRing& R) {
	if (c->ring_ctas_per_sm == 0) {
		int a = 0, b = 0;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, raster_ring_kernel<false>, 128, 0);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, raster_ring_kernel<true>, 128, 0);
		c->ring_ctas_per_sm = a < b ? a : b;
		if (c->ring_ctas_per_sm < 2) c->ring_ctas_per_sm = 2;
		if (c->ring_ctas_per_sm > 8) c->ring_ctas_per_sm = 8;
	}
	int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->cfg.device);
	const uint32_t layers = (uint32_t)c->ring_ctas_per_sm;
	uint32_t pl = c->ring_proc_layers ? c->ring_proc_layers : (layers >= 6 ? 2u : 1u);      // process CTAs per SM
	if (pl >= layers) pl = layers - 1;
	R.walk_ctas = (uint32_t)nsm * (layers - pl);
	R.nw_walk = R.walk_ctas * 4u; R.nw_proc = (uint32_t)nsm * pl * 4u;
	R.keep_items = KEEP ? 1u : 0u;
	static const uint32_t dbg = [] { const char* e = getenv("RAD_RING_DEBUG"); return e ? (uint32_t)atoi(e) : 0u; }();   // measurement knob: 1 walk CTAs do not wait for the ring (wrong results), 16 role time stamps
	R.debug = dbg;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((uint32_t)nsm * layers); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;     // all CTAs resident at once, or the launch fails
	cfg.attrs = at; cfg.numAttrs = 1;
	void* args[2] = { (void*)&D, (void*)&R };
	return cudaLaunchKernelExC(&cfg, (const void*)raster_ring_kernel<KEEP>, args);
}
static void launch_ring(rad_ctx* c, bool keep_items, const std::function<void(int)>& mark) {
	const RadDev& D0 = c->d;
	const uint32_t nslots = D0.h1 - D0.h0;
	const uint64_t slot_bytes = (uint64_t)D0.RES * 8ull;
	// stage = about 16 MB of keys, ring = what stays L2-resident next to the records and F (<= ~80 MB)
	uint32_t sg = c->ring_sg ? c->ring_sg : (uint32_t)((16ull << 20) / slot_bytes);
	if (sg < 1) sg = 1;
	if (sg > 8) sg = 8;
	uint32_t rs = c->ring_rs ? c->ring_rs : (slot_bytes * sg * 4ull <= (80ull << 20) ? 4u : (slot_bytes * sg * 3ull <= (80ull << 20) ? 3u : 2u));
	if (rs < 2) rs = 2;
	while (sg > 1 && sg * rs > D0.k) sg--;
	while (rs > 2 && sg * rs > D0.k) rs--;
	for (uint32_t g0 = 0; g0 < nslots; g0 += RAD_RING_SLOTS) {
		const uint32_t n = nslots - g0 < RAD_RING_SLOTS ? nslots - g0 : RAD_RING_SLOTS;
		RadDev D = D0;
		D.h0 = D0.h0 + g0; D.h1 = D.h0 + n; D.kbase = D.h0;
		D.inline_area = 0u;                                    // every triangle is parked: no key buffer exists at set-up time
		RadRing R = {};
		R.nslots = n; R.sg = sg; R.rs = rs; R.nst = (n + sg - 1) / sg;
		const uint32_t rounds = (R.nst + rs - 1) / rs;         // one epoch tag per trip round the ring
		if (c->epoch < rounds) rad_launch_clear_keys(c);
		R.tag0 = c->epoch; c->epoch -= rounds; c->d.tag = R.tag0 - (rounds - 1);
		R.cap_pairs = D.pairs_cap / n; R.cap_sm = D.q_sm_cap / n; R.cap_tri = D.q_tri_cap / n; R.cap_ent = D.q_ent_cap / n;
		cudaMemsetAsync(D.rc, 0, sizeof(RadRingCtl), c->stream);
		cudaMemsetAsync(&D.rc->t_start, 0xFF, 8, c->stream);
		raster_cull_kernel<true><<<dim3((D.P + 255) / 256, 1, n), 256, 0, c->stream>>>(D, R);
		launch_setup_kernel<true>(D, R, n, c->stream);
		c->launches += 2;
		if (mark) mark(1);
		const cudaError_t e = keep_items ? launch_ring_kernel<true>(c, D, R) : launch_ring_kernel<false>(c, D, R);
		if (e != cudaSuccess) { c->err = std::string("raster_ring_kernel launch: ") + cudaGetErrorString(e); c->ring_failed = true; }
		c->launches++;
		if (mark) mark(2);
		if (mark && getenv("RAD_RING_DEBUG") && (atoi(getenv("RAD_RING_DEBUG")) & 16)) {
			RadRingCtl* h = new RadRingCtl;
			cudaMemcpyAsync(h, D.rc, sizeof(RadRingCtl), cudaMemcpyDeviceToHost, c->stream); cudaStreamSynchronize(c->stream);
			uint32_t nrec = 0, nent = 0; for (uint32_t i = 0; i < n; i++) { nrec += h->slot[i].q_small; nent += h->slot[i].q_entries; }
			fprintf(stderr, "ring: walk %.1f us (first CTA start -> last walk CTA end), process end %.1f us, mean wait per CTA: walk %.1f us, process %.1f us; %u records, %u chunk entries, sg %u rs %u nst %u, %u walk CTAs\n",
			        (h->t_walk_end - h->t_start) * 1e-3, (h->t_proc_end - h->t_start) * 1e-3, h->t_walk_wait * 1e-3 / R.walk_ctas, h->t_proc_wait * 1e-3 / ((R.nw_proc / 4) ? (R.nw_proc / 4) : 1), nrec, nent, R.sg, R.rs, R.nst, R.walk_ctas);
			delete h;
		}
	}
	c->keys_dirty = false;
}

void rad_launch_raster_process_marked(rad_ctx* c, bool keep_items, const std::function<void(int)>& mark) {
	const uint32_t nslots = c->d.h1 - c->d.h0;
	if (nslots == 0) return;
	if (ring_eligible(c, nslots)) { launch_ring(c, keep_items, mark); return; }
	uint32_t L = mark ? 1u : c->lanes;        // the per-stage profile wants the stages back to back
	if (L > nslots) L = nslots;
	if (L < 1) L = 1;
	// group size and tags (every lane runs the same number of groups; group j of every lane shares tag j)
	uint32_t lane_slots = (nslots + L - 1) / L, g;
	{
		RadDev V = lane_view(c, 0, L); V.h1 = V.h0 + lane_slots;
		g = queue_group(V);
		const uint64_t cap = ((uint64_t)c->l2_group_mb << 20) / ((uint64_t)V.RES * 8ull);
		if (cap >= 1 && cap < g) g = (uint32_t)cap;
		if (g < (lane_slots + 253) / 254) g = (lane_slots + 253) / 254;   // at most 254 groups: one epoch tag each
	}
	uint32_t tags[RAD_MAX_HEMICUBES];
	const uint32_t ngroups = (lane_slots + g - 1) / g;
	// the groups of a batch recycle the same key buffers, so their tags must be strictly decreasing: if fewer tags are left
	// than the batch has groups, start over from cleared keys (a wrap in the middle would hand a later group a LARGER tag
	// and its keys would lose against the stale ones of an earlier group)
	if (c->epoch < ngroups) rad_launch_clear_keys(c);
	for (uint32_t j = 0; j < ngroups; j++) tags[j] = rad_next_tag(c);
	if (L > 1) cudaEventRecord(c->ev_fork, c->stream);
	for (uint32_t lane = 0; lane < L; lane++) {
		const uint32_t ls0 = (uint32_t)((uint64_t)nslots * lane / L), ls1 = (uint32_t)((uint64_t)nslots * (lane + 1) / L);
		cudaStream_t st = L > 1 ? c->lane_stream[lane] : c->stream;
		if (L > 1) cudaStreamWaitEvent(st, c->ev_fork, 0);
		RadDev V = lane_view(c, lane, L);
		RadTiles TV = c->tl;                                 // tile mode: the lane's share of the bins
		if (c->tile_mode) {
			TV.cnt += (size_t)ls0 * TV.T * 2u; TV.base += (size_t)ls0 * TV.T * 2u + lane;
			TV.refs_cap = c->tl.refs_cap / L; TV.refs += (size_t)lane * TV.refs_cap;
		}
		uint32_t j = 0;
		for (uint32_t s0 = ls0; s0 < ls1; s0 += g, j++) {
			const uint32_t n = ls1 - s0 < g ? ls1 - s0 : g;
			const uint32_t kbase = c->d.h0 + s0 - ls0;       // key buffer = lane's first buffer (ls0) + position in the group
			V.tag = tags[j]; c->d.tag = tags[j];
			if (c->tile_mode) {
				launch_setup(c, V, st, s0, n, kbase, true); if (mark) mark(1);
				rad_launch_tiles_view(c, V, TV, st, s0, n, keep_items, mark);
				continue;
			}
			launch_setup(c, V, st, s0, n, kbase); if (mark) mark(1);
			launch_chunks(c, V, st, kbase); if (mark) mark(2);
			rad_launch_process_view(c, V, st, s0, n, kbase, keep_items); if (mark) mark(4);
		}
		if (L > 1) { cudaEventRecord(c->ev_lane[lane], st); if (!c->defer_join) cudaStreamWaitEvent(c->stream, c->ev_lane[lane], 0); }
	}
	c->last_lanes = L;
}
void rad_join_lanes(rad_ctx* c) {
	if (c->last_lanes > 1) for (uint32_t lane = 0; lane < c->last_lanes; lane++) cudaStreamWaitEvent(c->stream, c->ev_lane[lane], 0);
}
void rad_launch_raster_process(rad_ctx* c, bool keep_items) { rad_launch_raster_process_marked(c, keep_items, nullptr); }

void rad_launch_resolve(rad_ctx* c, bool) {
	const RadDev& D = c->d;
	const uint32_t nslots = D.h1 - D.h0;
	if (nslots == 0) return;
	uint32_t bx = (D.RES + 255) / 256;
	if (bx > 148 * 8) bx = 148 * 8;
	resolve_kernel<<<dim3(bx, nslots), 256, 0, c->stream>>>(D);
	c->launches++;
}

void rad_launch_clear_keys(rad_ctx* c) {
	const RadDev& D = c->d;
	clear_keys_kernel<<<148 * 8, 256, 0, c->stream>>>(D.keys, (size_t)c->key_bufs * D.RES);   // every buffer: the epoch is shared by all of them
	c->launches++;
	c->keys_dirty = false;
	c->epoch = 254;
}

// Epoch tag of the next render into a key buffer: strictly decreasing, so that every stale key (larger tag) loses
// against any new one under atomicMin and reads back as "empty"; a real clear happens only when the tags run out.
uint32_t rad_next_tag(rad_ctx* c) {
	if (c->epoch == 0) rad_launch_clear_keys(c);
	return c->epoch--;
}

void rad_launch_read_depth(rad_ctx* c, uint32_t hi, uint32_t* d_out) {
	const RadDev& D = c->d;
	read_depth_kernel<<<(D.RES + 255) / 256, 256, 0, c->stream>>>(D.keys + (size_t)hi * D.RES, d_out, D.RES, D.tag);
	c->launches++;
}
