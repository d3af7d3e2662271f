// Lane-level coverage walks of the tile-binned rasteriser (raster_tiles.cu), written as host/device code so that the
// arithmetic can be checked on the CPU against a brute-force statement of the raster rules (tests/cpu/tile_walk_check.cpp)
// before it ever runs on a GPU.  Nothing here touches memory: a walk calls emit(offset inside the tile, key).
//
// Same rules and the same integers / float operations as raster.cu (and oracle/oracle.cpp): pixel centres, exact edge
// functions with the top-left bias, depth = (z0 + l1*dz1) + l2*dz2 from the unbiased edge values, clamped, RNE to 24 bit,
// LESS against 1.0.  Build with --fmad=false (nvcc) / -ffp-contract=off (g++).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#ifdef __CUDACC__
#define RAD_HD __host__ __device__ __forceinline__
#else
#define RAD_HD inline
#endif

#ifndef RAD_TILE_W
#define RAD_TILE_W 64             // a tile is RAD_TILE_W x RAD_TILE_H atlas pixels, its keys live in shared memory
#endif
#ifndef RAD_TILE_H
#define RAD_TILE_H 32
#endif
#define RAD_TILE_PIX (RAD_TILE_W * RAD_TILE_H)

namespace tw {

RAD_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
	return __uint_as_float(u);
#else
	float f; memcpy(&f, &u, 4); return f;
#endif
}
RAD_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
	return __float_as_uint(f);
#else
	uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RAD_HD uint32_t f2u_rn(float f) {          // float -> uint32, round to nearest even (f >= 0)
#ifdef __CUDA_ARCH__
	return __float2uint_rn(f);
#else
	return (uint32_t)nearbyintf(f);        // default rounding mode = RNE
#endif
}
RAD_HD int imin(int a, int b) { return a < b ? a : b; }
RAD_HD int imax(int a, int b) { return a > b ? a : b; }
RAD_HD int iabs(int a) { return a < 0 ? -a : a; }

// key of a covered fragment inside a tile: depth24 << 32 | id+1 (no epoch tag: a tile starts from cleared keys);
// returns false when the fragment fails GL_LESS against the cleared depth 1.0
RAD_HD bool frag_key(float z0, float l1, float dz1, float l2, float dz2, uint32_t id1, unsigned long long& key) {
	float z = (z0 + l1 * dz1) + l2 * dz2;
	z = fminf(fmaxf(z, 0.0f), 1.0f);
	const uint32_t dq = f2u_rn(z * 16777215.0f);
	key = ((unsigned long long)dq << 32) | id1;
	return dq < 0xFFFFFFu;
}

// the 56 bytes of a small-quad record (RadSmallQuad, rad_internal.cuh) that a walk reads, as four loaded words
struct RecWords { uint32_t a[4], b[4], c[4], d[2]; };

RAD_HD uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
// host-side twin of make_small (raster.cu), used by the CPU check to build records
RAD_HD RecWords make_record(int X0, int Y0, int X1, int Y1, int X2, int Y2, int X3, int Y3, float Z0, float Z1, float Z2, float Z3,
                            float invA, float invB, uint32_t id1, uint32_t slot, int px0, int py0, int bw, int bh) {
	const int cx = px0 * 256 + 128, cy = py0 * 256 + 128;
	RecWords r;
	r.a[0] = pack16(X0 - cx, Y0 - cy); r.a[1] = pack16(X1 - cx, Y1 - cy); r.a[2] = pack16(X2 - cx, Y2 - cy); r.a[3] = pack16(X3 - cx, Y3 - cy);
	r.b[0] = f2u(Z0); r.b[1] = f2u(Z1); r.b[2] = f2u(Z2); r.b[3] = f2u(Z3);
	r.c[0] = f2u(invA); r.c[1] = f2u(invB); r.c[2] = id1; r.c[3] = pack16((int)slot, (32768 + bw - 1) / bw);
	r.d[0] = pack16(px0, py0); r.d[1] = (uint32_t)bw | ((uint32_t)bh << 8);
	return r;
}

// biased edge function of (a -> b) at the origin (vertices origin-relative) and its per-pixel steps
RAD_HD void edge_origin(int ax, int ay, int bx, int by, int& e, int& sx, int& sy, int& bias) {
	const int dx = bx - ax, dy = by - ay;
	bias = (dy < 0 || (dy == 0 && dx < 0)) ? 0 : -1;
	e = dx * (-ay) - dy * (-ax) + bias;
	sx = -dy * 256; sy = dx * 256;
}

// Quarter-warp walk of one small-quad record restricted to one tile.  Lane l8 (0..7) visits the positions l8, l8 + 8, ...
// of the linear run of wi * ceil(hi / 2) two-row columns of (record bbox) n (tile); every position tests two pixels.
struct QuadWalk {
	int npx, w8, hh, q8, r8, x, y, idx;
	int a0, a1, a2, b0, b1;                               // biased edge values at the walk origin
	int a0x, a1x, a2x, b0x, b1x, a0y, a1y, a2y, b0y, b1y; // per-pixel steps
	int bA1, bA2, bB1, bB2;
	float Z0, dA1, dA2, dB2, invA, invB;
	uint32_t id1;
	int org;                                              // offset of the walk origin inside the tile

	RAD_HD int steps() const { return (npx + 7) >> 3; }
	RAD_HD void none() { npx = 0; w8 = 1; hh = 0; q8 = 0; r8 = 0; x = 0; y = 0; idx = 0; org = 0; }   // no record: every step is a no-op

	// (tx0, ty0) = atlas pixel of the tile's corner, (tw_, th_) = its size clipped to the atlas
	RAD_HD void init(const RecWords& r, int tx0, int ty0, int tw_, int th_, int l8) {
		const int x0 = (int)(r.a[0] << 16) >> 16, y0 = (int)r.a[0] >> 16, x1 = (int)(r.a[1] << 16) >> 16, y1 = (int)r.a[1] >> 16;
		const int x2 = (int)(r.a[2] << 16) >> 16, y2 = (int)r.a[2] >> 16, x3 = (int)(r.a[3] << 16) >> 16, y3 = (int)r.a[3] >> 16;
		int bA0, bB0;
		edge_origin(x1, y1, x2, y2, a0, a0x, a0y, bA0);      // A: edges (1->2), (2->0), (0->1)
		edge_origin(x2, y2, x0, y0, a1, a1x, a1y, bA1);
		edge_origin(x0, y0, x1, y1, a2, a2x, a2y, bA2);
		edge_origin(x2, y2, x3, y3, b0, b0x, b0y, bB0);      // B: edges (2->3), (3->0), (0->2)
		edge_origin(x3, y3, x0, y0, b1, b1x, b1y, bB1);
		{ const int dx = x2 - x0, dy = y2 - y0; bB2 = (dy < 0 || (dy == 0 && dx < 0)) ? 0 : -1; }
		(void)bA0; (void)bB0;
		Z0 = u2f(r.b[0]);
		dA1 = u2f(r.b[1]) - Z0; dA2 = u2f(r.b[2]) - Z0; dB2 = u2f(r.b[3]) - Z0;
		invA = u2f(r.c[0]); invB = u2f(r.c[1]);
		id1 = r.c[2];
		uint32_t rcpw = r.c[3] >> 16;
		const int px0 = (int)(r.d[0] & 0xFFFFu), py0 = (int)(r.d[0] >> 16);
		const int bw = (int)(r.d[1] & 0xFFu), bh = (int)((r.d[1] >> 8) & 0xFFu);
		// (record bbox) n (tile)
		const int ix0 = imax(px0, tx0), iy0 = imax(py0, ty0);
		const int wi = imin(px0 + bw, tx0 + tw_) - ix0, hi = imin(py0 + bh, ty0 + th_) - iy0;
		idx = l8;
		if (wi <= 0 || hi <= 0) { npx = 0; w8 = 1; hh = 0; q8 = 0; r8 = 0; x = 0; y = 0; org = 0; return; }
		const int ox = ix0 - px0, oy = iy0 - py0;          // move the origin of the edge functions to the intersection's corner
		a0 += ox * a0x + oy * a0y; a1 += ox * a1x + oy * a1y; a2 += ox * a2x + oy * a2y;
		b0 += ox * b0x + oy * b0y; b1 += ox * b1x + oy * b1y;
		if (wi != bw) rcpw = (uint32_t)((32768 + wi - 1) / wi);
		w8 = wi; hh = hi;
		npx = wi * ((hi + 1) >> 1);
		q8 = (int)((8u * rcpw) >> 15); r8 = 8 - q8 * wi;                  // 8 / wi, 8 % wi
		y = (int)(((uint32_t)l8 * rcpw) >> 15); x = l8 - y * wi;          // this lane's first position (y counts row pairs)
		org = (iy0 - ty0) * RAD_TILE_W + (ix0 - tx0);
	}

	template <class Emit>
	RAD_HD void pixel(int e0, int e1, int e2, int f0, int f1, int off, Emit& emit) const {
		const int e1u = e1 - bA1;
		const int f2 = bB2 - e1u;                            // B's edge (0->2) is minus A's unbiased edge (2->0)
		const bool inA = (e0 | e1 | e2) >= 0, inB = (f0 | f1 | f2) >= 0;
		if (inA || inB) {
			const float inv = inA ? invA : invB;
			const float l1 = (float)(inA ? e1u : f1 - bB1) * inv, l2 = (float)(inA ? e2 - bA2 : -e1u) * inv;
			unsigned long long key;
			if (frag_key(Z0, l1, inA ? dA1 : dA2, l2, inA ? dA2 : dB2, id1, key)) emit(off, key);
		}
	}

	// one step: this lane's position idx (if any), then on to idx + 8
	template <class Emit>
	RAD_HD void step(Emit& emit) {
		if (idx < npx) {
			const int e0 = a0 + x * a0x + 2 * y * a0y, e1 = a1 + x * a1x + 2 * y * a1y, e2 = a2 + x * a2x + 2 * y * a2y;
			const int f0 = b0 + x * b0x + 2 * y * b0y, f1 = b1 + x * b1x + 2 * y * b1y;
			const int off = org + 2 * y * RAD_TILE_W + x;
			pixel(e0, e1, e2, f0, f1, off, emit);
			if (2 * y + 1 < hh) pixel(e0 + a0y, e1 + a1y, e2 + a2y, f0 + b0y, f1 + b1y, off + RAD_TILE_W, emit);
		}
		idx += 8; x += r8; y += q8;
		if (x >= w8) { x -= w8; y++; }
	}
};

// ---- large triangles (RadBigTri): the CTA's warps share the 8x4-pixel steps of (triangle bbox) n (tile) ----------------
struct BigTri { int X0, Y0, X1, Y1, X2, Y2; float z0, dz1, dz2, inv_area; uint32_t id1; int px0, py0, px1, py1; };

RAD_HD long long edge_fn(int ax, int ay, int bx, int by, int cx, int cy) {
	return (long long)(bx - ax) * (long long)(cy - ay) - (long long)(by - ay) * (long long)(cx - ax);
}
RAD_HD int edge_bias(int ax, int ay, int bx, int by) {
	const int dx = bx - ax, dy = by - ay;
	return (dy < 0 || (dy == 0 && dx < 0)) ? 0 : -1;
}

struct BigWalk {
	long long e0, e1, e2;            // biased edge values at the centre of pixel (ix0, iy0)
	long long sx0, sx1, sx2, sy0, sy1, sy2;
	int b1, b2;
	int ix0, iy0, wi, hi, nsx, nsteps, org;
	bool narrow;                     // every edge value inside the intersection fits int32 (see fits32 in raster.cu)

	RAD_HD void init(const BigTri& t, int tx0, int ty0, int tw_, int th_) {
		ix0 = imax(t.px0, tx0); iy0 = imax(t.py0, ty0);
		wi = imin(t.px1 + 1, tx0 + tw_) - ix0; hi = imin(t.py1 + 1, ty0 + th_) - iy0;
		if (wi <= 0 || hi <= 0) { nsteps = 0; nsx = 1; org = 0; narrow = false; return; }
		nsx = (wi + 7) >> 3; nsteps = nsx * ((hi + 3) >> 2);
		org = (iy0 - ty0) * RAD_TILE_W + (ix0 - tx0);
		const int cx = ix0 * 256 + 128, cy = iy0 * 256 + 128;
		const int b0 = edge_bias(t.X1, t.Y1, t.X2, t.Y2);
		b1 = edge_bias(t.X2, t.Y2, t.X0, t.Y0); b2 = edge_bias(t.X0, t.Y0, t.X1, t.Y1);
		e0 = edge_fn(t.X1, t.Y1, t.X2, t.Y2, cx, cy) + b0;
		e1 = edge_fn(t.X2, t.Y2, t.X0, t.Y0, cx, cy) + b1;
		e2 = edge_fn(t.X0, t.Y0, t.X1, t.Y1, cx, cy) + b2;
		sx0 = -(long long)(t.Y2 - t.Y1) * 256; sy0 = (long long)(t.X2 - t.X1) * 256;
		sx1 = -(long long)(t.Y0 - t.Y2) * 256; sy1 = (long long)(t.X0 - t.X2) * 256;
		sx2 = -(long long)(t.Y1 - t.Y0) * 256; sy2 = (long long)(t.X1 - t.X0) * 256;
		const int r = imax(imax(imax(iabs(t.X0 - cx), iabs(t.Y0 - cy)), imax(iabs(t.X1 - cx), iabs(t.Y1 - cy))), imax(iabs(t.X2 - cx), iabs(t.Y2 - cy)));
		narrow = (long long)r * (long long)(r + (RAD_TILE_W + 8) * 256) < (1ll << 29);
	}

	// true when no pixel of the intersection can be covered: one edge has all four corner pixel centres outside
	RAD_HD bool rejects() const {
		if (nsteps == 0) return true;
		const long long dx = wi - 1, dy = hi - 1;
		return (e0 < 0 && e0 + dx * sx0 < 0 && e0 + dy * sy0 < 0 && e0 + dx * sx0 + dy * sy0 < 0) ||
		       (e1 < 0 && e1 + dx * sx1 < 0 && e1 + dy * sy1 < 0 && e1 + dx * sx1 + dy * sy1 < 0) ||
		       (e2 < 0 && e2 + dx * sx2 < 0 && e2 + dy * sy2 < 0 && e2 + dx * sx2 + dy * sy2 < 0);
	}

	// step s of the walk (8 x 4 pixels), lane 0..31
	template <class Emit>
	RAD_HD void step(const BigTri& t, int s, int lane, Emit& emit) const {
		const int cy = s / nsx, cx = s - cy * nsx;
		const int dx = 8 * cx + (lane & 7), dy = 4 * cy + (lane >> 3);
		if (dx >= wi || dy >= hi) return;
		float l1, l2;
		if (narrow) {
			const int f0 = (int)e0 + dx * (int)sx0 + dy * (int)sy0, f1 = (int)e1 + dx * (int)sx1 + dy * (int)sy1, f2 = (int)e2 + dx * (int)sx2 + dy * (int)sy2;
			if ((f0 | f1 | f2) < 0) return;
			l1 = (float)(f1 - b1) * t.inv_area; l2 = (float)(f2 - b2) * t.inv_area;
		} else {
			const long long f0 = e0 + dx * sx0 + dy * sy0, f1 = e1 + dx * sx1 + dy * sy1, f2 = e2 + dx * sx2 + dy * sy2;
			if ((f0 | f1 | f2) < 0) return;
			l1 = (float)(f1 - b1) * t.inv_area; l2 = (float)(f2 - b2) * t.inv_area;
		}
		unsigned long long key;
		if (frag_key(t.z0, l1, t.dz1, l2, t.dz2, t.id1, key)) emit(org + dy * RAD_TILE_W + dx, key);
	}
};

// tiles overlapped by an inclusive pixel bbox
RAD_HD void tile_range(int px0, int py0, int px1, int py1, int& tx0, int& ty0, int& tx1, int& ty1) {
	tx0 = px0 / RAD_TILE_W; tx1 = px1 / RAD_TILE_W; ty0 = py0 / RAD_TILE_H; ty1 = py1 / RAD_TILE_H;
}

} // namespace tw
