// CPU check of the lane-level walks of the tile-binned rasteriser (radiosity_b200/csrc/tile_walk.cuh): the header is
// compiled for the host and every walk is run lane by lane, tile by tile, against a brute-force statement of the raster
// rules (exact int64 edge functions, top-left rule, the depth formula of raster.cu / oracle.cpp).  Keys must be equal
// bit for bit.  Build: g++ -O1 -ffp-contract=off -std=c++17 tile_walk_check.cpp   (test infrastructure only)
#include "../../radiosity_b200/csrc/tile_walk.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 20); }
static int rndi(int lo, int hi) { return lo + (int)(rnd() % (uint32_t)(hi - lo + 1)); }
static float rndf() { return (float)(rnd() & 0xFFFFFF) / 16777216.0f; }

struct Atlas {
	int W, H; std::vector<unsigned long long> k;
	Atlas(int w, int h) : W(w), H(h), k((size_t)w * h, ~0ull) {}
	void put(int x, int y, unsigned long long key) { unsigned long long& d = k[(size_t)y * W + x]; if (key < d) d = key; }
};

struct Vtx { int X, Y; float Z; };

// brute force: one triangle, scissor = whole atlas
static bool brute_tri(Atlas& A, Vtx a, Vtx b, Vtx c, uint32_t id1, int* bbox /* px0,py0,px1,py1 */, float* inv_out) {
	const long long area2 = tw::edge_fn(a.X, a.Y, b.X, b.Y, c.X, c.Y);
	if (area2 <= 0) return false;
	const int minx = std::min(a.X, std::min(b.X, c.X)), maxx = std::max(a.X, std::max(b.X, c.X));
	const int miny = std::min(a.Y, std::min(b.Y, c.Y)), maxy = std::max(a.Y, std::max(b.Y, c.Y));
	const int px0 = std::max((minx - 128 + 255) >> 8, 0), px1 = std::min((maxx - 128) >> 8, A.W - 1);
	const int py0 = std::max((miny - 128 + 255) >> 8, 0), py1 = std::min((maxy - 128) >> 8, A.H - 1);
	if (px0 > px1 || py0 > py1) return false;
	const float inv = 1.0f / (float)area2;
	const float z0 = a.Z, dz1 = b.Z - a.Z, dz2 = c.Z - a.Z;
	const int b0 = tw::edge_bias(b.X, b.Y, c.X, c.Y), b1 = tw::edge_bias(c.X, c.Y, a.X, a.Y), b2 = tw::edge_bias(a.X, a.Y, b.X, b.Y);
	for (int py = py0; py <= py1; py++)
		for (int px = px0; px <= px1; px++) {
			const int cx = px * 256 + 128, cy = py * 256 + 128;
			const long long e0 = tw::edge_fn(b.X, b.Y, c.X, c.Y, cx, cy) + b0, e1 = tw::edge_fn(c.X, c.Y, a.X, a.Y, cx, cy) + b1, e2 = tw::edge_fn(a.X, a.Y, b.X, b.Y, cx, cy) + b2;
			if ((e0 | e1 | e2) < 0) continue;
			unsigned long long key;
			if (tw::frag_key(z0, (float)(e1 - b1) * inv, dz1, (float)(e2 - b2) * inv, dz2, id1, key)) A.put(px, py, key);
		}
	bbox[0] = px0; bbox[1] = py0; bbox[2] = px1; bbox[3] = py1; *inv_out = inv;
	return true;
}

struct TileEmit {
	Atlas* A; int tx0, ty0, tw_, th_; long* bad;
	void operator()(int off, unsigned long long key) {
		const int row = off / RAD_TILE_W, col = off % RAD_TILE_W;
		if (off < 0 || off >= RAD_TILE_PIX || col >= tw_ || row >= th_) { (*bad)++; return; }
		A->put(tx0 + col, ty0 + row, key);
	}
};

static int radius_about(int cx, int cy, int X, int Y) { return std::max(std::abs(X - cx), std::abs(Y - cy)); }

int main(int argc, char** argv) {
	const int cases = argc > 1 ? atoi(argv[1]) : 20000;
	const int W = 160, H = 80;      // 3 x 3 tiles, the last column 32 wide, the last row 16 high
	long bad = 0, mism = 0, quads = 0, lones = 0, bigs = 0, wide = 0, multi = 0;
	for (int it = 0; it < cases; it++) {
		Atlas ref(W, H), got(W, H);
		const uint32_t id1 = 1 + (rnd() % 1000000u);
		const int kind = it % 3;
		if (kind < 2) {
			// a small quad (kind 0) or a lone small triangle (kind 1) somewhere in / around the atlas
			const int sz = rndi(1, kind == 0 ? 22 : 30) * 256;
			const int ox = rndi(-20 * 256, (W + 4) * 256), oy = rndi(-20 * 256, (H + 4) * 256);
			Vtx v[4];
			// counter-clockwise around a centre, jittered
			v[0] = { ox + rndi(0, sz / 3), oy + rndi(0, sz / 3), rndf() };
			v[1] = { ox + sz - rndi(0, sz / 3), oy + rndi(0, sz / 3), rndf() };
			v[2] = { ox + sz - rndi(0, sz / 3), oy + sz - rndi(0, sz / 3), rndf() };
			v[3] = { ox + rndi(0, sz / 3), oy + sz - rndi(0, sz / 3), rndf() };
			if (it % 7 == 0) v[2].Z = 1.5f;                   // some fragments beyond the far plane (fail LESS)
			if (it % 2) for (int j = 0; j < 4; j++) { v[j].X &= ~127; v[j].Y &= ~127; }   // vertices on pixel centres / pixel edges: the tie rules decide
			if (kind == 1) v[3] = v[0];
			int ba[4], bb[4]; float invA = 0, invB = 0;
			const bool okA = brute_tri(ref, v[0], v[1], v[2], id1, ba, &invA);
			const bool okB = kind == 0 && brute_tri(ref, v[0], v[2], v[3], id1, bb, &invB);
			if (!okA || (kind == 0 && !okB)) continue;        // the set-up kernel parks such quads as lone triangles
			int px0 = ba[0], py0 = ba[1], px1 = ba[2], py1 = ba[3];
			if (kind == 0) { px0 = std::min(px0, bb[0]); py0 = std::min(py0, bb[1]); px1 = std::max(px1, bb[2]); py1 = std::max(py1, bb[3]); }
			const int bw = px1 - px0 + 1, bh = py1 - py0 + 1;
			if (bw * bh > 512 || bw > 255 || bh > 255) continue;
			const int cx = px0 * 256 + 128, cy = py0 * 256 + 128;
			int r = 0; for (int j = 0; j < 4; j++) r = std::max(r, radius_about(cx, cy, v[j].X, v[j].Y));
			if (!((long long)r * (long long)(r + (std::max(bw, bh) + 8) * 256) < (1ll << 29))) continue;
			const tw::RecWords rec = tw::make_record(v[0].X, v[0].Y, v[1].X, v[1].Y, v[2].X, v[2].Y, v[3].X, v[3].Y, v[0].Z, v[1].Z, v[2].Z, v[3].Z,
			                                         invA, kind == 0 ? invB : 0.0f, id1, 0, px0, py0, bw, bh);
			int t0x, t0y, t1x, t1y; tw::tile_range(px0, py0, px1, py1, t0x, t0y, t1x, t1y);
			if (t1x > t0x || t1y > t0y) multi++;
			for (int ty = t0y; ty <= t1y; ty++)
				for (int tx = t0x; tx <= t1x; tx++) {
					const int tx0 = tx * RAD_TILE_W, ty0 = ty * RAD_TILE_H;
					TileEmit em{ &got, tx0, ty0, std::min(RAD_TILE_W, W - tx0), std::min(RAD_TILE_H, H - ty0), &bad };
					int msteps = 0;
					for (int l8 = 0; l8 < 8; l8++) {
						tw::QuadWalk q; q.init(rec, tx0, ty0, em.tw_, em.th_, l8);
						msteps = std::max(msteps, q.steps());
						const int n = q.steps() + 3;              // a quarter warp keeps stepping while its neighbours are busy
						for (int s = 0; s < n; s++) q.step(em);
					}
					if (msteps == 0) bad++;                       // the bins only hold tiles the bbox overlaps
				}
			(kind == 0 ? quads : lones)++;
		} else {
			// a large triangle; every third one with far-away vertices (int64 walk)
			const bool far_ = (it % 9) == 2;
			const int span = far_ ? 60000 * 256 : 200 * 256;
			Vtx v[3];
			for (int j = 0; j < 3; j++) v[j] = { rndi(-span, span + W * 256), rndi(-span, span + H * 256), rndf() };
			if (it % 2) for (int j = 0; j < 3; j++) { v[j].X &= ~127; v[j].Y &= ~127; }
			if (tw::edge_fn(v[0].X, v[0].Y, v[1].X, v[1].Y, v[2].X, v[2].Y) < 0) std::swap(v[1], v[2]);
			int bx[4]; float inv = 0;
			if (!brute_tri(ref, v[0], v[1], v[2], id1, bx, &inv)) continue;
			tw::BigTri t{ v[0].X, v[0].Y, v[1].X, v[1].Y, v[2].X, v[2].Y, v[0].Z, v[1].Z - v[0].Z, v[2].Z - v[0].Z, inv, id1, bx[0], bx[1], bx[2], bx[3] };
			int t0x, t0y, t1x, t1y; tw::tile_range(bx[0], bx[1], bx[2], bx[3], t0x, t0y, t1x, t1y);
			for (int ty = t0y; ty <= t1y; ty++)
				for (int tx = t0x; tx <= t1x; tx++) {
					const int tx0 = tx * RAD_TILE_W, ty0 = ty * RAD_TILE_H;
					TileEmit em{ &got, tx0, ty0, std::min(RAD_TILE_W, W - tx0), std::min(RAD_TILE_H, H - ty0), &bad };
					tw::BigWalk w; w.init(t, tx0, ty0, em.tw_, em.th_);
					if (!w.narrow) wide++;
					if (w.rejects()) continue;                    // the bin kernel drops this (triangle, tile) pair
					for (int warp = 0; warp < 4; warp++)
						for (int s = warp; s < w.nsteps; s += 4)
							for (int lane = 0; lane < 32; lane++) w.step(t, s, lane, em);
				}
			bigs++;
		}
		for (size_t i = 0; i < ref.k.size(); i++) if (ref.k[i] != got.k[i]) mism++;
	}
	printf("{\"quads\": %ld, \"lone_triangles\": %ld, \"big_triangles\": %ld, \"int64_walks\": %ld, \"multi_tile_records\": %ld, \"bad_offsets\": %ld, \"mismatched_pixels\": %ld}\n",
	       quads, lones, bigs, wide, multi, bad, mism);
	return (bad || mism || quads < cases / 10 || lones < cases / 10 || bigs < cases / 10) ? 1 : 0;
}
