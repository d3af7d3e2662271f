// ============================================================================================
// TEST INFRASTRUCTURE — CPU ORACLE for the radiosity shooting loop of david-sabata/Radiosity.
//
// A plain, scalar C++ restatement of the reference's algorithm for the hot path named in
// BASELINE.json (shooter selection -> five-face hemicube item-buffer render -> ProcessHemicube
// delta-form-factor scatter-add -> radiosity/unshot update).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library; the product
// (radiosity_b200/) never links, loads or calls it.
//
// PINNING.  The reference ships no tests or golden vectors.  The host-side parts
// restated here (scene, subdivision, selection, camera/MVP, form-factor table, colour codec) are
// pinned bit-for-bit against the reference's own code compiled into oracle/_ref/libref_host.so
// (tests/test_golden_reference.py) and against tests/golden/ fixtures generated from it.  The
// ProcessHemicube kernel restated here is pinned on the reference's own kernel TEXT, compiled from
// the reference header and run on the CPU (oracle/ref_kernel.cpp): identical record streams.  The
// one piece whose arithmetic lives in a GPU driver that is not in /root/reference — the OpenGL
// rasteriser — is restated from the reference's call sites plus the OpenGL 3.3 rasterisation
// rules; for THAT parity is unpinned by any reference-run output (no GL stack exists in this
// image).  It is anchored on invariants (closed box: sum F == sum dFF, F[self] == 0, every decoded
// id < P, codec round trip) and pinned against an independent restatement of the same rules in
// plain Python (tests/test_oracle_raster_rules.py).
//
// All float arithmetic is IEEE single, evaluated in the order written, never contracted
// (compile with -ffp-contract=off); the CUDA path is written to the same operation order so that
// item buffers can be compared bit for bit.
//
// Citations are file:line under /root/reference/source/.
// ============================================================================================
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <list>
#include <string>
#include <fstream>
#include <sstream>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------------------------------
// Vector3f subset (Vector.h:334-549).  NOTE the reference's cross product is reversed:
// a.v_Cross(b) returns b x a (Vector.h:534-537).
// ------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
static inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
static inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline V3 mulf(V3 a, float t) { return v3(a.x * t, a.y * t, a.z * t); }
static inline V3 mulv(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 divf(V3 a, float t) { t = 1 / t; return v3(a.x * t, a.y * t, a.z * t); } // Vector.h:449-453
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }           // Vector.h:356-359
static inline float len(V3 a) { return (float)std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); } // Vector.h:351-354 (float overload)
static inline V3 refcross(V3 a, V3 b) {           // a.v_Cross(b)  (= b x a), Vector.h:534-537
	return v3(b.y * a.z - b.z * a.y, b.z * a.x - b.x * a.z, b.x * a.y - b.y * a.x);
}
static inline V3 normalized(V3 a) {               // Vector.h:390-400
	float t = len(a);
	if (t != 0) { t = 1 / t; a.x *= t; a.y *= t; a.z *= t; }
	return a;
}

// ------------------------------------------------------------------------------------------
// Matrix4f subset: column-major f[col][row] (Vector.h:890-892)
// ------------------------------------------------------------------------------------------
struct M4 { float f[4][4]; };

static void mat_product(M4& out, const M4& a, const M4& b) {   // Vector.cpp:445-458
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j)
			out.f[i][j] = a.f[0][j] * b.f[i][0] + a.f[1][j] * b.f[i][1] + a.f[2][j] * b.f[i][2] + a.f[3][j] * b.f[i][3];
}

static void mat_frustum(M4& m, float l, float r, float b, float t, float n, float f) { // Transform.cpp:26-46
	m.f[0][0] = 2 * n / (r - l); m.f[1][0] = 0; m.f[2][0] = (r + l) / (r - l); m.f[3][0] = 0;
	m.f[0][1] = 0; m.f[1][1] = 2 * n / (t - b); m.f[2][1] = (t + b) / (t - b); m.f[3][1] = 0;
	m.f[0][2] = 0; m.f[1][2] = 0; m.f[2][2] = -(f + n) / (f - n); m.f[3][2] = -2 * f * n / (f - n);
	m.f[0][3] = 0; m.f[1][3] = 0; m.f[2][3] = -1; m.f[3][3] = 0;
}

static const float kPi = 3.1415926535897932384626433832795028841971691075f;   // Vector.cpp:109

static void mat_perspective(M4& m, float fov, float aspect, float n, float f) {   // Transform.cpp:70-80
	float h = float(std::tan(fov * kPi / 180 * .5f)) * n;   // float argument -> float overload (tanf), as in the reference TU
	float w = h * aspect;
	mat_frustum(m, -w, w, -h, h, n, f);
}

static void mat_lookat(M4& m, V3 eye, V3 target, V3 up) {   // Transform.cpp:127-156
	V3 dir = normalized(sub(target, eye));
	V3 right = normalized(refcross(dir, up));
	up = refcross(right, dir);
	const float* R = &right.x; const float* U = &up.x; const float* D = &dir.x;
	for (int i = 0; i < 3; ++i) { m.f[i][0] = R[i]; m.f[i][1] = U[i]; m.f[i][2] = -D[i]; }
	for (int i = 0; i < 3; ++i) { m.f[i][3] = 0; m.f[3][i] = 0; }
	m.f[3][3] = 1;
	// Translate(-eye) == (*this) *= Translation  (Vector.cpp:325-330, 478-527)
	M4 t;
	for (int j = 0; j < 4; ++j) for (int i = 0; i < 3; ++i) t.f[i][j] = (float)(i == j);
	t.f[3][0] = -eye.x; t.f[3][1] = -eye.y; t.f[3][2] = -eye.z; t.f[3][3] = 1;
	M4 r;
	mat_product(r, m, t);   // same term order as operator*= : f[0][r]*T[c][0] + f[1][r]*T[c][1] + ...
	m = r;
}

// ------------------------------------------------------------------------------------------
// Patch geometry (Patch.cpp:253-276) and Camera::lookFromPatch (Camera.cpp:19-52)
// ------------------------------------------------------------------------------------------
enum { LOOK_FRONT = 0, LOOK_UP, LOOK_DOWN, LOOK_LEFT, LOOK_RIGHT };   // Camera.h:18-24
static const int kLookPerm[5] = { LOOK_UP, LOOK_DOWN, LOOK_LEFT, LOOK_RIGHT, LOOK_FRONT }; // Main.h:210-211

struct Quad { V3 v1, v2, v3, v4; };
static inline Quad quad_from(const float* p) {
	Quad q = { v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), v3(p[6], p[7], p[8]), v3(p[9], p[10], p[11]) };
	return q;
}
static inline V3 quad_center(const Quad& q) {
	return v3((q.v1.x + q.v2.x + q.v3.x + q.v4.x) / 4.0f, (q.v1.y + q.v2.y + q.v3.y + q.v4.y) / 4.0f,
	          (q.v1.z + q.v2.z + q.v3.z + q.v4.z) / 4.0f);
}
static inline V3 quad_normal(const Quad& q) { return refcross(sub(q.v2, q.v1), sub(q.v4, q.v1)); } // A.Cross(B)
static inline V3 quad_up(const Quad& q) { return sub(q.v4, q.v1); }

static void build_mvp(M4& mvp, const Quad& q, int look) {   // Main.cpp:1172-1183
	V3 eye = quad_center(q), normal = quad_normal(q), pup = quad_up(q), target, up;
	switch (look) {
	case LOOK_FRONT: target = normal; up = pup; break;
	case LOOK_UP: target = pup; up = neg(normal); break;
	case LOOK_DOWN: target = neg(pup); up = normal; break;
	case LOOK_LEFT: target = neg(refcross(normal, pup)); up = pup; break;
	default: target = refcross(normal, pup); up = pup; break;
	}
	M4 proj, mv;
	mat_perspective(proj, 90, 1.0f, 0.01f, 1000);
	mat_lookat(mv, eye, add(target, eye), up);   // Camera.cpp:97-103 (Identity() *= M is exact)
	mat_product(mvp, proj, mv);
}

// ------------------------------------------------------------------------------------------
// Scene: PrimitiveModel data (PrimitiveModel.cpp:92-213), Patch::divide (Patch.cpp:47-222),
// Model::subdivide (Model.cpp:27-60), ModelContainer::updateData (ModelContainer.cpp:81-155)
// ------------------------------------------------------------------------------------------
struct OPatch { Quad q; V3 color, illum, rad; };

static void divide(const OPatch& p, double area, std::vector<OPatch>& out) {
	V3 A = p.q.v1, B = p.q.v2, C = p.q.v3, D = p.q.v4;
	V3 u1 = sub(A, C), u2 = sub(B, D);
	double phi = std::acos(dot(u1, u2) / (len(u1) * len(u2)));   // float argument -> float overload, widened to double
	double S = 0.5f * len(u1) * len(u2) * std::sin(phi);
	if (S <= (area * 1.01)) { out.push_back(p); return; }
	double a = std::sqrt(area);
	unsigned kx = (unsigned)(std::ceil(len(sub(B, A)) / a));
	unsigned ky = (unsigned)(std::ceil(len(sub(C, B)) / a));
	V3 pCD = divf(sub(C, D), float(kx));
	V3 pAB = divf(sub(B, A), float(kx));
	for (unsigned i = 0; i < kx * ky; i++) {
		unsigned col = i % kx, row = i / kx;
		V3 bCD = add(mulf(pCD, float(col)), D), bAB = add(mulf(pAB, float(col)), A);
		V3 bCD1 = add(mulf(pCD, float(col + 1)), D), bAB1 = add(mulf(pAB, float(col + 1)), A);
		OPatch n = p;
		n.q.v1 = add(bAB, mulf(divf(sub(bCD, bAB), float(ky)), float(row)));
		n.q.v2 = add(bAB1, mulf(divf(sub(bCD1, bAB1), float(ky)), float(row)));
		n.q.v3 = add(bAB1, mulf(divf(sub(bCD1, bAB1), float(ky)), float(row + 1)));
		n.q.v4 = add(bAB, mulf(divf(sub(bCD, bAB), float(ky)), float(row + 1)));
		out.push_back(n);
	}
}

#define W3 1.0f, 1.0f, 1.0f
static const float kRoom[60] = {
	0.0f, 0.0f, 5.592f, 5.496f, 0.0f, 5.592f, 5.560f, 5.488f, 5.592f, 0.0f, 5.488f, 5.592f,
	0.0f, 5.488f, 0.0f, 0.0f, 5.488f, 5.592f, 5.560f, 5.488f, 5.592f, 5.560f, 5.488f, 0.0f,
	0.0f, 0.0f, 0.0f, 5.528f, 0.0f, 0.0f, 5.496f, 0.0f, 5.592f, 0.0f, 0.0f, 5.592f,
	0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 5.592f, 0.0f, 5.488f, 5.593f, 0.0f, 5.488f, 0.0f,
	5.496f, 0.0f, 5.592f, 5.528f, 0.0f, 0.0f, 5.560f, 5.488f, 0.0f, 5.560f, 5.488f, 5.592f };
static const float kRoomColors[15] = { W3, W3, W3, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f };
static const float kLight[12] = { 3.430f, 5.485f, 2.270f, 2.130f, 5.485f, 2.270f, 2.130f, 5.485f, 3.320f, 3.430f, 5.485f, 3.320f };
static const float kClosure[12] = { 5.528f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 5.488f, 0.0f, 5.560f, 5.488f, 0.0f };
static const float kCube[60] = {
	1.3f, 1.65f, 0.65f, 2.9f, 1.65f, 1.14f, 2.4f, 1.65f, 2.72f, 0.82f, 1.65f, 2.25f,
	2.9f, 0.0f, 1.14f, 2.4f, 0.0f, 2.72f, 2.4f, 1.65f, 2.72f, 2.9f, 1.65f, 1.14f,
	1.3f, 0.0f, 0.65f, 2.9f, 0.0f, 1.14f, 2.9f, 1.65f, 1.14f, 1.3f, 1.65f, 0.65f,
	0.82f, 0.0f, 2.25f, 1.3f, 0.0f, 0.65f, 1.3f, 1.65f, 0.65f, 0.82f, 1.65f, 2.25f,
	2.4f, 0.0f, 2.72f, 0.82f, 0.0f, 2.25f, 0.82f, 1.65f, 2.25f, 2.4f, 1.65f, 2.72f };
static const float kBlock[60] = {
	4.23f, 3.3f, 2.47f, 4.72f, 3.3f, 4.06f, 3.14f, 3.3f, 4.56f, 2.65f, 3.3f, 2.96f,
	4.23f, 0.0f, 2.47f, 4.72f, 0.0f, 4.06f, 4.72f, 3.3f, 4.06f, 4.23f, 3.3f, 2.47f,
	4.72f, 0.0f, 4.06f, 3.14f, 0.0f, 4.56f, 3.14f, 3.3f, 4.56f, 4.72f, 3.3f, 4.06f,
	3.14f, 0.0f, 4.56f, 2.65f, 0.0f, 2.96f, 2.65f, 3.3f, 2.96f, 3.14f, 3.3f, 4.56f,
	2.65f, 0.0f, 2.96f, 4.23f, 0.0f, 2.47f, 4.23f, 3.3f, 2.47f, 2.65f, 3.3f, 2.96f };

static void push_quads(std::vector<OPatch>& m, const float* v, int nquads, const float* colors) {
	for (int i = 0; i < nquads; i++) {
		OPatch p;
		p.q = quad_from(v + 12 * i);
		p.color = colors ? v3(colors[3 * i], colors[3 * i + 1], colors[3 * i + 2]) : v3(1.0f, 1.0f, 1.0f);
		p.illum = v3(0, 0, 0); p.rad = v3(0, 0, 0);
		m.push_back(p);
	}
}

static std::vector<OPatch> g_scene;

static void subdivide_into(const std::vector<OPatch>& model, double area, std::vector<OPatch>& scene) {
	for (size_t i = 0; i < model.size(); i++) {
		if (area > 0) divide(model[i], area, scene);   // PrimitiveModel.cpp:79-84: only when area > 0
		else scene.push_back(model[i]);
	}
}

static void build_cornell(double area) {
	g_scene.clear();
	std::vector<OPatch> room, closure, cube, block;
	push_quads(room, kRoom, 5, kRoomColors);
	{   // light: colour 1, illumination 1, radiosity 100 (PrimitiveModel.cpp:19-29)
		OPatch l; l.q = quad_from(kLight); l.color = v3(1.0f, 1.0f, 1.0f);
		V3 e = v3(1.0f, 1.0f, 1.0f); l.illum = e; l.rad = mulf(e, 100);
		room.push_back(l);
	}
	push_quads(closure, kClosure, 1, NULL);
	push_quads(cube, kCube, 5, NULL);
	push_quads(block, kBlock, 5, NULL);
	// addModel order: room, roomClosure, cube, block (ModelContainer.cpp:44-47)
	subdivide_into(room, area, g_scene);
	subdivide_into(closure, area, g_scene);
	subdivide_into(cube, area, g_scene);
	subdivide_into(block, area, g_scene);
}

// WaveFrontModel::parse (WaveFrontModel.cpp:15-144): "v " lines (/1000), "f " lines with 3 or 4
// indices (text after '/' ignored), triangles become degenerate quads, >4-gons dropped, bad indices
// skipped; patches are colourless and unlit (4-arg Patch ctor, Patch.cpp:8-13).
static bool build_obj(const char* path, double area) {
	g_scene.clear();
	std::ifstream f(path);
	if (!f.good()) return false;
	std::vector<V3> verts;
	std::vector<OPatch> model;
	std::string buffer;
	while (f.good()) {
		std::getline(f, buffer);
		if (buffer.find("v ") == 0) {
			buffer.erase(0, 2);
			std::vector<float> pts; std::istringstream str(buffer); float x;
			while (str >> x) pts.push_back(x);
			if (pts.size() < 3) return false;
			verts.push_back(v3(pts[0] / 1000, pts[1] / 1000, pts[2] / 1000));
			continue;
		}
		if (buffer.find("f ") == 0) {
			buffer.erase(0, 2);
			std::vector<unsigned> fv;
			while (buffer.size() > 0 && fv.size() <= 4) {
				size_t pos = buffer.find_first_of(' ');
				std::string tok = buffer.substr(0, pos);
				std::istringstream str(tok); int vi;
				if (str >> vi) fv.push_back((unsigned)(vi - 1));
				if (pos == std::string::npos) buffer.erase(); else buffer.erase(0, pos + 1);
			}
			if (fv.size() == 3) fv.push_back(fv.back());
			if (fv.size() != 4) continue;
			bool ok = true;
			for (int i = 0; i < 4; i++) if (fv[i] >= verts.size()) ok = false;
			if (!ok) continue;
			OPatch p; p.q.v1 = verts[fv[0]]; p.q.v2 = verts[fv[1]]; p.q.v3 = verts[fv[2]]; p.q.v4 = verts[fv[3]];
			p.color = v3(0, 0, 0); p.illum = v3(0, 0, 0); p.rad = v3(0, 0, 0);
			model.push_back(p);
		}
	}
	subdivide_into(model, area, g_scene);
	return true;
}

// ------------------------------------------------------------------------------------------
// Colour codec (Colors.cpp:31-110) — stateful like the reference
// ------------------------------------------------------------------------------------------
struct Codec { int shift[3]; unsigned mask[3], revMask[3], correction, range; };
static Codec g_codec;
static inline unsigned mask32(int b) { return ((1u << (b - 1)) - 1) | (1u << (b - 1)); }   // Colors.cpp:5

static void codec_setup(Codec& c, unsigned colors) {
	const int bits[3] = { 10, 10, 10 };
	++colors;
	short m = (short)int(ceil(log((double)colors) / log(2.0)));
	short totbits = bits[0] + bits[1] + bits[2];
	short z = totbits - m;
	if (z < 0) z = 0;
	c.range = std::min(colors, (unsigned)pow(2.0, totbits));
	short zr = z / 3, zg = zr, zb = z - 2 * zr;
	c.shift[0] = zr; c.shift[1] = zg + bits[0]; c.shift[2] = zb + bits[0] + bits[1];
	c.mask[0] = mask32(bits[0] - zr); c.mask[1] = mask32(bits[1] - zg); c.mask[2] = mask32(bits[2] - zb);
	c.mask[1] <<= bits[0] - zr;
	c.mask[2] <<= bits[0] - zr + bits[1] - zg;
	c.shift[1] -= bits[0] - zr;
	c.shift[2] -= bits[0] - zr + bits[1] - zg;
	for (int i = 0; i < 3; i++) c.revMask[i] = c.mask[i] << c.shift[i];
	unsigned tmp = 1 + (1 << bits[0]) + (1 << (bits[0] + bits[1]));
	c.correction = (1 << (zr - 1)) | (1 << (bits[0] + zg - 1)) | (1 << (bits[0] + bits[1] + zb - 1));
	c.correction -= tmp;
}
static inline unsigned codec_color(const Codec& c, unsigned idx) {
	return ((idx & c.mask[0]) << c.shift[0]) | ((idx & c.mask[1]) << c.shift[1]) | ((idx & c.mask[2]) << c.shift[2]);
}
static inline unsigned codec_unpack(const Codec& c, unsigned col) {   // macro injected at Main.cpp:463-467
	return ((col & c.revMask[0]) >> c.shift[0]) | ((col & c.revMask[1]) >> c.shift[1]) | ((col & c.revMask[2]) >> c.shift[2]);
}

// ------------------------------------------------------------------------------------------
// Atlas layout: viewports + scissors per face (Main.cpp:314-389), one hemicube (hi = 0)
// order = kLookPerm: UP, DOWN, LEFT, RIGHT, FRONT
// ------------------------------------------------------------------------------------------
struct Rect { int x, y, w, h; };
static void atlas_rects(int N, Rect vp[5], Rect sc[5]) {
	vp[0] = { 0, N, N, N };               sc[0] = { 0, N, N, N / 2 };
	vp[1] = { N, N / 2, N, N };           sc[1] = { N, N, N, N / 2 };
	vp[2] = { -1 * int(N / 2), 0, N, N }; sc[2] = { 0, 0, N / 2, N };
	vp[3] = { int(N * 1.5), 0, N, N };    sc[3] = { int(N * 1.5), 0, N / 2, N };
	vp[4] = { N / 2, 0, N, N };           sc[4] = { N / 2, 0, N, N };
}

// ------------------------------------------------------------------------------------------
// GL hemicube raster restated (Shaders.cpp:235-260; Main.cpp:12,1104,1148-1200,689-725 + GL rules)
//
//   vertex stage : clip = MVP * (pos,1), row r = ((m[0][r]*x + m[1][r]*y) + m[2][r]*z) + m[3][r]
//   clipping     : near plane only (z + w >= 0); the far plane (1000) is never reached by the
//                  scenes and x/y are handled by the scissor (guard-band style).  New vertices are
//                  always interpolated from the inside vertex towards the outside one, so two
//                  triangles sharing a clipped edge get the identical vertex.
//   viewport     : iw = 1/w, xw = (x*iw) * (N/2) + (vx + N/2), zw = (z*iw) * 0.5 + 0.5
//   snapping     : 8 sub-pixel bits, round to nearest even
//   culling      : GL_CULL_FACE, back, front = CCW (Main.cpp:12) -> keep signed area > 0
//   coverage     : pixel centres, exact integer edge functions, top-left rule on ties
//   depth        : linear interpolation of zw, quantised to 24 bit (GL_DEPTH_COMPONENT24,
//                  Main.cpp:283-289), GL_LESS against a buffer cleared to 1.0, draw order = patch id
// ------------------------------------------------------------------------------------------
struct CV { float x, y, z, w; };
static inline CV xform(const M4& m, V3 p) {
	CV c;
	c.x = ((m.f[0][0] * p.x + m.f[1][0] * p.y) + m.f[2][0] * p.z) + m.f[3][0];
	c.y = ((m.f[0][1] * p.x + m.f[1][1] * p.y) + m.f[2][1] * p.z) + m.f[3][1];
	c.z = ((m.f[0][2] * p.x + m.f[1][2] * p.y) + m.f[2][2] * p.z) + m.f[3][2];
	c.w = ((m.f[0][3] * p.x + m.f[1][3] * p.y) + m.f[2][3] * p.z) + m.f[3][3];
	return c;
}
static inline CV clip_lerp(const CV& in, const CV& out, float din, float dout) {
	float t = din / (din - dout);
	CV r;
	r.x = in.x + t * (out.x - in.x);
	r.y = in.y + t * (out.y - in.y);
	r.z = in.z + t * (out.z - in.z);
	r.w = in.w + t * (out.w - in.w);
	return r;
}
static inline int snap(float v) {
	float s = v * 256.0f;
	if (s > 536870912.0f) s = 536870912.0f;       // +-2^29 keeps coordinate differences inside int32
	if (s < -536870912.0f) s = -536870912.0f;
	return (int)lrintf(s);
}
static inline int64_t edge_fn(int ax, int ay, int bx, int by, int cx, int cy) {
	return (int64_t)(bx - ax) * (int64_t)(cy - ay) - (int64_t)(by - ay) * (int64_t)(cx - ax);
}
static inline int edge_bias(int ax, int ay, int bx, int by) {   // 0 for top/left edges, -1 otherwise
	int dx = bx - ax, dy = by - ay;
	return (dy < 0 || (dy == 0 && dx < 0)) ? 0 : -1;
}

static const uint64_t kClearKey = 0xFFFFFFFFFFFFFFFFull;

// optional work statistics (orc_raster_stats): [0] triangles reaching raster_tri, [1] after the frustum reject, [2] front-facing,
// [3] with a non-empty scissored bbox, [4] bbox pixels, [5] covered fragments, [6] fragments that won the depth test when drawn,
// [8..8+20) histogram of bbox area by floor(log2(area))
static uint64_t* g_stats = NULL;

static void raster_tri(const CV t[3], const Rect& vp, const Rect& sc, int W, uint32_t id1, uint64_t* keys) {
	// exact trivial reject: a triangle wholly beyond one viewport edge cannot produce a pixel inside the
	// scissor (x > w  =>  x/w >= 1  =>  every snapped X lies at or beyond the viewport edge); w > 0 here
	if (g_stats) g_stats[0]++;
	if ((t[0].x > t[0].w && t[1].x > t[1].w && t[2].x > t[2].w) || (t[0].x < -t[0].w && t[1].x < -t[1].w && t[2].x < -t[2].w) ||
	    (t[0].y > t[0].w && t[1].y > t[1].w && t[2].y > t[2].w) || (t[0].y < -t[0].w && t[1].y < -t[1].w && t[2].y < -t[2].w)) return;
	if (g_stats) g_stats[1]++;
	int X[3], Y[3]; float Z[3];
	float hw = (float)vp.w * 0.5f, hh = (float)vp.h * 0.5f;
	float ox = (float)vp.x + hw, oy = (float)vp.y + hh;
	for (int i = 0; i < 3; i++) {
		float iw = 1.0f / t[i].w;                               // perspective divide: reciprocal, then multiply
		float xn = t[i].x * iw, yn = t[i].y * iw, zn = t[i].z * iw;
		X[i] = snap(xn * hw + ox);
		Y[i] = snap(yn * hh + oy);
		Z[i] = zn * 0.5f + 0.5f;
	}
	int64_t area2 = edge_fn(X[0], Y[0], X[1], Y[1], X[2], Y[2]);
	if (area2 <= 0) return;
	if (g_stats) g_stats[2]++;
	int minx = std::min(X[0], std::min(X[1], X[2])), maxx = std::max(X[0], std::max(X[1], X[2]));
	int miny = std::min(Y[0], std::min(Y[1], Y[2])), maxy = std::max(Y[0], std::max(Y[1], Y[2]));
	int px0 = std::max((minx - 128 + 255) >> 8, sc.x), px1 = std::min((maxx - 128) >> 8, sc.x + sc.w - 1);
	int py0 = std::max((miny - 128 + 255) >> 8, sc.y), py1 = std::min((maxy - 128) >> 8, sc.y + sc.h - 1);
	if (px0 > px1 || py0 > py1) return;
	if (g_stats) { g_stats[3]++; uint64_t a = (uint64_t)(px1 - px0 + 1) * (py1 - py0 + 1); g_stats[4] += a; int l = 0; while ((a >> (l + 1)) != 0) l++; g_stats[8 + (l < 19 ? l : 19)]++; }
	int b0 = edge_bias(X[1], Y[1], X[2], Y[2]), b1 = edge_bias(X[2], Y[2], X[0], Y[0]), b2 = edge_bias(X[0], Y[0], X[1], Y[1]);
	float inv_area = 1.0f / (float)area2;
	float dz1 = Z[1] - Z[0], dz2 = Z[2] - Z[0];
	// exact integer edge functions, stepped incrementally (one pixel = 256 sub-pixel units)
	const int cx0 = px0 * 256 + 128, cy0 = py0 * 256 + 128;
	int64_t r0 = edge_fn(X[1], Y[1], X[2], Y[2], cx0, cy0) + b0;
	int64_t r1 = edge_fn(X[2], Y[2], X[0], Y[0], cx0, cy0) + b1;
	int64_t r2 = edge_fn(X[0], Y[0], X[1], Y[1], cx0, cy0) + b2;
	const int64_t sx0 = -(int64_t)(Y[2] - Y[1]) * 256, sy0 = (int64_t)(X[2] - X[1]) * 256;
	const int64_t sx1 = -(int64_t)(Y[0] - Y[2]) * 256, sy1 = (int64_t)(X[0] - X[2]) * 256;
	const int64_t sx2 = -(int64_t)(Y[1] - Y[0]) * 256, sy2 = (int64_t)(X[1] - X[0]) * 256;
	for (int py = py0; py <= py1; py++, r0 += sy0, r1 += sy1, r2 += sy2) {
		int64_t e0 = r0, e1 = r1, e2 = r2;
		uint64_t* krow = keys + (size_t)py * W;
		for (int px = px0; px <= px1; px++, e0 += sx0, e1 += sx1, e2 += sx2) {
			if ((e0 | e1 | e2) < 0) continue;               // biased edge values: inside iff all >= 0
			float l1 = (float)(e1 - b1) * inv_area, l2 = (float)(e2 - b2) * inv_area;
			float z = (Z[0] + l1 * dz1) + l2 * dz2;
			if (z < 0.0f) z = 0.0f;
			if (z > 1.0f) z = 1.0f;
			uint32_t d = (uint32_t)lrintf(z * 16777215.0f);
			if (d >= 0xFFFFFFu) continue;                 // GL_LESS against the cleared 1.0
			uint64_t key = ((uint64_t)d << 32) | id1;       // LESS + draw order == min over (depth, id)
			if (g_stats) { g_stats[5]++; if (key < krow[px]) g_stats[6]++; }
			if (key < krow[px]) krow[px] = key;
		}
	}
}

static void raster_quad_face(const M4& mvp, const Quad& q, const Rect& vp, const Rect& sc, int W, uint32_t id1, uint64_t* keys) {
	CV c[4] = { xform(mvp, q.v1), xform(mvp, q.v2), xform(mvp, q.v3), xform(mvp, q.v4) };
	static const int tri[2][3] = { { 0, 1, 2 }, { 0, 2, 3 } };   // ModelContainer.cpp:112-117
	for (int t = 0; t < 2; t++) {
		CV in[3] = { c[tri[t][0]], c[tri[t][1]], c[tri[t][2]] };
		float d[3]; int nin = 0;
		for (int i = 0; i < 3; i++) { d[i] = in[i].z + in[i].w; nin += d[i] >= 0.0f; }
		if (nin == 0) continue;
		if (nin == 3) { raster_tri(in, vp, sc, W, id1, keys); continue; }
		CV poly[4]; int n = 0;
		for (int i = 0; i < 3; i++) {
			int j = (i + 1) % 3;
			bool ii = d[i] >= 0.0f, ji = d[j] >= 0.0f;
			if (ii) poly[n++] = in[i];
			if (ii && !ji) poly[n++] = clip_lerp(in[i], in[j], d[i], d[j]);
			else if (!ii && ji) poly[n++] = clip_lerp(in[j], in[i], d[j], d[i]);
		}
		CV t0[3] = { poly[0], poly[1], poly[2] };
		raster_tri(t0, vp, sc, W, id1, keys);
		if (n == 4) { CV t1[3] = { poly[0], poly[2], poly[3] }; raster_tri(t1, vp, sc, W, id1, keys); }
	}
}

// Render the 5 faces of one hemicube seen from `shooter` into keys[W*H] (cleared here).
static void render_hemicube_keys(unsigned P, const float* verts, unsigned shooter, int N, uint64_t* keys, int threads) {
	const int W = 2 * N, H = (int)(N * 1.5);
	const size_t RES = (size_t)W * H;
	Rect vp[5], sc[5];
	atlas_rects(N, vp, sc);
	M4 mvp[5];
	Quad sq = quad_from(verts + 12 * (size_t)shooter);
	for (int f = 0; f < 5; f++) build_mvp(mvp[f], sq, kLookPerm[f]);
	for (size_t i = 0; i < RES; i++) keys[i] = kClearKey;
	if (threads <= 1) {
		for (int f = 0; f < 5; f++)
			for (unsigned p = 0; p < P; p++)
				raster_quad_face(mvp[f], quad_from(verts + 12 * (size_t)p), vp[f], sc[f], W, p + 1, keys);
		return;
	}
#ifdef _OPENMP
	// patch chunks rasterised into private key buffers, merged by min: identical to the sequential
	// LESS-in-draw-order result because the key orders by (depth, id).
	static std::vector<std::vector<uint64_t> > priv;
	if ((int)priv.size() < threads) priv.resize(threads);
	#pragma omp parallel num_threads(threads)
	{
		int t = omp_get_thread_num(), T = omp_get_num_threads();
		uint64_t* k = keys;
		if (t > 0) { priv[t].assign(RES, kClearKey); k = priv[t].data(); }
		unsigned p0 = (unsigned)((uint64_t)P * t / T), p1 = (unsigned)((uint64_t)P * (t + 1) / T);
		for (int f = 0; f < 5; f++)
			for (unsigned p = p0; p < p1; p++)
				raster_quad_face(mvp[f], quad_from(verts + 12 * (size_t)p), vp[f], sc[f], W, p + 1, k);
		#pragma omp barrier
		size_t i0 = RES * t / T, i1 = RES * (t + 1) / T;
		for (int o = 1; o < T; o++) {
			const uint64_t* s = priv[o].data();
			for (size_t i = i0; i < i1; i++) if (s[i] < keys[i]) keys[i] = s[i];
		}
	}
#endif
}

// ------------------------------------------------------------------------------------------
// Form-factor table (FormFactors.cpp:23-67, 280-339)
// ------------------------------------------------------------------------------------------
static void formfactors(unsigned N, unsigned hemicubes, float* out) {
	const float Pi = 3.1415926535897932384626433832795028841931f;   // FormFactors.cpp:3
	int n = (int)N;
	std::vector<float> top((size_t)n * n), side((size_t)n * n / 2);
	float half_px = (1.0f / n);
	float px_area = (2.0f / n);
	px_area *= px_area;
	for (int x = 0; x < n; x++)
		for (int y = 0; y < n; y++) {
			float dx = ((x - n / 2) / (n / 2.0f)) + half_px;
			float dy = ((y - n / 2) / (n / 2.0f)) + half_px;
			float f = (dx * dx + dy * dy + 1);
			f *= f * Pi;
			top[x + (size_t)y * n] = px_area / f;
		}
	for (int x = 0; x < n; x++)
		for (int y = 0; y < n / 2; y++) {
			float dx = (x - n / 2) / (n / 2.0f) + half_px;
			float dy = (n / 2 - 1 - y) / (n / 2.0f) + half_px;
			float f = (dx * dx + dy * dy + 1);
			f *= f * Pi;
			side[x + (size_t)y * n] = (px_area * (dy + half_px)) / f;
		}
	unsigned HW = N, HH = N, TW = (unsigned)(N * 2), TH = (unsigned)(N * 1.5), RES = TW * TH;
	for (unsigned i = 0; i < RES; i++) {
		unsigned x = i % TW, y = i / TW;
		if (x >= HW / 2 && x < HW * 1.5 && y < HH) out[i] = top[y * HW + (x - HW / 2)];
		else if (x < HW / 2 && y < HH) out[i] = side[(HH / 2 - x) * HW - y - 1];
		else if (x >= HW * 1.5 && y < HH) out[i] = side[(x - (unsigned)(HW * 1.5)) * HW + y];
		else if (x < HW && y >= HH) out[i] = side[(y - HH) * HW + x];
		else if (x >= HW && y >= HH) out[i] = side[((HH / 2 - 1) - (y - HH)) * HW + (x - HW)];
	}
	for (unsigned h = 1; h < hemicubes; h++)
		for (size_t i = (size_t)RES * h; i < (size_t)RES * (h + 1); i++) out[i] = out[i % RES];
}

// ------------------------------------------------------------------------------------------
// Shooter selection (ModelContainer.cpp:242-299)
// ------------------------------------------------------------------------------------------
struct EnergyLess {
	const float* rad;
	bool operator()(unsigned a, unsigned b) const {
		V3 A = v3(rad[3 * a], rad[3 * a + 1], rad[3 * a + 2]), B = v3(rad[3 * b], rad[3 * b + 1], rad[3 * b + 2]);
		return len2(A) < len2(B);
	}
};
static void select_reference(unsigned P, const float* rad, unsigned count, unsigned* ids, int* is_null) {
	std::list<unsigned> tops;
	EnergyLess c; c.rad = rad;
	for (unsigned pi = 0; pi < P; pi++) {
		float e = len2(v3(rad[3 * pi], rad[3 * pi + 1], rad[3 * pi + 2]));
		bool take = tops.empty();
		if (!take) {
			unsigned b = tops.back();
			take = e > 0 && len2(v3(rad[3 * b], rad[3 * b + 1], rad[3 * b + 2])) <= e;
		}
		if (take) {
			tops.push_back(pi);
			tops.sort(c);
			tops.reverse();
			if (tops.size() > count) {
				std::list<unsigned>::iterator it = tops.begin();
				for (unsigned i = 0; i < count; i++) it++;
				tops.erase(it, tops.end());
			}
		}
	}
	std::list<unsigned>::iterator it = tops.begin();
	for (unsigned i = 0; i < count; i++) {
		if (it == tops.end()) { ids[i] = 0; is_null[i] = 1; continue; }
		ids[i] = *it; is_null[i] = 0; it++;
	}
}

// Clean top-k (the documented DIVERGENT schedule used for batched/multi-GPU runs): the `count`
// patches with the largest |B|^2 > 0, ordered by (energy desc, id asc); fewer if fewer have energy.
static void select_topk(unsigned P, const float* rad, unsigned count, unsigned* ids, int* is_null) {
	std::vector<std::pair<float, unsigned> > v;
	for (unsigned pi = 0; pi < P; pi++) {
		float e = len2(v3(rad[3 * pi], rad[3 * pi + 1], rad[3 * pi + 2]));
		if (e > 0) v.push_back(std::make_pair(e, pi));
	}
	size_t k = std::min<size_t>(count, v.size());
	std::partial_sort(v.begin(), v.begin() + k, v.end(), [](const std::pair<float, unsigned>& a, const std::pair<float, unsigned>& b) {
		return a.first > b.first || (a.first == b.first && a.second < b.second); });
	for (unsigned i = 0; i < count; i++) {
		if (i < k) { ids[i] = v[i].second; is_null[i] = 0; } else { ids[i] = 0; is_null[i] = 1; }
	}
}

// ------------------------------------------------------------------------------------------
// Reference-format atlas: patch-view shader + RGBA8 framebuffer (Shaders.cpp:246-259, Main.cpp:173,277)
// ------------------------------------------------------------------------------------------
static void encode_atlas(const Codec& c, const uint32_t* ids, size_t n, uint8_t* rgba) {
	for (size_t i = 0; i < n; i++) {
		unsigned col = ids[i] ? codec_color(c, ids[i]) : 0;   // colour index = patch id + 1; cleared = black
		unsigned ch[3] = { col & 1023u, (col >> 10) & 1023u, (col >> 20) & 1023u };
		for (int k = 0; k < 3; k++) {
			float v = (float)ch[k] / 1024.0f;                 // v_color = v_col / 1024.0
			rgba[4 * i + k] = (uint8_t)lrintf(v * 255.0f);    // float -> UNORM8
		}
		rgba[4 * i + 3] = ids[i] ? 255 : 0;
	}
}

// Kernel_ProcessHemicube.h:23-69 executed work-item by work-item (gid1 = y outer, gid0 inner).
// Returns the final write index (== number of record slots used, padding included).
static unsigned process_hemicube_cl(const Codec& c, const uint8_t* rgba, const float* ffactors, unsigned n_width, unsigned n_height,
                                    unsigned n_hemicubes, unsigned workitems_x, uint32_t* p_hemicubes, uint32_t* p_ids, float* p_energies) {
	const int alloc_block = 4;
	unsigned write_index = 0;
	unsigned span = n_width / workitems_x;
	for (unsigned y = 0; y < n_height * n_hemicubes; y++)
		for (unsigned g0 = 0; g0 < workitems_x; g0++) {
			int x0 = (int)std::min(g0 * span, n_width - 1);
			int x1 = (int)std::min((unsigned)x0 + span, n_width);
			const float* ff = ffactors + (size_t)n_width * y;
			const uint8_t* row = rgba + 4 * (size_t)n_width * y;
			int space = 0; unsigned wid = 0;
			while (x0 < x1) {
				// read_imagef on a UNORM8 image returns c/255.0f; (uint)(f*1024) truncates
				#define PID(px) (1048576u * (uint32_t)((row[4*(px)+2] / 255.0f) * 1024) + 1024u * (uint32_t)((row[4*(px)+1] / 255.0f) * 1024) + (uint32_t)((row[4*(px)] / 255.0f) * 1024))
				uint32_t pid = PID(x0);
				float energy = 0;
				while (x0 < x1) {
					uint32_t act = PID(x0);
					if (act == pid) { energy += ff[x0]; ++x0; } else break;
				}
				#undef PID
				if (pid > 0) {
					if (!space) { wid = write_index; write_index += alloc_block; space = alloc_block - 1; }
					else { ++wid; --space; }
					p_hemicubes[wid] = y / n_height;
					p_ids[wid] = codec_unpack(c, pid + c.correction) - 1;
					p_energies[wid] = energy;
				}
			}
			for (; space > 0; --space) { ++wid; p_hemicubes[wid] = 0; p_ids[wid] = 0; p_energies[wid] = 0; }
		}
	return write_index;
}

// Direct form of the same thing on a decoded id atlas (id+1 per pixel): identical run structure
// and float summation order as process_hemicube_cl + the record gather of Main.cpp:1257-1269.
static void process_hemicube_ids(const uint32_t* ids, const float* ff, unsigned n_width, unsigned n_height,
                                 unsigned workitems_x, unsigned P, float* F) {
	unsigned span = n_width / workitems_x;
	for (unsigned y = 0; y < n_height; y++)
		for (unsigned g0 = 0; g0 < workitems_x; g0++) {
			unsigned x0 = std::min(g0 * span, n_width - 1), x1 = std::min(x0 + span, n_width);
			const uint32_t* row = ids + (size_t)n_width * y;
			const float* f = ff + (size_t)n_width * y;
			while (x0 < x1) {
				uint32_t pid = row[x0]; float energy = 0;
				while (x0 < x1 && row[x0] == pid) { energy += f[x0]; ++x0; }
				if (pid > 0 && pid - 1 < P) F[pid - 1] += energy;
			}
		}
}

} // namespace

// ============================================================================================
// extern "C" surface (ctypes)
// ============================================================================================
extern "C" {

unsigned orc_scene_cornell(double area) { build_cornell(area); return (unsigned)g_scene.size(); }
int orc_scene_obj(const char* path, double area) { return build_obj(path, area) ? (int)g_scene.size() : -1; }
unsigned orc_patch_count() { return (unsigned)g_scene.size(); }
void orc_scene_get(float* verts12, float* color3, float* rad3, float* illum3) {
	for (size_t i = 0; i < g_scene.size(); i++) {
		const OPatch& p = g_scene[i];
		if (verts12) memcpy(verts12 + 12 * i, &p.q, 48);
		if (color3) memcpy(color3 + 3 * i, &p.color, 12);
		if (rad3) memcpy(rad3 + 3 * i, &p.rad, 12);
		if (illum3) memcpy(illum3 + 3 * i, &p.illum, 12);
	}
}

// Config::freeze (Config.cpp:30-48)
void orc_config(unsigned side, unsigned hemicubes, unsigned* out9) {
	unsigned W = (unsigned)(side * 2), H = (unsigned)(side * 1.5);
	out9[0] = side; out9[1] = side; out9[2] = W; out9[3] = H; out9[4] = (unsigned)(W * H);
	out9[5] = std::min(4u, W); out9[6] = H * hemicubes; out9[7] = 500; out9[8] = hemicubes;
}

void orc_formfactors(unsigned side, unsigned hemicubes, float* out) { formfactors(side, hemicubes, out); }

void orc_colors_setup(unsigned patches, unsigned* out11) {
	codec_setup(g_codec, patches);
	if (!out11) return;
	for (int i = 0; i < 3; i++) { out11[i] = (unsigned)g_codec.shift[i]; out11[3 + i] = g_codec.revMask[i]; out11[6 + i] = g_codec.mask[i]; }
	out11[9] = g_codec.correction; out11[10] = g_codec.range;
}
unsigned orc_color(unsigned idx) { return codec_color(g_codec, idx); }
unsigned orc_color_index(unsigned color) { return codec_unpack(g_codec, color + g_codec.correction); }   // Colors.cpp:104-108

void orc_select(unsigned P, const float* rad3, unsigned count, int mode, unsigned* ids, int* is_null) {
	if (mode == 0) select_reference(P, rad3, count, ids, is_null); else select_topk(P, rad3, count, ids, is_null);
}

void orc_patch_geom(const float* verts12, float* center3, float* normal3, float* up3) {
	Quad q = quad_from(verts12);
	V3 c = quad_center(q), n = quad_normal(q), u = quad_up(q);
	memcpy(center3, &c, 12); memcpy(normal3, &n, 12); memcpy(up3, &u, 12);
}
void orc_mvp(const float* verts12, int look, float* out16) {
	M4 m; build_mvp(m, quad_from(verts12), look);
	memcpy(out16, m.f, 64);
}
void orc_projection(float* out16) { M4 m; mat_perspective(m, 90, 1.0f, 0.01f, 1000); memcpy(out16, m.f, 64); }

// 5-face hemicube item buffer seen from `shooter`: ids_out[W*H] = patch id + 1 (0 = nothing),
// depth_out (optional) = 24-bit depth (0xFFFFFF where empty).  Row 0 is the bottom row (GL).
void orc_render_hemicube(unsigned P, const float* verts, unsigned shooter, unsigned N, uint32_t* ids_out, uint32_t* depth_out, int threads) {
	size_t RES = (size_t)(2 * N) * (size_t)(unsigned)(N * 1.5);
	std::vector<uint64_t> keys(RES);
	render_hemicube_keys(P, verts, shooter, (int)N, keys.data(), threads);
	for (size_t i = 0; i < RES; i++) {
		bool empty = keys[i] == kClearKey;
		ids_out[i] = empty ? 0u : (uint32_t)(keys[i] & 0xFFFFFFFFu);
		if (depth_out) depth_out[i] = empty ? 0xFFFFFFu : (uint32_t)(keys[i] >> 32);
	}
}

void orc_encode_atlas(unsigned P, const uint32_t* ids, size_t n, uint8_t* rgba) {
	Codec c; codec_setup(c, P);
	encode_atlas(c, ids, n, rgba);
}

unsigned orc_process_hemicube_cl(unsigned P, const uint8_t* rgba, const float* ff, unsigned width, unsigned height, unsigned hemicubes,
                                 unsigned workitems_x, uint32_t* hem_out, uint32_t* ids_out, float* en_out) {
	Codec c; codec_setup(c, P);
	return process_hemicube_cl(c, rgba, ff, width, height, hemicubes, workitems_x, hem_out, ids_out, en_out);
}

// record gather of Main.cpp:1257-1269 for hemicube `hi`; returns the number of out-of-range ids
unsigned orc_gather_records(unsigned P, unsigned n_records, const uint32_t* hem, const uint32_t* ids, const float* en, unsigned hi, float* F) {
	unsigned bad = 0;
	for (unsigned i = 0; i < n_records; i++) {
		if (ids[i] >= P) { bad++; continue; }
		if (hem[i] != hi) continue;
		F[ids[i]] += en[i];
	}
	return bad;
}

void orc_process_hemicube_ids(const uint32_t* ids, const float* ff, unsigned N, unsigned workitems_x, unsigned P, float* F) {
	process_hemicube_ids(ids, ff, 2 * N, (unsigned)(N * 1.5), workitems_x, P, F);
}

// The shooting loop, Main.cpp:1137-1309 (S1..S6 of SURVEY.md §8a).  State arrays are updated in place.
//   select_mode 0 = reference list semantics, 1 = clean top-k
//   via_codec   1 = go through the RGBA8 colour atlas + literal kernel restatement (slow, faithful)
//               0 = decoded-id fast form (same sums, same order)
//   stop_test   1 = honour the |lastEnergy| < 0.1 termination (Main.cpp:1297-1300)
//   schedule    optional [n_batches*k] log of emitter ids (0xFFFFFFFF for NULL)
// Returns the number of batches executed.
unsigned orc_shoot(unsigned P, const float* verts, const float* color3, float* rad3, float* illum3, unsigned N, unsigned k,
                   unsigned n_batches, int select_mode, int via_codec, int stop_test, int threads, uint32_t* schedule, float* last_energy_len) {
	const unsigned W = 2 * N, H = (unsigned)(N * 1.5), RES = W * H;
	const float reflectivity = 0.3f;   // Patch.h:13
	// the reference replicates the one-hemicube table k times (FormFactors.cpp:317-323); only the literal kernel
	// restatement needs the copies, every other consumer indexes the first one
	// work buffers live across calls (a caller that shoots batch by batch must not pay table construction and page
	// faults every time — this is the CPU baseline of bench.py)
	static std::vector<float> ff; static unsigned ff_N = 0, ff_k = 0;
	const unsigned ffk = via_codec ? k : 1;
	if (ff_N != N || ff_k != ffk) { ff.resize((size_t)RES * ffk); formfactors(N, ffk, ff.data()); ff_N = N; ff_k = ffk; }
	static std::vector<uint64_t> keys; keys.resize(RES);
	static std::vector<std::vector<uint64_t> > tkeys;
	const size_t nt = threads > 1 && k > 1 ? (size_t)threads : 0;
	if (tkeys.size() < nt) tkeys.resize(nt);
	for (size_t t = 0; t < nt; t++) tkeys[t].resize(RES);
	// the k-hemicube atlas is only needed by the reference-format path; otherwise every thread keeps one hemicube of ids
	static std::vector<uint32_t> atlas; atlas.resize((size_t)RES * (via_codec ? k : std::max<size_t>(1, nt)));
	static std::vector<float> F, Fall;
	F.assign(P, 0.0f);
	if (!via_codec) Fall.assign((size_t)P * k, 0.0f);
	std::vector<unsigned> em(k); std::vector<int> isnull(k);
	std::vector<V3> snap_rad(k);
	Codec codec; codec_setup(codec, P);
	std::vector<uint8_t> rgba; std::vector<uint32_t> rec_h, rec_i; std::vector<float> rec_e;
	if (via_codec) { rgba.resize((size_t)RES * k * 4); rec_h.resize((size_t)RES * k + 8); rec_i.resize((size_t)RES * k + 8); rec_e.resize((size_t)RES * k + 8); }
	unsigned wx = std::min(4u, W);
	unsigned done = 0;
	float last_len = 0;
	for (unsigned shoot = 0; shoot < n_batches; shoot++) {
		orc_select(P, rad3, k, select_mode, em.data(), isnull.data());                                    // S1
		if (via_codec)
			for (unsigned hi = 0; hi < k; hi++)                                                           // glClear (rendered slots are fully rewritten below)
				if (isnull[hi]) std::fill(atlas.begin() + (size_t)RES * hi, atlas.begin() + (size_t)RES * (hi + 1), 0u);
		for (unsigned hi = 0; hi < k; hi++) {                                                             // S2 (snapshots)
			if (schedule) schedule[(size_t)shoot * k + hi] = isnull[hi] ? 0xFFFFFFFFu : em[hi];
			if (!isnull[hi]) snap_rad[hi] = v3(rad3[3 * em[hi]], rad3[3 * em[hi] + 1], rad3[3 * em[hi] + 2]);
		}
		// the k hemicubes of a batch are independent given the snapshots (Main.cpp:1155-1198): with several host
		// threads each thread renders whole hemicubes; with k == 1 the threads split the patches of the one hemicube
		const bool par_hemi = threads > 1 && k > 1;
		#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (par_hemi)
		for (int hi = 0; hi < (int)k; hi++) {
			if (isnull[hi]) continue;
			uint64_t* kk = keys.data();
			size_t tid = 0;
#ifdef _OPENMP
			if (par_hemi) { tid = (size_t)omp_get_thread_num(); kk = tkeys[tid].data(); }
#endif
			render_hemicube_keys(P, verts, em[hi], (int)N, kk, par_hemi ? 1 : threads);
			uint32_t* a = atlas.data() + (size_t)RES * (via_codec ? (size_t)hi : tid);
			for (unsigned i = 0; i < RES; i++) a[i] = kk[i] == kClearKey ? 0u : (uint32_t)(kk[i] & 0xFFFFFFFFu);
			if (!via_codec) process_hemicube_ids(a, ff.data(), W, H, wx, P, Fall.data() + (size_t)P * hi);
		}
		unsigned nrec = 0;
		if (via_codec) {
			encode_atlas(codec, atlas.data(), (size_t)RES * k, rgba.data());
			nrec = process_hemicube_cl(codec, rgba.data(), ff.data(), W, H, k, wx, rec_h.data(), rec_i.data(), rec_e.data());
		}
		for (unsigned hi = 0; hi < k; hi++) {                                                             // S3 + S4
			if (isnull[hi]) continue;
			float* Fh = F.data();
			if (via_codec) orc_gather_records(P, nrec, rec_h.data(), rec_i.data(), rec_e.data(), hi, F.data());
			else Fh = Fall.data() + (size_t)P * hi;
			V3 ec = v3(color3[3 * em[hi]], color3[3 * em[hi] + 1], color3[3 * em[hi] + 2]);
			for (unsigned i = 0; i < P; i++) {
				// p->radiosity += p_tmp_radiosities[hi] * p_tmp_formfactors[i] * p->getReflectivity() * p_emitters[hi]->getColor();
				V3 d = mulv(mulf(mulf(snap_rad[hi], Fh[i]), reflectivity), ec);
				rad3[3 * i] += d.x; rad3[3 * i + 1] += d.y; rad3[3 * i + 2] += d.z;
			}
			std::fill(Fh, Fh + P, 0.0f);
		}
		V3 last = v3(0, 0, 0);                                                                            // S5
		for (unsigned hi = 0; hi < k; hi++) {
			if (isnull[hi]) continue;
			unsigned e = em[hi];
			last = v3(rad3[3 * e], rad3[3 * e + 1], rad3[3 * e + 2]);
			illum3[3 * e] += snap_rad[hi].x; illum3[3 * e + 1] += snap_rad[hi].y; illum3[3 * e + 2] += snap_rad[hi].z;
			rad3[3 * e] -= snap_rad[hi].x; rad3[3 * e + 1] -= snap_rad[hi].y; rad3[3 * e + 2] -= snap_rad[hi].z;
		}
		last_len = len(last);
		done++;
		if (stop_test && last_len < 0.1) break;                                                           // S6
	}
	if (last_energy_len) *last_energy_len = last_len;
	return done;
}

// work statistics of one hemicube (single thread); out[28]
void orc_raster_stats(unsigned P, const float* verts, unsigned shooter, unsigned N, uint64_t* out) {
	size_t RES = (size_t)(2 * N) * (size_t)(unsigned)(N * 1.5);
	std::vector<uint64_t> keys(RES);
	memset(out, 0, 28 * sizeof(uint64_t));
	g_stats = out;
	render_hemicube_keys(P, verts, shooter, (int)N, keys.data(), 1);
	g_stats = NULL;
}

// Display stage: Colors::smoothShadePatch (Colors.cpp:198-261) for every patch.  Vertex colour = mean over
// the patch and three of its 8 neighbours of colour (.) (I + B); output order lb, rb, rt, lt (Colors.cpp:256-259).
// neighbours: int[P*8], index 0 = top-left, clockwise (Patch.h:50).
void orc_smooth_shade(unsigned P, const float* color3, const float* rad3, const float* illum3, const int* nb8, float* out12) {
	static const int corner[4][3] = { { 5, 6, 7 }, { 3, 4, 5 }, { 1, 2, 3 }, { 7, 0, 1 } };   // lb, rb, rt, lt — in the reference's summation order
	for (unsigned i = 0; i < P; i++) {
		for (int c = 0; c < 4; c++) {
			V3 acc = mulv(v3(color3[3 * i], color3[3 * i + 1], color3[3 * i + 2]),
			              add(v3(illum3[3 * i], illum3[3 * i + 1], illum3[3 * i + 2]), v3(rad3[3 * i], rad3[3 * i + 1], rad3[3 * i + 2])));
			for (int j = 0; j < 3; j++) {
				const unsigned n = (unsigned)nb8[8 * (size_t)i + corner[c][j]];
				V3 t = mulv(v3(color3[3 * n], color3[3 * n + 1], color3[3 * n + 2]),
				            add(v3(illum3[3 * n], illum3[3 * n + 1], illum3[3 * n + 2]), v3(rad3[3 * n], rad3[3 * n + 1], rad3[3 * n + 2])));
				acc = add(acc, t);
			}
			acc = divf(acc, 4);
			out12[12 * (size_t)i + 3 * c] = acc.x; out12[12 * (size_t)i + 3 * c + 1] = acc.y; out12[12 * (size_t)i + 3 * c + 2] = acc.z;
		}
	}
}

int orc_max_threads() {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

} // extern "C"
