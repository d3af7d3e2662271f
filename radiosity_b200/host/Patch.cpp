#include "Patch.h"

static const Vector3f kZero(0.0f, 0.0f, 0.0f);

Patch::Patch() : radiosity(kZero), illumination(kZero), vec1(kZero), vec2(kZero), vec3(kZero), vec4(kZero), color(kZero) {
	for (int i = 0; i < 8; i++) { neighbours[i] = NULL; relativeNeighbours[i] = 0; }
}
Patch::Patch(Vector3f a, Vector3f b, Vector3f c, Vector3f d, Vector3f col, Vector3f illum, Vector3f rad)
	: radiosity(rad), illumination(illum), vec1(a), vec2(b), vec3(c), vec4(d), color(col) {
	for (int i = 0; i < 8; i++) { neighbours[i] = NULL; relativeNeighbours[i] = 0; }
}
Patch::Patch(Vector3f a, Vector3f b, Vector3f c, Vector3f d) : Patch(a, b, c, d, kZero, kZero, kZero) {}
Patch::Patch(Vector3f a, Vector3f b, Vector3f c, Vector3f d, Vector3f col) : Patch(a, b, c, d, col, kZero, kZero) {}
Patch::Patch(Vector3f a, Vector3f b, Vector3f c, Vector3f d, Vector3f col, Vector3f illum) : Patch(a, b, c, d, col, illum, kZero) {}
Patch::~Patch() {}

// Reference: Patch.cpp:47-222.  Area from the diagonals, S = 1/2 |u1| |u2| sin(phi); split counts from the
// edge lengths A-B and B-C; children on a bilinear grid; then the 8-neighbourhood is wired.
std::vector<Patch*>* Patch::divide(double area) {
	const Vector3f A = vec1, B = vec2, C = vec3, D = vec4;
	const Vector3f u1 = A - C, u2 = B - D;
	const double phi = std::acos(u1.f_Dot(u2) / (u1.f_Length() * u2.f_Length()));   // float overload, as in the reference
	const double S = 0.5f * u1.f_Length() * u2.f_Length() * std::sin(phi);
	if (S <= (area * 1.01)) return NULL;          // 1 % tolerance against re-splitting rounding noise

	const double side = std::sqrt(area);
	const unsigned int kx = (unsigned int)(std::ceil((B - A).f_Length() / side));
	const unsigned int ky = (unsigned int)(std::ceil((C - B).f_Length() / side));
	const Vector3f stepTop = (C - D) / float(kx);
	const Vector3f stepBottom = (B - A) / float(kx);

	std::vector<Patch*>* out = new std::vector<Patch*>;
	out->reserve((size_t)kx * ky);
	for (unsigned int i = 0; i < kx * ky; i++) {
		const unsigned int col = i % kx, row = i / kx;
		// grid lines at columns col and col+1, interpolated between the bottom (A-B) and top (D-C) edges
		const Vector3f top0 = stepTop * float(col) + D, bottom0 = stepBottom * float(col) + A;
		const Vector3f top1 = stepTop * float(col + 1) + D, bottom1 = stepBottom * float(col + 1) + A;
		const Vector3f rise0 = (top0 - bottom0) / float(ky), rise1 = (top1 - bottom1) / float(ky);
		out->push_back(new Patch(bottom0 + rise0 * float(row), bottom1 + rise1 * float(row),
		                         bottom1 + rise1 * float(row + 1), bottom0 + rise0 * float(row + 1),
		                         color, illumination, radiosity));
	}

	// neighbours: index j = 0..7 starts top-left and runs clockwise.  A neighbour outside the grid falls
	// back to the patch itself; a diagonal that leaves the grid over the top/bottom slides to the
	// horizontal neighbour, one that leaves over the left/right slides to the vertical neighbour.
	static const int dcol[8] = { -1, 0, 1, 1, 1, 0, -1, -1 };
	static const int drow[8] = { 1, 1, 1, 0, -1, -1, -1, 0 };
	for (unsigned int i = 0; i < out->size(); i++) {
		Patch* p = (*out)[i];
		const int col = (int)(i % kx), row = (int)(i / kx);
		for (int j = 0; j < 8; j++) {
			int c = col + dcol[j], r = row + drow[j];
			const bool c_ok = c >= 0 && c < (int)kx, r_ok = r >= 0 && r < (int)ky;
			if (dcol[j] != 0 && drow[j] != 0) {          // diagonal
				if (!r_ok) { r = row; }                 // over the top / bottom: same row
				else if (!c_ok) { c = col; }            // over the side: same column
			} else if (!c_ok || !r_ok) { c = col; r = row; }
			if (c < 0 || c >= (int)kx || r < 0 || r >= (int)ky) { c = col; r = row; }
			p->neighbours[j] = (*out)[(size_t)r * kx + c];
		}
	}
	return out;
}

std::vector<float> Patch::getVerticesCoords() {
	const Vector3f* v[4] = { &vec1, &vec2, &vec3, &vec4 };
	std::vector<float> out;
	out.reserve(12);
	for (int i = 0; i < 4; i++) { out.push_back(v[i]->x); out.push_back(v[i]->y); out.push_back(v[i]->z); }
	return out;
}

Vector3f Patch::getUp() { return Vector3f(vec4 - vec1); }

Vector3f Patch::getCenter() {
	return Vector3f((vec1.x + vec2.x + vec3.x + vec4.x) / 4.0f, (vec1.y + vec2.y + vec3.y + vec4.y) / 4.0f,
	                (vec1.z + vec2.z + vec3.z + vec4.z) / 4.0f);
}

Vector3f Patch::getNormal() {
	Vector3f a = vec2 - vec1;
	return a.Cross(vec4 - vec1);   // reversed cross product: (v4 - v1) x (v2 - v1)
}
