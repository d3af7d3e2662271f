#include "Colors.h"

static inline Vector3f lit(Patch* p) { return p->getColor() * (p->illumination + p->radiosity); }

void Colors::smoothShadePatch(float* colors, Patch* p) {
	// corner -> the three neighbours sharing it, in the reference's summation order; output order lb, rb, rt, lt
	static const int corner[4][3] = { { 5, 6, 7 }, { 3, 4, 5 }, { 1, 2, 3 }, { 7, 0, 1 } };
	for (int c = 0; c < 4; c++) {
		Vector3f acc = lit(p);
		for (int j = 0; j < 3; j++) acc += lit(p->neighbours[corner[c][j]]);
		acc = acc / 4;
		colors[3 * c] = acc.x; colors[3 * c + 1] = acc.y; colors[3 * c + 2] = acc.z;
	}
}
