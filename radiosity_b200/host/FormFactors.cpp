#include "FormFactors.h"
#include "Config.h"
#include <vector>

#define Pi (3.1415926535897932384626433832795028841931f)   // FormFactors.cpp:3 (note: not Vector.cpp's f_pi literal)

// Top face: dF = dA / (pi (x^2 + y^2 + 1)^2); side faces: dF = dA * z / (pi (x^2 + z^2 + 1)^2) where the
// reference's numerator carries an extra half pixel (FormFactors.cpp:65) — kept, the table is API output.
void computeHemicubeFormFactors(unsigned int side, float* out) {
	const int n = (int)side;
	std::vector<float> top((size_t)n * n), flank((size_t)n * n / 2);
	const float halfPixel = (1.0f / n);
	float pixelArea = (2.0f / n);
	pixelArea *= pixelArea;
	for (int x = 0; x < n; x++) {
		for (int y = 0; y < n; y++) {
			const float dx = ((x - n / 2) / (n / 2.0f)) + halfPixel;
			const float dy = ((y - n / 2) / (n / 2.0f)) + halfPixel;
			float f = (dx * dx + dy * dy + 1);
			f *= f * Pi;
			top[x + (size_t)y * n] = pixelArea / f;
		}
	}
	for (int x = 0; x < n; x++) {
		for (int y = 0; y < n / 2; y++) {
			const float dx = (x - n / 2) / (n / 2.0f) + halfPixel;
			const float dy = (n / 2 - 1 - y) / (n / 2.0f) + halfPixel;
			float f = (dx * dx + dy * dy + 1);
			f *= f * Pi;
			flank[x + (size_t)y * n] = (pixelArea * (dy + halfPixel)) / f;
		}
	}
	// atlas: rows [0,N) = LEFT half | FRONT | RIGHT half, rows [N,1.5N) = UP half | DOWN half
	const unsigned int N = side, W = (unsigned int)(N * 2), H = (unsigned int)(N * 1.5);
	for (unsigned int y = 0; y < H; y++) {
		for (unsigned int x = 0; x < W; x++) {
			float v;
			if (y < N) {
				if (x < N / 2) v = flank[(N / 2 - x) * N - y - 1];                     // LEFT
				else if (x < N * 1.5) v = top[y * N + (x - N / 2)];                    // FRONT
				else v = flank[(x - (unsigned int)(N * 1.5)) * N + y];                 // RIGHT
			} else {
				if (x < N) v = flank[(y - N) * N + x];                                 // UP
				else v = flank[((N / 2 - 1) - (y - N)) * N + (x - N)];                 // DOWN
			}
			out[(size_t)y * W + x] = v;
		}
	}
}

float* precomputeHemicubeFormFactors() {
	const unsigned int res = Config::PATCHVIEW_TEX_RES(), k = Config::HEMICUBES_CNT();
	float* ff = new float[(size_t)res * k];
	computeHemicubeFormFactors(Config::HEMICUBE_W(), ff);
	for (unsigned int h = 1; h < k; h++)
		for (unsigned int i = 0; i < res; i++) ff[(size_t)res * h + i] = ff[i];
	return ff;
}
