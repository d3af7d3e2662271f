#include "Transform.h"

const float f_pi = 3.1415926535897932384626433832795028841971691075f;

void CGLTransform::Frustum(Matrix4f& m, float l, float r, float b, float t, float n, float f) {
	const float w = r - l, h = t - b, d = f - n;
	// column 0..3, rows 0..3 — the standard glFrustum matrix
	m[0][0] = 2 * n / w;   m[0][1] = 0;           m[0][2] = 0;              m[0][3] = 0;
	m[1][0] = 0;           m[1][1] = 2 * n / h;   m[1][2] = 0;              m[1][3] = 0;
	m[2][0] = (r + l) / w; m[2][1] = (t + b) / h; m[2][2] = -(f + n) / d;   m[2][3] = -1;
	m[3][0] = 0;           m[3][1] = 0;           m[3][2] = -2 * f * n / d; m[3][3] = 0;
}

void CGLTransform::Perspective(Matrix4f& m, float fov, float aspect, float n, float f) {
	// half extent of the near rectangle; the float overload of tan is what the reference resolves to
	const float half_h = float(std::tan(fov * f_pi / 180 * .5f)) * n;
	const float half_w = half_h * aspect;
	Frustum(m, -half_w, half_w, -half_h, half_h, n, f);
}

void CGLTransform::LookAt(Matrix4f& m, Vector3f eye, Vector3f target, Vector3f up) {
	Vector3f dir(target - eye);
	dir.Normalize();
	Vector3f right(dir.v_Cross(up));
	right.Normalize();
	up = right.v_Cross(dir);
	for (int i = 0; i < 3; ++i) {
		m[i][0] = right[i]; m[i][1] = up[i]; m[i][2] = -dir[i];
		m[i][3] = 0; m[3][i] = 0;
	}
	m[3][3] = 1;
	m.Translate(-eye.x, -eye.y, -eye.z);
}
