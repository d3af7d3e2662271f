// Minimal float32 vector / matrix types with the method names the reference's hot path uses
// (reference: Vector.h:334-549 Vector3<T>, Vector.h:886-1100 + Vector.cpp:120-160,325-330,445-527 Matrix4f).
// Only what the shooting path touches is provided; quaternions, planes, polygons and the other
// utility types of the reference's bundled math library are out of scope.
//
// Every expression is written in the reference's evaluation order: results are bit-identical to the
// reference when compiled without FMA contraction (-ffp-contract=off).
#pragma once
#include <cmath>

struct Vector3f {
	float x, y, z;

	Vector3f() {}
	Vector3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}

	float f_Length2() const { return x * x + y * y + z * z; }
	float f_Length() const { return (float)std::sqrt(x * x + y * y + z * z); }
	float f_Dot(const Vector3f& o) const { return x * o.x + y * o.y + z * o.z; }

	float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
	float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }

	void Normalize() {
		float t = f_Length();
		if (t != 0) { t = 1 / t; x *= t; y *= t; z *= t; }
	}

	bool operator==(const Vector3f& o) const { return x == o.x && y == o.y && z == o.z; }
	Vector3f operator+(const Vector3f& o) const { return Vector3f(x + o.x, y + o.y, z + o.z); }
	Vector3f operator-(const Vector3f& o) const { return Vector3f(x - o.x, y - o.y, z - o.z); }
	Vector3f operator*(const Vector3f& o) const { return Vector3f(x * o.x, y * o.y, z * o.z); }   // component-wise
	Vector3f operator-() const { return Vector3f(-x, -y, -z); }
	Vector3f operator*(float t) const { return Vector3f(x * t, y * t, z * t); }
	Vector3f operator/(float t) const { t = 1 / t; return Vector3f(x * t, y * t, z * t); }          // reciprocal, then scale
	Vector3f& operator+=(const Vector3f& o) { x += o.x; y += o.y; z += o.z; return *this; }
	Vector3f& operator-=(const Vector3f& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }

	// NOTE: as in the reference (Vector.h:534-549) the cross product is REVERSED: a.v_Cross(b) == b x a.
	// Normals, the LEFT/RIGHT hemicube faces and LookAt all rely on it.
	Vector3f v_Cross(const Vector3f& o) const { return Vector3f(o.y * z - o.z * y, o.z * x - o.x * z, o.x * y - o.y * x); }
	Vector3f Cross(const Vector3f& o) { *this = v_Cross(o); return *this; }   // in place
};

// 4x4, column-major storage f[column][row] (OpenGL compatible)
struct Matrix4f {
	float f[4][4];

	float* operator[](int col) { return f[col]; }
	const float* operator[](int col) const { return f[col]; }

	void Identity() {
		for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) f[c][r] = (float)(c == r);
	}
	void Translation(float tx, float ty, float tz) {
		Identity();
		f[3][0] = tx; f[3][1] = ty; f[3][2] = tz;
	}
	// this = a * b; each entry is summed over k = 0..3 in order (Vector.cpp:445-458)
	void ProductOf(const Matrix4f& a, const Matrix4f& b) {
		for (int c = 0; c < 4; ++c)
			for (int r = 0; r < 4; ++r)
				f[c][r] = a.f[0][r] * b.f[c][0] + a.f[1][r] * b.f[c][1] + a.f[2][r] * b.f[c][2] + a.f[3][r] * b.f[c][3];
	}
	Matrix4f operator*(const Matrix4f& o) const { Matrix4f m; m.ProductOf(*this, o); return m; }
	Matrix4f& operator*=(const Matrix4f& o) { Matrix4f m; m.ProductOf(*this, o); *this = m; return *this; }
	void Translate(float tx, float ty, float tz) { Matrix4f t; t.Translation(tx, ty, tz); *this *= t; }
};

extern const float f_pi;
