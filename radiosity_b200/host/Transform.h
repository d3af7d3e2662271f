// Projection / view matrices used by the hemicube cameras (reference: Transform.h:302,337,457,
// Transform.cpp:26-46 Frustum, :70-80 Perspective, :127-156 LookAt).  The GL matrix-stack emulation
// CGL3TransformState of the reference is viewer-only and not provided.
#pragma once
#include "Vector.h"

class CGLTransform {
public:
	static void Frustum(Matrix4f& m, float left, float right, float bottom, float top, float z_near, float z_far);
	static void Perspective(Matrix4f& m, float fov_degrees, float aspect, float z_near, float z_far);
	// mirrored with respect to gluLookAt because of the reversed cross product (see Vector.h)
	static void LookAt(Matrix4f& m, Vector3f eye, Vector3f target, Vector3f up);
};
