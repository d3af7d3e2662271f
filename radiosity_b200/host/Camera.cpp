#include "Camera.h"

Camera::Camera() : eye(2.78f, 2.73f, 0.025f), target(0.0f, 0.0f, 1.0f), up(0.0f, 1.0f, 0.0f) {}

void Camera::lookFromPatch(Patch* p, PatchLook dir) {
	eye = p->getCenter();
	const Vector3f n = p->getNormal(), u = p->getUp();
	// side directions come from the reversed cross product: n.v_Cross(u) == u x n
	switch (dir) {
	case PATCH_LOOK_FRONT: target = n; up = u; break;
	case PATCH_LOOK_UP: target = u; up = -n; break;
	case PATCH_LOOK_DOWN: target = -u; up = n; break;
	case PATCH_LOOK_LEFT: target = -n.v_Cross(u); up = u; break;
	case PATCH_LOOK_RIGHT: target = n.v_Cross(u); up = u; break;
	}
}

Matrix4f Camera::GetMatrix() {
	Matrix4f m;
	CGLTransform::LookAt(m, eye, target + eye, up);
	return m;
}
