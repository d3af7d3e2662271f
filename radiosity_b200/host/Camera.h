// Hemicube camera: looks out of a patch along one of the five hemicube directions
// (reference: Camera.h:7-35, Camera.cpp:19-52 lookFromPatch, :97-103 GetMatrix).  The free-fly viewer
// camera (Move / Aim / Reset) of the reference is out of scope.
#pragma once
#include "Vector.h"
#include "Transform.h"
#include "Patch.h"

class Camera {
public:
	enum PatchLook { PATCH_LOOK_FRONT = 0, PATCH_LOOK_UP, PATCH_LOOK_DOWN, PATCH_LOOK_LEFT, PATCH_LOOK_RIGHT };

	Camera();
	void lookFromPatch(Patch* p, PatchLook dir);
	Matrix4f GetMatrix();          // LookAt(eye, eye + target, up)

private:
	Vector3f eye, target, up;
};
