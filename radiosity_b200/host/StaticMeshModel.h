// Adapter for "static 3DS export" headers such as the reference's TestModel.h.
//
// TestModel.h is dead code in the reference (nothing includes it; TestModel.h:64-75 declares, :224-464 defines four
// triangle-mesh boxes as interleaved GL_T2F_N3F_V3F vertex arrays + index arrays + material ranges, TestModel.h:20-35),
// so there is no reference behaviour to be in parity with.  The adapter reads such a header AS TEXT at run time
// (nothing of the reference is compiled or copied in) and turns it into patches the rest of the path understands:
//   - every triangle becomes a degenerate quad (last vertex repeated) — the reference's own convention for triangles,
//     WaveFrontModel.cpp:98-99;
//   - positions are the last three floats of every 8-float vertex, multiplied by `scale` (the export is in scene
//     units of about a centimetre; the built-in scene is in metres);
//   - `flip` reverses the winding: the export is counter-clockwise seen from outside (OpenGL), a patch shoots along
//     (v4 - v1) x (v2 - v1) (Patch.cpp:272-276);
//   - material ranges (TMatRange {material, first index, index count}) pick a colour from a fixed palette; the faces of
//     `emissiveMaterial` (-1: none) get B = (100,100,100), I = (1,1,1) like the built-in light (PrimitiveModel.cpp:26-28).
#pragma once
#include <string>
#include "Model.h"

class StaticMeshModel : public Model {
public:
	StaticMeshModel(const std::string& headerPath, float scale = 0.01f, bool flip = false, int emissiveMaterial = -1);
	bool parse(const std::string& headerPath, float scale, bool flip, int emissiveMaterial);
	std::vector<Patch*>* getPatches(double area = 0);
	unsigned int objectCount() const { return objects; }
	unsigned int triangleCount() const { return triangles; }
private:
	unsigned int objects = 0, triangles = 0;
};
