// Quad patch: geometry, surface colour and the two energies of progressive radiosity
// (reference: Patch.h:16-72, Patch.cpp).  `radiosity` is the UNSHOT energy B, `illumination` the
// energy already shot I; both are public and mutated by the solver, as in the reference.
#pragma once
#include <vector>
#include <cmath>
#include "Vector.h"

#define REFLECTIVITY 0.3f   // surface reflectivity, constant for every patch (Patch.h:13)

class Patch {
public:
	Patch();
	Patch(Vector3f vec1, Vector3f vec2, Vector3f vec3, Vector3f vec4);
	Patch(Vector3f vec1, Vector3f vec2, Vector3f vec3, Vector3f vec4, Vector3f color);
	Patch(Vector3f vec1, Vector3f vec2, Vector3f vec3, Vector3f vec4, Vector3f color, Vector3f illumination);
	Patch(Vector3f vec1, Vector3f vec2, Vector3f vec3, Vector3f vec4, Vector3f color, Vector3f illumination, Vector3f radiosity);
	~Patch();

	// uniform-grid split into patches of at most `area` (one pass); NULL when the patch is small enough.
	// Children inherit colour, I and B unchanged.  The caller owns the returned vector and its patches.
	std::vector<Patch*>* divide(double area);
	std::vector<float> getVerticesCoords();   // 12 floats: v1 v2 v3 v4

	Vector3f getCenter();          // mean of the four vertices (hemicube eye)
	Vector3f getNormal();          // (v4 - v1) x (v2 - v1), not normalised
	Vector3f getUp();              // v4 - v1
	Vector3f getColor() { return color; }
	float getReflectivity() { return REFLECTIVITY; }
	Patch** getNeighbours() { return neighbours; }

	// extension: lets loaders that carry materials colour a patch after construction
	void setColor(const Vector3f& c) { color = c; }

	Vector3f radiosity;            // B
	Vector3f illumination;         // I
	unsigned int relativeNeighbours[8];   // neighbour ids, only meaningful in saved files
	Patch* neighbours[8];          // 8-neighbourhood, starting top-left, clockwise; self where there is none

protected:
	Vector3f vec1, vec2, vec3, vec4;
	Vector3f color;
};
