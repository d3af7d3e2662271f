// Minimal Wavefront OBJ loader (reference: WaveFrontModel.h, WaveFrontModel.cpp:15-144).
// Parsing behaviour kept from the reference: "v x y z" in millimetres (divided by 1000), "f" with 3 or 4
// vertex references (anything after '/' ignored), triangles become degenerate quads (last vertex
// repeated), faces with more than 4 vertices are dropped, faces with bad indices are skipped silently.
//
// Extension: in the reference an OBJ scene is black and unlit because its patches get
// no colour and no emitter.  Two comment directives — ignored by the reference's parser, so geometry
// parity is untouched — set the material of the faces that follow:
//     #@color r g b          surface colour
//     #@emit  r g b          unshot radiosity B of an emitter (its illumination I becomes 1,1,1)
//     #@emit  0 0 0          back to non-emitting
#pragma once
#include <string>
#include "Model.h"

class WaveFrontModel : public Model {
public:
	explicit WaveFrontModel(std::string filename);
	bool parse(std::string filename);
	std::vector<Patch*>* getPatches(double area = 0);
};
