// The built-in Cornell-style box (reference: PrimitiveModel.h, PrimitiveModel.cpp:9-213):
// ROOM = 5 walls + the light quad (B = 100, I = 1), ROOMCLOSURE = front wall, CUBE, BLOCK.
#pragma once
#include "Model.h"

class PrimitiveModel : public Model {
public:
	enum { ROOM, ROOMCLOSURE, CUBE, BLOCK };
	explicit PrimitiveModel(int type);
	~PrimitiveModel();
	std::vector<Patch*>* getPatches(double area = 0);

private:
	void addQuads(const float* coords, int nquads, const float* colors);
	int type;
};
