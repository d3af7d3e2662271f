// Pseudo-model that adopts already-subdivided patches, e.g. from a saved run
// (reference: LoadingModel.h, LoadingModel.cpp:4-25).  Never subdivides.
#pragma once
#include "Model.h"

class LoadingModel : public Model {
public:
	LoadingModel(Patch* data, unsigned long count);   // copies; relativeNeighbours[] -> neighbours[]
	std::vector<Patch*>* getPatches(double area);
};
