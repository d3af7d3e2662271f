#include "LoadingModel.h"

LoadingModel::LoadingModel(Patch* data, unsigned long count) {
	patches->reserve(count);
	for (unsigned long i = 0; i < count; i++) patches->push_back(new Patch(data[i]));
	for (unsigned long i = 0; i < count; i++) {
		Patch* p = (*patches)[i];
		for (int n = 0; n < 8; n++) {
			const unsigned int r = p->relativeNeighbours[n];
			p->neighbours[n] = r < count ? (*patches)[r] : p;
		}
	}
}

std::vector<Patch*>* LoadingModel::getPatches(double) { return patches; }
