#include "ModelContainer.h"
#include <list>

ModelContainer::ModelContainer()
	: maxPatchArea(0), needRefresh(false), patches(NULL), patchesCount(0), vertices(NULL), verticesCount(0), indices(NULL), indicesCount(0) {}

ModelContainer::~ModelContainer() {
	for (size_t i = 0; i < models.size(); i++) delete models[i];
	delete[] vertices; delete[] indices; delete[] patches;
}

void ModelContainer::load() {
	// id order of the reference (ModelContainer.cpp:44-47): room (with the light), closure, cube, block
	addModel(new PrimitiveModel(PrimitiveModel::ROOM));
	addModel(new PrimitiveModel(PrimitiveModel::ROOMCLOSURE));
	addModel(new PrimitiveModel(PrimitiveModel::CUBE));
	addModel(new PrimitiveModel(PrimitiveModel::BLOCK));
}

bool ModelContainer::load(const std::string& objPath) {
	WaveFrontModel* m = new WaveFrontModel(objPath);
	addModel(m);
	return m->getPatches(0)->size() > 0;
}

bool ModelContainer::loadStaticMesh(const std::string& headerPath, float scale, bool flip, int emissiveMaterial) {
	StaticMeshModel* m = new StaticMeshModel(headerPath, scale, flip, emissiveMaterial);
	addModel(m);
	return m->triangleCount() > 0;
}

int ModelContainer::addModel(Model* m) { needRefresh = true; models.push_back(m); return (int)models.size() - 1; }

void ModelContainer::removeModel(int i) {
	needRefresh = true;
	delete models[i];
	models.erase(models.begin() + i);
}

void ModelContainer::updateData() {
	std::vector<Patch*> all;
	for (size_t m = 0; m < models.size(); m++) {
		std::vector<Patch*>* mp = models[m]->getPatches(maxPatchArea);
		all.insert(all.end(), mp->begin(), mp->end());
	}
	patchesCount = (unsigned int)all.size();
	verticesCount = patchesCount * 4 * 3;
	indicesCount = patchesCount * 6;
	delete[] vertices; delete[] indices; delete[] patches;
	vertices = new float[verticesCount];
	indices = new int[indicesCount];
	patches = new Patch*[patchesCount];
	for (unsigned int p = 0; p < patchesCount; p++) {
		Patch* patch = all[p];
		const std::vector<float> c = patch->getVerticesCoords();
		for (int i = 0; i < 12; i++) vertices[12 * (size_t)p + i] = c[i];
		const int base = 4 * (int)p;                        // two triangles per quad: (0,1,2) and (0,2,3)
		int* ix = indices + 6 * (size_t)p;
		ix[0] = base; ix[1] = base + 1; ix[2] = base + 2; ix[3] = base; ix[4] = base + 2; ix[5] = base + 3;
		for (int j = 0; j < 8; j++) if (patch->neighbours[j] == NULL) patch->neighbours[j] = patch;
		patches[p] = patch;
	}
	needRefresh = false;
}

float* ModelContainer::getVertices() { if (needRefresh) updateData(); return vertices; }
unsigned int ModelContainer::getVerticesCount() { if (needRefresh) updateData(); return verticesCount; }
int* ModelContainer::getIndices() { if (needRefresh) updateData(); return indices; }
unsigned int ModelContainer::getIndicesCount() { if (needRefresh) updateData(); return indicesCount; }
Patch** ModelContainer::getPatches() { if (needRefresh) updateData(); return patches; }
unsigned int ModelContainer::getPatchesCount() { if (needRefresh) updateData(); return patchesCount; }

unsigned int ModelContainer::getHighestRadiosityPatchId() {
	if (needRefresh) updateData();
	unsigned int best = 0;
	for (unsigned int i = 1; i < patchesCount; i++)
		if (patches[i]->radiosity.f_Length2() > patches[best]->radiosity.f_Length2()) best = i;
	return best;
}

// Shooter selection with the reference's exact list behaviour (ModelContainer.cpp:259-299): walk the
// patches in id order keeping a list sorted by |B|^2 descending.  A patch enters when the list is empty
// (so patch 0 is always seeded) or when it has energy and at least as much as the current tail — hence the
// list can stay short while it is not full.  Each insertion re-sorts with a stable ascending sort followed
// by a reversal, which flips the order inside every group of equal energies; the list is then cut to `count`.
// The device implementation (select_update.cu) reproduces exactly this.
void ModelContainer::getHighestRadiosityPatchesId(unsigned int count, Patch** p_emitters, unsigned int* p_emitters_ids) {
	if (needRefresh) updateData();
	Patch** pp = patches;
	auto less_energy = [pp](unsigned int a, unsigned int b) { return pp[a]->radiosity.f_Length2() < pp[b]->radiosity.f_Length2(); };
	std::list<unsigned int> tops;
	for (unsigned int pi = 0; pi < patchesCount; pi++) {
		const float e = patches[pi]->radiosity.f_Length2();
		if (!tops.empty() && !(e > 0 && patches[tops.back()]->radiosity.f_Length2() <= e)) continue;
		tops.push_back(pi);
		tops.sort(less_energy);
		tops.reverse();
		while (tops.size() > count) tops.pop_back();
	}
	std::list<unsigned int>::const_iterator it = tops.begin();
	for (unsigned int i = 0; i < count; i++) {
		if (it == tops.end()) { p_emitters_ids[i] = 0; p_emitters[i] = NULL; continue; }
		p_emitters_ids[i] = *it;
		p_emitters[i] = patches[*it];
		++it;
	}
}
