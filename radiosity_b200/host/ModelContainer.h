// The scene: a list of models flattened into the arrays the solver consumes
// (reference: ModelContainer.h:10-54, ModelContainer.cpp).  Patch id == index into these arrays.
#pragma once
#include <vector>
#include <string>
#include <stdint.h>
#include "PrimitiveModel.h"
#include "WaveFrontModel.h"
#include "StaticMeshModel.h"
#include "Vector.h"

class ModelContainer {
public:
	ModelContainer();
	~ModelContainer();

	void load();                            // the built-in Cornell box: room, closure, cube, block
	bool load(const std::string& objPath);  // extension: a Wavefront OBJ scene (see WaveFrontModel.h)
	bool loadStaticMesh(const std::string& headerPath, float scale = 0.01f, bool flip = false, int emissiveMaterial = -1);   // TestModel.h-style export (StaticMeshModel.h)

	int addModel(Model* m);                 // takes ownership; returns its index
	void removeModel(int i);
	unsigned int getModelsCount() const { return (unsigned int)models.size(); }
	void updateData();                      // (re)builds vertices / indices / patches

	float* getVertices();                   // float[P*12], 4 unshared vertices per patch
	unsigned int getVerticesCount();        // P*12
	int* getIndices();                      // int[P*6]: (4p, 4p+1, 4p+2, 4p, 4p+2, 4p+3)
	unsigned int getIndicesCount();         // P*6
	Patch** getPatches();
	unsigned int getPatchesCount();

	unsigned int getHighestRadiosityPatchId();   // plain argmax of |B|^2 (first maximum)
	// the `count` shooters of a batch, reference list semantics (see ModelContainer.cpp)
	void getHighestRadiosityPatchesId(unsigned int count, Patch** p_emitters, unsigned int* p_emitters_ids);

	double maxPatchArea;                    // > 0: subdivide every model once to this area

protected:
	bool needRefresh;
	std::vector<Model*> models;
	Patch** patches; unsigned int patchesCount;
	float* vertices; unsigned int verticesCount;
	int* indices; unsigned int indicesCount;
};
