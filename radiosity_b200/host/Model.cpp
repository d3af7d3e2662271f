#include "Model.h"

Model::Model() : patches(new std::vector<Patch*>()) {}

Model::~Model() {
	for (size_t i = 0; i < patches->size(); i++) delete (*patches)[i];
	delete patches;
}

// Reference: Model.cpp:27-60 — one pass; an undivided patch is copied and becomes its own neighbour.
void Model::subdivide(double area) {
	std::vector<Patch*>* next = new std::vector<Patch*>();
	for (size_t i = 0; i < patches->size(); i++) {
		Patch* p = (*patches)[i];
		std::vector<Patch*>* children = p->divide(area);
		if (children == NULL) {
			Patch* copy = new Patch(*p);
			for (int j = 0; j < 8; j++) if (copy->neighbours[j] == NULL) copy->neighbours[j] = copy;
			next->push_back(copy);
		} else {
			next->insert(next->end(), children->begin(), children->end());
			delete children;
		}
		delete p;
	}
	delete patches;
	patches = next;
}
