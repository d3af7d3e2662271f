// Abstract scene object: owns its patches, can subdivide them (reference: Model.h:9-22, Model.cpp).
#pragma once
#include <vector>
#include "Patch.h"

class Model {
public:
	Model();
	virtual ~Model();
	virtual std::vector<Patch*>* getPatches(double area = 0) = 0;   // area > 0 triggers one subdivision pass

protected:
	void subdivide(double area);      // replaces every patch by its Patch::divide() children
	std::vector<Patch*>* patches;     // owned
};
