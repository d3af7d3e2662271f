#include "PrimitiveModel.h"

namespace {
// scene data of the reference (PrimitiveModel.cpp:92-213): 4 vertices x (x, y, z) per quad
const float kRoom[5 * 12] = {
	0.0f, 0.0f, 5.592f,    5.496f, 0.0f, 5.592f,   5.560f, 5.488f, 5.592f,  0.0f, 5.488f, 5.592f,    // back wall
	0.0f, 5.488f, 0.0f,    0.0f, 5.488f, 5.592f,   5.560f, 5.488f, 5.592f,  5.560f, 5.488f, 0.0f,    // ceiling
	0.0f, 0.0f, 0.0f,      5.528f, 0.0f, 0.0f,     5.496f, 0.0f, 5.592f,    0.0f, 0.0f, 5.592f,      // floor
	0.0f, 0.0f, 0.0f,      0.0f, 0.0f, 5.592f,     0.0f, 5.488f, 5.593f,    0.0f, 5.488f, 0.0f,      // x = 0 wall (red)
	5.496f, 0.0f, 5.592f,  5.528f, 0.0f, 0.0f,     5.560f, 5.488f, 0.0f,    5.560f, 5.488f, 5.592f,  // x = 5.5 wall (green)
};
const float kRoomColors[5 * 3] = { 1, 1, 1,  1, 1, 1,  1, 1, 1,  1, 0, 0,  0, 1, 0 };
const float kLight[12] = { 3.430f, 5.485f, 2.270f,  2.130f, 5.485f, 2.270f,  2.130f, 5.485f, 3.320f,  3.430f, 5.485f, 3.320f };
const float kClosure[12] = { 5.528f, 0.0f, 0.0f,  0.0f, 0.0f, 0.0f,  0.0f, 5.488f, 0.0f,  5.560f, 5.488f, 0.0f };
const float kCube[5 * 12] = {
	1.3f, 1.65f, 0.65f,   2.9f, 1.65f, 1.14f,   2.4f, 1.65f, 2.72f,   0.82f, 1.65f, 2.25f,
	2.9f, 0.0f, 1.14f,    2.4f, 0.0f, 2.72f,    2.4f, 1.65f, 2.72f,   2.9f, 1.65f, 1.14f,
	1.3f, 0.0f, 0.65f,    2.9f, 0.0f, 1.14f,    2.9f, 1.65f, 1.14f,   1.3f, 1.65f, 0.65f,
	0.82f, 0.0f, 2.25f,   1.3f, 0.0f, 0.65f,    1.3f, 1.65f, 0.65f,   0.82f, 1.65f, 2.25f,
	2.4f, 0.0f, 2.72f,    0.82f, 0.0f, 2.25f,   0.82f, 1.65f, 2.25f,  2.4f, 1.65f, 2.72f,
};
const float kBlock[5 * 12] = {
	4.23f, 3.3f, 2.47f,   4.72f, 3.3f, 4.06f,   3.14f, 3.3f, 4.56f,   2.65f, 3.3f, 2.96f,
	4.23f, 0.0f, 2.47f,   4.72f, 0.0f, 4.06f,   4.72f, 3.3f, 4.06f,   4.23f, 3.3f, 2.47f,
	4.72f, 0.0f, 4.06f,   3.14f, 0.0f, 4.56f,   3.14f, 3.3f, 4.56f,   4.72f, 3.3f, 4.06f,
	3.14f, 0.0f, 4.56f,   2.65f, 0.0f, 2.96f,   2.65f, 3.3f, 2.96f,   3.14f, 3.3f, 4.56f,
	2.65f, 0.0f, 2.96f,   4.23f, 0.0f, 2.47f,   4.23f, 3.3f, 2.47f,   2.65f, 3.3f, 2.96f,
};
inline Vector3f at(const float* p, int v) { return Vector3f(p[3 * v], p[3 * v + 1], p[3 * v + 2]); }
}

void PrimitiveModel::addQuads(const float* coords, int nquads, const float* colors) {
	for (int q = 0; q < nquads; q++) {
		const float* c = coords + 12 * q;
		const Vector3f col = colors ? Vector3f(colors[3 * q], colors[3 * q + 1], colors[3 * q + 2]) : Vector3f(1.0f, 1.0f, 1.0f);
		patches->push_back(new Patch(at(c, 0), at(c, 1), at(c, 2), at(c, 3), col));
	}
}

PrimitiveModel::PrimitiveModel(int type_) : type(type_) {
	switch (type) {
	case ROOM: {
		addQuads(kRoom, 5, kRoomColors);
		// the light: white, already "lit" (I = 1) and carrying B = 100 per channel (PrimitiveModel.cpp:19-29)
		const Vector3f white(1.0f, 1.0f, 1.0f), energy(1.0f, 1.0f, 1.0f);
		patches->push_back(new Patch(at(kLight, 0), at(kLight, 1), at(kLight, 2), at(kLight, 3), white, energy, energy * 100));
		break;
	}
	case ROOMCLOSURE: addQuads(kClosure, 1, NULL); break;
	case CUBE: addQuads(kCube, 5, NULL); break;
	case BLOCK: addQuads(kBlock, 5, NULL); break;
	}
}

PrimitiveModel::~PrimitiveModel() {}

std::vector<Patch*>* PrimitiveModel::getPatches(double area) {
	if (area > 0) subdivide(area);
	return patches;
}
