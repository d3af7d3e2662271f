#include "WaveFrontModel.h"
#include <fstream>
#include <sstream>
#include <iostream>

WaveFrontModel::WaveFrontModel(std::string filename) {
	if (!parse(filename)) std::cerr << "Unable to load model '" << filename << "'" << std::endl;
}

bool WaveFrontModel::parse(std::string filename) {
	std::ifstream in(filename.c_str());
	if (!in.is_open()) return false;
	std::vector<Vector3f> verts;
	Vector3f color(0.0f, 0.0f, 0.0f), emit(0.0f, 0.0f, 0.0f);
	bool emitting = false;
	std::string line;
	unsigned int lineno = 0;
	while (std::getline(in, line)) {
		lineno++;
		if (line.compare(0, 2, "v ") == 0) {
			std::istringstream ss(line.substr(2));
			std::vector<float> c; float x;
			while (ss >> x) c.push_back(x);
			if (c.size() < 3) { std::cerr << "Bad vertex definition in '" << filename << "', line " << lineno << std::endl; return false; }
			if (c.size() > 3) std::cerr << "Warning: Ignoring [w] coordinate in '" << filename << "', line " << lineno << std::endl;
			verts.push_back(Vector3f(c[0] / 1000, c[1] / 1000, c[2] / 1000));      // millimetres -> metres
		} else if (line.compare(0, 2, "f ") == 0) {
			std::string rest = line.substr(2);
			std::vector<unsigned int> idx;
			while (!rest.empty() && idx.size() <= 4) {                               // reads at most 5 to detect n-gons
				const size_t sp = rest.find_first_of(' ');
				std::istringstream tok(rest.substr(0, sp));
				int v;
				if (tok >> v) idx.push_back((unsigned int)(v - 1));                  // "12/3/4" -> 12; OBJ is 1-based
				if (sp == std::string::npos) rest.clear(); else rest.erase(0, sp + 1);
			}
			if (idx.size() == 3) idx.push_back(idx.back());                          // triangle -> degenerate quad
			if (idx.size() != 4) continue;                                           // n-gons are dropped
			bool ok = true;
			for (int i = 0; i < 4; i++) ok = ok && idx[i] < verts.size();
			if (!ok) continue;
			Patch* p = new Patch(verts[idx[0]], verts[idx[1]], verts[idx[2]], verts[idx[3]]);
			p->setColor(color);
			if (emitting) { p->radiosity = emit; p->illumination = Vector3f(1.0f, 1.0f, 1.0f); }
			patches->push_back(p);
		} else if (line.compare(0, 8, "#@color ") == 0) {
			std::istringstream ss(line.substr(8));
			ss >> color.x >> color.y >> color.z;
		} else if (line.compare(0, 7, "#@emit ") == 0) {
			std::istringstream ss(line.substr(7));
			ss >> emit.x >> emit.y >> emit.z;
			emitting = emit.f_Length2() > 0;
		}
	}
	return true;
}

std::vector<Patch*>* WaveFrontModel::getPatches(double area) {
	if (area > 0) subdivide(area);
	return patches;
}
