#include "StaticMeshModel.h"
#include <fstream>
#include <sstream>
#include <iostream>
#include <map>
#include <cstdlib>
#include <cctype>

namespace {

// numbers between the braces that follow `<name>[...] = {` (nested braces flattened); pos is advanced past the block
bool read_array(const std::string& text, size_t at, std::vector<double>& out, size_t* end) {
	size_t open = text.find('{', at);
	if (open == std::string::npos) return false;
	int depth = 0; size_t i = open;
	std::string tok;
	for (; i < text.size(); i++) {
		const char ch = text[i];
		if (ch == '{') depth++;
		else if (ch == '}') { depth--; }
		const bool num = std::isdigit((unsigned char)ch) || ch == '-' || ch == '+' || ch == '.' || ch == 'e' || ch == 'E';
		if (num) tok += ch;
		else {
			if (!tok.empty() && (std::isdigit((unsigned char)tok[0]) || ((tok[0] == '-' || tok[0] == '+' || tok[0] == '.') && tok.size() > 1))) out.push_back(std::strtod(tok.c_str(), NULL));
			tok.clear();                                    // a trailing 'f' of a float literal ends the token here
		}
		if (depth == 0 && ch == '}') break;
	}
	if (end) *end = i;
	return depth == 0;
}

const float kPalette[8][3] = { {0.75f, 0.75f, 0.75f}, {0.75f, 0.25f, 0.25f}, {0.25f, 0.75f, 0.25f}, {0.25f, 0.25f, 0.75f},
                               {0.75f, 0.75f, 0.25f}, {0.75f, 0.25f, 0.75f}, {0.25f, 0.75f, 0.75f}, {0.5f, 0.5f, 0.5f} };

} // namespace

StaticMeshModel::StaticMeshModel(const std::string& headerPath, float scale, bool flip, int emissiveMaterial) {
	if (!parse(headerPath, scale, flip, emissiveMaterial)) std::cerr << "Unable to load static mesh '" << headerPath << "'" << std::endl;
}

bool StaticMeshModel::parse(const std::string& headerPath, float scale, bool flip, int emissiveMaterial) {
	std::ifstream in(headerPath.c_str());
	if (!in.is_open()) return false;
	std::stringstream buf; buf << in.rdbuf();
	const std::string text = buf.str();
	// definitions look like `... p_object_<n>_vertices[8 * 26] = { ... };` (declarations have no '=')
	std::map<int, std::vector<double> > verts, idx, mats;
	size_t pos = 0;
	while ((pos = text.find("p_object_", pos)) != std::string::npos) {
		size_t p = pos + 9;
		int n = 0; bool any = false;
		while (p < text.size() && std::isdigit((unsigned char)text[p])) { n = n * 10 + (text[p] - '0'); p++; any = true; }
		pos = p;
		if (!any || p >= text.size() || text[p] != '_') continue;
		const size_t name_end = text.find_first_of("[ \t\n;,)", p);
		if (name_end == std::string::npos) break;
		const std::string kind = text.substr(p + 1, name_end - p - 1);
		const size_t close = text.find(']', name_end), semi = text.find(';', name_end), eq = text.find('=', name_end);
		if (text[name_end] != '[' || close == std::string::npos || eq == std::string::npos || eq > semi) continue;   // declaration or a use
		std::vector<double> vals; size_t end = 0;
		if (!read_array(text, eq, vals, &end)) return false;
		if (kind == "vertices") verts[n] = vals; else if (kind == "indices") idx[n] = vals; else if (kind == "materials") mats[n] = vals;
		pos = end;
	}
	for (std::map<int, std::vector<double> >::const_iterator it = idx.begin(); it != idx.end(); ++it) {
		const int n = it->first;
		const std::vector<double>& ix = it->second;
		if (!verts.count(n)) continue;
		const std::vector<double>& vv = verts[n];
		const size_t nv = vv.size() / 8;
		const std::vector<double>* mr = mats.count(n) ? &mats[n] : NULL;
		objects++;
		for (size_t t = 0; t + 2 < ix.size(); t += 3) {
			size_t a = (size_t)ix[t], b = (size_t)ix[t + 1], c = (size_t)ix[t + 2];
			if (a >= nv || b >= nv || c >= nv) continue;                      // bad index: skipped silently, like the OBJ loader
			if (flip) std::swap(b, c);
			Vector3f P[3]; const size_t id[3] = { a, b, c };
			for (int k = 0; k < 3; k++) P[k] = Vector3f((float)vv[id[k] * 8 + 5] * scale, (float)vv[id[k] * 8 + 6] * scale, (float)vv[id[k] * 8 + 7] * scale);
			int material = 0;
			if (mr) for (size_t r = 0; r + 2 < mr->size(); r += 3)
				if ((double)t >= (*mr)[r + 1] && (double)t < (*mr)[r + 1] + (*mr)[r + 2]) material = (int)(*mr)[r];
			Patch* p = new Patch(P[0], P[1], P[2], P[2]);                     // triangle -> degenerate quad
			const float* col = kPalette[((material % 8) + 8) % 8];
			p->setColor(Vector3f(col[0], col[1], col[2]));
			if (material == emissiveMaterial) { p->radiosity = Vector3f(100.0f, 100.0f, 100.0f); p->illumination = Vector3f(1.0f, 1.0f, 1.0f); }
			patches->push_back(p);
			triangles++;
		}
	}
	return triangles > 0;
}

std::vector<Patch*>* StaticMeshModel::getPatches(double area) {
	if (area > 0) subdivide(area);
	return patches;
}
