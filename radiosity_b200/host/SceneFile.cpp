#include "SceneFile.h"
#include "LoadingModel.h"
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>
#include <stdint.h>

namespace {

struct Rec { float v[12], color[3], rad[3], illum[3]; uint32_t nb[8]; };

// reference record layouts (offsets of Patch fields, Patch.h:46-54): radiosity, illumination, relativeNeighbours[8],
// neighbours[8] (pointers, written as NULL), vec1..vec4, color
struct Layout { size_t count_bytes, rec_bytes, off_rad, off_illum, off_rel, off_vec, off_color; };
const Layout kWin32 = { 4, 148, 0, 12, 24, 24 + 32 + 8 * 4, 24 + 32 + 8 * 4 + 48 };
const Layout kLP64 = { 8, 184, 0, 12, 24, 24 + 32 + 8 * 8, 24 + 32 + 8 * 8 + 48 };
const Layout kWin64 = { 4, 184, 0, 12, 24, 24 + 32 + 8 * 8, 24 + 32 + 8 * 8 + 48 };   // LLP64: 8-byte pointers, 4-byte unsigned long count

// does a file of `size` bytes hold exactly `count` records of layout L after its count field?  (count comes from the file:
// derived from the size instead of multiplied, so that a crafted count cannot wrap the product)
bool holds(uint64_t size, uint64_t hdr, uint64_t rec_bytes, uint64_t count) {
	return size >= hdr && (size - hdr) % rec_bytes == 0 && (size - hdr) / rec_bytes == count;
}

void gather(ModelContainer& scene, std::vector<Rec>& out) {
	const unsigned P = scene.getPatchesCount();
	Patch** pp = scene.getPatches();
	const float* verts = scene.getVertices();
	std::map<Patch*, uint32_t> index;
	for (unsigned i = 0; i < P; i++) index[pp[i]] = i;
	out.resize(P);
	for (unsigned i = 0; i < P; i++) {
		Rec& r = out[i];
		memcpy(r.v, verts + 12 * (size_t)i, 48);
		const Vector3f c = pp[i]->getColor();
		r.color[0] = c.x; r.color[1] = c.y; r.color[2] = c.z;
		r.rad[0] = pp[i]->radiosity.x; r.rad[1] = pp[i]->radiosity.y; r.rad[2] = pp[i]->radiosity.z;
		r.illum[0] = pp[i]->illumination.x; r.illum[1] = pp[i]->illumination.y; r.illum[2] = pp[i]->illumination.z;
		for (int j = 0; j < 8; j++) {
			std::map<Patch*, uint32_t>::const_iterator it = index.find(pp[i]->neighbours[j]);
			r.nb[j] = it == index.end() ? i : it->second;       // Main.cpp:1607-1620
		}
	}
}

bool adopt(const std::vector<Rec>& recs, ModelContainer& scene) {
	const size_t n = recs.size();
	std::vector<Patch> data;
	data.reserve(n);
	for (size_t i = 0; i < n; i++) {
		const Rec& r = recs[i];
		Patch p(Vector3f(r.v[0], r.v[1], r.v[2]), Vector3f(r.v[3], r.v[4], r.v[5]), Vector3f(r.v[6], r.v[7], r.v[8]), Vector3f(r.v[9], r.v[10], r.v[11]),
		        Vector3f(r.color[0], r.color[1], r.color[2]), Vector3f(r.illum[0], r.illum[1], r.illum[2]), Vector3f(r.rad[0], r.rad[1], r.rad[2]));
		for (int j = 0; j < 8; j++) { p.relativeNeighbours[j] = r.nb[j] < n ? r.nb[j] : (unsigned int)i; p.neighbours[j] = NULL; }
		data.push_back(p);
	}
	// scene = ModelContainer(); scene.addModel(new LoadingModel(data, count))   (Main.cpp:1513-1519)
	const double area = scene.maxPatchArea;
	while (scene.getModelsCount() > 0) scene.removeModel(0);
	scene.maxPatchArea = area;
	scene.addModel(new LoadingModel(n ? &data[0] : NULL, (unsigned long)n));
	return true;
}

} // namespace

bool SaveToFile(const std::string& path, ModelContainer& scene, RRFormat format) {
	std::vector<Rec> recs;
	gather(scene, recs);
	FILE* fp = fopen(path.c_str(), "wb");
	if (!fp) return false;
	bool ok = true;
	if (format == RR_PORTABLE) {
		const uint32_t version = 1; const uint64_t count = recs.size();
		ok = fwrite("RRB2", 1, 4, fp) == 4 && fwrite(&version, 4, 1, fp) == 1 && fwrite(&count, 8, 1, fp) == 1;
		if (ok && count) ok = fwrite(&recs[0], sizeof(Rec), recs.size(), fp) == recs.size();
	} else {
		const Layout& L = format == RR_REFERENCE_WIN32 ? kWin32 : kLP64;
		const uint64_t count = recs.size();
		ok = fwrite(&count, L.count_bytes, 1, fp) == 1;          // little endian: the low bytes are the 32-bit count
		std::vector<unsigned char> buf(L.rec_bytes);
		for (size_t i = 0; ok && i < recs.size(); i++) {
			const Rec& r = recs[i];
			memset(&buf[0], 0, L.rec_bytes);                      // neighbours[] = NULL, padding = 0
			memcpy(&buf[L.off_rad], r.rad, 12); memcpy(&buf[L.off_illum], r.illum, 12);
			memcpy(&buf[L.off_rel], r.nb, 32); memcpy(&buf[L.off_vec], r.v, 48); memcpy(&buf[L.off_color], r.color, 12);
			ok = fwrite(&buf[0], 1, L.rec_bytes, fp) == L.rec_bytes;
		}
	}
	return fclose(fp) == 0 && ok;
}

bool LoadFromFile(const std::string& path, ModelContainer& scene) {
	FILE* fp = fopen(path.c_str(), "rb");
	if (!fp) return false;
	fseek(fp, 0, SEEK_END);
	const long size = ftell(fp);
	fseek(fp, 0, SEEK_SET);
	std::vector<unsigned char> data(size > 0 ? (size_t)size : 0);
	const bool rd = size > 0 && fread(&data[0], 1, (size_t)size, fp) == (size_t)size;
	fclose(fp);
	if (!rd) return false;
	std::vector<Rec> recs;
	if (size >= 16 && memcmp(&data[0], "RRB2", 4) == 0) {
		uint32_t version; uint64_t count;
		memcpy(&version, &data[4], 4); memcpy(&count, &data[8], 8);
		if (version != 1 || !holds((uint64_t)size, 16, sizeof(Rec), count)) return false;
		recs.resize((size_t)count);
		if (count) memcpy(&recs[0], &data[16], (size_t)count * sizeof(Rec));
	} else {
		const Layout* L = NULL; uint64_t count = 0;
		uint32_t c32 = 0; uint64_t c64 = 0;
		if (size >= 4) memcpy(&c32, &data[0], 4);
		if (size >= 8) memcpy(&c64, &data[0], 8);
		if (size >= 8 && holds((uint64_t)size, 8, kLP64.rec_bytes, c64)) { L = &kLP64; count = c64; }
		else if (size >= 4 && holds((uint64_t)size, 4, kWin32.rec_bytes, c32)) { L = &kWin32; count = c32; }
		else if (size >= 4 && holds((uint64_t)size, 4, kWin64.rec_bytes, c32)) { L = &kWin64; count = c32; }
		if (!L) return false;
		recs.resize((size_t)count);
		for (size_t i = 0; i < recs.size(); i++) {
			const unsigned char* b = &data[L->count_bytes + i * L->rec_bytes];
			Rec& r = recs[i];
			memcpy(r.rad, b + L->off_rad, 12); memcpy(r.illum, b + L->off_illum, 12);
			memcpy(r.nb, b + L->off_rel, 32); memcpy(r.v, b + L->off_vec, 48); memcpy(r.color, b + L->off_color, 12);
		}
	}
	return adopt(recs, scene);
}
