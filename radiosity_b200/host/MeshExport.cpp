#include "MeshExport.h"
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {
unsigned char to_byte(float v) {          // GL colour clamp, then 8-bit UNORM
	if (!(v > 0.0f)) return 0;            // also catches NaN
	if (v >= 1.0f) return 255;
	return (unsigned char)std::lrintf(v * 255.0f);
}
}

bool ExportPly(const std::string& path, ModelContainer& scene, const float* colors12, float exposure) {
	const unsigned int P = scene.getPatchesCount();
	FILE* f = std::fopen(path.c_str(), "wb");
	if (!f) return false;
	std::fprintf(f, "ply\nformat binary_little_endian 1.0\ncomment radiosity-b200: Colors::smoothShadePatch vertex colours\n"
	                "element vertex %u\nproperty float x\nproperty float y\nproperty float z\n"
	                "property uchar red\nproperty uchar green\nproperty uchar blue\n"
	                "element face %u\nproperty list uchar int vertex_indices\nend_header\n", 4u * P, P);
	const float* v = scene.getVertices();                 // 12 floats per patch, 4 unshared vertices (ModelContainer.cpp:100-107)
	std::vector<unsigned char> buf;
	buf.resize((size_t)P * 4 * 15);
	unsigned char* o = buf.data();
	for (size_t i = 0; i < (size_t)P * 4; i++) {
		std::memcpy(o, v + 3 * i, 12); o += 12;
		for (int c = 0; c < 3; c++) *o++ = to_byte(colors12[3 * i + c] * exposure);
	}
	bool ok = std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
	buf.resize((size_t)P * 17);
	o = buf.data();
	for (uint32_t p = 0; p < P; p++) {
		*o++ = 4;
		for (int32_t k = 0; k < 4; k++) { const int32_t idx = (int32_t)(4 * p + k); std::memcpy(o, &idx, 4); o += 4; }
	}
	ok = ok && std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
	return std::fclose(f) == 0 && ok;
}
