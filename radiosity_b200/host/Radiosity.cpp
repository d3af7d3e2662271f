#include "Radiosity.h"
#include "Config.h"
#include "FormFactors.h"
#include "Transform.h"
#include <vector>
#include <cstring>
#include <map>
#include <stdint.h>

RadiositySolver::RadiositySolver() : passCounter(0), computeRadiosity(true), scene(NULL), ctx(NULL) {}
RadiositySolver::~RadiositySolver() { if (ctx) rad_destroy(ctx); }

bool RadiositySolver::fail(const char* where) {
	err = std::string(where) + ": " + rad_last_error(ctx);
	return false;
}

static void gatherState(ModelContainer& s, std::vector<float>& color, std::vector<float>& rad, std::vector<float>& illum) {
	const unsigned int P = s.getPatchesCount();
	Patch** pp = s.getPatches();
	color.resize(3 * (size_t)P); rad.resize(3 * (size_t)P); illum.resize(3 * (size_t)P);
	for (unsigned int i = 0; i < P; i++) {
		const Vector3f c = pp[i]->getColor();
		color[3 * i] = c.x; color[3 * i + 1] = c.y; color[3 * i + 2] = c.z;
		rad[3 * i] = pp[i]->radiosity.x; rad[3 * i + 1] = pp[i]->radiosity.y; rad[3 * i + 2] = pp[i]->radiosity.z;
		illum[3 * i] = pp[i]->illumination.x; illum[3 * i + 1] = pp[i]->illumination.y; illum[3 * i + 2] = pp[i]->illumination.z;
	}
}

bool RadiositySolver::init(ModelContainer& s, int device, unsigned int selectMode, unsigned int flags) {
	if (!Config::isFrozen()) { err = "RadiositySolver::init: Config::freeze() first"; return false; }
	scene = &s;
	const unsigned int P = s.getPatchesCount();
	if (P == 0) { err = "RadiositySolver::init: empty scene"; return false; }
	rad_config cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.hemicube_side = Config::HEMICUBE_W();
	cfg.hemicubes = Config::HEMICUBES_CNT();
	cfg.max_patches = P;
	cfg.device = device;
	cfg.select_mode = selectMode;
	cfg.reflectivity = REFLECTIVITY;
	cfg.flags = flags;
	Matrix4f proj;
	CGLTransform::Perspective(proj, 90, 1.0f, 0.01f, 1000);   // Main.cpp:1172 (aspect 1: hemicube faces are square)
	memcpy(cfg.projection, &proj[0][0], sizeof(cfg.projection));
	if (ctx) { rad_destroy(ctx); ctx = NULL; }
	if (rad_create(&ctx, &cfg) != RAD_OK) { err = std::string("rad_create: ") + rad_last_error(NULL); ctx = NULL; return false; }

	float* ff = precomputeHemicubeFormFactors();              // Main.cpp:298; the device keeps one hemicube's worth
	const int r = rad_set_formfactors(ctx, ff, Config::PATCHVIEW_TEX_RES());
	delete[] ff;
	if (r != RAD_OK) return fail("rad_set_formfactors");

	std::vector<float> color, rad, illum;
	gatherState(s, color, rad, illum);
	if (rad_upload_scene(ctx, s.getVertices(), color.data(), rad.data(), illum.data(), P) != RAD_OK) return fail("rad_upload_scene");
	// neighbour ids for the display stage (Patch::neighbours -> scene indices)
	{
		std::map<Patch*, int> index;
		Patch** pp = s.getPatches();
		for (unsigned int i = 0; i < P; i++) index[pp[i]] = (int)i;
		std::vector<int32_t> nb(8 * (size_t)P);
		for (unsigned int i = 0; i < P; i++)
			for (int j = 0; j < 8; j++) {
				std::map<Patch*, int>::const_iterator it = index.find(pp[i]->neighbours[j]);
				nb[8 * (size_t)i + j] = it == index.end() ? (int)i : it->second;
			}
		if (rad_upload_neighbours(ctx, nb.data(), P) != RAD_OK) return fail("rad_upload_neighbours");
	}
	passCounter = 0;
	computeRadiosity = true;
	return true;
}

bool RadiositySolver::shadeVertices(float* colors12) {
	if (!ctx) { err = "RadiositySolver::shadeVertices: not initialised"; return false; }
	if (rad_shade_vertices(ctx, colors12, NULL) != RAD_OK) return fail("rad_shade_vertices");
	return true;
}

bool RadiositySolver::shoot(unsigned int batches, bool stopTest, rad_stats* stats) {
	if (!ctx) { err = "RadiositySolver::shoot: not initialised"; return false; }
	rad_stats st;
	memset(&st, 0, sizeof(st));
	if (rad_shoot(ctx, batches, stopTest ? 1 : 0, &st) != RAD_OK) return fail("rad_shoot");
	passCounter += st.batches_done;
	if (stopTest && st.stopped) computeRadiosity = false;
	if (stats) *stats = st;
	return true;
}

bool RadiositySolver::syncToScene() {
	if (!ctx || !scene) { err = "RadiositySolver::syncToScene: not initialised"; return false; }
	const unsigned int P = scene->getPatchesCount();
	std::vector<float> rad(3 * (size_t)P), illum(3 * (size_t)P);
	if (rad_download_state(ctx, rad.data(), illum.data()) != RAD_OK) return fail("rad_download_state");
	Patch** pp = scene->getPatches();
	for (unsigned int i = 0; i < P; i++) {
		pp[i]->radiosity = Vector3f(rad[3 * i], rad[3 * i + 1], rad[3 * i + 2]);
		pp[i]->illumination = Vector3f(illum[3 * i], illum[3 * i + 1], illum[3 * i + 2]);
	}
	return true;
}

bool RadiositySolver::syncFromScene() {
	if (!ctx || !scene) { err = "RadiositySolver::syncFromScene: not initialised"; return false; }
	std::vector<float> color, rad, illum;
	gatherState(*scene, color, rad, illum);
	if (rad_upload_state(ctx, rad.data(), illum.data()) != RAD_OK) return fail("rad_upload_state");
	return true;
}
