// TEST INFRASTRUCTURE — thin extern "C" driver around the reference's OWN host classes
// (compiled from /root/reference/source by oracle/ref_build.sh into oracle/_ref/libref_host.so).
// It contains no algorithm of its own: every number it returns is produced by reference code
// (ModelContainer, Patch, Camera, CGLTransform, Colors, FormFactors, Config).  Used by tests/ to
// pin oracle/oracle.cpp and the product host library, and by tests/golden/make_golden.py to
// generate the committed fixtures.  Never linked or loaded by radiosity_b200/.
#include <iostream>
#include <iomanip>
#include <sstream>
#include <fstream>
#include <vector>
#include <deque>
#include <list>
#include <map>
#include <string>
#include <limits>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#define private public      // probe needs Config::frozen and Patch::vec1..4 / color
#define protected public
#include "Config.h"
#include "ModelContainer.h"
#include "Camera.h"
#include "Transform.h"
#include "Colors.h"
#include "LoadingModel.h"
#undef private
#undef protected
#include <map>
#include <string>

float* precomputeHemicubeFormFactors();   // FormFactors.cpp:280

static ModelContainer* g_scene = nullptr;

extern "C" {

// ModelContainer::load() (ModelContainer.cpp:35-56) + maxPatchArea (Main.cpp:842-843)
unsigned refp_scene_build(double area) {
	delete g_scene;
	g_scene = new ModelContainer();
	g_scene->load();
	g_scene->maxPatchArea = area;
	return g_scene->getPatchesCount();
}

// WaveFrontModel (WaveFrontModel.cpp:15-144) as the only model of a scene
unsigned refp_scene_build_obj(const char* path, double area) {
	delete g_scene;
	g_scene = new ModelContainer();
	g_scene->addModel(new WaveFrontModel(std::string(path)));
	g_scene->maxPatchArea = area;
	return g_scene->getPatchesCount();
}

unsigned refp_patch_count() { return g_scene ? g_scene->getPatchesCount() : 0; }

// flat arrays exactly as the reference hands them to GL (ModelContainer.cpp:81-155) + patch state
void refp_scene_get(float* verts12, int* indices6, float* color3, float* rad3, float* illum3) {
	unsigned P = g_scene->getPatchesCount();
	float* v = g_scene->getVertices();
	int* ix = g_scene->getIndices();
	Patch** pp = g_scene->getPatches();
	if (verts12) memcpy(verts12, v, sizeof(float) * 12 * P);
	if (indices6) memcpy(indices6, ix, sizeof(int) * 6 * P);
	for (unsigned i = 0; i < P; i++) {
		Vector3f c = pp[i]->getColor();
		if (color3) { color3[3*i] = c.x; color3[3*i+1] = c.y; color3[3*i+2] = c.z; }
		if (rad3) { rad3[3*i] = pp[i]->radiosity.x; rad3[3*i+1] = pp[i]->radiosity.y; rad3[3*i+2] = pp[i]->radiosity.z; }
		if (illum3) { illum3[3*i] = pp[i]->illumination.x; illum3[3*i+1] = pp[i]->illumination.y; illum3[3*i+2] = pp[i]->illumination.z; }
	}
}

void refp_scene_set_radiosity(const float* rad3) {
	unsigned P = g_scene->getPatchesCount();
	Patch** pp = g_scene->getPatches();
	for (unsigned i = 0; i < P; i++)
		pp[i]->radiosity = Vector3f(rad3[3*i], rad3[3*i+1], rad3[3*i+2]);
}

// neighbour pointers -> scene indices (Patch.cpp:129-218, Model.cpp:39-47)
void refp_neighbours(int* out8) {
	unsigned P = g_scene->getPatchesCount();
	Patch** pp = g_scene->getPatches();
	std::map<Patch*, int> idx;
	for (unsigned i = 0; i < P; i++) idx[pp[i]] = (int)i;
	for (unsigned i = 0; i < P; i++)
		for (int j = 0; j < 8; j++) {
			std::map<Patch*, int>::iterator it = idx.find(pp[i]->neighbours[j]);
			out8[8*i+j] = it == idx.end() ? -1 : it->second;
		}
}

// ModelContainer::getHighestRadiosityPatchesId (ModelContainer.cpp:259-299)
void refp_select(unsigned count, unsigned* ids, int* is_null) {
	Patch** em = new Patch*[count];
	g_scene->getHighestRadiosityPatchesId(count, em, ids);
	for (unsigned i = 0; i < count; i++) is_null[i] = em[i] == NULL;
	delete[] em;
}

// ModelContainer::getHighestRadiosityPatchId (ModelContainer.cpp:222-237)
unsigned refp_select_single() { return g_scene->getHighestRadiosityPatchId(); }

// Patch::getCenter/getNormal/getUp (Patch.cpp:253-276)
void refp_patch_geom(unsigned patch, float* center3, float* normal3, float* up3) {
	Patch* p = g_scene->getPatches()[patch];
	Vector3f c = p->getCenter(), n = p->getNormal(), u = p->getUp();
	center3[0] = c.x; center3[1] = c.y; center3[2] = c.z;
	normal3[0] = n.x; normal3[1] = n.y; normal3[2] = n.z;
	up3[0] = u.x; up3[1] = u.y; up3[2] = u.z;
}

// the MVP exactly as OnIdle builds it (Main.cpp:1172-1183): Perspective(90,1,0.01,1000) * (I *= LookAt)
// look: Camera::PatchLook value (FRONT=0, UP, DOWN, LEFT, RIGHT).  out[c*4+r] = m[c][r] (column-major).
void refp_mvp(unsigned patch, int look, float* out16) {
	Camera cam;
	Matrix4f t_projection;
	CGLTransform::Perspective(t_projection, 90, 1.0f, 0.01f, 1000);
	Matrix4f t_modelview;
	t_modelview.Identity();
	cam.lookFromPatch(g_scene->getPatches()[patch], (Camera::PatchLook)look);
	t_modelview *= cam.GetMatrix();
	Matrix4f t_mvp = t_projection * t_modelview;
	for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) out16[c*4+r] = t_mvp[c][r];
}

void refp_projection(float* out16) {
	Matrix4f t_projection;
	CGLTransform::Perspective(t_projection, 90, 1.0f, 0.01f, 1000);
	for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) out16[c*4+r] = t_projection[c][r];
}

// Config::freeze derived sizes (Config.cpp:30-48).  out: HEMICUBE_W, HEMICUBE_H, TEX_W, TEX_H, TEX_RES,
// OCL_WORKITEMS_X, OCL_WORKITEMS_Y, SHOOTS_PER_CYCLE, HEMICUBES_CNT
void refp_config(unsigned side, unsigned hemicubes, unsigned* out9) {
	Config::frozen = false;
	Config::setHemicubeSide(side);
	Config::setHemicubesCount(hemicubes);
	Config::freeze();
	out9[0] = Config::HEMICUBE_W(); out9[1] = Config::HEMICUBE_H();
	out9[2] = Config::PATCHVIEW_TEX_W(); out9[3] = Config::PATCHVIEW_TEX_H(); out9[4] = Config::PATCHVIEW_TEX_RES();
	out9[5] = Config::OCL_WORKITEMS_X(); out9[6] = Config::OCL_WORKITEMS_Y();
	out9[7] = Config::SHOOTS_PER_CYCLE(); out9[8] = Config::HEMICUBES_CNT();
}

// precomputeHemicubeFormFactors (FormFactors.cpp:280-339); out has TEX_RES * hemicubes floats
void refp_formfactors(unsigned side, unsigned hemicubes, float* out) {
	unsigned c[9];
	refp_config(side, hemicubes, c);
	float* ff = precomputeHemicubeFormFactors();
	memcpy(out, ff, sizeof(float) * (size_t)c[4] * hemicubes);
	delete[] ff;
}

// Colors codec (Colors.cpp:31-110).  out: shift[3], revMask[3], mask[3], correction, range
void refp_colors_setup(unsigned patches, unsigned* out11) {
	Colors::setNeededColors(patches);
	short* sh = Colors::getShifts();
	unsigned* rm = Colors::getRevMasks();
	for (int i = 0; i < 3; i++) { out11[i] = (unsigned)sh[i]; out11[3+i] = rm[i]; out11[6+i] = Colors::mask[i]; }
	out11[9] = Colors::getCorrection();
	out11[10] = Colors::range;
}
unsigned refp_color(unsigned colorIndex) { return Colors::color(colorIndex); }
unsigned refp_color_index(unsigned color) { return (unsigned)Colors::index(color); }

// display stage: Colors::smoothShadePatch for every patch (Colors.cpp:198-261, called from Main.cpp:1323-1341)
void refp_smooth_shade(float* out12) {
	unsigned P = g_scene->getPatchesCount();
	Patch** pp = g_scene->getPatches();
	for (unsigned i = 0; i < P; i++) Colors::smoothShadePatch(out12 + 12 * (size_t)i, pp[i]);
}
void refp_scene_set_illumination(const float* il3) {
	unsigned P = g_scene->getPatchesCount();
	Patch** pp = g_scene->getPatches();
	for (unsigned i = 0; i < P; i++) pp[i]->illumination = Vector3f(il3[3*i], il3[3*i+1], il3[3*i+2]);
}

// a `.rr` file exactly as the reference's SaveToFile body writes it (Main.cpp:1596-1632), with the reference's own Patch
// class and this build's ABI (LP64: 8-byte count, sizeof(Patch) = 184)
int refp_save_rr(const char* path) {
	FILE* fp = fopen(path, "wb");
	if (!fp) return 0;
	unsigned long count = g_scene->getPatchesCount();
	Patch** patches = g_scene->getPatches();
	Patch* data = new Patch[count];
	for (unsigned long i = 0; i < count; i++) data[i] = (*patches[i]);
	std::map<Patch*, unsigned long> where;
	for (unsigned long j = 0; j < count; j++) where[patches[j]] = j;
	for (unsigned long i = 0; i < count; i++)
		for (unsigned int n = 0; n < 8; n++) {
			std::map<Patch*, unsigned long>::iterator it = where.find(patches[i]->neighbours[n]);
			if (it != where.end()) { data[i].relativeNeighbours[n] = it->second; data[i].neighbours[n] = NULL; }
		}
	size_t w = fwrite(&count, sizeof(unsigned long), 1, fp);
	w += fwrite(data, sizeof(Patch), count, fp);
	fclose(fp);
	delete[] data;
	return w == 1 + count;
}
// LoadingModel round trip with the reference's own classes (LoadFromFile body, Main.cpp:1492-1519)
unsigned refp_load_rr(const char* path) {
	FILE* fp = fopen(path, "rb");
	if (!fp) return 0;
	unsigned long count = 0;
	if (fread(&count, sizeof(unsigned long), 1, fp) != 1) { fclose(fp); return 0; }
	Patch* data = new Patch[count];
	size_t rd = fread(data, sizeof(Patch), count, fp);
	fclose(fp);
	if (rd != count) { delete[] data; return 0; }
	delete g_scene;
	g_scene = new ModelContainer();
	g_scene->addModel(new LoadingModel(data, count));
	delete[] data;
	return g_scene->getPatchesCount();
}

unsigned refp_sizeof_patch() { return (unsigned)sizeof(Patch); }

// ---- the CPU tail of a batch, run on the reference's OWN text ------------------------------------------------------
// oracle/ref_build.sh prints three pieces of Main.cpp into its temp dir (nothing is copied into the repo):
//   main_tail_snapshot.inc   the statement `p_tmp_radiosities[hi] = p_emitters[hi]->radiosity;`            (Main.cpp:1161)
//   main_tail_transfer.inc   record gather into p_tmp_formfactors + energy transfer, per hemicube          (Main.cpp:1251-1279)
//   main_tail_update.inc     emitter update, lastEnergy, stop test                                         (Main.cpp:1284-1303)
// with the MARK(...) profiling lines dropped.  This function only declares the variables that text uses, under the
// names Main.cpp gives them (Main.cpp:301,311,605-607,1115-1116,1140,1232), and runs it on the probe's scene.
//   em_ids / em_null   the emitter list of the batch as getHighestRadiosityPatchesId returned it
//   rec_*              the kernel's record stream (hemicube, patch id, energy), n_records = write index
// Returns computeRadiosity (0 after the stop test fired); *last_len = lastEnergy.f_Length().
#define MARK(x)
int refp_main_tail(unsigned k, const unsigned* em_ids, const int* em_null, unsigned n_records,
                   const unsigned* rec_hemicubes, const unsigned* rec_ids, const float* rec_energies, float* last_len) {
	Config::frozen = false;
	Config::setHemicubesCount(k);
	Config::freeze();
	ModelContainer& scene = *g_scene;
	Patch** scenePatches = scene.getPatches();
	unsigned int scenePatchesCount = scene.getPatchesCount();
	std::vector<Patch*> emitters(k);
	for (unsigned i = 0; i < k; i++) emitters[i] = em_null[i] ? NULL : scenePatches[em_ids[i]];
	Patch** p_emitters = &emitters[0];
	std::vector<Vector3f> tmp_radiosities(k);
	Vector3f* p_tmp_radiosities = &tmp_radiosities[0];
	std::vector<float> tmp_formfactors(scenePatchesCount, 0.0f);
	float* p_tmp_formfactors = &tmp_formfactors[0];
	const unsigned int* p_ocl_hemicubes = rec_hemicubes; const unsigned int* p_ocl_pids = rec_ids; const float* p_ocl_energies = rec_energies;
	unsigned int n_last_index = n_records;
	bool computeRadiosity = true, debugOutput = false;
	unsigned int passCounter = 0, shoot = 0;
	struct { double f_Time() const { return 0.0; } } timer; double t_start = 0.0;
	std::streambuf* quiet = std::cout.rdbuf(NULL);          // the "Done in ..." line of the stop test
	for (unsigned int hi = 0; hi < Config::HEMICUBES_CNT(); hi++) {
		if (p_emitters[hi] == NULL) continue;               // Main.cpp:1157-1158
#include "main_tail_snapshot.inc"
	}
	{
#include "main_tail_transfer.inc"
	}
#include "main_tail_update.inc"
	std::cout.rdbuf(quiet);
	(void)passCounter; (void)shoot; (void)debugOutput; (void)t_start;
	*last_len = lastEnergy.f_Length();
	return computeRadiosity ? 1 : 0;
}
// patch state back out (after refp_main_tail)
void refp_scene_get_state(float* rad3, float* illum3) { refp_scene_get(NULL, NULL, NULL, rad3, illum3); }

} // extern "C"
