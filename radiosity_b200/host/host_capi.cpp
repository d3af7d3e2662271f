// extern "C" view of the host library for ctypes (tests, bench.py).  Thin: every function forwards to the
// C++ classes that mirror the reference's host API.
#include <map>
#include <cstring>
#include "Config.h"
#include "ModelContainer.h"
#include "LoadingModel.h"
#include "FormFactors.h"
#include "Camera.h"
#include "Transform.h"
#include "Radiosity.h"
#include "Colors.h"
#include "SceneFile.h"
#include "MeshExport.h"

extern "C" {

void* radhost_scene_new() { return new ModelContainer(); }
void radhost_scene_free(void* s) { delete (ModelContainer*)s; }
void radhost_scene_load_cornell(void* s) { ((ModelContainer*)s)->load(); }
int radhost_scene_load_obj(void* s, const char* path) { try { return ((ModelContainer*)s)->load(std::string(path)) ? 1 : 0; } catch (...) { return 0; } }
int radhost_scene_load_static_mesh(void* s, const char* path, float scale, int flip, int emissive_material) {
	return ((ModelContainer*)s)->loadStaticMesh(std::string(path), scale, flip != 0, emissive_material) ? 1 : 0;
}
void radhost_scene_set_area(void* s, double area) { ((ModelContainer*)s)->maxPatchArea = area; }
unsigned radhost_scene_patch_count(void* s) { return ((ModelContainer*)s)->getPatchesCount(); }

void radhost_scene_get(void* sv, float* verts12, int* indices6, float* color3, float* rad3, float* illum3) {
	ModelContainer* s = (ModelContainer*)sv;
	const unsigned P = s->getPatchesCount();
	if (verts12) memcpy(verts12, s->getVertices(), sizeof(float) * 12 * (size_t)P);
	if (indices6) memcpy(indices6, s->getIndices(), sizeof(int) * 6 * (size_t)P);
	Patch** pp = s->getPatches();
	for (unsigned i = 0; i < P; i++) {
		const Vector3f c = pp[i]->getColor();
		if (color3) { color3[3 * i] = c.x; color3[3 * i + 1] = c.y; color3[3 * i + 2] = c.z; }
		if (rad3) { rad3[3 * i] = pp[i]->radiosity.x; rad3[3 * i + 1] = pp[i]->radiosity.y; rad3[3 * i + 2] = pp[i]->radiosity.z; }
		if (illum3) { illum3[3 * i] = pp[i]->illumination.x; illum3[3 * i + 1] = pp[i]->illumination.y; illum3[3 * i + 2] = pp[i]->illumination.z; }
	}
}
void radhost_scene_set_state(void* sv, const float* rad3, const float* illum3) {
	ModelContainer* s = (ModelContainer*)sv;
	Patch** pp = s->getPatches();
	for (unsigned i = 0; i < s->getPatchesCount(); i++) {
		if (rad3) pp[i]->radiosity = Vector3f(rad3[3 * i], rad3[3 * i + 1], rad3[3 * i + 2]);
		if (illum3) pp[i]->illumination = Vector3f(illum3[3 * i], illum3[3 * i + 1], illum3[3 * i + 2]);
	}
}
void radhost_scene_neighbours(void* sv, int* out8) {
	ModelContainer* s = (ModelContainer*)sv;
	const unsigned P = s->getPatchesCount();
	Patch** pp = s->getPatches();
	std::map<Patch*, int> idx;
	for (unsigned i = 0; i < P; i++) idx[pp[i]] = (int)i;
	for (unsigned i = 0; i < P; i++)
		for (int j = 0; j < 8; j++) {
			std::map<Patch*, int>::iterator it = idx.find(pp[i]->neighbours[j]);
			out8[8 * (size_t)i + j] = it == idx.end() ? -1 : it->second;
		}
}
void radhost_scene_select(void* sv, unsigned count, unsigned* ids, int* is_null) {
	ModelContainer* s = (ModelContainer*)sv;
	Patch** em = new Patch*[count];
	s->getHighestRadiosityPatchesId(count, em, ids);
	for (unsigned i = 0; i < count; i++) is_null[i] = em[i] == NULL;
	delete[] em;
}
unsigned radhost_scene_select_single(void* s) { return ((ModelContainer*)s)->getHighestRadiosityPatchId(); }

void radhost_patch_geom(void* sv, unsigned patch, float* c3, float* n3, float* u3) {
	Patch* p = ((ModelContainer*)sv)->getPatches()[patch];
	const Vector3f c = p->getCenter(), n = p->getNormal(), u = p->getUp();
	c3[0] = c.x; c3[1] = c.y; c3[2] = c.z; n3[0] = n.x; n3[1] = n.y; n3[2] = n.z; u3[0] = u.x; u3[1] = u.y; u3[2] = u.z;
}
// P * MV exactly as the reference's OnIdle composes it (Main.cpp:1172-1183); out[c*4+r]
void radhost_mvp(void* sv, unsigned patch, int look, float* out16) {
	Camera cam;
	Matrix4f proj, mv;
	CGLTransform::Perspective(proj, 90, 1.0f, 0.01f, 1000);
	mv.Identity();
	cam.lookFromPatch(((ModelContainer*)sv)->getPatches()[patch], (Camera::PatchLook)look);
	mv *= cam.GetMatrix();
	const Matrix4f mvp = proj * mv;
	memcpy(out16, &mvp.f[0][0], 64);
}
void radhost_projection(float* out16) {
	Matrix4f proj;
	CGLTransform::Perspective(proj, 90, 1.0f, 0.01f, 1000);
	memcpy(out16, &proj.f[0][0], 64);
}

void radhost_config(unsigned side, unsigned hemicubes, unsigned shoots, double area, unsigned* out9) {
	Config::unfreeze();
	Config::setHemicubeSide(side); Config::setHemicubesCount(hemicubes);
	if (shoots) Config::setShootsPerCycle(shoots);
	if (area > 0) Config::setMaxPatchArea(area);
	Config::freeze();
	if (!out9) return;
	out9[0] = Config::HEMICUBE_W(); out9[1] = Config::HEMICUBE_H(); out9[2] = Config::PATCHVIEW_TEX_W();
	out9[3] = Config::PATCHVIEW_TEX_H(); out9[4] = Config::PATCHVIEW_TEX_RES(); out9[5] = Config::OCL_WORKITEMS_X();
	out9[6] = Config::OCL_WORKITEMS_Y(); out9[7] = Config::SHOOTS_PER_CYCLE(); out9[8] = Config::HEMICUBES_CNT();
}
int radhost_config_set_when_frozen() {   // error behaviour check: the setter must refuse and keep the value
	const unsigned before = Config::HEMICUBE_W();
	Config::setHemicubeSide(before + 16);
	Config::freeze();
	return Config::HEMICUBE_W() == before;
}
void radhost_formfactors(unsigned side, unsigned hemicubes, float* out) {
	radhost_config(side, hemicubes, 0, 0, NULL);
	float* ff = precomputeHemicubeFormFactors();
	memcpy(out, ff, sizeof(float) * (size_t)Config::PATCHVIEW_TEX_RES() * hemicubes);
	delete[] ff;
}

void* radhost_solver_new(void* scene, int device, unsigned select_mode, unsigned flags, char* err, unsigned errlen) {
	RadiositySolver* s = new RadiositySolver();
	if (!s->init(*(ModelContainer*)scene, device, select_mode, flags)) {
		if (err && errlen) { strncpy(err, s->error().c_str(), errlen - 1); err[errlen - 1] = 0; }
		delete s;
		return NULL;
	}
	return s;
}
void radhost_solver_free(void* s) { delete (RadiositySolver*)s; }
int radhost_solver_shoot(void* s, unsigned batches, int stop_test, rad_stats* st) { return ((RadiositySolver*)s)->shoot(batches, stop_test != 0, st) ? 0 : -1; }
int radhost_solver_sync_to_scene(void* s) { return ((RadiositySolver*)s)->syncToScene() ? 0 : -1; }
int radhost_solver_sync_from_scene(void* s) { return ((RadiositySolver*)s)->syncFromScene() ? 0 : -1; }
rad_ctx* radhost_solver_ctx(void* s) { return ((RadiositySolver*)s)->context(); }
const char* radhost_solver_error(void* s) { return ((RadiositySolver*)s)->error().c_str(); }
unsigned radhost_solver_pass_counter(void* s) { return ((RadiositySolver*)s)->passCounter; }
int radhost_solver_running(void* s) { return ((RadiositySolver*)s)->computeRadiosity ? 1 : 0; }

// Colors::smoothShadePatch over the whole scene on the host (reference API) and on the device (RadiositySolver)
void radhost_scene_smooth_shade(void* sv, float* out12) {
	ModelContainer* s = (ModelContainer*)sv;
	Patch** pp = s->getPatches();
	for (unsigned i = 0; i < s->getPatchesCount(); i++) Colors::smoothShadePatch(out12 + 12 * (size_t)i, pp[i]);
}
int radhost_solver_shade(void* s, float* out12) { return ((RadiositySolver*)s)->shadeVertices(out12) ? 0 : -1; }

// (file contents are untrusted input: nothing may throw across the C ABI)
int radhost_scene_save(void* s, const char* path, int format) { try { return SaveToFile(std::string(path), *(ModelContainer*)s, (RRFormat)format) ? 1 : 0; } catch (...) { return 0; } }
int radhost_scene_load(void* s, const char* path) { try { return LoadFromFile(std::string(path), *(ModelContainer*)s) ? 1 : 0; } catch (...) { return 0; } }

int radhost_scene_export_ply(void* s, const float* colors12, const char* path, float exposure) { return ExportPly(std::string(path), *(ModelContainer*)s, colors12, exposure) ? 1 : 0; }

unsigned radhost_sizeof_patch() { return (unsigned)sizeof(Patch); }

} // extern "C"
