// Headless `radiosity` driver: the reference's main() (Main.cpp:761-911) without the window.
// Same bare "key value" arguments: area, hemicube, shoots, hemicubes (Main.cpp:769-782); extra keys:
//   shots <n>     total batches to run (default: until the stop test fires, checked every `shoots` batches)
//   select <reference|topk>     shooter selection semantics (include/rad_cuda.h)
//   device <n>    CUDA device ordinal
//   obj <path>    load a Wavefront OBJ scene instead of the built-in Cornell box
//   mesh <path>   load a "static 3DS export" header such as the reference's TestModel.h (StaticMeshModel.h);
//                 mesh_scale <f> (default 0.01), mesh_flip <0|1>, mesh_emit <material index> go with it
//   dump <path>   write "id Bx By Bz Ix Iy Iz" per patch when done
//   save <path>   checkpoint the scene + energies when done (portable .rr, SceneFile.h; the reference's Ctrl+S)
//   ply <path>    write the shaded scene (vertex colours of the display stage, computed on the GPU) as a binary PLY;
//                 exposure <f> scales the colours before the [0, 1] clamp (default 1)
//   load <path>   resume from a checkpoint instead of building a scene (the reference's Ctrl+O; also reads its raw dumps)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <iostream>
#include <iomanip>
#include <fstream>
#include "Config.h"
#include "ModelContainer.h"
#include "Radiosity.h"
#include "SceneFile.h"
#include "MeshExport.h"
#include <vector>

int main(int argc, const char** argv) {
	if ((argc - 1) % 2 > 0) { std::cerr << "Wrong number of arguments (expected key value pairs)" << std::endl; return -1; }
	long shots = -1; int device = 0; unsigned int select = RAD_SELECT_REFERENCE;
	const char* obj = NULL; const char* dump = NULL; const char* save = NULL; const char* load = NULL;
	const char* ply = NULL; float exposure = 1.0f;
	const char* mesh = NULL; float mesh_scale = 0.01f; bool mesh_flip = false; int mesh_emit = -1;
	for (int i = 1; i < argc; i += 2) {
		const char* k = argv[i]; const char* v = argv[i + 1];
		if (!strcmp(k, "area")) Config::setMaxPatchArea(atof(v));
		else if (!strcmp(k, "hemicube")) Config::setHemicubeSide(atoi(v));
		else if (!strcmp(k, "shoots")) Config::setShootsPerCycle(atoi(v));
		else if (!strcmp(k, "hemicubes")) Config::setHemicubesCount(atoi(v));
		else if (!strcmp(k, "shots")) shots = atol(v);
		else if (!strcmp(k, "device")) device = atoi(v);
		else if (!strcmp(k, "select")) select = !strcmp(v, "topk") ? RAD_SELECT_TOPK : RAD_SELECT_REFERENCE;
		else if (!strcmp(k, "obj")) obj = v;
		else if (!strcmp(k, "mesh")) mesh = v;
		else if (!strcmp(k, "mesh_scale")) mesh_scale = (float)atof(v);
		else if (!strcmp(k, "mesh_flip")) mesh_flip = atoi(v) != 0;
		else if (!strcmp(k, "mesh_emit")) mesh_emit = atoi(v);
		else if (!strcmp(k, "dump")) dump = v;
		else if (!strcmp(k, "save")) save = v;
		else if (!strcmp(k, "load")) load = v;
		else if (!strcmp(k, "ply")) ply = v;
		else if (!strcmp(k, "exposure")) exposure = (float)atof(v);
	}
	Config::freeze();

	ModelContainer scene;
	if (load) { if (!LoadFromFile(std::string(load), scene)) { std::cerr << "Unable to load checkpoint '" << load << "'" << std::endl; return -1; } }
	else if (obj) { if (!scene.load(std::string(obj))) { std::cerr << "Unable to load '" << obj << "'" << std::endl; return -1; } }
	else if (mesh) { if (!scene.loadStaticMesh(std::string(mesh), mesh_scale, mesh_flip, mesh_emit)) { std::cerr << "Unable to load '" << mesh << "'" << std::endl; return -1; } }
	else scene.load();
	if (!load) scene.maxPatchArea = Config::MAX_PATCH_AREA();
	std::cout << "patches: " << scene.getPatchesCount() << ", hemicube " << Config::HEMICUBE_W() << ", atlas "
	          << Config::PATCHVIEW_TEX_W() << "x" << Config::PATCHVIEW_TEX_H() << " x " << Config::HEMICUBES_CNT() << std::endl;

	RadiositySolver solver;
	if (!solver.init(scene, device, select)) { std::cerr << solver.error() << std::endl; return -1; }

	const auto t0 = std::chrono::steady_clock::now();
	double gpu_ms = 0; unsigned long cycles = 0;
	while (solver.computeRadiosity && (shots < 0 || (long)solver.passCounter < shots)) {
		unsigned int n = Config::SHOOTS_PER_CYCLE();
		if (shots >= 0 && (long)(solver.passCounter + n) > shots) n = (unsigned int)(shots - solver.passCounter);
		rad_stats st;
		if (!solver.shoot(n, shots < 0, &st)) { std::cerr << solver.error() << std::endl; return -1; }
		gpu_ms += st.gpu_ms; cycles += st.shots_done;
	}
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	std::cout << "Done in " << secs << " seconds, " << cycles << " cycles" << std::endl;   // Main.cpp:1299
	std::cout << "gpu time " << gpu_ms << " ms, " << std::setprecision(6) << (cycles / (gpu_ms * 1e-3)) << " hemicubes/s" << std::endl;
	if (!solver.syncToScene()) { std::cerr << solver.error() << std::endl; return -1; }
	if (save && !SaveToFile(std::string(save), scene)) { std::cerr << "Unable to write '" << save << "'" << std::endl; return -1; }
	if (ply) {     // display stage on the device (K5), then the mesh the reference would have drawn
		std::vector<float> colors((size_t)scene.getPatchesCount() * 12);
		if (!solver.shadeVertices(colors.data())) { std::cerr << solver.error() << std::endl; return -1; }
		if (!ExportPly(std::string(ply), scene, colors.data(), exposure)) { std::cerr << "Unable to write '" << ply << "'" << std::endl; return -1; }
	}
	if (dump) {
		std::ofstream out(dump);
		Patch** pp = scene.getPatches();
		out << std::setprecision(9);
		for (unsigned int i = 0; i < scene.getPatchesCount(); i++)
			out << i << " " << pp[i]->radiosity.x << " " << pp[i]->radiosity.y << " " << pp[i]->radiosity.z << " "
			    << pp[i]->illumination.x << " " << pp[i]->illumination.y << " " << pp[i]->illumination.z << "\n";
	}
	return 0;
}
