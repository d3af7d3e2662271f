// Headless progressive-refinement driver: the compute half of the reference's OnIdle
// (Main.cpp:1124-1309) with the OpenGL hemicube render and the OpenCL ProcessHemicube kernel replaced by
// calls into librad_cuda.so (include/rad_cuda.h).  Window, input handling and drawing are not part of it.
#pragma once
#include <string>
#include "ModelContainer.h"
#include "../../include/rad_cuda.h"

class RadiositySolver {
public:
	RadiositySolver();
	~RadiositySolver();

	// InitGLObjects + InitCLObjects (Main.cpp:11-397, 405-611): takes N, k from the frozen Config, builds the dFF
	// table (precomputeHemicubeFormFactors), the projection (Perspective(90, 1, 0.01, 1000)) and uploads the scene.
	bool init(ModelContainer& scene, int device = 0, unsigned int selectMode = RAD_SELECT_REFERENCE, unsigned int flags = 0);
	// `batches` iterations of the shooting loop (Main.cpp:1137-1309); one batch shoots HEMICUBES_CNT patches.
	bool shoot(unsigned int batches, bool stopTest, rad_stats* stats = NULL);
	// copy B / I back into Patch::radiosity / Patch::illumination
	bool syncToScene();
	// push Patch::radiosity / Patch::illumination to the device again (after editing patches on the host)
	bool syncFromScene();
	// display stage on the device: Colors::smoothShadePatch for every patch (Main.cpp:1323-1341); out = float[P*12]
	bool shadeVertices(float* colors12);

	rad_ctx* context() { return ctx; }
	const std::string& error() const { return err; }
	unsigned int passCounter;      // batches executed so far (Main.h passCounter)
	bool computeRadiosity;         // cleared when the stop test fires (Main.cpp:1300)

private:
	bool fail(const char* where);
	ModelContainer* scene;
	rad_ctx* ctx;
	std::string err;
};
