// Run parameters of the radiosity solver — same static interface as the reference's Config
// (Config.h:6-38): setters until freeze(), derived atlas sizes afterwards (Config.cpp:30-48).
#pragma once
#include <iostream>

class Config {
public:
	static void setHemicubeSide(unsigned int n);     // hemicube edge in pixels (reference default 16)
	static void setOCLWorkitemsX(unsigned int n);    // kept for API compatibility: spans per atlas row of the reference's OpenCL kernel
	static void setMaxPatchArea(double n);           // subdivision target area (default 0.5)
	static void setShootsPerCycle(unsigned int n);   // batches per OnIdle (default 500)
	static void setHemicubesCount(unsigned int n);   // emitters per batch, k (default 10)

	static void freeze();
	static bool isFrozen() { return frozen; }
	static void unfreeze();                          // extension: lets a long-lived host reconfigure (tests, benches)

	static unsigned int HEMICUBE_W();
	static unsigned int HEMICUBE_H();
	static unsigned int PATCHVIEW_TEX_W();           // 2N
	static unsigned int PATCHVIEW_TEX_H();           // 1.5N
	static unsigned int PATCHVIEW_TEX_RES();         // 3N^2
	static double MAX_PATCH_AREA();
	static unsigned int OCL_WORKITEMS_X();
	static unsigned int OCL_WORKITEMS_Y();
	static unsigned int SHOOTS_PER_CYCLE();
	static unsigned int HEMICUBES_CNT();

private:
	static bool guard(const char* what);
	static bool frozen;
	static unsigned int hemicubeSide, oclWorkitemsX, shootsPerCycle, hemicubesCount;
	static double maxPatchArea;
};
