#include "Config.h"
#include <algorithm>

bool Config::frozen = false;
unsigned int Config::hemicubeSide = 16;
unsigned int Config::oclWorkitemsX = 4;
unsigned int Config::shootsPerCycle = 500;
double Config::maxPatchArea = 0.5;
unsigned int Config::hemicubesCount = 10;

namespace {
struct Derived { unsigned w, h, tex_w, tex_h, tex_res, wi_x, wi_y; double area; } g_d = { 0, 0, 0, 0, 0, 0, 0, 0 };
}

bool Config::guard(const char*) {
	if (!frozen) return true;
	std::cerr << "Error: Trying to modify frozen configuration" << std::endl;   // same behaviour as Config.cpp:54-112
	return false;
}

void Config::setHemicubeSide(unsigned int n) { if (guard("hemicube")) hemicubeSide = n; }
void Config::setOCLWorkitemsX(unsigned int n) { if (guard("workitems")) oclWorkitemsX = n; }
void Config::setMaxPatchArea(double n) { if (guard("area")) maxPatchArea = n; }
void Config::setShootsPerCycle(unsigned int n) { if (guard("shoots")) shootsPerCycle = n; }
void Config::setHemicubesCount(unsigned int n) { if (guard("hemicubes")) hemicubesCount = n; }

void Config::freeze() {
	frozen = true;
	g_d.w = g_d.h = hemicubeSide;
	g_d.tex_w = (unsigned int)(g_d.w * 2);
	g_d.tex_h = (unsigned int)(g_d.h * 1.5);
	g_d.tex_res = (unsigned int)(g_d.tex_w * g_d.tex_h);
	g_d.area = maxPatchArea;
	g_d.wi_x = std::min(oclWorkitemsX, g_d.tex_w);
	g_d.wi_y = g_d.tex_h * hemicubesCount;
}
void Config::unfreeze() { frozen = false; }

unsigned int Config::HEMICUBE_W() { return g_d.w; }
unsigned int Config::HEMICUBE_H() { return g_d.h; }
unsigned int Config::PATCHVIEW_TEX_W() { return g_d.tex_w; }
unsigned int Config::PATCHVIEW_TEX_H() { return g_d.tex_h; }
unsigned int Config::PATCHVIEW_TEX_RES() { return g_d.tex_res; }
double Config::MAX_PATCH_AREA() { return g_d.area; }
unsigned int Config::OCL_WORKITEMS_X() { return g_d.wi_x; }
unsigned int Config::OCL_WORKITEMS_Y() { return g_d.wi_y; }
unsigned int Config::SHOOTS_PER_CYCLE() { return shootsPerCycle; }
unsigned int Config::HEMICUBES_CNT() { return hemicubesCount; }
