// Display helper of the reference kept for API completeness (reference: Colors.h:35-52, Colors.cpp:198-261).
// Only smoothShadePatch is provided: the patch-id <-> colour codec of the reference (Colors::color / index / ...) exists to
// survive an RGBA8 framebuffer; the CUDA path stores ids directly, so the codec lives in the test oracle only.
#pragma once
#include "Patch.h"

class Colors {
public:
	// 12 floats: vertex colours lb, rb, rt, lt = mean over the patch and three of its neighbours of colour (.) (I + B)
	static void smoothShadePatch(float* colors, Patch* p);
};
