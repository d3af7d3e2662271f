// Delta form factors of the hemicube, laid out like the atlas (reference: FormFactors.h:20,
// FormFactors.cpp:23-67 Calc_HemicubeFormFactors, :280-339 precomputeHemicubeFormFactors).
#pragma once

// Returns a new[]-allocated array of PATCHVIEW_TEX_RES * HEMICUBES_CNT floats (the one-hemicube table
// repeated HEMICUBES_CNT times, as the reference uploads it, Main.cpp:576).  Uses the frozen Config.
float* precomputeHemicubeFormFactors();
// One hemicube only (3*N*N floats) into caller storage — what rad_set_formfactors() wants.
void computeHemicubeFormFactors(unsigned int hemicubeSide, float* out);
