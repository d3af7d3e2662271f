// Shaded-mesh export: what the reference shows in its window after a run — per-vertex colours from
// Colors::smoothShadePatch (OnIdle, Main.cpp:1318-1366: the colour VBO is refilled after every cycle; OpenGL clamps the
// float colours to [0, 1]) on the scene's quads — written as a binary little-endian PLY so that a headless run can be
// looked at (MeshLab, Blender, ...): 4 vertices per patch (x y z float, red green blue uchar), one quad face per patch.
#pragma once
#include <string>
#include "ModelContainer.h"

// colors12: 12 floats per patch (r g b of the patch's four vertices), as produced by Colors::smoothShadePatch on the
// host or by RadiositySolver::shadeVertices on the device.  exposure scales the colours before the clamp.
// Returns false on I/O error.
bool ExportPly(const std::string& path, ModelContainer& scene, const float* colors12, float exposure = 1.0f);
