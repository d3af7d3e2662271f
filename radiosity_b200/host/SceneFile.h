// Checkpoint / resume of a radiosity run: the reference's Ctrl+S / Ctrl+O `.rr` files
// (SaveToFile Main.cpp:1562-1647, LoadFromFile Main.cpp:1460-1555, LoadingModel.cpp:4-17) without the Win32 dialogs.
//
// The reference writes `unsigned long count` followed by a raw dump of `Patch[count]`, neighbour pointers replaced by scene
// indices in relativeNeighbours[] — a layout that depends on the ABI it was compiled for:
//     RR_REFERENCE_WIN32   4-byte count, 148-byte records (32-bit pointers: the reference's own target)
//     RR_REFERENCE_LP64    8-byte count, 184-byte records (what the reference produces when built on 64-bit Linux)
// Both can be read and written here field by field.  RR_PORTABLE is a versioned little-endian format of our own:
//     "RRB2" u32 version(1) u64 count, then per patch 12 f32 vertices, 3 f32 colour, 3 f32 radiosity, 3 f32 illumination, 8 u32 neighbours.
#pragma once
#include <string>
#include "ModelContainer.h"

enum RRFormat { RR_PORTABLE = 0, RR_REFERENCE_WIN32 = 1, RR_REFERENCE_LP64 = 2 };

// Writes every patch of the scene (geometry, colour, B, I, neighbour indices).  Returns false on I/O error.
bool SaveToFile(const std::string& path, ModelContainer& scene, RRFormat format = RR_PORTABLE);
// Replaces the content of `scene` by the patches of the file (through a LoadingModel, as the reference does); the
// format is detected from the magic / the file size.  Returns false (scene untouched) if the file is not recognised.
bool LoadFromFile(const std::string& path, ModelContainer& scene);
