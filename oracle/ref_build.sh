#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the *unmodified-in-logic* host subset of the reference
# (david-sabata/Radiosity, /root/reference/source) into oracle/_ref/libref_host.so so that
# the CPU oracle restatement (oracle/oracle.cpp) and the product host library can be pinned
# against the reference's own code.  Nothing under radiosity_b200/ may link or load this.
#
# The reference is MSVC-dialect C++ (no build files ship with it).  We never copy its sources
# into the repo: they are streamed through a handful of purely syntactic `sed` fixes into a
# throw-away temp dir, compiled there with a 4-header shim, and only the resulting .so is kept
# (oracle/_ref/ is git-ignored but travels to the GPU box with the gpurun snapshot).
#
# Units built: Vector Transform Patch Model PrimitiveModel WaveFrontModel
# LoadingModel ModelContainer Config Colors FormFactors Camera.   Units that cannot be built
# (need Win32/WGL/GL/CL): Main OpenGL30Drv FrameBuffer Shaders — the GL raster and the OpenCL
# kernel are therefore restated in oracle/oracle.cpp ("port").  The OpenCL kernel's TEXT, however, is plain C apart from
# a few built-ins: it is compiled here too and run on the CPU (oracle/ref_kernel.cpp) to pin the restatement.
set -euo pipefail
REF=${REF_SRC:-/root/reference/source}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "ref_build: $REF not present (GPU box?) — keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/src" "$TMP/shim/GL"

UNITS="Vector Transform Patch Model PrimitiveModel WaveFrontModel LoadingModel ModelContainer Config Colors FormFactors Camera"
HDRS="Vector.h Transform.h Patch.h Model.h PrimitiveModel.h WaveFrontModel.h LoadingModel.h ModelContainer.h Config.h Colors.h FormFactors.h Camera.h Timer.h OpenGL30Drv.h"

fix() {  # syntactic MSVC -> g++ fixes only; no logic is touched
  sed -E \
    -e 's/unsigned int\(/(unsigned int)(/g' \
    -e 's/bool Comparator::operator\(\)/bool operator()/' \
    -e 's/const enum \{/enum {/' \
    -e 's/static enum PatchLook/enum PatchLook/' "$1"
}
for h in $HDRS; do [ -f "$REF/$h" ] && fix "$REF/$h" > "$TMP/src/$h"; done
for u in $UNITS; do fix "$REF/$u.cpp" > "$TMP/src/$u.cpp"; done
# in-class extra qualification "Patch::Patch(" (header only)
sed -i -E 's/Patch::Patch\(/Patch(/g' "$TMP/src/Patch.h"
sed -i -E 's/bool WaveFrontModel::parse/bool parse/' "$TMP/src/WaveFrontModel.h"
sed -i -E 's/bool ModelContainer::operator\(\)/bool operator()/' "$TMP/src/ModelContainer.h"

cat > "$TMP/shim/crtdbg.h" <<'H'
#pragma once
#include <cassert>
#define _ASSERT(x) assert(x)
#define _ASSERTE(x) assert(x)
H
cat > "$TMP/shim/windows.h" <<'H'
#pragma once
#include <cstdint>
#include <cstring>
typedef void *HDC, *HGLRC, *HWND;
typedef unsigned char byte;
struct PIXELFORMATDESCRIPTOR { int d; };
#define __int32 int
H
printf '#pragma once\n#include "Model.h"\n' > "$TMP/shim/model.h"
printf '#pragma once\ntypedef unsigned int GLuint, GLenum; typedef int GLint;\n' > "$TMP/shim/GL/glew.h"

CXXFLAGS="-std=c++17 -O2 -fPIC -fpermissive -w -ffp-contract=off -include cstdint -include cstring -include cassert -include crtdbg.h -include algorithm -include cmath -I$TMP/shim -I$TMP/src"
OBJS=""
for u in $UNITS; do
  g++ $CXXFLAGS -c "$TMP/src/$u.cpp" -o "$TMP/$u.o"
  OBJS="$OBJS $TMP/$u.o"
done
# the CPU tail of a batch (Main.cpp:1161, 1251-1279, 1284-1303) as TEXT for oracle/ref_probe.cpp's refp_main_tail: printed from
# the reference's Main.cpp into the temp dir, profiling MARK lines dropped; Main.cpp itself cannot be compiled (Win32/GL/CL)
sed -n '/p_tmp_radiosities\[hi\] = p_emitters\[hi\]->radiosity;/p' "$REF/Main.cpp" > "$TMP/src/main_tail_snapshot.inc"
sed -n '/MARK("clEnqueueReleaseGLObjects");/,/MARK("energies update");/p' "$REF/Main.cpp" | sed -e '/MARK(/d' > "$TMP/src/main_tail_transfer.inc"
sed -n '/Vector3f lastEnergy;/,/MARK("emitters update");/p' "$REF/Main.cpp" | sed -e '/MARK(/d' > "$TMP/src/main_tail_update.inc"
for f in snapshot transfer update; do [ -s "$TMP/src/main_tail_$f.inc" ] || { echo "ref_build: could not extract main_tail_$f from Main.cpp" >&2; exit 1; }; done
g++ $CXXFLAGS -c "$HERE/ref_probe.cpp" -o "$TMP/ref_probe.o"
# the reference's OpenCL kernel TEXT (Kernel_ProcessHemicube.h), printed from the reference header into the temp dir and
# compiled as C++ with the built-ins of oracle/ref_kernel.cpp; one syntactic fix: the vector literal (int2)(x, y)
printf '#include <cstdio>\n#include "Kernel_ProcessHemicube.h"\nint main() { fputs(kernel_processHemicube, stdout); return 0; }\n' > "$TMP/kgen.cpp"
g++ -w -I"$REF" "$TMP/kgen.cpp" -o "$TMP/kgen"
"$TMP/kgen" | sed -e 's/(int2)(/make_int2(/g' -e '/#pragma OPENCL/d' > "$TMP/src/kernel_body.inc"
g++ $CXXFLAGS -c "$HERE/ref_kernel.cpp" -o "$TMP/ref_kernel.o"
g++ -shared -o "$OUT/libref_host.so" $OBJS "$TMP/ref_probe.o" "$TMP/ref_kernel.o"
echo "ref_build: wrote $OUT/libref_host.so"
