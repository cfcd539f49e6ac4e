#!/bin/sh
# Build oracle/_ref/libref.so: the reference's UNMODIFIED hot-path sources, compiled where they
# lie under /root/reference, against the header stand-ins in oracle/glm_stub (glm / text-csv are
# un-vendored Conan packages of the reference and cannot be fetched offline).  TEST INFRASTRUCTURE
# ONLY.  Outputs go to oracle/_ref/ (git-ignored, travels with gpurun).  The reference's own build
# (CMake + Conan) is not run.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then echo "build_ref: $REF absent; keeping prebuilt $OUT" >&2; exit 0; fi
mkdir -p "$OUT"
CXXFLAGS="-std=c++11 -O2 -ffp-contract=off -fPIC -Dcimg_display=0 -w -I$HERE/glm_stub -I$REF/headers -I$REF/vendor/cimg -I$REF/vendor/tinyobjloader"
pids=""
for f in geometry drawing shading material fileloader; do
  g++ $CXXFLAGS -c "$REF/$f.cpp" -o "$OUT/$f.o" & pids="$pids $!"
done
g++ $CXXFLAGS -c "$HERE/ref_driver.cpp" -o "$OUT/ref_driver.o" & pids="$pids $!"
for p in $pids; do wait $p; done
g++ -shared -o "$OUT/libref.so" "$OUT"/geometry.o "$OUT"/drawing.o "$OUT"/shading.o "$OUT"/material.o "$OUT"/fileloader.o "$OUT"/ref_driver.o -lpthread
rm -f "$OUT"/*.o
echo "built $OUT/libref.so"
