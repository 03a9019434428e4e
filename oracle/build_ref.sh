#!/bin/bash
# Builds ONE variant of the reference kernels + harness into oracle/_ref/ (TEST INFRASTRUCTURE; called by oracle/Makefile).
#   build_ref.sh <out.so> <kind: ref|dropin> <variant: 4096|256|256lod>
# The unmodified reference sources are compiled where they lie under $REF. Variants with other compile-time constants are
# compiled from a THROW-AWAY copy of $REF made under a temporary directory, with variables.h patched by sed (constants only);
# the copy is deleted after the build, so that no reference source ever lives in this repository's tree -- only the built .so.
set -euo pipefail
OUT=$1; KIND=$2; VARIANT=$3
REF=${REF:-/root/reference/src}
NVCC=${NVCC:-nvcc}; CXX=${CXX:-g++}
HERE=$(cd "$(dirname "$0")" && pwd)
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVFLAGS="-std=c++20 $ARCH -rdc=true --expt-relaxed-constexpr -Xcudafe --diag_suppress=20012 -Xcudafe --diag_suppress=177 -Xcompiler -fPIC -lineinfo -include $HERE/shim/prelude.h -I$HERE/shim"
CXXFLAGS="-std=c++20 -O2 -mbmi2 -fPIC -w -include $HERE/shim/prelude.h -I$HERE/shim -I/usr/local/cuda/include"
TMP=$(mktemp -d /tmp/brickmap_ref_XXXXXX); trap 'rm -rf "$TMP"' EXIT
SRC=$REF
if [ "$VARIANT" != "4096" ]; then
	SRC=$TMP/src; mkdir -p "$SRC"; cp -r "$REF"/. "$SRC"/; chmod -R u+w "$SRC"
	sed -i -e 's/grid_size = 4096;/grid_size = 256;/' -e 's/grid_height = 512;/grid_height = 256;/' \
	       -e 's/ray_queue_buffer_size = 2 \* 1.048.576;/ray_queue_buffer_size = 262144;/' "$SRC/variables.h"
	grep -q 'grid_size = 256;' "$SRC/variables.h" && grep -q 'ray_queue_buffer_size = 262144;' "$SRC/variables.h"
	if [ "$VARIANT" = "256lod" ]; then
		sed -i -e "s/lod_distance_8x8x8 = 600.000;/lod_distance_8x8x8 = 300;/" -e "s/lod_distance_2x2x2 = 100.000;/lod_distance_2x2x2 = 60;/" "$SRC/variables.h"
		grep -q 'lod_distance_8x8x8 = 300;' "$SRC/variables.h" && grep -q 'lod_distance_2x2x2 = 60;' "$SRC/variables.h"
	fi
fi
OBJ=$TMP/obj; mkdir -p "$OBJ" "$(dirname "$OUT")"
if [ "$KIND" = "ref" ]; then
	$NVCC $NVFLAGS -I"$SRC" -c "$HERE/ref_harness.cu" -o "$OBJ/ref_harness.o"
	for f in kernel.cu sunsky.cu; do $NVCC $NVFLAGS -I"$SRC" -c "$SRC/$f" -o "$OBJ/$f.o"; done
else
	$NVCC $NVFLAGS -DBM_DROPIN -I"$SRC" -c "$HERE/ref_harness.cu" -o "$OBJ/ref_harness.o"
	$NVCC $NVFLAGS -I"$SRC" -I"$HERE/../include" -x cu -c "$HERE/../integration/launch_kernels_dropin.cpp" -o "$OBJ/launch_kernels_dropin.o"
fi
for f in Scene.cpp SimplexNoise.cpp assert_cuda.cpp variables.cpp; do $CXX $CXXFLAGS -I"$SRC" -c "$SRC/$f" -o "$OBJ/$f.o"; done
if [ "$KIND" = "ref" ]; then
	$NVCC $ARCH -shared -o "$OUT" "$OBJ"/*.o
else
	$NVCC $ARCH -shared -o "$OUT" "$OBJ"/*.o -L"$HERE/../brickmap_b200" -lbrickmap_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../brickmap_b200'
fi
