#!/bin/bash
# Compiles the reference's own GPU bundle-adjustment backend - the generated Caspar solver under
# /root/reference/src/thirdparty/Symforce-Caspar/generated/f32 (241 .cu + solver.cc, compiled from where they lie,
# never copied) - plus oracle/caspar_harness.cu into oracle/_ref/libcaspar_ref.so (git-ignored; travels to the GPU box).
# Flags follow the generated CMakeLists.txt (--use_fast_math, -O3); the architecture is sm_100 SASS instead of upstream's
# sm_75..89 + PTX list so that the baseline is not handicapped by a JIT of 240 kernels on Blackwell.
# Exit 0 and do nothing if /root/reference is absent (GPU box) - the prebuilt file is used there.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference/src
GEN="$REF/thirdparty/Symforce-Caspar/generated/f32"
[ -d "$GEN" ] || { echo "no /root/reference: keeping prebuilt oracle/_ref"; exit 0; }
OBJ="$HERE/_ref/caspar_obj"
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100,code=sm_100 -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC -w -I$GEN -I$REF"
export NVCC FLAGS OBJ
compile_one() {
  src="$1"; base="$(basename "$src")"; obj="$OBJ/${base%.*}.o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then "$NVCC" $FLAGS -x cu -c "$src" -o "$obj"; fi
}
export -f compile_one
ls "$GEN"/*.cu "$GEN"/*.cc | grep -v pybind | xargs -P "$(nproc)" -I{} bash -c 'compile_one {}'
"$NVCC" $FLAGS -c "$HERE/caspar_harness.cu" -o "$OBJ/_harness.o"
"$NVCC" -gencode arch=compute_100,code=sm_100 -shared -o "$HERE/_ref/libcaspar_ref.so" "$OBJ"/*.o
echo "built $HERE/_ref/libcaspar_ref.so"
