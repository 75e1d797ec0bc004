#!/bin/bash
# Compiles the reference's own PatchMatch CUDA sources from where they lie under /root/reference (never copied)
# against the stub headers in oracle/ref_stubs/, with the reference's flags (--use_fast_math, per-thread default
# stream) and, like upstream on Blackwell (mvs/CMakeLists.txt:148-276), as compute_90 PTX that the driver JITs.
# Output: oracle/_ref/libpm_ref.so (git-ignored; travels to the GPU box).  Exit 0 and do nothing if /root/reference
# is absent (GPU box) — the prebuilt file is used there.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference/src
[ -d "$REF" ] || { echo "no /root/reference: keeping prebuilt oracle/_ref"; exit 0; }
mkdir -p "$HERE/_ref"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
"$NVCC" -gencode arch=compute_90,code=compute_90 -O3 -std=c++17 --use_fast_math --default-stream per-thread \
  -DCOLMAP_CUDA_ENABLED -Xcompiler -fPIC -w -shared -I"$HERE/ref_stubs" -I"$REF" \
  -o "$HERE/_ref/libpm_ref.so" \
  "$HERE/ref_harness.cu" "$REF/colmap/mvs/patch_match_cuda.cu" "$REF/colmap/mvs/gpu_mat_prng.cu" \
  "$REF/colmap/mvs/gpu_mat_ref_image.cu" "$REF/colmap/util/cudacc.cc" "$REF/colmap/util/cuda.cc"
echo "built $HERE/_ref/libpm_ref.so"
