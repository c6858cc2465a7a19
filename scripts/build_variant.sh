#!/bin/bash
# Developer helper: build a kernel variant of libsrt.so (production-mode instantiations only) for A/B timing.
#   scripts/build_variant.sh NAME [-DFOO=1 ...]   ->  build/variants/libsrt_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC -shared \
     -DSRT_DEV_MINIMAL "$@" -o build/variants/libsrt_$name.so spectral_raytracer_b200/csrc/srt_api.cu 2>&1 | grep -v "warning #177\|V3 sub\|\^\|^$\|Remark" || true
