#!/bin/bash
# Build a tuning variant of the library: tools/build_variant.sh <name> [extra nvcc flags...]  -> variants/lib_<name>.so
# (git-ignored, travels with gpurun; select it at run time with SRUKF_LIB_PATH=variants/lib_<name>.so)
name=$1; shift
cd "$(dirname "$0")/../cv_monoslam_b200/csrc" || exit 1
mkdir -p ../../variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -Xptxas=-v "$@" \
  -o ../../variants/lib_$name.so srukf_kernels.cu srukf_capi.cu > ../../variants/ptxas_$name.log 2>&1
echo "variant $name rc=$?"
