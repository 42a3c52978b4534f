#!/bin/bash
# Builds an A/B variant of libvkrt_cuda.so with extra -D flags for wavefront.cu: tools/build_variant.sh NAME "-DTRACE_SMEM_STACK=12"
# -> variants/NAME/libvkrt_cuda.so (git-ignored; selected at run time with VKRT_CUDA_LIB=variants/NAME/libvkrt_cuda.so)
set -e
cd "$(dirname "$0")/../vkrt_b200"
NAME=$1; shift
D=../variants/$NAME
mkdir -p $D
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC,-fvisibility=hidden,-Wall -Xptxas -v --expt-relaxed-constexpr -rdc=false -use_fast_math -ftz=false "$@" -c csrc/wavefront.cu -o $D/wavefront.o 2> $D/wavefront.ptxas.log
nvcc $ARCH -shared -o $D/libvkrt_cuda.so build/api.o build/bvh_build.o $D/wavefront.o -ldl
grep -A1 "k_trace" $D/wavefront.ptxas.log | grep -E "registers|spill" | head -8
