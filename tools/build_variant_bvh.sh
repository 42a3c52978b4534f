#!/bin/bash
# A/B variant of libvkrt_cuda.so with extra -D flags for bvh_build.cu: tools/build_variant_bvh.sh NAME "-DVKRT_PLOC_RADIUS=25"
# -> variants/NAME/libvkrt_cuda.so (git-ignored; selected at run time with VKRT_CUDA_LIB=variants/NAME/libvkrt_cuda.so)
set -e
cd "$(dirname "$0")/../vkrt_b200"
NAME=$1; shift
D=../variants/$NAME
mkdir -p $D
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC,-fvisibility=hidden,-Wall -Xptxas -v --expt-relaxed-constexpr -rdc=false "$@" -c csrc/bvh_build.cu -o $D/bvh_build.o 2> $D/bvh_build.ptxas.log
nvcc $ARCH -shared -o $D/libvkrt_cuda.so build/api.o $D/bvh_build.o build/wavefront.o -ldl
