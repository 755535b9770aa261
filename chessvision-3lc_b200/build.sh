#!/usr/bin/env bash
# Build libchessvision_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
mkdir -p build
$NVCC $FLAGS -c csrc/conv_tc.cu -o build/conv_tc.o &
$NVCC $FLAGS -c csrc/stem.cu -o build/stem.o &
$NVCC $FLAGS -Xptxas -v -c csrc/stem_tc.cu -o build/stem_tc.o 2> build/stem_tc.ptxas.log &
$NVCC $FLAGS -fmad=false -c csrc/geometry.cu -o build/geometry.o &
$NVCC $FLAGS -c csrc/api.cu -o build/api.o &
$NVCC $FLAGS -c csrc/wgrad_tc.cu -o build/wgrad_tc.o &
$NVCC $FLAGS -c csrc/train_kernels.cu -o build/train_kernels.o &
$NVCC $FLAGS -c csrc/train.cu -o build/train.o &
$NVCC $FLAGS -c csrc/consumers.cu -o build/consumers.o &
$NVCC $FLAGS -c csrc/train_cls.cu -o build/train_cls.o &
$NVCC $FLAGS -c csrc/jpeg.cu -o build/jpeg.o &
# a failed compile must fail the build (a bare `wait` returns 0 and the link would reuse a stale object)
for job in $(jobs -p); do wait "$job"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o libchessvision_b200.so build/conv_tc.o build/stem.o build/stem_tc.o build/geometry.o build/api.o build/wgrad_tc.o build/train_kernels.o build/train.o build/consumers.o build/train_cls.o build/jpeg.o -lpthread
echo "built $(pwd)/libchessvision_b200.so"
