#!/bin/bash
# Compiles the CUDA library in-tree: dqc_b200/libb200qc.so (sm_100a only).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
    -ccbin /usr/bin/g++ $EXTRA_NVCC_FLAGS -o ../libb200qc.so b200qc.cu
echo "built $(realpath ../libb200qc.so)"
