#!/bin/bash
# Compiles the CUDA library in-tree: dqc_b200/libb200qc.so (sm_100a only).  Two translation units compiled
# in parallel (b200qc.cu is the slow one: the Rys integral classes); objects are cached in csrc/build/.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ $EXTRA_NVCC_FLAGS"
mkdir -p build
newer() {  # newer <object> <sources...>: 0 when the object is older than any source (or missing)
    local o=$1; shift
    [ -f "$o" ] || return 0
    for f in "$@"; do [ "$f" -nt "$o" ] && return 0; done
    return 1
}
PIDS=""
if newer build/b200qc.o b200qc.cu common.cuh tables.cuh ao_eval.cuh becke.cuh xc.cuh gemm_f64.cuh rho.cuh vxc.cuh \
        sb_common.cuh xc_sb.cuh rys.cuh ints.cuh jk.cuh dfj.cuh ../../include/b200qc.h build.sh; then
    $NVCC $FLAGS -c -o build/b200qc.o b200qc.cu & PIDS="$PIDS $!"
fi
if newer build/b200qc_tc.o b200qc_tc.cu common.cuh sb_common.cuh vxc_i8.cuh rho_i8.cuh rho_i8_ps.cuh gemm_i8.cuh peak_i8.cuh \
        ../../include/b200qc.h build.sh; then
    $NVCC $FLAGS -c -o build/b200qc_tc.o b200qc_tc.cu & PIDS="$PIDS $!"
fi
if newer build/b200qc_jk.o b200qc_jk.cu common.cuh jk_reg.cuh ../../include/b200qc.h build.sh; then
    $NVCC $FLAGS -c -o build/b200qc_jk.o b200qc_jk.cu & PIDS="$PIDS $!"
fi
for p in $PIDS; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o ../libb200qc.so build/b200qc.o build/b200qc_tc.o build/b200qc_jk.o
echo "built $(realpath ../libb200qc.so)"
