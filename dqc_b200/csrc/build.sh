#!/bin/bash
# Compiles the CUDA library in-tree: dqc_b200/libb200qc.so (sm_100a only).  Translation units compiled in parallel
# (b200qc.cu: integrals and fp64 kernels; b200qc_tc.cu: tcgen05 kernels; b200qc_jk.cu x JKR_NUNITS: the unrolled
# J/K quartet kernels); objects are cached in csrc/build/.
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
# the register-resident J/K engine: JKR_NUNITS small translation units (csrc/jk_reg_units.inc, tools/gen_jk_units.py)
NUNITS=$(sed -n 's/^#define JKR_NUNITS \([0-9]*\)$/\1/p' jk_reg_units.inc)
JKOBJS=""
JKTODO=""
for u in $(seq 0 $((NUNITS - 1))); do
    JKOBJS="$JKOBJS build/b200qc_jk_$u.o"
    if newer build/b200qc_jk_$u.o b200qc_jk.cu common.cuh jk_reg.cuh jk_reg_units.inc ../../include/b200qc.h build.sh; then
        JKTODO="$JKTODO $u"
    fi
done
rm -f build/b200qc_jk_*.failed
if [ -n "$JKTODO" ]; then
    ( echo $JKTODO | tr ' ' '\n' | xargs -P ${JOBS:-$(nproc)} -I{} sh -c \
        "$NVCC $FLAGS -DJKR_UNIT={} -c -o build/b200qc_jk_{}.o b200qc_jk.cu || touch build/b200qc_jk_{}.failed" ) & PIDS="$PIDS $!"
fi
for p in $PIDS; do wait $p; done
if ls build/b200qc_jk_*.failed >/dev/null 2>&1; then echo "J/K unit compile failed" >&2; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o ../libb200qc.so build/b200qc.o build/b200qc_tc.o $JKOBJS
echo "built $(realpath ../libb200qc.so)"
