// b200qc -- primary translation unit of the B200-native Fock-build library (see include/b200qc.h): integrals,
// fp64 kernels, constant tables, shared host state.  The tcgen05 kernels are in b200qc_tc.cu.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include "common.cuh"
#include "tables.cuh"
#include "ao_eval.cuh"
#include "becke.cuh"
#include "xc.cuh"
#include "gemm_f64.cuh"
#include "rho.cuh"
#include "vxc.cuh"
#include "xc_sb.cuh"
#include "rys.cuh"
#include "ints.cuh"
#include "jk.cuh"
#include "dfj.cuh"
