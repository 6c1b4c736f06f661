// b200qc -- secondary translation unit: the tcgen05 (5th-generation tensor core) kernels.  Shares the host
// state (error string, launch counter, profiler records) defined in b200qc.cu.
#define B200QC_TU_SECONDARY
#include "common.cuh"
#include "sb_common.cuh"
#include "vxc_i8.cuh"
#include "rho_i8.cuh"
#include "rho_i8_ps.cuh"
#include "gemm_i8.cuh"
#include "peak_i8.cuh"
