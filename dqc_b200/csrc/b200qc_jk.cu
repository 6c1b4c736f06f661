// b200qc -- third translation unit: the register-resident quartet engine of the direct J/K build (jk_reg.cuh).
// Shares the host state (error string, launch counter, profiler records) defined in b200qc.cu; the plan that
// feeds it lives in jk.cuh (primary unit).
#define B200QC_TU_SECONDARY
#include "common.cuh"
#include "jk_reg.cuh"
