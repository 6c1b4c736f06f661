// temporary stubs until the integral kernels land
#pragma once
#include "common.cuh"
#ifndef B200QC_WITH_INTS
extern "C" int b200qc_int1e(const b200qc_basis *, int, const int *, const double *, double *, void *) { b200qc_set_error("not built"); return 3; }
extern "C" int b200qc_int2c2e(const b200qc_basis *, const int *, double *, void *) { b200qc_set_error("not built"); return 3; }
extern "C" int b200qc_int3c2e(const b200qc_basis *, const int *, double *, void *) { b200qc_set_error("not built"); return 3; }
extern "C" int b200qc_int2e(const b200qc_basis *, const int *, double *, void *) { b200qc_set_error("not built"); return 3; }
extern "C" int b200qc_jk_direct(const b200qc_basis *, int, int, const double *, int, double *, double *, void *) { b200qc_set_error("not built"); return 3; }
extern "C" int64_t b200qc_dfj_worksize(int64_t, int64_t) { return 0; }
extern "C" int b200qc_dfj(const double *, int64_t, int64_t, const double *, const double *, double *, double *, void *) { b200qc_set_error("not built"); return 3; }
extern "C" int b200qc_pack_tril(const double *, int64_t, int64_t, double *, void *) { b200qc_set_error("not built"); return 3; }
#endif
