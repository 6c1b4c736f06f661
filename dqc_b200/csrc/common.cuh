// Shared declarations for the b200qc CUDA library (sm_100a only).  The library is a single
// translation unit (b200qc.cu includes every *.cuh), so __constant__ tables need no -rdc.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/b200qc.h"

#define ATM_SLOTS 6
#define BAS_SLOTS 8
#define B200QC_LMAX 4                       // highest angular momentum handled (g)
#define NCART(l) (((l) + 1) * ((l) + 2) / 2)
#define NUM_SMS 148

// ---- error plumbing ----------------------------------------------------------------------
static void b200qc_set_error(const std::string &msg);
static int64_t g_launch_count = 0;

#define QC_CHECK(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            char buf__[512];                                                                    \
            snprintf(buf__, sizeof(buf__), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,        \
                     cudaGetErrorString(e__));                                                  \
            b200qc_set_error(buf__);                                                            \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

#define QC_REQUIRE(cond, msg)                                                                   \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            b200qc_set_error(std::string(__func__) + ": " + (msg));                             \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

#define QC_LAUNCHED(n)                                                                          \
    do {                                                                                        \
        g_launch_count += (n);                                                                  \
        QC_CHECK(cudaGetLastError());                                                           \
    } while (0)

// ---- device-resident basis ---------------------------------------------------------------
// Per-shell record unpacked from (atm, bas, env) so kernels do one 48-byte read per shell.
struct ShellRec {
    double x, y, z;
    int l, nprim, ptr_exp, ptr_coef;  // env offsets
    int ao_off, atom;                 // first spherical AO of the shell (absolute), owning atom
};

struct b200qc_basis {
    int natm, nbas, nenv;
    std::vector<int> h_atm, h_bas, h_ao_loc;
    std::vector<double> h_env;
    std::vector<ShellRec> h_shells;
    int *d_atm = nullptr, *d_bas = nullptr, *d_ao_loc = nullptr;
    double *d_env = nullptr;
    ShellRec *d_shells = nullptr;
};

// cart -> real-spherical coefficient tables (libcint convention), one per l, in constant memory.
// c2s_l[m * ncart + c]; filled once by qc_init_tables() from the same closed formula the oracle
// uses (independent code, cross-checked by tests).
struct C2STables {
    double s0[1];
    double s1[9];
    double s2[5 * 6];
    double s3[7 * 10];
    double s4[9 * 15];
};
static __constant__ C2STables c_c2s;
static __constant__ signed char c_cart_pow[B200QC_LMAX + 1][15][3];  // (lx, ly, lz) of each cartesian
static int qc_init_tables();  // idempotent; returns 0 on success

__device__ __forceinline__ const double *c2s_ptr(int l) {
    switch (l) {
        case 0: return c_c2s.s0;
        case 1: return c_c2s.s1;
        case 2: return c_c2s.s2;
        case 3: return c_c2s.s3;
        default: return c_c2s.s4;
    }
}

// ---- Rys table ---------------------------------------------------------------------------
#define RYS_NMAX 7
struct RysTable {
    int nmax, deg, nint;
    double h, xmax;
    const double *coef[RYS_NMAX];  // device, (nint, 2n, deg+1)
    double herm[RYS_NMAX][2][RYS_NMAX];
};
static __constant__ RysTable c_rys;
static bool g_rys_ready = false;

static inline cudaStream_t as_stream(void *s) { return (cudaStream_t)s; }
