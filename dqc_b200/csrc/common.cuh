// Shared declarations for the b200qc CUDA library (sm_100a only).  The library is a single
// translation unit (b200qc.cu includes every *.cuh), so __constant__ tables need no -rdc.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/b200qc.h"

#define ATM_SLOTS 6
#define BAS_SLOTS 8
#define B200QC_LMAX 4                       // highest angular momentum handled (g)
#define NCART(l) (((l) + 1) * ((l) + 2) / 2)
#define NUM_SMS 148

// ---- error plumbing ----------------------------------------------------------------------
// The library is two translation units (b200qc.cu: integrals, fp64 kernels, tables, plumbing; b200qc_tc.cu: the
// tcgen05 kernels, -DB200QC_TU_SECONDARY).  Host state shared by both is defined once, in the primary one.
#define QC_HIDDEN __attribute__((visibility("hidden")))
#ifdef B200QC_TU_SECONDARY
#define QC_SHARED(decl, init) extern QC_HIDDEN decl
#else
#define QC_SHARED(decl, init) QC_HIDDEN decl init
#endif
QC_HIDDEN void b200qc_set_error(const std::string &msg);
QC_SHARED(int64_t g_launch_count, = 0);

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
    int device = -1;   // the CUDA device its arrays (and the constant tables uploaded with it) live on
    int cart = 0;      // 1: the integral kernels return RAW CARTESIAN blocks x^a y^b z^c sum_p c_p e^(-a_p r^2) for every
                       // shell of this basis (no real-spherical transform; ao_loc counts (l+1)(l+2)/2 per shell) -- the
                       // form derivative integrals are assembled from (b200qc_basis_set_cartesian)
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
static bool g_rys_ready[64] = {};   // per device (tables.cuh)

static inline cudaStream_t as_stream(void *s) { return (cudaStream_t)s; }

// ---- direct J/K, register-resident quartet engine (jk_reg.cuh, third translation unit b200qc_jk.cu) ----
// The plan (jk.cuh) hands it shell-pair records with their primitive-pair data precomputed once per geometry.
struct JKPair {            // one shell pair, canonical order l(i) >= l(j)
    double ax, ay, az;     // centre of shell i
    double abx, aby, abz;  // A - B
    int ish, jsh;          // shell indices (symmetry factors only)
    int ao_i, ao_j;        // first AO of each shell relative to the plan's slice
    int npp, pp_off;       // primitive pairs [pp_off, pp_off + npp) of the plan's JKPrim array
};
struct JKPrim {            // one primitive pair (a_i, a_j) of a shell pair
    double p, hp;          // a_i + a_j, 1 / (2 p)
    double px, py, pz;     // P = (a_i A + a_j B) / p
    double c;              // c_i c_j exp(-a_i a_j |AB|^2 / p) / p  x  the s / p real-spherical constants of both shells
};
#define JKR_MAXROOTS 5
struct JKRArgs {
    int l[4];                     // (li lj | lk ll), li >= lj, lk >= ll
    const JKPair *bra, *ket;      // class slices of the plan's pair records
    const JKPrim *prims;
    const int2 *items;            // work items: (bra index, first ket index); kets [k0, min(k0 + JKR_CHUNK, nket_of_bra))
    const int *nket_of_bra;
    int64_t nitems;
    int item0, item_stride;       // this rank digests items item0, item0 + item_stride, ...
    const double *dm;             // (nao, nao) symmetric
    double *vj, *vk;              // (nao, nao) accumulators or NULL
    int nao;
    const double *rys_coef;       // device, (nint, 2 nroots, deg + 1) of this class's root count
    int rys_nint, rys_deg;
    double rys_h, rys_xmax;
    double herm_u[JKR_MAXROOTS], herm_w[JKR_MAXROOTS];   // large-x rule: u_r = herm_u[r] / x, w_r = herm_w[r] / sqrt(x)
    double c2s_d[30];             // cart -> real-spherical matrix of d shells, [m][c] (tables.cuh)
};
#define JKR_CHUNK 128      // kets per work item (four per lane)
#define JKR_MAXPP 81       // primitive pairs per shell pair the engine stages (9 x 9: the contracted s shells of cc-pVDZ)
QC_HIDDEN int jkr_supported(const int l[4]);
QC_HIDDEN int jkr_launch(const JKRArgs &A, cudaStream_t st);

// ---- optional per-kernel timing (bench.py's roofline numbers) ------------------------------
// When enabled every major launch is bracketed by CUDA events on its own stream; reading the
// profile synchronises once and returns, per kernel id, launch count and summed device time.
enum {
    PROF_AO_EVAL = 0, PROF_BECKE, PROF_RHO, PROF_XC, PROF_VXC_VB, PROF_VXC_GEMM, PROF_VXC_REDUCE,
    PROF_DFJ_PASS1, PROF_DFJ_PASS2, PROF_DFJ_SMALL, PROF_JK, PROF_INTS, PROF_PEAK, PROF_SB_GATHER, PROF_GEMV, PROF_I8_SLICE, PROF_GEMM_I8, PROF_N
};
static const char *const g_prof_names[PROF_N] = {
    "ao_eval_kernel", "becke_weights_kernel", "rho_kernel", "xc_kernel", "vxc_vb_kernel", "vxc_gemm_kernel",
    "slab_reduce_kernel", "dfj_pass1_kernel", "dfj_pass2_kernel", "dfj_small_kernels", "jk_kernel",
    "int_dense_kernel", "dmma_peak_kernel", "sb_gather_dm_kernel", "gemv_rows_kernel", "sb_slice_kernel",
    "gemm_i8_kernel"};
struct ProfRec {
    int id;
    cudaEvent_t a, b;
};
QC_SHARED(bool g_prof_on, = false);
QC_SHARED(std::vector<ProfRec> g_prof, );
static inline void prof_begin(int id, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfRec r;
    r.id = id;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
}
static inline void prof_end(cudaStream_t st) {
    if (!g_prof_on || g_prof.empty()) return;
    cudaEventRecord(g_prof.back().b, st);
}
