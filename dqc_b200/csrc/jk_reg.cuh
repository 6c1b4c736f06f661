// K8r -- direct Coulomb / exchange build, register-resident quartet engine.  Same contract as jk_kernel (jk.cuh):
//   J_kl = sum_ij D_ij (ij|kl)            dqc/hamilton/hcgto.py:204-211
//   K_jk = sum_il D_il (ij|kl)            dqc/hamilton/hcgto.py:224-241
// from unique shell quartets that are digested as they are produced (the reference materialises nao^4 doubles).
//
// Work decomposition.  One launch = one class pair (li lj | lk ll) x (primitive-count buckets), so control flow is
// warp-uniform.  A WARP owns one bra shell pair and a chunk of JKR_CHUNK ket pairs of the Schwarz-sorted ket list;
// a LANE owns one contracted quartet at a time, entirely in registers:
//   * the bra pair's primitive data (p, 1/2p, P, c_i c_j e^(-a_i a_j |AB|^2 / p) / p: precomputed once per geometry by
//     the plan) and its density tile D_ij are staged in shared memory by the warp and read as broadcasts; the ket pair's
//     primitive data come from the same plan array, one pair per lane;
//   * Rys roots and weights by Clenshaw summation of the piecewise Chebyshev table (rys.cuh), which the CTA keeps in
//     shared memory TRANSPOSED -- [coefficient][function][interval], so lanes in different intervals hit different
//     banks (from global memory every lane would touch its own cache line: 2 n x 14 wavefronts per primitive quartet);
//   * 2-D recurrence tables (vertical + both horizontal transfers) and the sum over roots are compile-time unrolled
//     per angular-momentum class -- every table entry and every cartesian component is a register;
//   * digestion: the lane contracts its block with the six density tiles; the bra-tile J_ij stays in registers across
//     the lane's kets, is summed over the warp by shuffles and leaves as ONE atomic per element per work item; the
//     ket-side J_kl and the four K tiles go out with fp64 atomics (RED.E.ADD.F64.STRONG.GPU).
// Shells with l <= 1 only: their real-spherical transform is a constant per shell that the plan folds into the
// primitive coefficients, so the cartesian block IS the spherical one.  Classes with d shells and above stay on the
// shared-memory engine of jk.cuh.
#pragma once
#include "common.cuh"
#include <utility>

#define JKR_THREADS 128
#define JKR_WARPS (JKR_THREADS / 32)
#define JKR_NINT 64
#define JKR_NCOEF 14
#define JKR_WSM (JKR_MAXPP * 6 + 40)   // doubles of shared memory per warp: bra primitive pairs + D_ij tile

namespace jkr {

__host__ __device__ constexpr int ncart(int l) { return (l + 1) * (l + 2) / 2; }

// power of x (d = 0), y (1) or z (2) in cartesian component c of a shell (libcint order: lx descending, then ly)
__host__ __device__ constexpr int cart_pow(int l, int c, int d) {
    int k = 0;
    for (int x = l; x >= 0; x--)
        for (int y = l - x; y >= 0; y--) {
            if (k == c) return d == 0 ? x : (d == 1 ? y : l - x - y);
            k++;
        }
    return 0;
}

template <int LI, int LJ, int LK, int LL>
struct Shape {
    static constexpr int NI = ncart(LI), NJ = ncart(LJ), NK = ncart(LK), NL = ncart(LL);
    static constexpr int NIJ = NI * NJ, NKL = NK * NL, NC = NIJ * NKL;
    static constexpr int NR = (LI + LJ + LK + LL) / 2 + 1;
    static constexpr int TS = (LI + 1) * (LJ + 1) * (LK + 1) * (LL + 1);   // entries of one 2-D table
    // index of component c's factor in the table of dimension d
    __host__ __device__ static constexpr int comp_index(int c, int d) {
        const int cl = c % NL, ck = (c / NL) % NK, cj = (c / NKL) % NJ, ci = c / (NKL * NJ);
        return ((cart_pow(LI, ci, d) * (LJ + 1) + cart_pow(LJ, cj, d)) * (LK + 1) + cart_pow(LK, ck, d)) * (LL + 1) +
               cart_pow(LL, cl, d);
    }
};

// One (root, dimension) table T[i][j][k][l] = I(i, j, k, l): vertical recurrence on (n, m), then the horizontal
// transfers n -> (i, j) with AB and m -> (k, l) with CD.
template <int LI, int LJ, int LK, int LL>
__device__ __forceinline__ void build_table(double (&T)[Shape<LI, LJ, LK, LL>::TS], double w0, double c00, double c01,
                                            double b10, double b01, double b00, double ab, double cd) {
    constexpr int NIJ = LI + LJ + 1, NKL = LK + LL + 1;
    double W[NIJ][NKL];
    W[0][0] = w0;
#pragma unroll
    for (int n = 0; n < NIJ - 1; n++) {
        double v = c00 * W[n][0];
        if (n > 0) v += (n * b10) * W[n - 1][0];
        W[n + 1][0] = v;
    }
#pragma unroll
    for (int m = 0; m < NKL - 1; m++) {
#pragma unroll
        for (int n = 0; n < NIJ; n++) {
            double v = c01 * W[n][m];
            if (m > 0) v += (m * b01) * W[n][m - 1];
            if (n > 0) v += (n * b00) * W[n - 1][m];
            W[n][m + 1] = v;
        }
    }
#pragma unroll
    for (int j = 0; j <= LJ; j++) {
        if (j > 0) {
#pragma unroll
            for (int n = 0; n < NIJ - j; n++) {
#pragma unroll
                for (int m = 0; m < NKL; m++) W[n][m] = W[n + 1][m] + ab * W[n][m];
            }
        }
#pragma unroll
        for (int i = 0; i <= LI; i++) {
            double line[NKL];
#pragma unroll
            for (int m = 0; m < NKL; m++) line[m] = W[i][m];
#pragma unroll
            for (int l = 0; l <= LL; l++) {
                if (l > 0) {
#pragma unroll
                    for (int m = 0; m < NKL - l; m++) line[m] = line[m + 1] + cd * line[m];
                }
#pragma unroll
                for (int k = 0; k <= LK; k++) T[((i * (LJ + 1) + j) * (LK + 1) + k) * (LL + 1) + l] = line[k];
            }
        }
    }
}

template <class S, int C>
__device__ __forceinline__ void acc_one(double (&acc)[S::NC], const double (&Tx)[S::TS], const double (&Ty)[S::TS],
                                        const double (&Tz)[S::TS]) {
    constexpr int ix = S::comp_index(C, 0), iy = S::comp_index(C, 1), iz = S::comp_index(C, 2);
    acc[C] += Tx[ix] * Ty[iy] * Tz[iz];
}
template <class S, int... Cs>
__device__ __forceinline__ void acc_all(double (&acc)[S::NC], const double (&Tx)[S::TS], const double (&Ty)[S::TS],
                                        const double (&Tz)[S::TS], std::integer_sequence<int, Cs...>) {
    (acc_one<S, Cs>(acc, Tx, Ty, Tz), ...);
}

// nodes u_r = t_r^2 and weights of the NR-point Rys rule at x: same arithmetic as rys_eval (rys.cuh), all 2 NR
// Clenshaw recurrences interleaved, table rt[(k * 2 NR + f) * JKR_NINT + interval] in shared memory
template <int NR>
__device__ __forceinline__ void rys_roots(double x, const double *__restrict__ rt, const JKRArgs &A, double ih,
                                          double (&u)[NR], double (&w)[NR]) {
    if (x >= A.rys_xmax) {
        const double s = 1.0 / x, rs = sqrt(s);
#pragma unroll
        for (int r = 0; r < NR; r++) {
            u[r] = A.herm_u[r] * s;
            w[r] = A.herm_w[r] * rs;
        }
        return;
    }
    int it = (int)(x * ih);   // ih = 1 / h; the shipped table has h = 1, where this is the same arithmetic as rys_eval
    if (it > JKR_NINT - 1) it = JKR_NINT - 1;
    const double t = 2.0 * (x - it * A.rys_h) * ih - 1.0;
    const double t2 = 2.0 * t;
    const double *c = rt + it;
    double b1[2 * NR], b2[2 * NR];
#pragma unroll
    for (int f = 0; f < 2 * NR; f++) b1[f] = b2[f] = 0.0;
#pragma unroll
    for (int k = JKR_NCOEF - 1; k >= 1; k--) {
#pragma unroll
        for (int f = 0; f < 2 * NR; f++) {
            const double b0 = c[(k * 2 * NR + f) * JKR_NINT] + t2 * b1[f] - b2[f];
            b2[f] = b1[f];
            b1[f] = b0;
        }
    }
#pragma unroll
    for (int f = 0; f < 2 * NR; f++) {
        const double v = c[f * JKR_NINT] + t * b1[f] - b2[f];
        if (f < NR) u[f] = v; else w[f - NR] = v;
    }
}

template <int LI, int LJ, int LK, int LL, bool DOJ, bool DOK>
__global__ void __launch_bounds__(JKR_THREADS) jk_reg_kernel(const JKRArgs A) {
    using S = Shape<LI, LJ, LK, LL>;
    constexpr int NI = S::NI, NJ = S::NJ, NK = S::NK, NL = S::NL, NR = S::NR, NC = S::NC, NKL = S::NKL, NIJ = S::NIJ;
    extern __shared__ __align__(16) double jkr_smem[];
    double *rt = jkr_smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *wprim = jkr_smem + JKR_NCOEF * 2 * NR * JKR_NINT + warp * JKR_WSM;
    double *wdij = wprim + JKR_MAXPP * 6;
    for (int e = threadIdx.x; e < JKR_NINT * 2 * NR * JKR_NCOEF; e += JKR_THREADS) {
        const int it = e / (2 * NR * JKR_NCOEF), rem = e - it * (2 * NR * JKR_NCOEF);
        const int f = rem / JKR_NCOEF, k = rem - f * JKR_NCOEF;
        rt[(k * 2 * NR + f) * JKR_NINT + it] = A.rys_coef[e];
    }
    __syncthreads();
    const int nao = A.nao;
    const double ih = 1.0 / A.rys_h;
    const double *__restrict__ D = A.dm;
    const int64_t wstride = (int64_t)gridDim.x * JKR_WARPS;
    for (int64_t slot = (int64_t)blockIdx.x * JKR_WARPS + warp;; slot += wstride) {
        const int64_t item = A.item0 + slot * A.item_stride;
        if (item >= A.nitems) break;
        const int2 wi = A.items[item];
        const int k0 = wi.y, k1 = min(k0 + JKR_CHUNK, A.nket_of_bra[wi.x]);
        const JKPair bp = A.bra[wi.x];
        __syncwarp();
        {
            const double *src = reinterpret_cast<const double *>(A.prims + bp.pp_off);
            for (int e = lane; e < bp.npp * 6; e += 32) wprim[e] = src[e];
            for (int e = lane; e < NIJ; e += 32) wdij[e] = D[(int64_t)(bp.ao_i + e / NJ) * nao + bp.ao_j + e % NJ];
        }
        __syncwarp();
        const int ai = bp.ao_i, aj = bp.ao_j;
        double jij[NIJ];
#pragma unroll
        for (int e = 0; e < NIJ; e++) jij[e] = 0.0;
        for (int kk0 = k0; kk0 < k1; kk0 += 32) {
            const int kk = kk0 + lane;
            if (kk < k1) {
                const JKPair kp = A.ket[kk];
                double acc[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) acc[c] = 0.0;
                const JKPrim *__restrict__ kpr = A.prims + kp.pp_off;
                for (int pk = 0; pk < kp.npp; pk++) {
                    const JKPrim kq = kpr[pk];
                    const double qcx = kq.px - kp.ax, qcy = kq.py - kp.ay, qcz = kq.pz - kp.az;
                    for (int pb = 0; pb < bp.npp; pb++) {
                        const double p = wprim[pb * 6], hp = wprim[pb * 6 + 1];
                        const double px = wprim[pb * 6 + 2], py = wprim[pb * 6 + 3], pz = wprim[pb * 6 + 4];
                        const double pq = p + kq.p;
                        const double rs = rsqrt(pq);
                        const double ipq = rs * rs;
                        const double dx = px - kq.px, dy = py - kq.py, dz = pz - kq.pz;
                        const double x = p * kq.p * ipq * (dx * dx + dy * dy + dz * dz);
                        const double pref = wprim[pb * 6 + 5] * kq.c * 34.98683665524972497 /* 2 pi^2.5 */ * rs;
                        const double a0 = kq.p * ipq, a1 = p * ipq;
                        double u[NR], w[NR];
                        rys_roots<NR>(x, rt, A, ih, u, w);
                        const double pax = px - bp.ax, pay = py - bp.ay, paz = pz - bp.az;
#pragma unroll
                        for (int r = 0; r < NR; r++) {
                            const double a0u = a0 * u[r], a1u = a1 * u[r];
                            const double b10 = (1.0 - a0u) * hp, b01 = (1.0 - a1u) * kq.hp, b00 = 0.5 * u[r] * ipq;
                            double Tx[S::TS], Ty[S::TS], Tz[S::TS];
                            build_table<LI, LJ, LK, LL>(Tx, 1.0, pax - a0u * dx, qcx + a1u * dx, b10, b01, b00, bp.abx,
                                                        kp.abx);
                            build_table<LI, LJ, LK, LL>(Ty, 1.0, pay - a0u * dy, qcy + a1u * dy, b10, b01, b00, bp.aby,
                                                        kp.aby);
                            build_table<LI, LJ, LK, LL>(Tz, w[r] * pref, paz - a0u * dz, qcz + a1u * dz, b10, b01, b00,
                                                        bp.abz, kp.abz);
                            acc_all<S>(acc, Tx, Ty, Tz, std::make_integer_sequence<int, NC>());
                        }
                    }
                }
                // ---- digestion (same sums as jk_kernel; D symmetric, the caller adds the transpose at the end) ----
                double f = 1.0;
                if (bp.ish == bp.jsh) f *= 0.5;
                if (kp.ish == kp.jsh) f *= 0.5;
                if (bp.ish == kp.ish && bp.jsh == kp.jsh) f *= 0.5;
                const int ak = kp.ao_i, al = kp.ao_j;
                if (DOJ) {
                    double *__restrict__ J = A.vj;
                    double t[NKL];
#pragma unroll
                    for (int c = 0; c < NKL; c++) t[c] = D[(int64_t)(ak + c / NL) * nao + al + c % NL];
                    const double f2 = 2.0 * f;
#pragma unroll
                    for (int e = 0; e < NIJ; e++) {
                        double v = 0.0;
#pragma unroll
                        for (int c = 0; c < NKL; c++) v += acc[e * NKL + c] * t[c];
                        jij[e] += f2 * v;
                    }
#pragma unroll
                    for (int c = 0; c < NKL; c++) {
                        double v = 0.0;
#pragma unroll
                        for (int e = 0; e < NIJ; e++) v += acc[e * NKL + c] * wdij[e];
                        atomicAdd(J + (int64_t)(ak + c / NL) * nao + al + c % NL, f2 * v);
                    }
                }
                if (DOK) {
                    double *__restrict__ Kx = A.vk;
                    {   // K_ik += f sum_jl B D_jl
                        double t[NJ * NL];
#pragma unroll
                        for (int e = 0; e < NJ * NL; e++) t[e] = D[(int64_t)(aj + e / NL) * nao + al + e % NL];
#pragma unroll
                        for (int a = 0; a < NI; a++) {
#pragma unroll
                            for (int c = 0; c < NK; c++) {
                                double v = 0.0;
#pragma unroll
                                for (int b = 0; b < NJ; b++) {
#pragma unroll
                                    for (int d = 0; d < NL; d++) v += acc[((a * NJ + b) * NK + c) * NL + d] * t[b * NL + d];
                                }
                                atomicAdd(Kx + (int64_t)(ai + a) * nao + ak + c, f * v);
                            }
                        }
                    }
                    {   // K_jk += f sum_il B D_il
                        double t[NI * NL];
#pragma unroll
                        for (int e = 0; e < NI * NL; e++) t[e] = D[(int64_t)(ai + e / NL) * nao + al + e % NL];
#pragma unroll
                        for (int b = 0; b < NJ; b++) {
#pragma unroll
                            for (int c = 0; c < NK; c++) {
                                double v = 0.0;
#pragma unroll
                                for (int a = 0; a < NI; a++) {
#pragma unroll
                                    for (int d = 0; d < NL; d++) v += acc[((a * NJ + b) * NK + c) * NL + d] * t[a * NL + d];
                                }
                                atomicAdd(Kx + (int64_t)(aj + b) * nao + ak + c, f * v);
                            }
                        }
                    }
                    {   // K_il += f sum_jk B D_jk
                        double t[NJ * NK];
#pragma unroll
                        for (int e = 0; e < NJ * NK; e++) t[e] = D[(int64_t)(aj + e / NK) * nao + ak + e % NK];
#pragma unroll
                        for (int a = 0; a < NI; a++) {
#pragma unroll
                            for (int d = 0; d < NL; d++) {
                                double v = 0.0;
#pragma unroll
                                for (int b = 0; b < NJ; b++) {
#pragma unroll
                                    for (int c = 0; c < NK; c++) v += acc[((a * NJ + b) * NK + c) * NL + d] * t[b * NK + c];
                                }
                                atomicAdd(Kx + (int64_t)(ai + a) * nao + al + d, f * v);
                            }
                        }
                    }
                    {   // K_jl += f sum_ik B D_ik
                        double t[NI * NK];
#pragma unroll
                        for (int e = 0; e < NI * NK; e++) t[e] = D[(int64_t)(ai + e / NK) * nao + ak + e % NK];
#pragma unroll
                        for (int b = 0; b < NJ; b++) {
#pragma unroll
                            for (int d = 0; d < NL; d++) {
                                double v = 0.0;
#pragma unroll
                                for (int a = 0; a < NI; a++) {
#pragma unroll
                                    for (int c = 0; c < NK; c++) v += acc[((a * NJ + b) * NK + c) * NL + d] * t[a * NK + c];
                                }
                                atomicAdd(Kx + (int64_t)(aj + b) * nao + al + d, f * v);
                            }
                        }
                    }
                }
            }
        }
        if (DOJ) {
            // J_ij of the bra tile: one sum over the warp, one atomic per element per work item
#pragma unroll
            for (int e = 0; e < NIJ; e++) {
                double v = jij[e];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == (e & 31)) atomicAdd(A.vj + (int64_t)(ai + e / NJ) * nao + aj + e % NJ, v);
            }
        }
    }
}

template <int LI, int LJ, int LK, int LL, bool DOJ, bool DOK>
static int launch_t(const JKRArgs &A, cudaStream_t st) {
    using S = Shape<LI, LJ, LK, LL>;
    static_assert(S::NR <= JKR_MAXROOTS && S::NIJ <= 40, "class outside the staged sizes");
    const size_t smem = sizeof(double) * ((size_t)JKR_NCOEF * 2 * S::NR * JKR_NINT + (size_t)JKR_WARPS * JKR_WSM);
    auto kern = jk_reg_kernel<LI, LJ, LK, LL, DOJ, DOK>;
    QC_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    QC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, JKR_THREADS, smem));
    QC_REQUIRE(occ >= 1, "J/K register kernel does not fit on an SM");
    const int64_t mine = (A.nitems - A.item0 + A.item_stride - 1) / A.item_stride;   // items item0, item0 + stride, ...
    if (mine <= 0) return 0;
    const int64_t nblk = std::min<int64_t>((mine + JKR_WARPS - 1) / JKR_WARPS, (int64_t)NUM_SMS * occ);
    prof_begin(PROF_JK, st);
    kern<<<(unsigned)nblk, JKR_THREADS, smem, st>>>(A);
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

template <int LI, int LJ, int LK, int LL>
static int launch_mode(const JKRArgs &A, cudaStream_t st) {
    if (A.vj && A.vk) return launch_t<LI, LJ, LK, LL, true, true>(A, st);
    if (A.vj) return launch_t<LI, LJ, LK, LL, true, false>(A, st);
    if (A.vk) return launch_t<LI, LJ, LK, LL, false, true>(A, st);
    return 0;
}

}  // namespace jkr

// classes the engine is instantiated for: li >= lj, lk >= ll, (li, lj) >= (lk, ll), all l <= 1
#define JKR_CLASSES(X) X(0, 0, 0, 0) X(1, 0, 0, 0) X(1, 0, 1, 0) X(1, 1, 0, 0) X(1, 1, 1, 0) X(1, 1, 1, 1)

int jkr_supported(const int l[4]) {
#define JKR_X(a, b, c, d) if (l[0] == a && l[1] == b && l[2] == c && l[3] == d) return 1;
    JKR_CLASSES(JKR_X)
#undef JKR_X
    return 0;
}

int jkr_launch(const JKRArgs &A, cudaStream_t st) {
    QC_REQUIRE(A.rys_nint == JKR_NINT && A.rys_deg + 1 == JKR_NCOEF, "Rys table shape differs from the compiled one");
#define JKR_X(a, b, c, d) \
    if (A.l[0] == a && A.l[1] == b && A.l[2] == c && A.l[3] == d) return jkr::launch_mode<a, b, c, d>(A, st);
    JKR_CLASSES(JKR_X)
#undef JKR_X
    b200qc_set_error("jkr_launch: class not instantiated");
    return 2;
}
