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
//   * Rys roots and weights by Clenshaw summation of a piecewise Chebyshev table -- the refined form of the base table
//     of rys.cuh (tables.cuh, b200qc_rys_refine: intervals of width 1/2, degree 9: 10 steps instead of 14 at the same
//     truncation error) -- which the CTA keeps in shared memory TRANSPOSED, [coefficient][function][interval], so lanes
//     in different intervals hit different banks (from global memory every lane would touch its own cache line: 2 n x
//     10 wavefronts per primitive quartet);
//   * 2-D recurrence tables (vertical + both horizontal transfers) and the sum over roots are compile-time unrolled
//     per angular-momentum class -- every table entry and every cartesian component is a register;
//   * digestion: the lane contracts its block with the six density tiles; the bra-tile J_ij is summed over the warp by
//     shuffles (once per work item for the single-pass classes, whose lanes keep it in registers across their kets; once
//     per round of 32 kets otherwise), collected in the warp's shared memory and leaves as ONE atomic per element per
//     work item; the ket-side J_kl and the four K tiles go out with fp64 atomics (REDG.E.ADD.F64.RN.STRONG.GPU).
// s, p and d shells.  For l <= 1 the real-spherical transform is a constant per shell that the plan folds into the
// primitive coefficients; d shells are digested in CARTESIAN components: the density tiles are expanded with the 5 x 6
// c2s matrix when they are loaded and the J / K tiles contracted with it before they are added (J = C J_cart C^T with
// D_cart = C^T D C), so the 6-component block never needs transforming.  Classes whose block exceeds the registers
// ((dp|pp) and up) run as several passes over slices of the bra components, each pass recomputing roots and tables and
// keeping <= 72 integrals per lane (the unused table entries of a pass are dead code).  Classes with f shells and
// above stay on the shared-memory engine of jk.cuh.
#pragma once
#include "common.cuh"
#include <utility>

#define JKR_THREADS 128
#define JKR_WARPS (JKR_THREADS / 32)
#define JKR_NINT 128   // the refined root table (tables.cuh: b200qc_rys_refine): intervals of width 1/2,
#define JKR_NCOEF 10   // Chebyshev degree 9
#define JKR_WSM (JKR_MAXPP * 6 + 36 + 36 + 32)   // doubles of shared memory per warp: bra primitive pairs, D_ij and J_ij
                                                // tiles (cartesian), D_ij as stored

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

// ---- real-spherical <-> cartesian on the AO side ---------------------------------------------------------------
// l <= 1: the engine's cartesian functions ARE the AOs (constant folded into the coefficients); l == 2: the five real
// d functions, c2s matrix (5 x 6, rows m = -2..2 = xy, yz, z^2, xz, x^2 - y^2; tables.cuh) in JKRArgs::c2s_d.
template <int L> struct Sph { static constexpr int n = L <= 1 ? ncart(L) : 2 * L + 1; };
__host__ __device__ constexpr bool d_nz(int m, int c) {   // cartesian order xx xy xz yy yz zz
    return m == 0 ? c == 1 : m == 1 ? c == 4 : m == 2 ? (c == 0 || c == 3 || c == 5) : m == 3 ? c == 2 : (c == 0 || c == 3);
}
template <int L> __host__ __device__ constexpr bool c2s_nz(int m, int c) { return L <= 1 ? m == c : d_nz(m, c); }
template <int L> __device__ __forceinline__ double c2s_v(const JKRArgs &A, int m, int c) {
    return L <= 1 ? 1.0 : A.c2s_d[m * 6 + c];
}

// t[ca][cb] (cartesian) of the density block with first AOs (ra, cb0)
template <int LA, int LB>
__device__ __forceinline__ void load_tile(const double *__restrict__ D, int ra, int cb0, int nao, const JKRArgs &A,
                                          double (&t)[ncart(LA) * ncart(LB)]) {
    constexpr int NSA = Sph<LA>::n, NSB = Sph<LB>::n, NCA = ncart(LA), NCB = ncart(LB);
    double s[NSA * NSB];
#pragma unroll
    for (int e = 0; e < NSA * NSB; e++) s[e] = D[(int64_t)(ra + e / NSB) * nao + cb0 + e % NSB];
    double h[NCA * NSB];
    if constexpr (LA <= 1) {
#pragma unroll
        for (int e = 0; e < NCA * NSB; e++) h[e] = s[e];
    } else {
#pragma unroll
        for (int ca = 0; ca < NCA; ca++) {
#pragma unroll
            for (int mb = 0; mb < NSB; mb++) {
                double v = 0.0;
#pragma unroll
                for (int ma = 0; ma < NSA; ma++)
                    if (c2s_nz<LA>(ma, ca)) v += c2s_v<LA>(A, ma, ca) * s[ma * NSB + mb];
                h[ca * NSB + mb] = v;
            }
        }
    }
    if constexpr (LB <= 1) {
#pragma unroll
        for (int e = 0; e < NCA * NCB; e++) t[e] = h[e];
    } else {
#pragma unroll
        for (int ca = 0; ca < NCA; ca++) {
#pragma unroll
            for (int cb = 0; cb < NCB; cb++) {
                double v = 0.0;
#pragma unroll
                for (int mb = 0; mb < NSB; mb++)
                    if (c2s_nz<LB>(mb, cb)) v += c2s_v<LB>(A, mb, cb) * h[ca * NSB + mb];
                t[ca * NCB + cb] = v;
            }
        }
    }
}

struct AllRows { __host__ __device__ static constexpr bool has(int) { return true; } };

// M[ra + ma][cb0 + mb] += scale * (c2s_A t c2s_B^T)[ma][mb]; cartesian rows of t outside RM::has are absent (zero):
// only the AO rows they reach are touched
template <int LA, int LB, class RM>
__device__ __forceinline__ void add_tile(double *__restrict__ M, int ra, int cb0, int nao, const JKRArgs &A,
                                         const double (&t)[ncart(LA) * ncart(LB)], double scale) {
    constexpr int NSA = Sph<LA>::n, NSB = Sph<LB>::n, NCA = ncart(LA), NCB = ncart(LB);
#pragma unroll
    for (int ma = 0; ma < NSA; ma++) {
        bool touched = false;
#pragma unroll
        for (int ca = 0; ca < NCA; ca++) touched = touched || (RM::has(ca) && c2s_nz<LA>(ma, ca));
        if (!touched) continue;
        double h[NCB];
#pragma unroll
        for (int cb = 0; cb < NCB; cb++) {
            double v = 0.0;
#pragma unroll
            for (int ca = 0; ca < NCA; ca++)
                if (RM::has(ca) && c2s_nz<LA>(ma, ca)) v += c2s_v<LA>(A, ma, ca) * t[ca * NCB + cb];
            h[cb] = v;
        }
#pragma unroll
        for (int mb = 0; mb < NSB; mb++) {
            double v = 0.0;
#pragma unroll
            for (int cb = 0; cb < NCB; cb++)
                if (c2s_nz<LB>(mb, cb)) v += c2s_v<LB>(A, mb, cb) * h[cb];
            atomicAdd(M + (int64_t)(ra + ma) * nao + cb0 + mb, scale * v);
        }
    }
}

// A pass = the bra components e = ci * NJ + cj in [E0, E1): their NE x NKL integrals are the register block.
template <class S, int E0_, int E1_>
struct Pass {
    static constexpr int E0 = E0_, E1 = E1_, NE = E1_ - E0_, NACC = NE * S::NKL;
    static constexpr int A0 = E0_ / S::NJ, A1 = (E1_ - 1) / S::NJ;
    __host__ __device__ static constexpr bool has(int a, int b) { return a * S::NJ + b >= E0_ && a * S::NJ + b < E1_; }
    struct RowsA { __host__ __device__ static constexpr bool has(int a) { return a >= A0 && a <= A1; } };
    struct RowsB {
        __host__ __device__ static constexpr bool has(int b) {
            return A1 > A0 + 1 ? true
                 : A1 == A0 ? (b >= E0_ % S::NJ && b <= (E1_ - 1) % S::NJ)
                            : (b >= E0_ % S::NJ || b <= (E1_ - 1) % S::NJ);
        }
    };
};

template <class S, class P, int C>
__device__ __forceinline__ void acc_one(double (&acc)[P::NACC], const double (&Tx)[S::TS], const double (&Ty)[S::TS],
                                        const double (&Tz)[S::TS]) {
    constexpr int cg = P::E0 * S::NKL + C;   // component index in the full block
    constexpr int ix = S::comp_index(cg, 0), iy = S::comp_index(cg, 1), iz = S::comp_index(cg, 2);
    acc[C] += Tx[ix] * Ty[iy] * Tz[iz];
}
template <class S, class P, int... Cs>
__device__ __forceinline__ void acc_all(double (&acc)[P::NACC], const double (&Tx)[S::TS], const double (&Ty)[S::TS],
                                        const double (&Tz)[S::TS], std::integer_sequence<int, Cs...>) {
    (acc_one<S, P, Cs>(acc, Tx, Ty, Tz), ...);
}

// a[r] without run-time indexing of a register array
template <int N> __device__ __forceinline__ double pick(const double (&a)[N], int r) {
    double v = a[0];
#pragma unroll
    for (int i = 1; i < N; i++) v = r == i ? a[i] : v;
    return v;
}

// One pass over one contracted quartet per lane (valid lanes): integrals of the pass's bra components, digestion.
// Every lane of the warp calls it (the J_ij partial sums are reduced over the warp at the end).
// DEFER (single-pass classes with small bra tiles): the lane keeps its J_ij partial sums in registers (jacc) across its
// kets and the warp sum happens once per work item instead of once per round of 32 kets.
template <int LI, int LJ, int LK, int LL, class P, bool DOJ, bool DOK, bool DEFER>
__device__ __forceinline__ void quartet_pass(bool valid, const JKPair &bp, const JKPair &kp, const double *wprim,
                                             const double *wdij, double *wjij, const double *__restrict__ rt,
                                             const JKRArgs &A, double ih, int lane,
                                             double (&jacc)[DEFER ? Shape<LI, LJ, LK, LL>::NIJ : 1]) {
    using S = Shape<LI, LJ, LK, LL>;
    constexpr int NI = S::NI, NJ = S::NJ, NK = S::NK, NL = S::NL, NR = S::NR, NKL = S::NKL;
    constexpr int E0 = P::E0, NE = P::NE;
    double jc[NE];
#pragma unroll
    for (int e = 0; e < NE; e++) jc[e] = 0.0;
    if (valid) {
        const int nao = A.nao;
        const double *__restrict__ D = A.dm;
        double acc[P::NACC];
#pragma unroll
        for (int c = 0; c < P::NACC; c++) acc[c] = 0.0;
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
                // d classes with three and more roots: the loop over roots stays rolled (a quarter of the code)
#pragma unroll((LI >= 2 && NR >= 3) ? 1 : NR)
                for (int r = 0; r < NR; r++) {
                    const double ur = pick<NR>(u, r), wr = pick<NR>(w, r);
                    const double a0u = a0 * ur, a1u = a1 * ur;
                    const double b10 = (1.0 - a0u) * hp, b01 = (1.0 - a1u) * kq.hp, b00 = 0.5 * ur * ipq;
                    double Tx[S::TS], Ty[S::TS], Tz[S::TS];
                    build_table<LI, LJ, LK, LL>(Tx, 1.0, pax - a0u * dx, qcx + a1u * dx, b10, b01, b00, bp.abx, kp.abx);
                    build_table<LI, LJ, LK, LL>(Ty, 1.0, pay - a0u * dy, qcy + a1u * dy, b10, b01, b00, bp.aby, kp.aby);
                    build_table<LI, LJ, LK, LL>(Tz, wr * pref, paz - a0u * dz, qcz + a1u * dz, b10, b01, b00, bp.abz,
                                                kp.abz);
                    acc_all<S, P>(acc, Tx, Ty, Tz, std::make_integer_sequence<int, P::NACC>());
                }
            }
        }
        // ---- digestion (the sums of jk_kernel, in cartesian components; D symmetric, the caller adds the transpose) ----
        double f = 1.0;
        if (bp.ish == bp.jsh) f *= 0.5;
        if (kp.ish == kp.jsh) f *= 0.5;
        if (bp.ish == kp.ish && bp.jsh == kp.jsh) f *= 0.5;
        const int ai = bp.ao_i, aj = bp.ao_j, ak = kp.ao_i, al = kp.ao_j;
        if (DOJ) {
            const double f2 = 2.0 * f;
            {   // J_ij += 2f sum_kl B D_kl
                double t[NKL];
                load_tile<LK, LL>(D, ak, al, nao, A, t);
#pragma unroll
                for (int e = 0; e < NE; e++) {
                    double v = 0.0;
#pragma unroll
                    for (int c = 0; c < NKL; c++) v += acc[e * NKL + c] * t[c];
                    jc[e] = f2 * v;
                }
            }
            {   // J_kl += 2f sum_ij B D_ij
                double t[NKL];
#pragma unroll
                for (int c = 0; c < NKL; c++) {
                    double v = 0.0;
#pragma unroll
                    for (int e = 0; e < NE; e++) v += acc[e * NKL + c] * wdij[E0 + e];
                    t[c] = v;
                }
                add_tile<LK, LL, AllRows>(A.vj, ak, al, nao, A, t, f2);
            }
        }
        if (DOK) {
            {   // K_ik += f sum_jl B D_jl
                double d[NJ * NL], t[NI * NK];
                load_tile<LJ, LL>(D, aj, al, nao, A, d);
#pragma unroll
                for (int a = 0; a < NI; a++) {
                    if (!P::RowsA::has(a)) continue;
#pragma unroll
                    for (int c = 0; c < NK; c++) {
                        double v = 0.0;
#pragma unroll
                        for (int b = 0; b < NJ; b++) {
                            if (!P::has(a, b)) continue;
#pragma unroll
                            for (int dd = 0; dd < NL; dd++)
                                v += acc[((a * NJ + b - E0) * NK + c) * NL + dd] * d[b * NL + dd];
                        }
                        t[a * NK + c] = v;
                    }
                }
                add_tile<LI, LK, typename P::RowsA>(A.vk, ai, ak, nao, A, t, f);
            }
            {   // K_jk += f sum_il B D_il
                double d[NI * NL], t[NJ * NK];
                load_tile<LI, LL>(D, ai, al, nao, A, d);
#pragma unroll
                for (int b = 0; b < NJ; b++) {
                    if (!P::RowsB::has(b)) continue;
#pragma unroll
                    for (int c = 0; c < NK; c++) {
                        double v = 0.0;
#pragma unroll
                        for (int a = 0; a < NI; a++) {
                            if (!P::has(a, b)) continue;
#pragma unroll
                            for (int dd = 0; dd < NL; dd++)
                                v += acc[((a * NJ + b - E0) * NK + c) * NL + dd] * d[a * NL + dd];
                        }
                        t[b * NK + c] = v;
                    }
                }
                add_tile<LJ, LK, typename P::RowsB>(A.vk, aj, ak, nao, A, t, f);
            }
            {   // K_il += f sum_jk B D_jk
                double d[NJ * NK], t[NI * NL];
                load_tile<LJ, LK>(D, aj, ak, nao, A, d);
#pragma unroll
                for (int a = 0; a < NI; a++) {
                    if (!P::RowsA::has(a)) continue;
#pragma unroll
                    for (int dd = 0; dd < NL; dd++) {
                        double v = 0.0;
#pragma unroll
                        for (int b = 0; b < NJ; b++) {
                            if (!P::has(a, b)) continue;
#pragma unroll
                            for (int c = 0; c < NK; c++)
                                v += acc[((a * NJ + b - E0) * NK + c) * NL + dd] * d[b * NK + c];
                        }
                        t[a * NL + dd] = v;
                    }
                }
                add_tile<LI, LL, typename P::RowsA>(A.vk, ai, al, nao, A, t, f);
            }
            {   // K_jl += f sum_ik B D_ik
                double d[NI * NK], t[NJ * NL];
                load_tile<LI, LK>(D, ai, ak, nao, A, d);
#pragma unroll
                for (int b = 0; b < NJ; b++) {
                    if (!P::RowsB::has(b)) continue;
#pragma unroll
                    for (int dd = 0; dd < NL; dd++) {
                        double v = 0.0;
#pragma unroll
                        for (int a = 0; a < NI; a++) {
                            if (!P::has(a, b)) continue;
#pragma unroll
                            for (int c = 0; c < NK; c++)
                                v += acc[((a * NJ + b - E0) * NK + c) * NL + dd] * d[a * NK + c];
                        }
                        t[b * NL + dd] = v;
                    }
                }
                add_tile<LJ, LL, typename P::RowsB>(A.vk, aj, al, nao, A, t, f);
            }
        }
    }
    if (DOJ && DEFER) {
#pragma unroll
        for (int e = 0; e < NE; e++) jacc[E0 + e] += jc[e];
    } else if (DOJ) {
        // bra-tile J_ij: summed over the warp, kept (cartesian) in the warp's shared memory until the work item ends
#pragma unroll
        for (int e = 0; e < NE; e++) {
            double v = jc[e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) wjij[E0 + e] += v;
        }
    }
}

template <int LI, int LJ, int LK, int LL, int NPASS, int P0, bool DOJ, bool DOK, bool DEFER, int... Ps>
__device__ __forceinline__ void quartet_passes(bool valid, const JKPair &bp, const JKPair &kp, const double *wprim,
                                               const double *wdij, double *wjij, const double *__restrict__ rt,
                                               const JKRArgs &A, double ih, int lane,
                                               double (&jacc)[DEFER ? Shape<LI, LJ, LK, LL>::NIJ : 1],
                                               std::integer_sequence<int, Ps...>) {
    using S = Shape<LI, LJ, LK, LL>;
    (quartet_pass<LI, LJ, LK, LL, Pass<S, (P0 + Ps) * S::NIJ / NPASS, (P0 + Ps + 1) * S::NIJ / NPASS>, DOJ, DOK, DEFER>(
         valid, bp, kp, wprim, wdij, wjij, rt, A, ih, lane, jacc), ...);
}

// c2s coefficient with run-time indices (staging code only): c2sd = the d matrix in shared memory
__device__ __forceinline__ double c2s_rt(int l, int m, int c, const double *c2sd) {
    return l <= 1 ? (m == c ? 1.0 : 0.0) : c2sd[m * 6 + c];
}

// passes [P0, P1) of the NPASS passes over the bra components (the many-pass classes are split over several kernels)
template <int LI, int LJ, int LK, int LL, int NPASS, int P0, int P1, bool DOJ, bool DOK>
__global__ void __launch_bounds__(JKR_THREADS) jk_reg_kernel(const JKRArgs A) {
    using S = Shape<LI, LJ, LK, LL>;
    constexpr int NI = S::NI, NJ = S::NJ, NR = S::NR, NIJ = S::NIJ;
    constexpr int NSI = Sph<LI>::n, NSJ = Sph<LJ>::n;
    static_assert(NIJ % NPASS == 0, "passes must divide the bra components");
    extern __shared__ __align__(16) double jkr_smem[];
    double *rt = jkr_smem;
    double *c2sd = jkr_smem + JKR_NCOEF * 2 * NR * JKR_NINT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *wprim = c2sd + 32 + warp * JKR_WSM;
    double *wdij = wprim + JKR_MAXPP * 6;      // D_ij, cartesian
    double *wjij = wdij + 36;                  // J_ij of the work item, cartesian
    double *wsph = wjij + 36;                  // D_ij as stored (spherical AOs)
    for (int e = threadIdx.x; e < JKR_NINT * 2 * NR * JKR_NCOEF; e += JKR_THREADS) {
        const int it = e / (2 * NR * JKR_NCOEF), rem = e - it * (2 * NR * JKR_NCOEF);
        const int f = rem / JKR_NCOEF, k = rem - f * JKR_NCOEF;
        rt[(k * 2 * NR + f) * JKR_NINT + it] = A.rys_coef[e];
    }
    if (threadIdx.x < 30) c2sd[threadIdx.x] = A.c2s_d[threadIdx.x];
    __syncthreads();
    const int nao = A.nao;
    const double ih = 1.0 / A.rys_h;
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
            for (int e = lane; e < NSI * NSJ; e += 32)
                wsph[e] = A.dm[(int64_t)(bp.ao_i + e / NSJ) * nao + bp.ao_j + e % NSJ];
        }
        __syncwarp();
        for (int e = lane; e < NIJ; e += 32) {
            const int ca = e / NJ, cb = e - ca * NJ;
            double v = 0.0;
            for (int ma = 0; ma < NSI; ma++)
                for (int mb = 0; mb < NSJ; mb++)
                    v += c2s_rt(LI, ma, ca, c2sd) * c2s_rt(LJ, mb, cb, c2sd) * wsph[ma * NSJ + mb];
            wdij[e] = v;
            wjij[e] = 0.0;
        }
        __syncwarp();
        constexpr bool DEFER = DOJ && NPASS == 1 && S::NC <= 54;
        double jacc[DEFER ? NIJ : 1];
#pragma unroll
        for (int e = 0; e < (DEFER ? NIJ : 1); e++) jacc[e] = 0.0;
        for (int kk0 = k0; kk0 < k1; kk0 += 32) {
            const int kk = kk0 + lane;
            const bool valid = kk < k1;
            const JKPair kp = A.ket[valid ? kk : k0];
            quartet_passes<LI, LJ, LK, LL, NPASS, P0, DOJ, DOK, DEFER>(valid, bp, kp, wprim, wdij, wjij, rt, A, ih, lane,
                                                                       jacc, std::make_integer_sequence<int, P1 - P0>());
        }
        if (DEFER) {
#pragma unroll
            for (int e = 0; e < (DEFER ? NIJ : 1); e++) {
                double v = jacc[e];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) wjij[e] = v;
            }
        }
        if (DOJ) {
            __syncwarp();
            for (int e = lane; e < NSI * NSJ; e += 32) {
                const int ma = e / NSJ, mb = e - ma * NSJ;
                double v = 0.0;
                for (int ca = 0; ca < NI; ca++)
                    for (int cb = 0; cb < NJ; cb++)
                        v += c2s_rt(LI, ma, ca, c2sd) * c2s_rt(LJ, mb, cb, c2sd) * wjij[ca * NJ + cb];
                atomicAdd(A.vj + (int64_t)(bp.ao_i + ma) * nao + bp.ao_j + mb, v);
            }
        }
    }
}

template <int LI, int LJ, int LK, int LL, int NPASS, int P0, int P1, bool DOJ, bool DOK>
static int launch_t(const JKRArgs &A, cudaStream_t st) {
    using S = Shape<LI, LJ, LK, LL>;
    static_assert(S::NR <= JKR_MAXROOTS && S::NIJ <= 36, "class outside the staged sizes");
    const size_t smem = sizeof(double) * ((size_t)JKR_NCOEF * 2 * S::NR * JKR_NINT + 32 + (size_t)JKR_WARPS * JKR_WSM);
    auto kern = jk_reg_kernel<LI, LJ, LK, LL, NPASS, P0, P1, DOJ, DOK>;
    // attribute + occupancy once per kernel and device: a build is hundreds of launches, and on small systems the
    // host side of a launch is what bounds it
    static int occ_of_device[64] = {};
    int dev = 0;
    QC_CHECK(cudaGetDevice(&dev));
    QC_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
    if (occ_of_device[dev] == 0) {
        QC_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int o = 0;
        QC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, JKR_THREADS, smem));
        QC_REQUIRE(o >= 1, "J/K register kernel does not fit on an SM");
        occ_of_device[dev] = o;
    }
    const int occ = occ_of_device[dev];
    const int64_t mine = (A.nitems - A.item0 + A.item_stride - 1) / A.item_stride;   // items item0, item0 + stride, ...
    if (mine <= 0) return 0;
    const int64_t nblk = std::min<int64_t>((mine + JKR_WARPS - 1) / JKR_WARPS, (int64_t)NUM_SMS * occ);
    prof_begin(PROF_JK, st);
    kern<<<(unsigned)nblk, JKR_THREADS, smem, st>>>(A);
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

}  // namespace jkr

#if !defined(JKR_UNIT) || JKR_UNIT == 0
// Measured issue-rate peak of the plain fp64 pipe (DFMA, 8 independent chains per thread): the roofline denominator of
// the J/K quartet kernels (MEASURED_PEAKS.json has no fp64 entry).
__global__ void __launch_bounds__(256, 2) dfma_peak_kernel(int iters, double *out) {
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = 1e-3 * (threadIdx.x + i);
    const double a = 1.0 - 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x + 1);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i];
    if (s == 123.456) out[0] = s;   // keeps the loop alive
}
extern "C" int b200qc_peak_fp64_fma(int iters, double *scratch, double *tflops, void *stream) {
    cudaStream_t st = as_stream(stream);
    const int nblk = NUM_SMS * 8;
    cudaEvent_t e0, e1;
    QC_CHECK(cudaEventCreate(&e0));
    QC_CHECK(cudaEventCreate(&e1));
    dfma_peak_kernel<<<nblk, 256, 0, st>>>(iters / 8 + 1, scratch);   // warm-up
    QC_CHECK(cudaEventRecord(e0, st));
    dfma_peak_kernel<<<nblk, 256, 0, st>>>(iters, scratch);
    QC_CHECK(cudaEventRecord(e1, st));
    QC_LAUNCHED(2);
    QC_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    QC_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *tflops = 2.0 * 8.0 * iters * 256.0 * nblk / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}
#endif

// Class table, entry definitions per translation unit and the dispatcher: generated by tools/gen_jk_units.py
#ifndef JKR_UNIT
#define JKR_UNIT 0
#endif
#include "jk_reg_units.inc"
