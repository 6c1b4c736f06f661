// K10 / K11 -- generic fp64-accurate GEMM on the 5th-generation tensor cores, the engine of the density-fitted
// exact-exchange build:    C[b][m][n] (=, +=) alpha * sum_k A[b][m][k] B[b][n][k]        ("NT", both K-major)
// with both fp64 operands cut into S int8 slices per row (error-free Ozaki scheme, exactly as in vxc_i8.cuh /
// rho_i8.cuh: S (S + 1) / 2 tcgen05.mma.kind::i8 products per K step, one int32 TMEM accumulator per
// anti-diagonal, exact int64 recombination, power-of-two row scales applied in fp64).
//
// Operand planes live in global memory ALREADY in the no-swizzle UMMA core-matrix order, so a pipeline stage
// is two contiguous blocks fetched by one thread with two cp.async.bulk copies (mbarrier complete_tx):
//   rows-per-tile W = 128 (A operand) or 64 (B operand), K step 32:
//   [batch][row tile = r / W][k tile = k / 32][slice][(k % 32) / 16][(r % W) / 8][r % 8][k % 16]   (bytes)
// A "batch" is either an independent problem (stage 1 of DF-K: one per AO index i) or a K chunk of one long
// contraction (stage 2: split-K, every chunk with its own row scales, fp64 atomics into the same C).
//
// Replaces: nothing in the reference -- dqc raises for exact exchange with density fitting
// (dqc/hamilton/hcgto.py:229-230); SURVEY section 8a defines the extension
// K^DF_ij = sum_PQ (ik|P) (P|Q)^-1 (Q|jl) D_kl, validated against the 4-centre K (a6).
#pragma once
#include "rho_i8.cuh"

#define GI8_STAGES 5

// ---- slicers -------------------------------------------------------------------------------------------
// Source element (batch b, row r, column k) at src[b * sb + r * sr + k * sk]; pair mode (the packed (ij|P)
// tensor): batch = AO i, k = AO j, r = aux P at src[tri(i, j) * pair_ld + r].
struct I8SliceArgs {
    const double *src;
    int64_t sb, sr, sk, pair_ld;
    int pair_mode;
    int nbatch, R, K, K_last;   // valid rows; valid columns per batch (K_last for the last batch)
    int Rpad, Kpad;             // multiples of W and 32
    signed char *planes;
    double *scales;             // [batch][Rpad]
};

__device__ __forceinline__ int64_t i8_src_index(const I8SliceArgs &a, int b, int r, int k) {
    if (a.pair_mode) {
        const int64_t i = b, j = k;
        const int64_t t = i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i;
        return t * a.pair_ld + r;
    }
    return (int64_t)b * a.sb + (int64_t)r * a.sr + (int64_t)k * a.sk;
}

// 16 consecutive k of one row -> S x 16 bytes
template <int S>
__device__ __forceinline__ void i8_quantise16(const double (&x)[16], double inv, uint4 (&out)[S]) {
    unsigned int w[S][4];
#pragma unroll
    for (int s = 0; s < S; s++) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        double y = x[j] * inv;
#pragma unroll
        for (int s = 0; s < S; s++) {
            w[s][j >> 2] |= ((unsigned int)slice_digit(y) & 0xffu) << (8 * (j & 3));
        }
    }
#pragma unroll
    for (int s = 0; s < S; s++) out[s] = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
}

// Row-contiguous sources (consecutive rows are adjacent in memory): thread = row, loops over k.
template <int S, int W>
__global__ void __launch_bounds__(128)
i8_slice_rc_kernel(I8SliceArgs a) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * 128 + threadIdx.x;
    if (r >= a.Rpad) return;
    const int Kb = (b == a.nbatch - 1) ? a.K_last : a.K;
    const bool rv = r < a.R;
    double m = 0.0;
    if (rv)
        for (int k = 0; k < Kb; k++) m = fmax(m, fabs(a.src[i8_src_index(a, b, r, k)]));
    int e = 0;
    if (m > 0.0) frexp(m, &e);
    const double inv = ldexp(64.0, -e);
    a.scales[(int64_t)b * a.Rpad + r] = ldexp(1.0, e);
    constexpr int PL = 32 * W;
    const int nk = a.Kpad / 32;
    signed char *P = a.planes + ((int64_t)b * (a.Rpad / W) + r / W) * nk * S * PL + ((r % W) >> 3) * 128 + (r & 7) * 16;
    for (int kg = 0; kg < a.Kpad / 16; kg++) {
        double x[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int k = kg * 16 + j;
            x[j] = (rv && k < Kb) ? a.src[i8_src_index(a, b, r, k)] : 0.0;
        }
        uint4 q[S];
        i8_quantise16<S>(x, inv, q);
        // B role (W = 64): [K chunk][slice][row] so that consecutive slices form one K-major operand (see gemm_i8_kernel)
        constexpr int KC = (W == I8_BN) ? S * W * 16 : W * 16, SL = (W == I8_BN) ? W * 16 : PL;
        signed char *Q = P + (int64_t)(kg >> 1) * S * PL + (kg & 1) * KC;
#pragma unroll
        for (int s = 0; s < S; s++) *reinterpret_cast<uint4 *>(Q + s * SL) = q[s];
    }
}

// K-contiguous sources (sk == 1): warp = row, a lane owns 16 consecutive k at a time.
template <int S, int W>
__global__ void __launch_bounds__(256)
i8_slice_kc_kernel(I8SliceArgs a) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= a.Rpad) return;
    const int Kb = (b == a.nbatch - 1) ? a.K_last : a.K;
    const bool rv = r < a.R;
    const double *X = a.src + (int64_t)b * a.sb + (int64_t)r * a.sr;
    double m = 0.0;
    if (rv)
        for (int k = lane; k < Kb; k += 32) m = fmax(m, fabs(X[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    int e = 0;
    if (m > 0.0) frexp(m, &e);
    const double inv = ldexp(64.0, -e);
    if (lane == 0) a.scales[(int64_t)b * a.Rpad + r] = ldexp(1.0, e);
    constexpr int PL = 32 * W;
    const int nk = a.Kpad / 32;
    signed char *P = a.planes + ((int64_t)b * (a.Rpad / W) + r / W) * nk * S * PL + ((r % W) >> 3) * 128 + (r & 7) * 16;
    for (int kg = lane; kg < a.Kpad / 16; kg += 32) {
        double x[16];
        const int k0 = kg * 16;
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] = (rv && k0 + j < Kb) ? X[k0 + j] : 0.0;
        uint4 q[S];
        i8_quantise16<S>(x, inv, q);
        // B role (W = 64): [K chunk][slice][row] so that consecutive slices form one K-major operand (see gemm_i8_kernel)
        constexpr int KC = (W == I8_BN) ? S * W * 16 : W * 16, SL = (W == I8_BN) ? W * 16 : PL;
        signed char *Q = P + (int64_t)(kg >> 1) * S * PL + (kg & 1) * KC;
#pragma unroll
        for (int s = 0; s < S; s++) *reinterpret_cast<uint4 *>(Q + s * SL) = q[s];
    }
}

// K-contiguous source with KNOWN row maxima (from the producing GEMM's epilogue): one pass, both operand forms.
// Source element (batch b, row r, k) at src[b * sb + r * sr + k]; rowmax[b_row * ... ] is indexed [r][b] with
// stride rm_ld (the producing GEMM had the roles of batch and row swapped).  A CTA = 8 consecutive rows x one
// 512-column segment per iteration: coalesced fp64 loads (lane = column), bytes staged in shared memory by
// column, then 16-byte chunks written so that 8 rows x 16 bytes form one contiguous 128-byte line of a plane.
struct I8Slice2Args {
    const double *src;
    int64_t sb, sr;
    const double *rowmax;      // non-negative doubles, [r * rm_ld + b]
    int64_t rm_ld;
    int nbatch, R, K, K_last, Kpad;
    int RpadA, RpadB;          // multiples of 128 / 64
    signed char *planesA, *planesB;
    double *scalesA, *scalesB; // [batch][RpadA], [batch][RpadB]
};

template <int S>
__global__ void __launch_bounds__(256)
i8_slice_dual_kernel(I8Slice2Args a) {
    __shared__ __align__(16) signed char sm[8][S][512];
    __shared__ double sinv[8];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * 8;
    const int r = r0 + warp;
    const int Kb = (b == a.nbatch - 1) ? a.K_last : a.K;
    const bool rv = r < a.R;
    if (lane == 0) {
        const double m = rv ? a.rowmax[(int64_t)r * a.rm_ld + b] : 0.0;
        int e = 0;
        if (m > 0.0) frexp(m, &e);
        sinv[warp] = ldexp(64.0, -e);
        const double sc = ldexp(1.0, e);
        if (r < a.RpadA) a.scalesA[(int64_t)b * a.RpadA + r] = sc;
        if (r < a.RpadB) a.scalesB[(int64_t)b * a.RpadB + r] = sc;
    }
    __syncwarp();
    const double inv = sinv[warp];
    const double *X = a.src + (int64_t)b * a.sb + (int64_t)r * a.sr;
    const bool aligned4 = ((a.sb | a.sr) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.src) & 31) == 0;
    const int nk = a.Kpad / 32;
    const bool wa = r0 < a.RpadA, wb = r0 < a.RpadB;     // 8-row groups never straddle a padded size
    // writer mapping: 8 consecutive lanes = the 8 rows of one 16-column group
    const int wrow = threadIdx.x & 7, wgrp = threadIdx.x >> 3;          // 32 groups per 512-column segment
    const int rr = r0 + wrow;
    signed char *PA = a.planesA + ((int64_t)b * (a.RpadA / 128) + rr / 128) * nk * S * 4096 + ((rr % 128) >> 3) * 128 + (rr & 7) * 16;
    signed char *PB = a.planesB + ((int64_t)b * (a.RpadB / 64) + rr / 64) * nk * S * 2048 + ((rr % 64) >> 3) * 128 + (rr & 7) * 16;
    for (int seg = 0; seg * 512 < a.Kpad; seg++) {
        const int k0 = seg * 512;
        // a lane takes 4 consecutive columns at a time: one 256-bit load (the source rows are 32-byte aligned when sr
        // and sb are multiples of 4 -- else scalar loads), one 32-bit shared-memory store per slice
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int k = k0 + t * 128 + 4 * lane;
            double y[4] = {0.0, 0.0, 0.0, 0.0};
            if (rv && k + 3 < Kb && aligned4) {
                asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];"
                             : "=d"(y[0]), "=d"(y[1]), "=d"(y[2]), "=d"(y[3]) : "l"(X + k));
            } else if (rv) {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (k + j < Kb) y[j] = __ldcs(X + k + j);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) y[j] *= inv;
#pragma unroll
            for (int s = 0; s < S; s++) {
                unsigned int w = 0u;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    w |= ((unsigned int)slice_digit(y[j]) & 0xffu) << (8 * j);
                }
                *reinterpret_cast<unsigned int *>(&sm[warp][s][t * 128 + 4 * lane]) = w;
            }
        }
        __syncthreads();
        const int kg = (k0 >> 4) + wgrp;                 // 16-column group of this writer
        if (kg * 16 < a.Kpad) {
#pragma unroll
            for (int s = 0; s < S; s++) {
                const uint4 v = *reinterpret_cast<const uint4 *>(&sm[wrow][s][wgrp * 16]);
                if (wa) *reinterpret_cast<uint4 *>(PA + (int64_t)(kg >> 1) * S * 4096 + s * 4096 + (kg & 1) * 2048) = v;
                if (wb) *reinterpret_cast<uint4 *>(PB + (int64_t)(kg >> 1) * S * 2048 + (kg & 1) * (S * 1024) + s * 1024) = v;
            }
        }
        __syncthreads();
    }
}

// ---- the GEMM ------------------------------------------------------------------------------------------
struct I8GemmArgs {
    const signed char *A, *B;
    const double *ascale, *bscale;
    int64_t a_bstride, b_bstride;      // bytes between batches (0 = operand shared by all batches)
    int64_t as_bstride, bs_bstride;    // scale entries between batches
    int nbatch, mtiles, ntiles, nk, nk_last;
    int M, N;                          // valid rows / columns of C
    double *C;
    int64_t c_bstride, ldc;
    int mode;                          // 0 store, 1 atomic add, 2 atomic add of the lower-triangle tiles only
    double alpha;
    int units_per_batch;
    // optional (mode 0): running maxima of |C| per (batch, row / rm_div), as the bit patterns of non-negative
    // doubles (atomicMax on them is monotone) -- the row scales of a following slicing pass, for free
    unsigned long long *rowmax;
    int64_t rm_bstride;
    int rm_div;
};

// unit -> (batch, M tile, N tile); mode 2 enumerates only tiles with tn * 64 < (tm + 1) * 128
__device__ __forceinline__ void gi8_locate(const I8GemmArgs &g, int u, int &b, int &tm, int &tn) {
    b = u / g.units_per_batch;
    int r = u - b * g.units_per_batch;
    if (g.mode != 2) {
        tm = r / g.ntiles;
        tn = r - tm * g.ntiles;
        return;
    }
    tm = 0;
    for (;;) {
        const int c = min(g.ntiles, 2 * tm + 2);
        if (r < c) break;
        r -= c;
        tm++;
    }
    tn = r;
}

template <int S>
__global__ void __launch_bounds__(I8_THREADS, 1)
gemm_i8_kernel(I8GemmArgs g) {
    extern __shared__ __align__(1024) unsigned char i8_smem[];
    constexpr int A_STAGE = S * I8_A_PLANE, B_STAGE = S * I8_B_PLANE, STAGE = A_STAGE + B_STAGE;
    __shared__ uint64_t full_bar[GI8_STAGES], empty_bar[GI8_STAGES], accum_full, accum_empty;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nunits = g.nbatch * g.units_per_batch;

    if (tid == 0) {
        for (int i = 0; i < GI8_STAGES; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(&accum_full, 1);
        mbar_init(&accum_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    const uint32_t sbase = smem_u32(i8_smem);

    if (warp == 0) {
        // ===== producer: two contiguous blocks per stage =====
        if (lane == 0) {
            int it = 0;
            for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
                int b, tm, tn;
                gi8_locate(g, u, b, tm, tn);
                const int ks = (b == g.nbatch - 1) ? g.nk_last : g.nk;
                const signed char *A = g.A + b * g.a_bstride + (int64_t)tm * g.nk * A_STAGE;
                const signed char *B = g.B + b * g.b_bstride + (int64_t)tn * g.nk * B_STAGE;
                for (int kt = 0; kt < ks; kt++, it++) {
                    const int slot = it % GI8_STAGES;
                    mbar_wait(&empty_bar[slot], ((it / GI8_STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full_bar[slot], STAGE);
                    bulk_g2s(sbase + slot * STAGE, A + (int64_t)kt * A_STAGE, A_STAGE, &full_bar[slot]);
                    bulk_g2s(sbase + slot * STAGE + A_STAGE, B + (int64_t)kt * B_STAGE, B_STAGE, &full_bar[slot]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: K-major, no swizzle (LBO = stride between the two 16-byte K chunks, SBO = 128) =====
        if (lane == 0) {
            // B stage: [K chunk][slice][64 rows][16 B] -- the slices t0 .. t0 + n - 1 are ONE operand of 64 n rows
            const uint64_t da0 = umma_desc(sbase, 2048, 128), db0 = umma_desc(sbase + A_STAGE, S * 1024, 128);
            constexpr uint32_t IDESC0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_BM >> 4) << 24);
            int it = 0, nt = 0;
            for (int u = blockIdx.x; u < nunits; u += gridDim.x, nt++) {
                const int b = u / g.units_per_batch;
                const int ks = (b == g.nbatch - 1) ? g.nk_last : g.nk;
                mbar_wait(&accum_empty, (nt & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kt = 0; kt < ks; kt++, it++) {
                    const int slot = it % GI8_STAGES;
                    mbar_wait(&full_bar[slot], (it / GI8_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = da0 + (uint64_t)((slot * STAGE) >> 4), db = db0 + (uint64_t)((slot * STAGE) >> 4);
                    // A slice s2 against the B slices t0 .. t0 + n - 1 (N = 64 n <= 256) -> the adjacent accumulators
                    // s2 + t0 ..: 8 MMAs per K step instead of 21 at S = 6, 40 % fewer shared-memory operand bytes
#pragma unroll
                    for (int s2 = 0; s2 < S; s2++)
#pragma unroll
                        for (int t0 = 0; t0 < S - s2; t0 += 4) {
                            const int n = (S - s2 - t0) < 4 ? (S - s2 - t0) : 4;
                            umma_i8(tmem + (s2 + t0) * I8_BN, da + (uint64_t)((s2 * I8_A_PLANE) >> 4),
                                    db + (uint64_t)((t0 * 1024) >> 4), IDESC0 | ((uint32_t)((n * I8_BN) >> 3) << 17),
                                    (kt > 0 || s2 > 0) ? 1u : 0u);
                        }
                    umma_commit(&empty_bar[slot]);
                }
                umma_commit(&accum_full);
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue warps: thread = row of the tile (TMEM lane) =====
        const int lg = warp & 3;
        const int r = lg * 32 + lane;
        int nt = 0;
        for (int u = blockIdx.x; u < nunits; u += gridDim.x, nt++) {
            int b, tm, tn;
            gi8_locate(g, u, b, tm, tn);
            const int m = tm * I8_BM + r, n0 = tn * I8_BN;
            // scales were written for every padded row, so these reads are always in range
            const double sa_ = g.alpha * g.ascale[b * g.as_bstride + m];
            const double *cs = g.bscale + b * g.bs_bstride + n0;
            mbar_wait(&accum_full, nt & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            double x[64];
#pragma unroll
            for (int ch = 0; ch < 4; ch++) {
                double acc[16];
                i8_recombine16<S>(tmem + ((uint32_t)(lg * 32) << 16), ch * 16, acc);
#pragma unroll
                for (int j = 0; j < 16; j++) x[ch * 16 + j] = acc[j];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");          // all four warps have drained TMEM
            if (warp == 4 && lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&accum_empty)) : "memory");
            if (m >= g.M) continue;
            double *crow = g.C + b * g.c_bstride + (int64_t)m * g.ldc + n0;
            if (g.mode == 0) {
                const bool vec = ((g.ldc | g.c_bstride) & 1) == 0 && n0 + 64 <= g.N;
                double vmax = 0.0;
                if (vec) {
#pragma unroll
                    for (int j = 0; j < 64; j += 2) {
                        const double2 s2 = *reinterpret_cast<const double2 *>(cs + j);
                        const double v0 = x[j] * sa_ * s2.x, v1 = x[j + 1] * sa_ * s2.y;
                        vmax = fmax(vmax, fmax(fabs(v0), fabs(v1)));
                        *reinterpret_cast<double2 *>(crow + j) = make_double2(v0, v1);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 64; j++)
                        if (n0 + j < g.N) {
                            const double v0 = x[j] * sa_ * cs[j];
                            vmax = fmax(vmax, fabs(v0));
                            crow[j] = v0;
                        }
                }
                if (g.rowmax) atomicMax(g.rowmax + b * g.rm_bstride + m / g.rm_div, (unsigned long long)__double_as_longlong(vmax));
            } else {
#pragma unroll
                for (int j = 0; j < 64; j++)
                    if (n0 + j < g.N) atomicAdd(crow + j, x[j] * sa_ * cs[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- C ABI ---------------------------------------------------------------------------------------------
template <int S, int W>
static int i8_slice_launch(const I8SliceArgs &a, int kcontig, cudaStream_t st) {
    if (kcontig) {
        dim3 grid((unsigned)((a.Rpad + 7) / 8), (unsigned)a.nbatch);
        i8_slice_kc_kernel<S, W><<<grid, 256, 0, st>>>(a);
    } else {
        dim3 grid((unsigned)((a.Rpad + 127) / 128), (unsigned)a.nbatch);
        i8_slice_rc_kernel<S, W><<<grid, 128, 0, st>>>(a);
    }
    return 0;
}

extern "C" int b200qc_i8_slice(const double *src, int nbatch, int64_t sb, int64_t sr, int64_t sk, int pair_mode,
                               int64_t pair_ld, int R, int K, int K_last, int Rpad, int Kpad, int tile_rows,
                               int nslice, signed char *planes, double *scales, void *stream) {
    QC_REQUIRE(nslice == 5 || nslice == 6, "nslice must be 5 or 6");
    QC_REQUIRE(tile_rows == 128 || tile_rows == 64, "tile_rows must be 128 (A operand) or 64 (B operand)");
    QC_REQUIRE(Rpad % tile_rows == 0 && Kpad % 32 == 0 && R <= Rpad && K <= Kpad && K_last <= K && K_last >= 0,
               "padded sizes must be multiples of the tile (rows) and of 32 (K)");
    QC_REQUIRE(nbatch >= 1 && nbatch <= 65535, "1..65535 batches");
    I8SliceArgs a;
    a.src = src; a.sb = sb; a.sr = sr; a.sk = sk; a.pair_ld = pair_ld; a.pair_mode = pair_mode;
    a.nbatch = nbatch; a.R = R; a.K = K; a.K_last = K_last; a.Rpad = Rpad; a.Kpad = Kpad;
    a.planes = planes; a.scales = scales;
    const int kcontig = (!pair_mode && sk == 1) ? 1 : 0;
    cudaStream_t st = as_stream(stream);
    prof_begin(PROF_I8_SLICE, st);
    if (nslice == 5) {
        if (tile_rows == 128) i8_slice_launch<5, 128>(a, kcontig, st); else i8_slice_launch<5, 64>(a, kcontig, st);
    } else {
        if (tile_rows == 128) i8_slice_launch<6, 128>(a, kcontig, st); else i8_slice_launch<6, 64>(a, kcontig, st);
    }
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

extern "C" int b200qc_i8_slice_dual(const double *src, int nbatch, int64_t sb, int64_t sr, const double *rowmax,
                                    int64_t rm_ld, int R, int K, int K_last, int Kpad, int nslice,
                                    signed char *planesA, double *scalesA, signed char *planesB, double *scalesB,
                                    void *stream) {
    QC_REQUIRE(nslice == 5 || nslice == 6, "nslice must be 5 or 6");
    QC_REQUIRE(Kpad % 32 == 0 && K <= Kpad && K_last <= K && K_last >= 0, "Kpad must be a multiple of 32");
    QC_REQUIRE(nbatch >= 1 && nbatch <= 65535, "1..65535 batches");
    I8Slice2Args a;
    a.src = src; a.sb = sb; a.sr = sr; a.rowmax = rowmax; a.rm_ld = rm_ld;
    a.nbatch = nbatch; a.R = R; a.K = K; a.K_last = K_last; a.Kpad = Kpad;
    a.RpadA = (R + 127) / 128 * 128; a.RpadB = (R + 63) / 64 * 64;
    a.planesA = planesA; a.planesB = planesB; a.scalesA = scalesA; a.scalesB = scalesB;
    cudaStream_t st = as_stream(stream);
    dim3 grid((unsigned)(a.RpadA / 8), (unsigned)nbatch);
    prof_begin(PROF_I8_SLICE, st);
    if (nslice == 5) i8_slice_dual_kernel<5><<<grid, 256, 0, st>>>(a);
    else i8_slice_dual_kernel<6><<<grid, 256, 0, st>>>(a);
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

extern "C" int b200qc_gemm_i8(const signed char *aplanes, const double *ascale, int64_t a_bstride, int64_t as_bstride,
                              const signed char *bplanes, const double *bscale, int64_t b_bstride, int64_t bs_bstride,
                              int nbatch, int mtiles, int ntiles, int nk, int nk_last, int nslice, int M, int N,
                              double alpha, double *C, int64_t c_bstride, int64_t ldc, int mode, double *rowmax,
                              int64_t rm_bstride, int rm_div, void *stream) {
    QC_REQUIRE(nslice == 5 || nslice == 6, "nslice must be 5 or 6");
    QC_REQUIRE(rowmax == nullptr || (mode == 0 && rm_div >= 1), "rowmax needs mode 0 and rm_div >= 1");
    QC_REQUIRE(mode >= 0 && mode <= 2, "mode: 0 store, 1 atomic add, 2 atomic add of lower-triangle tiles");
    QC_REQUIRE(nk >= 1 && nk_last >= 1 && nk_last <= nk, "every batch needs at least one K step");
    // exactness: |anti-diagonal sum| <= S * 64 * 64 * K < 2^31 and the int64 merge needs K * 2^(12 + 7 (S - 1)) < 2^63
    QC_REQUIRE((int64_t)nk * 32 <= (nslice == 6 ? 32768 : 65536), "K per batch too long for exact integer accumulation");
    QC_REQUIRE(M <= mtiles * I8_BM && N <= ntiles * I8_BN, "valid sizes exceed the tiled sizes");
    I8GemmArgs g;
    g.A = aplanes; g.B = bplanes; g.ascale = ascale; g.bscale = bscale;
    g.a_bstride = a_bstride; g.b_bstride = b_bstride; g.as_bstride = as_bstride; g.bs_bstride = bs_bstride;
    g.nbatch = nbatch; g.mtiles = mtiles; g.ntiles = ntiles; g.nk = nk; g.nk_last = nk_last;
    g.M = M; g.N = N; g.C = C; g.c_bstride = c_bstride; g.ldc = ldc; g.mode = mode; g.alpha = alpha;
    g.rowmax = reinterpret_cast<unsigned long long *>(rowmax); g.rm_bstride = rm_bstride; g.rm_div = rm_div > 0 ? rm_div : 1;
    if (mode == 2) {
        int c = 0;
        for (int tm = 0; tm < mtiles; tm++) c += std::min(ntiles, 2 * tm + 2);
        g.units_per_batch = c;
    } else {
        g.units_per_batch = mtiles * ntiles;
    }
    QC_REQUIRE((int64_t)nbatch * g.units_per_batch < (1LL << 31), "too many tiles");
    if (nbatch == 0 || g.units_per_batch == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const size_t smem = (size_t)GI8_STAGES * nslice * (I8_A_PLANE + I8_B_PLANE);
    const int nunits = nbatch * g.units_per_batch;
    const int grid = nunits < NUM_SMS ? nunits : NUM_SMS;
    prof_begin(PROF_GEMM_I8, st);
    if (nslice == 5) {
        QC_CHECK(cudaFuncSetAttribute(gemm_i8_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_i8_kernel<5><<<grid, I8_THREADS, smem, st>>>(g);
    } else {
        QC_CHECK(cudaFuncSetAttribute(gemm_i8_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_i8_kernel<6><<<grid, I8_THREADS, smem, st>>>(g);
    }
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}
