// K2 on the 5th-generation tensor cores: X = phi_sb D_sb as an error-free sliced int8 GEMM (same Ozaki
// scheme and pipeline as vxc_i8.cuh), with the density contraction fused into the epilogue:
//     rho_g = sum_nu X_g,nu phi_g,nu,     grad_d rho_g = 2 sum_nu X_g,nu d_d phi_g,nu.
// GEMM roles: M = 128 grid points (TMEM lanes), N = 64 AOs per tile, K = the kept AOs of the superblock.
// Both operands are K-major: A = phi rows (sliced once at set-up, scale per grid point), B = rows of the
// gathered symmetric density D_sb (gathered + sliced every call, scale per row).  A work unit is one
// (superblock, 128-row block); its CTA walks the N tiles, the epilogue warps drain the S int32
// accumulators to fp64 registers (freeing TMEM for the next N tile at once), then multiply with the fp64
// AO values of the same rows (prefetched into L2 while the MMAs of the tile still run, read with 256-bit
// streaming loads) and keep the four row sums in registers across N tiles -- X never leaves the SM, exactly
// like the DMMA kernel (rho.cuh), and no atomics are needed.  S = 5 by default (E_xc to 3e-10 Ha of the fp64
// kernel at C60).  The kernel is HBM-bound at C60: the A tile of a unit is streamed once per N tile and does
// not survive in L2 in between (ncu: 46.8 GB at 94 % of the HBM peak against 23.5 GB algorithmic; the
// variants tried against that are listed in DESIGN.md section 7).
#pragma once
#include "vxc_i8.cuh"

#define RI8_STAGES 5      // no staging tile here: 5-stage ring (3..5 measure the same; 7 starves L1)

// ---- phi rows -> S int8 planes, K-major tiles of RT grid rows; scale per grid row ----
// RT = 128: the A operand (M tile) of rho_i8_kernel
//   out (per SB, bytes): [row tile = g / 128][k tile = mu / 32][slice][(mu % 32) / 16][(g % 128) / 8][g % 8][mu % 16]
// RT = 64: the stationary B operand (N tile) of rho_i8_ps_kernel, slices INSIDE the K chunk so that consecutive slices
// of the 64 rows form one K-major operand of 64 n rows (rho_i8_ps.cuh, "concatenated B slices")
//   out (per SB, bytes): [row tile = g / 64][k tile = mu / 32][(mu % 32) / 16][slice][(g % 64) / 8][g % 8][mu % 16]
template <int S, int RT>
__global__ void __launch_bounds__(256)
sb_slice_rows_kernel(const SBDesc *__restrict__ sbd, const double *__restrict__ ao, int sbp,
                     const int64_t *__restrict__ p_off, signed char *__restrict__ planes, double *__restrict__ rscale) {
    const int sb = blockIdx.y;
    const SBDesc d = sbd[sb];
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);     // one warp per grid row
    if (g >= sbp) return;
    const int64_t ld = d.nsp;
    const double *X = ao + d.ao_off + (int64_t)g * ld;
    double m = 0.0;
    for (int c = lane; c < d.nsp; c += 32) m = fmax(m, fabs(X[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    int e = 0;
    if (m > 0.0) frexp(m, &e);
    const double inv = ldexp(64.0, -e);
    if (lane == 0) rscale[(int64_t)sb * sbp + g] = ldexp(1.0, e);
    const int nkt = d.nsp / I8_KT;
    constexpr int PLANE = I8_KT * RT;
    signed char *P = planes + p_off[sb] + (int64_t)(g / RT) * nkt * S * PLANE + ((g % RT) >> 3) * 128 + (g & 7) * 16;
    for (int c = lane; c < d.nsp; c += 32) {
        double y = X[c] * inv;
        // byte strides of the K chunk and of the slice inside a K tile
        constexpr int KC = (RT == 64) ? S * RT * 16 : RT * 16, SL = (RT == 64) ? RT * 16 : PLANE;
        signed char *Q = P + (int64_t)(c >> 5) * S * PLANE + ((c & 31) >> 4) * KC + (c & 15);
#pragma unroll
        for (int s = 0; s < S; s++) {
            Q[s * SL] = (signed char)slice_digit(y);
        }
    }
}

// ---- B operand: D_sb[nu][mu] = D[idx_nu][idx_mu] gathered and sliced, K-major tiles; scale per row nu ----
// out (per SB, bytes): [n tile = nu / 64][k tile = mu / 32][slice][(mu % 32) / 16][(nu % 64) / 8][nu % 8][mu % 16]
// (Measured alternatives at C60, all slower than this plain two-pass, lane-per-column form at 1.9 ms: 16-byte stores
// staged through shared memory and a gathered row kept in registers between the passes (round 1: 2.5 - 3.9 ms); four
// columns per lane with 4-byte stores (2.2 ms); the whole row in registers with every gather in flight at once (2.4 ms).
// The kernel moves 3 GB of planes out and ~10 GB of L2-resident D in; neither LSU instructions nor latency bound it.)
template <int S, int BN>
__global__ void __launch_bounds__(256)
sb_gather_slice_dm_kernel(const SBDesc *__restrict__ sbd, const int *__restrict__ idx, const double *__restrict__ dm,
                          int nao, const int64_t *__restrict__ p_off, signed char *__restrict__ planes,
                          double *__restrict__ cscale) {
    const int sb = blockIdx.y;
    const SBDesc d = sbd[sb];
    const int lane = threadIdx.x & 31;
    const int nu = blockIdx.x * 8 + (threadIdx.x >> 5);    // one warp per row of D_sb
    if (nu >= d.nsp || d.dsb_idx_off != d.idx_off) return;  // same AO list as an earlier superblock: its planes are used
    const int *ix = idx + d.idx_off;
    const int a = ix[nu];
    const double *row = dm + (int64_t)(a < nao ? a : 0) * nao;
    double m = 0.0;
    for (int c = lane; c < d.nsp; c += 32) {
        const int b = ix[c];
        if (a < nao && b < nao) m = fmax(m, fabs(row[b]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    int e = 0;
    if (m > 0.0) frexp(m, &e);
    const double inv = ldexp(64.0, -e);
    if (lane == 0) cscale[d.idx_off + nu] = ldexp(1.0, e);
    const int nkt = d.nsp / I8_KT;
    constexpr int B_PLANE = I8_KT * BN;      // N tiles of BN rows (the last one zero-padded by the caller's memset)
    // BN = 128 (the A operand of rho_i8_ps_kernel): row nu = 32 lg + 4 q + r of a tile sits at position (TMEM lane)
    // 32 lg + q + 8 r, so that the four lanes q, q + 8, q + 16, q + 24 a thread of the epilogue drains with
    // tcgen05.ld.16x256b are four consecutive AO columns (one 256-bit load of the fp64 AO values)
    const int pos = (BN == I8_BM) ? ((nu % BN) & ~31) + (((nu & 31) >> 2) + 8 * (nu & 3)) : nu % BN;
    signed char *P = planes + p_off[sb] + (int64_t)(nu / BN) * nkt * S * B_PLANE + (pos >> 3) * 128 + (pos & 7) * 16;
    for (int c = lane; c < d.nsp; c += 32) {
        const int b = ix[c];
        double y = (a < nao && b < nao) ? row[b] * inv : 0.0;
        signed char *Q = P + (int64_t)(c >> 5) * S * B_PLANE + ((c & 31) >> 4) * (BN * 16) + (c & 15);
#pragma unroll
        for (int s = 0; s < S; s++) Q[s * B_PLANE] = (signed char)slice_digit(y);
    }
}

template <int S, int BN>
static void sb_gather_slice_dm_launch(const SBDesc *sbd, int nsb, int max_nsp, const int *idx, const double *dm, int nao,
                                      const int64_t *p_off, signed char *planes, double *cscale, cudaStream_t st) {
    dim3 gg((unsigned)(max_nsp / 8), (unsigned)nsb);
    sb_gather_slice_dm_kernel<S, BN><<<gg, 256, 0, st>>>(sbd, idx, dm, nao, p_off, planes, cscale);
}

// instruction descriptor: D = S32, A = B = signed int8, both K-major, N = 64, M = 128
#define RI8_IDESC ((2u << 4) | (1u << 7) | (1u << 10) | ((I8_BN >> 3) << 17) | ((I8_BM >> 4) << 24))

// MC = 1: launched in 2-CTA clusters; both CTAs work on the same (superblock, 128-row block), CTA r takes the N
// tiles 2 p + r, every A stage is fetched half by each CTA and multicast to both, and the two partial row sums
// are added atomically into the zero-initialised outputs (two addends: the sum does not depend on the order).
__device__ __forceinline__ void ldcs_f64x4(const double *p, double &a, double &b, double &c, double &d) {
    // one full 32-byte sector per lane (rows of different lanes are 4 KB apart: nothing else to coalesce)
    asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// EH = number of column halves the epilogue splits an N tile into (1: 4 epilogue warps x 64 columns,
// 2: 8 warps x 32 columns); the CTA has 128 + 128 EH threads.
// BN = N tile: 64, or 96 with S = 5 and EH = 2 (480 TMEM columns): the A tile of a unit is then streamed
// ceil(nsp / 96) instead of nsp / 64 times -- a third less of the HBM traffic that bounds this kernel.
// NCACHE > 0 (MC = 0 only): the first NCACHE K steps of the unit's A tile are fetched once, with the first N tile,
// into a shared-memory region behind a 4-stage ring and reused by the other N tiles -- the A tile is what this
// kernel re-reads from HBM (it does not survive in L2 between N tiles).
template <int S, int NCOMP, int MC, int EH, int BN, int NCACHE>
__global__ void __launch_bounds__(128 + 128 * EH, 1)
rho_i8_kernel(const SBDesc *__restrict__ sbd, int nsb, int sbp, const double *__restrict__ ao,
              const signed char *__restrict__ aplanes, const int64_t *__restrict__ a_off,
              const signed char *__restrict__ bplanes, const int64_t *__restrict__ b_off,
              const double *__restrict__ rscale, const double *__restrict__ cscale, int64_t ngrid_ld,
              double *__restrict__ rho, double *__restrict__ grad, int l2hint) {
    // (l2hint: bit 0 = L2 evict_last on the A planes; bits 8.. = timing-experiment variant)
    extern __shared__ __align__(1024) unsigned char i8_smem[];
    static_assert(S * BN <= 512 && BN % (8 * EH) == 0, "accumulators exceed the tensor memory");
    constexpr int B_PLANE = I8_KT * BN;
    constexpr int A_STAGE = S * I8_A_PLANE, B_STAGE = S * B_PLANE, STAGE = A_STAGE + B_STAGE;
    // instruction descriptor: D = S32, A = B = signed int8, both K-major, N = BN, M = 128
    constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);
    // ring depth (run-time, 2..8): what the ring does not take of the 228 KB stays L1 cache for the epilogue's AO reads
    const int NST = NCACHE ? 4 : ((l2hint >> 16) & 15);
    static_assert(NCACHE == 0 || MC == 0, "the A cache is not combined with the multicast pairs");
    __shared__ uint64_t full_bar[8], empty_bar[8], accum_full, accum_empty;
    __shared__ uint32_t tmem_base_smem;
    __shared__ double comb[I8_BM][NCOMP];     // row sums of the second column half, handed to the first
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mtiles = sbp / I8_BM;
    const int nunits = nsb * mtiles;
    const int crank = MC ? (int)cluster_ctarank() : 0;
    const int u0 = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, ustep = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int A_HALF = A_STAGE / 2;

    if (tid == 0) {
        for (int i = 0; i < NST; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], MC ? 2 : 1);
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
    if (MC) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    const uint32_t sbase = smem_u32(i8_smem);
    const uint32_t cbase = sbase + NST * STAGE;      // A cache (NCACHE x A_STAGE bytes), behind the ring
    const int tstep = MC ? 2 : 1;        // N tiles per step of the (pair of) CTA(s)
    // K steps of a unit served from the cache.  A cache slot is rewritten by the first N tile of the next unit, nkt
    // stages after its last reader: the ring's empty barrier (NST stages back) covers that only if nkt >= NST.
    auto ncached = [&](int nkt) { return (NCACHE && nkt >= NST) ? min(NCACHE, nkt) : 0; };
    const int variant = (l2hint >> 8) & 255;     // timing experiments (b200qc_i8_debug_variant): 1 no AO loads, 2 no MMAs, 3 no epilogue work
    l2hint &= 1;

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            int it = 0;
            const uint64_t pol = l2_policy_evict_last();
            for (int u = u0; u < nunits; u += ustep) {
                const int sb = u / mtiles, mt = u - sb * mtiles;
                const int nsp = sbd[sb].nsp, nkt = nsp / I8_KT, ntn = (nsp + BN - 1) / BN;
                const signed char *A = aplanes + a_off[sb] + (int64_t)mt * nkt * A_STAGE;
                const signed char *B = bplanes + b_off[sb];
                const int nc = ncached(nkt);
                for (int t0 = 0; t0 < ntn; t0 += tstep) {
                    const int tn = t0 + crank;
                    const bool active = tn < ntn;
                    for (int kt = 0; kt < nkt; kt++, it++) {
                        const int slot = it % NST;
                        mbar_wait(&empty_bar[slot], ((it / NST) & 1) ^ 1);
                        if (NCACHE && kt < nc) {
                            // cached K step: the A stage travels only with the first N tile, into the cache
                            if (t0 == 0) {
                                mbar_expect_tx(&full_bar[slot], STAGE);
                                bulk_g2s(cbase + kt * A_STAGE, A + (int64_t)kt * A_STAGE, A_STAGE, &full_bar[slot]);
                            } else {
                                mbar_expect_tx(&full_bar[slot], B_STAGE);
                            }
                            bulk_g2s(sbase + slot * STAGE + A_STAGE, B + ((int64_t)tn * nkt + kt) * B_STAGE, B_STAGE,
                                     &full_bar[slot]);
                        } else if (MC) {
                            mbar_expect_tx(&full_bar[slot], A_STAGE + (active ? B_STAGE : 0));
                            bulk_g2s_mc(sbase + slot * STAGE + crank * A_HALF, A + (int64_t)kt * A_STAGE + crank * A_HALF,
                                        A_HALF, &full_bar[slot], (uint16_t)3);
                            if (active)
                                bulk_g2s(sbase + slot * STAGE + A_STAGE, B + ((int64_t)tn * nkt + kt) * B_STAGE, B_STAGE,
                                         &full_bar[slot]);
                        } else {
                            mbar_expect_tx(&full_bar[slot], STAGE);
                            // the A tile of a unit is read once per N tile: ask L2 to keep it (mode bit 1)
                            if (l2hint) bulk_g2s_hint(sbase + slot * STAGE, A + (int64_t)kt * A_STAGE, A_STAGE, &full_bar[slot], pol);
                            else bulk_g2s(sbase + slot * STAGE, A + (int64_t)kt * A_STAGE, A_STAGE, &full_bar[slot]);
                            bulk_g2s(sbase + slot * STAGE + A_STAGE, B + ((int64_t)tn * nkt + kt) * B_STAGE, B_STAGE,
                                     &full_bar[slot]);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // K-major, no swizzle: LBO = stride between the two 16-byte K chunks of a K = 32 step
            // (A: 2048 B, B: 1024 B), SBO = stride between 8-row groups (128 B)
            const uint64_t da0 = umma_desc(sbase, 2048, 128), db0 = umma_desc(sbase + A_STAGE, BN * 16, 128);
            int it = 0, nt = 0;
            for (int u = u0; u < nunits; u += ustep) {
                const int sb = u / mtiles;
                const int nsp = sbd[sb].nsp, nkt = nsp / I8_KT, ntn = (nsp + BN - 1) / BN;
                const int nc = ncached(nkt);
                for (int t0 = 0; t0 < ntn; t0 += tstep) {
                    const bool active = t0 + crank < ntn;
                    if (active) {
                        mbar_wait(&accum_empty, (nt & 1) ^ 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    for (int kt = 0; kt < nkt; kt++, it++) {
                        const int slot = it % NST;
                        mbar_wait(&full_bar[slot], (it / NST) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t db = db0 + (uint64_t)((slot * STAGE) >> 4);
                        const uint64_t da = (NCACHE && kt < nc) ? umma_desc(cbase + kt * A_STAGE, 2048, 128)
                                                                : da0 + (uint64_t)((slot * STAGE) >> 4);
                        if (active && variant != 2) {
#pragma unroll
                            for (int dd = 0; dd < S; dd++)
#pragma unroll
                                for (int s2 = 0; s2 <= dd; s2++)
                                    umma_i8(tmem + dd * BN, da + (uint64_t)((s2 * I8_A_PLANE) >> 4),
                                            db + (uint64_t)(((dd - s2) * B_PLANE) >> 4), IDESC,
                                            (kt > 0 || s2 > 0) ? 1u : 0u);
                        }
                        if (MC) umma_commit_mc(&empty_bar[slot], (uint16_t)3); else umma_commit(&empty_bar[slot]);
                    }
                    if (active) {
                        umma_commit(&accum_full);
                        nt++;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; thread = grid row (TMEM lane quarter = warp % 4) x one 32-column half of the N tile.
        // The dot products stream the fp64 AO values straight from HBM: twice the warps = twice the loads in flight.
        constexpr int NC = BN / EH;             // columns of the N tile per thread
        const int lg = warp & 3, half = (warp - 4) >> 2;
        const int r = lg * 32 + lane;
        const int c0 = half * NC;
        int nt = 0;
        for (int u = u0; u < nunits; u += ustep) {
            const int sb = u / mtiles, mt = u - sb * mtiles;
            const SBDesc d = sbd[sb];
            const int ntn = (d.nsp + BN - 1) / BN;
            const int64_t ld = d.nsp;
            const int grow = mt * I8_BM + r;                             // row inside the superblock
            const double *phi = ao + d.ao_off + (int64_t)grow * ld;      // component 0, this row
            double part[NCOMP];
#pragma unroll
            for (int c = 0; c < NCOMP; c++) part[c] = 0.0;
            for (int tn = crank; tn < ntn; tn += tstep, nt++) {
                // columns of this thread that exist (the last N tile may reach past nsp; multiples of 16)
                const int nvalid = min(NC, d.nsp - (tn * BN + c0));
                // the AO values this thread will need for the tile: pull them into L2 while the MMAs still run
                if (variant != 1 && variant != 3) {
#pragma unroll
                    for (int c = 0; c < NCOMP; c++)
#pragma unroll
                        for (int j = 0; j < NC; j += 16)
                            if (j < nvalid)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(phi + (int64_t)c * sbp * ld + tn * BN + c0 + j));
                }
                mbar_wait(&accum_full, nt & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                double x[NC];
                if (variant == 3) {
#pragma unroll
                    for (int j = 0; j < NC; j++) x[j] = 1.0;
                } else {
#pragma unroll
                    for (int ch = 0; ch < NC / 8; ch++) {
                        double acc[8];
                        i8_recombine8<S>(tmem + ((uint32_t)(lg * 32) << 16), BN, c0 + ch * 8, acc);
#pragma unroll
                        for (int j = 0; j < 8; j++) x[ch * 8 + j] = acc[j];
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(128 * EH) : "memory");   // all epilogue warps have drained TMEM
                if (warp == 4 && lane == 0)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&accum_empty)) : "memory");
                // column scales, then the row dots with the fp64 AO values of this row
                const int n0 = tn * BN + c0;
                const double *cs = cscale + d.dsb_idx_off + n0;
                if (variant == 1 || variant == 3) {
#pragma unroll
                    for (int j = 0; j < NC; j++) part[0] += x[j];
                    continue;
                }
#pragma unroll
                for (int j = 0; j < NC; j += 2) {
                    if (j < nvalid) {
                        const double2 s2 = *reinterpret_cast<const double2 *>(cs + j);
                        x[j] *= s2.x;
                        x[j + 1] *= s2.y;
                    }
                }
#pragma unroll
                for (int c = 0; c < NCOMP; c++) {
                    const double *row = phi + (int64_t)c * sbp * ld + n0;
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < NC; j += 4) {
                        // streaming, zero-reuse 256-bit reads (kept from evicting the re-used int8 planes from L2)
                        if (j < nvalid) {
                            double v0, v1, v2, v3;
                            ldcs_f64x4(row + j, v0, v1, v2, v3);
                            s += x[j] * v0 + x[j + 1] * v1 + x[j + 2] * v2 + x[j + 3] * v3;
                        }
                    }
                    part[c] += s;
                }
            }
            // the two column halves of a row meet in shared memory (the next write of `comb` is ordered behind this
            // read by the bar.sync 1 of the next unit's first N tile)
            if (EH == 2) {
                if (half == 1) {
#pragma unroll
                    for (int c = 0; c < NCOMP; c++) comb[r][c] = part[c];
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            if (half == 0) {
                if (EH == 2) {
#pragma unroll
                    for (int c = 0; c < NCOMP; c++) part[c] += comb[r][c];
                }
                const double sa_ = rscale[(int64_t)sb * sbp + grow];
                const int64_t g = (int64_t)sb * sbp + grow;
                if (MC) {
                    atomicAdd(rho + g, sa_ * part[0]);
                    if (NCOMP == 4) {
#pragma unroll
                        for (int dd = 0; dd < 3; dd++) atomicAdd(grad + (int64_t)dd * ngrid_ld + g, 2.0 * sa_ * part[dd + 1]);
                    }
                } else {
                    rho[g] = sa_ * part[0];
                    if (NCOMP == 4) {
#pragma unroll
                        for (int dd = 0; dd < 3; dd++) grad[(int64_t)dd * ngrid_ld + g] = 2.0 * sa_ * part[dd + 1];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (MC) cluster_sync_all();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// Slices the static AO values (component 0) of every superblock once, row-wise, into K-major tiles of `row_tile`
// grid rows (128: A operand of rho_i8_kernel; 64: stationary B operand of rho_i8_ps_kernel):
// aplanes = sum_sb nslice * sbp * nsp bytes at a_off[sb]; rscale = nsb * sbp doubles.
extern "C" int b200qc_rho_i8_prepare(const void *sbdesc, int nsb, int sbp, int nslice, int row_tile, const double *ao,
                                     const int64_t *a_off, signed char *aplanes, double *rscale, void *stream) {
    QC_REQUIRE(nslice == 5 || nslice == 6, "nslice must be 5 or 6");
    QC_REQUIRE(row_tile == 128 || row_tile == 64, "row tile must be 128 or 64");
    QC_REQUIRE(sbp % I8_BM == 0, "superblock size must be a multiple of 128");
    if (nsb == 0) return 0;
    dim3 grid((unsigned)(sbp / 8), (unsigned)nsb);
    const SBDesc *sbd = (const SBDesc *)sbdesc;
    cudaStream_t st = as_stream(stream);
    if (nslice == 5 && row_tile == 128) sb_slice_rows_kernel<5, 128><<<grid, 256, 0, st>>>(sbd, ao, sbp, a_off, aplanes, rscale);
    else if (nslice == 5) sb_slice_rows_kernel<5, 64><<<grid, 256, 0, st>>>(sbd, ao, sbp, a_off, aplanes, rscale);
    else if (row_tile == 128) sb_slice_rows_kernel<6, 128><<<grid, 256, 0, st>>>(sbd, ao, sbp, a_off, aplanes, rscale);
    else sb_slice_rows_kernel<6, 64><<<grid, 256, 0, st>>>(sbd, ao, sbp, a_off, aplanes, rscale);
    QC_LAUNCHED(1);
    return 0;
}

template <int S, int BN>
static int rho_i8_run(const SBDesc *sbd, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                      const double *dm, int nao, const signed char *aplanes, const int64_t *a_off,
                      const double *rscale, signed char *bplanes, const int64_t *b_off, double *cscale, double *rho,
                      double *grad, cudaStream_t st) {
    prof_begin(PROF_SB_GATHER, st);
    sb_gather_slice_dm_launch<S, BN>(sbd, nsb, max_nsp, idx, dm, nao, b_off, bplanes, cscale, st);
    prof_end(st);
    QC_LAUNCHED(1);
    // mode bit 16: A cache (4-stage ring + 5 (S = 5) or 3 (S = 6) cached K steps of the A tile), 128 x 64 tiles only
    constexpr int NCA = (BN == 64) ? (S == 5 ? 5 : 3) : 0;
    const bool acache = (g_i8_mode & 16) != 0 && NCA > 0 && (g_i8_mode & 4) == 0;
    // ring depth: B200QC_I8_MODE bits 8..11 when set (experiments), else 5 stages
    int nst = (g_i8_mode >> 8) & 15;
    if (nst < 2 || nst > 8 || (size_t)nst * S * (I8_A_PLANE + I8_KT * BN) > 227 * 1024 - 2048) nst = RI8_STAGES;
    const size_t smem = acache ? (size_t)4 * S * (I8_A_PLANE + I8_KT * BN) + (size_t)NCA * S * I8_A_PLANE
                               : (size_t)nst * S * (I8_A_PLANE + I8_KT * BN);
    const int64_t ngl = (int64_t)nsb * sbp;
    const bool mc = (g_i8_mode & 4) != 0 && BN == 64;
    const int l2hint = (g_i8_mode & 1) | ((g_i8_variant & 255) << 8) | (nst << 16);
    if (mc) {   // partial row sums of the two CTAs of a pair are added atomically
        QC_CHECK(cudaMemsetAsync(rho, 0, sizeof(double) * ngl, st));
        if (grad) QC_CHECK(cudaMemsetAsync(grad, 0, sizeof(double) * 3 * ngl, st));
    }
    prof_begin(PROF_RHO, st);
#define RHO_I8_LAUNCH(NCOMP_, MC_, EH_) RHO_I8_LAUNCH_C(NCOMP_, MC_, EH_, 0)
#define RHO_I8_LAUNCH_C(NCOMP_, MC_, EH_, NC_)                                                                               \
    do {                                                                                                                    \
        QC_CHECK(cudaFuncSetAttribute(rho_i8_kernel<S, NCOMP_, MC_, EH_, BN, NC_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                      (int)smem));                                                                          \
        if (MC_)                                                                                                            \
            QC_CHECK(launch_cluster2(rho_i8_kernel<S, NCOMP_, MC_, EH_, BN, NC_>, NUM_SMS, 128 + 128 * EH_, smem, st, sbd, nsb, sbp, ao, \
                                     aplanes, a_off, (const signed char *)bplanes, b_off, rscale, (const double *)cscale,   \
                                     ngl, rho, grad, l2hint));                                                              \
        else                                                                                                                \
            rho_i8_kernel<S, NCOMP_, MC_, EH_, BN, NC_><<<NUM_SMS, 128 + 128 * EH_, smem, st>>>(sbd, nsb, sbp, ao, aplanes, a_off, \
                                                                                     bplanes, b_off, rscale, cscale, ngl,   \
                                                                                     rho, grad, l2hint);                    \
    } while (0)
    // EH = 1: four epilogue warps.  (Eight warps x 32 columns measured the same 8.5 ms at C60 once the AO rows
    // are prefetched into L2 and read with 256-bit loads, so the smaller CTA is used.)
    // (BN = 96: 48 columns per thread need the eight-warp epilogue.)
    constexpr int EHB = (BN == 96) ? 2 : 1;
    if (acache) {
        if (grad) RHO_I8_LAUNCH_C(4, 0, 1, NCA); else RHO_I8_LAUNCH_C(1, 0, 1, NCA);
    } else if (grad) {
        if (mc) RHO_I8_LAUNCH(4, (BN == 64 ? 1 : 0), 1); else RHO_I8_LAUNCH(4, 0, EHB);
    } else {
        if (mc) RHO_I8_LAUNCH(1, (BN == 64 ? 1 : 0), 1); else RHO_I8_LAUNCH(1, 0, EHB);
    }
#undef RHO_I8_LAUNCH
#undef RHO_I8_LAUNCH_C
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

