// K2, point-stationary form (the default): the same error-free sliced int8 density GEMM as rho_i8.cuh with the
// operand roles swapped so that every int8 plane of phi is read from HBM exactly once.
//
//     Xt[nu][g] = sum_mu D_sb[nu][mu] phi[g][mu]        (D symmetric: Xt = (phi D_sb)^T)
//     rho_g = sum_nu Xt[nu][g] phi[g][nu],   grad_d rho_g = 2 sum_nu Xt[nu][g] d_d phi[g][nu]
//
// GEMM roles: M = 128 AO rows nu of D_sb (TMEM lanes), N = 64 grid points (TMEM columns), K = the kept AOs mu.
// A work unit is one (superblock, 64-point tile).  Its B operand -- the S int8 planes of the 64 phi rows, all K
// steps: S * 64 * nsp bytes = 160 KB at nsp = 512 -- is fetched ONCE into a shared-memory cache and stays there
// while the CTA walks the nsp / 128 M tiles; the A operand (D_sb tiles, 1.3 MB per superblock, shared by the 8
// point tiles of the superblock and by the CTAs working on them, hence L2-resident) streams through a 3-stage
// ring.  rho_i8_kernel streamed the 128-row phi tile once per 64-wide N tile (8 times at C60) and those re-reads
// missed L2: 46.8 GB of DRAM reads against 23.5 GB algorithmic (profiles/r01_final_c60_ncu_full.txt).
//
// Superblocks with more than NBC * 32 kept AOs keep the first NBC - 3 K steps of B stationary and stream the rest
// through the last three cache slots once per M tile (from L2: the tile was just read).
//
// Epilogue (8 warps = lane quarter x column half): the accumulators are drained with tcgen05.ld.16x256b, whose
// fragment layout gives a thread 4 AO rows x 8 points (instead of the 1 row x 32 points of .32x32b), merged exactly and
// multiplied with the fp64 AO values phi_c[g][nu] straight into 8 x NCOMP per-thread accumulators that live across the
// M tiles of the unit: no cross-lane traffic per M tile at all (the first version reduced 32 x 32 values over the warp
// for every M tile -- 868 of its 2 600 instructions per thread and tile, and the kernel was bound by exactly that
// epilogue), one butterfly reduce-scatter (fixed order, no atomics) per unit.  TMEM is released to the next M tile as
// soon as it is in registers.
//
// Two things keep the tensor pipe fed (round 2):
//  * CL-CTA clusters (CL = 1, 2, 4): the CL CTAs of a cluster work on CL consecutive point tiles of the SAME superblock,
//    so they need the same D_sb stream; each CTA fetches 1 / CL of every A stage and multicasts it to the whole
//    cluster (cp.async.bulk ... .multicast::cluster), the stage is released cluster-wide by a multicast
//    tcgen05.commit.  L2 -> SM traffic of the A stream (25 GB of the 48 GB the kernel pulls at C60) divides by CL.
//  * concatenated B slices: tcgen05.mma re-reads both operands from shared memory for every instruction.  The S
//    planes of a B slot are laid out [K chunk][slice][row], so the slices t0 .. t0 + n - 1 of the 64 points form ONE
//    K-major operand of N = 64 n rows and one MMA (A slice s, N = 64 n <= 256) updates the n accumulators
//    s + t0 .. s + t0 + n - 1, which are adjacent TMEM columns: 6 MMAs per K step instead of 15 at S = 5, the same
//    tensor time, 40 % fewer shared-memory operand bytes (54 KB instead of 90 KB per K step).
#pragma once
#include "rho_i8.cuh"

#define RPS_BN 64          // grid points per tile (MMA N)
#define RPS_MAXST 8        // A-operand (D_sb) ring: up to 8 stages (run-time depth: what the B cache leaves)
#define RPS_MAXSLOT 16     // B cache slots (one K step of the 64-point tile each)
#define RPS_STREAM 3       // slots that turn into a ring when the tile does not fit
#define RPS_CLUSTER 2      // default cluster size of the multicast A stream
#define RPS_DEFAULT_SLOTS 16  // default B-cache slots

// exact int32 -> fp64 on the ALU + fp64 pipes (I2F.F64 is an XU instruction): 2^52 + 2^31 + v has v + 2^31 in its
// low mantissa word
__device__ __forceinline__ double i32_to_f64(int v) {
    return __hiloint2double(0x43300000, v ^ 0x80000000) - 4503601774854144.0;
}
// 16 TMEM lanes x 16 columns as two m16n8 fragments: regs 4 j + {0, 1}: (lane q, columns 8 j + 2 p + {0, 1}),
// regs 4 j + {2, 3}: (lane q + 8, same columns), q = laneid / 4, p = laneid % 4
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int S, int NCOMP, int CL>
__global__ void __launch_bounds__(384, 1)
rho_i8_ps_kernel(const SBDesc *__restrict__ sbd, int nsb, int sbp, const double *__restrict__ ao,
                 const signed char *__restrict__ pplanes, const int64_t *__restrict__ p_off,
                 const signed char *__restrict__ dplanes, const int64_t *__restrict__ d_off,
                 const double *__restrict__ rscale, const double *__restrict__ cscale, int64_t ngrid_ld,
                 double *__restrict__ rho, double *__restrict__ grad, int nbc, int nst, int variant) {
    extern __shared__ __align__(1024) unsigned char i8_smem[];
    constexpr int A_STAGE = S * I8_A_PLANE, A_PART = A_STAGE / CL, B_ROWS = RPS_BN * 16, B_SLOT = S * I8_KT * RPS_BN;
    static_assert(A_STAGE % (16 * CL) == 0, "the A stage splits into 16-byte aligned parts");
    // instruction descriptor without N: D = S32, A = B = signed int8, both K-major, M = 128
    constexpr uint32_t IDESC0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_BM >> 4) << 24);
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
    constexpr int NV = 8 * NCOMP, NVL = NV / 8;   // per-thread accumulators (8 points x components); values left per lane
    __shared__ uint64_t afull[RPS_MAXST], aempty[RPS_MAXST], bfull[RPS_MAXSLOT], bempty[RPS_MAXSLOT], accum_full, accum_empty;
    __shared__ uint32_t tmem_base_smem;
    __shared__ double comb[3][2][NVL][32];        // partial sums of lane quarters 1..3, handed to quarter 0
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ptiles = sbp / RPS_BN;
    // a cluster walks groups of CL consecutive point tiles of one superblock; CTA r of the cluster owns tile r of the group
    const int crank = (CL > 1) ? (int)cluster_ctarank() : 0;
    const int gps = ptiles / CL, ngroups = nsb * gps;
    const int g0 = (int)blockIdx.x / CL, gstep = (int)gridDim.x / CL;

    if (tid == 0) {
        for (int i = 0; i < RPS_MAXST; i++) {
            mbar_init(&afull[i], 1);
            mbar_init(&aempty[i], CL);            // every CTA of the cluster releases a stage
        }
        for (int i = 0; i < RPS_MAXSLOT; i++) {
            mbar_init(&bfull[i], 1);
            mbar_init(&bempty[i], 1);
        }
        mbar_init(&accum_full, 1);
        mbar_init(&accum_empty, 8);               // every epilogue warp releases TMEM on its own
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();               // the peers' barriers exist before anything is multicast to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    const uint32_t abase = smem_u32(i8_smem);                 // A ring
    const uint32_t bbase = abase + nst * A_STAGE;             // B cache: nbc slots
    // K steps of a unit whose B slot stays for the whole unit (the rest cycle through the last RPS_STREAM slots)
    auto ncached = [&](int nkt) { return nkt <= nbc ? nkt : nbc - RPS_STREAM; };

    // register budget: the producer / MMA warpgroup hands 112 registers per thread to the two epilogue warpgroups
    // (128 x 56 + 256 x 224 = 64512), which keep two chunks of AO values in flight beside the drained accumulators
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            int ait = 0;
            uint32_t pph = 0;          // per B slot: parity of the next "slot is free" wait
            for (int ug = g0; ug < ngroups; ug += gstep) {
                const int sb = ug / gps, pt = (ug - sb * gps) * CL + crank;
                const int nsp = sbd[sb].nsp, nkt = nsp / I8_KT, ntm = (nsp + I8_BM - 1) / I8_BM;
                const signed char *A = dplanes + d_off[sb];
                const signed char *B = pplanes + p_off[sb] + (int64_t)pt * nkt * B_SLOT;
                const int nc = ncached(nkt);
                int sit = 0;
                for (int mt = 0; mt < ntm; mt++) {
                    for (int kt = 0; kt < nkt; kt++, ait++) {
                        int bs = -1;
                        if (kt < nc) {
                            if (mt == 0) bs = kt;
                        } else {
                            bs = nc + sit % RPS_STREAM;
                            sit++;
                        }
                        if (bs >= 0) {
                            mbar_wait(&bempty[bs], ((pph >> bs) & 1u) ^ 1u);
                            pph ^= 1u << bs;
                            mbar_expect_tx(&bfull[bs], B_SLOT);
                            bulk_g2s(bbase + bs * B_SLOT, B + (int64_t)kt * B_SLOT, B_SLOT, &bfull[bs]);
                        }
                        const int slot = ait % nst;
                        mbar_wait(&aempty[slot], ((ait / nst) & 1) ^ 1);
                        mbar_expect_tx(&afull[slot], A_STAGE);
                        const signed char *src = A + ((int64_t)mt * nkt + kt) * A_STAGE + crank * A_PART;
                        if (CL > 1) bulk_g2s_mc(abase + slot * A_STAGE + crank * A_PART, src, A_PART, &afull[slot], CMASK);
                        else bulk_g2s(abase + slot * A_STAGE, src, A_STAGE, &afull[slot]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // K-major, no swizzle: LBO = stride between the two 16-byte K chunks of a K = 32 step
            // (A: 128 rows x 16 B = 2048; B: S slices x 64 rows x 16 B), SBO = stride between 8-row groups (128 B)
            const uint64_t da0 = umma_desc(abase, 2048, 128), db0 = umma_desc(bbase, S * B_ROWS, 128);
            int ait = 0, nt = 0;
            uint32_t cph = 0;          // per B slot: parity of the next "slot is filled" wait
            for (int ug = g0; ug < ngroups; ug += gstep) {
                const int sb = ug / gps;
                const int nsp = sbd[sb].nsp, nkt = nsp / I8_KT, ntm = (nsp + I8_BM - 1) / I8_BM;
                const int nc = ncached(nkt);
                int sit = 0;
                for (int mt = 0; mt < ntm; mt++, nt++) {
                    mbar_wait(&accum_empty, (nt & 1) ^ 1);           // the epilogue has drained the previous M tile
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kt = 0; kt < nkt; kt++, ait++) {
                        int bs;
                        bool first, last;
                        if (kt < nc) {
                            bs = kt;
                            first = mt == 0;
                            last = mt == ntm - 1;
                        } else {
                            bs = nc + sit % RPS_STREAM;
                            sit++;
                            first = last = true;
                        }
                        if (first) {
                            mbar_wait(&bfull[bs], (cph >> bs) & 1u);
                            cph ^= 1u << bs;
                        }
                        const int slot = ait % nst;
                        mbar_wait(&afull[slot], (ait / nst) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t da = da0 + (uint64_t)((slot * A_STAGE) >> 4), db = db0 + (uint64_t)((bs * B_SLOT) >> 4);
                        if (variant != 2) {
                            // A slice s2 against the B slices t0 .. t0 + n - 1 (one operand of 64 n rows) -> accumulators
                            // s2 + t0 .. s2 + t0 + n - 1 (adjacent TMEM columns); n <= 4 keeps N <= 256
#pragma unroll
                            for (int s2 = 0; s2 < S; s2++)
#pragma unroll
                                for (int t0 = 0; t0 < S - s2; t0 += 4) {
                                    const int n = (S - s2 - t0) < 4 ? (S - s2 - t0) : 4;
                                    umma_i8(tmem + (s2 + t0) * RPS_BN, da + (uint64_t)((s2 * I8_A_PLANE) >> 4),
                                            db + (uint64_t)((t0 * B_ROWS) >> 4), IDESC0 | ((uint32_t)((n * RPS_BN) >> 3) << 17),
                                            (kt > 0 || s2 > 0) ? 1u : 0u);
                                }
                        }
                        if (CL > 1) umma_commit_mc(&aempty[slot], CMASK); else umma_commit(&aempty[slot]);
                        if (last) umma_commit(&bempty[bs]);
                    }
                    umma_commit(&accum_full);
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ===== epilogue: TMEM lane quarter lg = warp % 4 (32 AO rows), column half = 32 of the 64 grid points =====
        // Fragment layout of tcgen05.ld.16x256b (measured: tools/tmem_layout.cu), q = lane / 4, p = lane % 4:
        //   TMEM lanes L = 32 lg + q + 8 r,          r = 0..3   (r = 0, 1: the load at lane base 0; 2, 3: lane base 16)
        //   points     g = 32 half + 8 k + 2 p + c,  k = 0..3, c = 0, 1
        // sb_gather_slice_dm_kernel stores row nu = 32 lg + 4 q + r of a D_sb tile at lane L (RPS_ROW_PERM), so the four
        // rows of a thread are four CONSECUTIVE AO columns: one 256-bit load per (component, point) instead of four
        // 64-bit ones (the LSU request queue, not bandwidth, was what the epilogue waited on: stall_lg 18 %, long
        // scoreboard 30 % in profiles/r02_k2_ps_v2_ncu.txt)
        const int lg = warp & 3, half = (warp - 4) >> 2;
        const int q = lane >> 2, p = lane & 3;
        const int et = tid - 128;                                    // 0..255: (component, point row) of the L2 prefetch
        const uint32_t tq = tmem + ((uint32_t)(lg * 32) << 16) + half * 32;
        int nt = 0;
        for (int ug = g0; ug < ngroups; ug += gstep) {
            const int sb = ug / gps, pt = (ug - sb * gps) * CL + crank;
            const SBDesc d = sbd[sb];
            const int ntm = (d.nsp + I8_BM - 1) / I8_BM;
            const int64_t ld = d.nsp, cstride = (int64_t)sbp * ld;
            // acc[c4 * 8 + k * 2 + c]: this thread's part (its 4 AO rows per M tile) of component c4 at point (k, c)
            double acc[NV];
#pragma unroll
            for (int i = 0; i < NV; i++) acc[i] = 0.0;
            // first AO row / point of this thread (component 0)
            const double *a0 = ao + d.ao_off + (int64_t)(pt * RPS_BN + half * 32 + 2 * p) * ld + lg * 32 + 4 * q;
            for (int mt = 0; mt < ntm; mt++, nt++) {
                const bool live = mt * I8_BM + lg * 32 < d.nsp;      // (nsp is a multiple of 64: a warp's 32 rows all exist or none)
                // optional: the fp64 AO values of the NEXT M tile (or of the first M tile of the next unit) start their way
                // from HBM into L2 now, one 1 KB row segment per thread, a whole MMA phase before they are read
                // (measured: off is faster -- 6.1 against 7.1 ms at C60 -- and the prefetched lines were often evicted again
                // before their use, 34 GB of DRAM reads instead of 24; kept as timing variant 5)
                if (variant == 5 && et < 64 * NCOMP) {
                    const int prow = et & 63, pc = et >> 6;
                    if (mt + 1 < ntm) {
                        const int w = min(I8_BM, d.nsp - (mt + 1) * I8_BM);
                        l2_prefetch_bulk(ao + d.ao_off + pc * cstride + (int64_t)(pt * RPS_BN + prow) * ld + (mt + 1) * I8_BM, w * 8);
                    } else if (ug + gstep < ngroups) {
                        const int ug2 = ug + gstep, sb2 = ug2 / gps, pt2 = (ug2 - sb2 * gps) * CL + crank;
                        const SBDesc d2 = sbd[sb2];
                        l2_prefetch_bulk(ao + d2.ao_off + (int64_t)pc * sbp * d2.nsp + (int64_t)(pt2 * RPS_BN + prow) * d2.nsp,
                                         min(I8_BM, d2.nsp) * 8);
                    }
                }
                // row scales of D_sb times the weight 2^(-12 - 7 (S - 1)) of the last anti-diagonal
                double cs[4];
#pragma unroll
                for (int r = 0; r < 4; r++)
                    cs[r] = live ? ldexp(cscale[d.dsb_idx_off + mt * I8_BM + lg * 32 + 4 * q + r], -12 - 7 * (S - 1)) : 0.0;
                mbar_wait(&accum_full, nt & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                double x[32];                                        // x[(k * 4 + r) * 2 + c]
                if (variant == 3) {
#pragma unroll
                    for (int j = 0; j < 32; j++) x[j] = 1.0;
                } else {
#pragma unroll
                    for (int kp = 0; kp < 2; kp++) {                 // 16 columns (two column groups k) at a time
                        uint32_t v[S][2][8];
#pragma unroll
                        for (int dd = 0; dd < S; dd++)
#pragma unroll
                            for (int bb = 0; bb < 2; bb++)
                                tmem_ld_16x256b_x2(tq + ((uint32_t)(bb * 16) << 16) + dd * RPS_BN + kp * 16, v[dd][bb]);
                        tmem_wait_ld();
#pragma unroll
                        for (int bb = 0; bb < 2; bb++)
#pragma unroll
                            for (int rg = 0; rg < 8; rg++) {
                                const int k = kp * 2 + (rg >> 2), r = bb * 2 + ((rg >> 1) & 1), c = rg & 1;
                                // exact merge of the S anti-diagonal sums, Horner in fp64 (every partial result is an
                                // integer below 2^53 for the K limits the host enforces)
                                double t = i32_to_f64((int)v[0][bb][rg]);
#pragma unroll
                                for (int dd = 1; dd < S; dd++) t = fma(t, 128.0, i32_to_f64((int)v[dd][bb][rg]));
                                x[(k * 4 + r) * 2 + c] = t * cs[r];
                            }
                    }
                }
                // this warp's part of TMEM is in registers: no CTA-wide barrier here, the eight epilogue warps run out of
                // phase with each other (one drains while another waits on its AO loads)
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&accum_empty)) : "memory");
                if (variant == 1 || variant == 3) {
#pragma unroll
                    for (int j = 0; j < 32; j++) acc[0] += x[j];
                    continue;
                }
                if (!live) continue;
                // row dots: 4 AO rows x 8 points x NCOMP components, straight FMAs into the per-thread accumulators
                // (lanes of equal p read 8 consecutive AO values = two full sectors of one grid row)
                const double *am = a0 + mt * I8_BM;
#pragma unroll
                for (int c4 = 0; c4 < NCOMP; c4++) {
                    double a[32];
#pragma unroll
                    for (int k = 0; k < 4; k++)
#pragma unroll
                        for (int c = 0; c < 2; c++)
                            ldcs_f64x4(am + c4 * cstride + (int64_t)(8 * k + c) * ld, a[(k * 4 + 0) * 2 + c], a[(k * 4 + 1) * 2 + c],
                                       a[(k * 4 + 2) * 2 + c], a[(k * 4 + 3) * 2 + c]);
#pragma unroll
                    for (int k = 0; k < 4; k++)
#pragma unroll
                        for (int c = 0; c < 2; c++)
#pragma unroll
                            for (int r = 0; r < 4; r++)
                                acc[c4 * 8 + k * 2 + c] = fma(x[(k * 4 + r) * 2 + c], a[(k * 4 + r) * 2 + c], acc[c4 * 8 + k * 2 + c]);
                }
            }
            // one reduction per unit: over the 8 lanes of equal p (butterfly reduce-scatter over lane bits 4, 3, 2; fixed
            // order, no atomics), then over the four lane quarters through shared memory
#pragma unroll
            for (int rd = 0; rd < 3; rd++) {
                const int lm = 16 >> rd, hv = NV >> (rd + 1);
                const bool upper = (lane & lm) != 0;
#pragma unroll
                for (int i = 0; i < hv; i++) {
                    const double send = upper ? acc[i] : acc[i + hv];
                    const double keep = upper ? acc[i + hv] : acc[i];
                    acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, lm);
                }
            }
            // this lane now holds the values i = q * NVL + j, j < NVL
            if (lg > 0) {
#pragma unroll
                for (int j = 0; j < NVL; j++) comb[lg - 1][half][j][lane] = acc[j];
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (lg == 0) {
#pragma unroll
                for (int j = 0; j < NVL; j++) {
                    const double t = ((acc[j] + comb[0][half][j][lane]) + comb[1][half][j][lane]) + comb[2][half][j][lane];
                    const int i = q * NVL + j, c4 = i >> 3, k = (i & 7) >> 1, c = i & 1;
                    const int64_t g = (int64_t)sb * sbp + pt * RPS_BN + half * 32 + 8 * k + 2 * p + c;
                    const double sa_ = rscale[g];
                    if (c4 == 0) rho[g] = sa_ * t;
                    else grad[(int64_t)(c4 - 1) * ngrid_ld + g] = 2.0 * sa_ * t;
                }
            }
            asm volatile("bar.sync 3, 256;" ::: "memory");          // `comb` is free for the next unit
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();               // no CTA leaves while a peer may still multicast into it
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// shared memory: 227 KB minus the static part (barriers, comb: < 7 KB) = A ring + B cache
template <int S>
static constexpr int rps_slots() {      // most B-cache slots beside a 3-stage ring
    return (227 * 1024 - 7 * 1024 - 3 * S * I8_A_PLANE) / (S * I8_KT * RPS_BN) < RPS_MAXSLOT
               ? (227 * 1024 - 7 * 1024 - 3 * S * I8_A_PLANE) / (S * I8_KT * RPS_BN)
               : RPS_MAXSLOT;
}
template <int S>
static int rps_stages(int nbc) {        // ring stages beside nbc B-cache slots
    const int n = (227 * 1024 - 7 * 1024 - nbc * S * I8_KT * RPS_BN) / (S * I8_A_PLANE);
    return n < RPS_MAXST ? n : RPS_MAXST;
}

template <typename K, typename... Args>
static cudaError_t launch_cluster(K kernel, int cl, int grid, int threads, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// clusters of `cl` CTAs (one CTA per SM) that can be resident at once
template <typename K>
static int max_active_clusters(K kernel, int cl, int threads, size_t smem) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(NUM_SMS / cl * cl));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

template <int S, int NCOMP, int CL>
static int rho_i8_ps_launch(const SBDesc *sbd, int nsb, int sbp, const double *ao, const signed char *pplanes,
                            const int64_t *p_off, const signed char *dplanes, const int64_t *d_off, const double *rscale,
                            const double *cscale, double *rho, double *grad, int nbc, cudaStream_t st) {
    auto kern = rho_i8_ps_kernel<S, NCOMP, CL>;
    const int64_t ngl = (int64_t)nsb * sbp;
    // the cache takes what the static shared memory (barriers, comb) leaves of the 227 KB: shrink it by a slot if
    // the driver does not accept the launch configuration
    size_t smem = 0;
    int nst = 0;
    for (;; nbc--) {
        QC_REQUIRE(nbc > RPS_STREAM, "rho_i8_ps_kernel does not fit in shared memory");
        // the A ring takes what the B cache leaves (the MMA issuer waits on the latency of the D_sb stream, not on its
        // bandwidth: bytes in flight are what counts)
        nst = rps_stages<S>(nbc);
        if (nst < 2) continue;
        smem = (size_t)nst * S * I8_A_PLANE + (size_t)nbc * S * I8_KT * RPS_BN;
        int nblk = 0;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, kern, 384, smem) == cudaSuccess && nblk >= 1)
            break;
        (void)cudaGetLastError();
    }
    int grid = NUM_SMS;
    if (CL > 1) {
        // persistent clusters: as many as can be resident at once (a GPC that cannot place another whole cluster
        // leaves its last SMs idle)
        static int ncl = -1;                        // (per instantiation and process; the device kind does not change)
        if (ncl < 0) ncl = max_active_clusters(kern, CL, 384, smem);
        QC_REQUIRE(ncl > 0, "no cluster of rho_i8_ps_kernel fits on the device");
        grid = (ncl < NUM_SMS / CL ? ncl : NUM_SMS / CL) * CL;
    }
    prof_begin(PROF_RHO, st);
    if (CL > 1)
        QC_CHECK(launch_cluster(kern, CL, grid, 384, smem, st, sbd, nsb, sbp, ao, pplanes, p_off, dplanes, d_off, rscale, cscale,
                                ngl, rho, grad, nbc, nst, g_i8_variant));
    else
        kern<<<grid, 384, smem, st>>>(sbd, nsb, sbp, ao, pplanes, p_off, dplanes, d_off, rscale, cscale, ngl, rho, grad, nbc,
                                      nst, g_i8_variant);
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

template <int S>
static int rho_i8_ps_run(const SBDesc *sbd, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                         const double *dm, int nao, const signed char *pplanes, const int64_t *p_off,
                         const double *rscale, signed char *dplanes, const int64_t *d_off, double *cscale, double *rho,
                         double *grad, cudaStream_t st) {
    prof_begin(PROF_SB_GATHER, st);
    sb_gather_slice_dm_launch<S, I8_BM>(sbd, nsb, max_nsp, idx, dm, nao, d_off, dplanes, cscale, st);
    prof_end(st);
    QC_LAUNCHED(1);
    // B200QC_I8_MODE bits 12..16: number of B cache slots (experiments: what the cache does not take stays L1);
    // bits 17..18: cluster size of the multicast A stream (0 = default, 1 = no clusters, 2, 3 = clusters of 4)
    int nbc = (g_i8_mode >> 12) & 31;
    if (nbc < RPS_STREAM + 1 || nbc > rps_slots<S>()) nbc = RPS_DEFAULT_SLOTS < rps_slots<S>() ? RPS_DEFAULT_SLOTS : rps_slots<S>();
    int cl = (g_i8_mode >> 17) & 3;
    cl = cl == 0 ? RPS_CLUSTER : (cl == 3 ? 4 : cl);
    while (cl > 1 && (sbp / RPS_BN) % cl != 0) cl >>= 1;
#define RPS_GO(NCOMP_, CL_) \
    rho_i8_ps_launch<S, NCOMP_, CL_>(sbd, nsb, sbp, ao, pplanes, p_off, dplanes, d_off, rscale, cscale, rho, grad, nbc, st)
    if (grad) return cl == 4 ? RPS_GO(4, 4) : cl == 2 ? RPS_GO(4, 2) : RPS_GO(4, 1);
    return cl == 4 ? RPS_GO(1, 4) : cl == 2 ? RPS_GO(1, 2) : RPS_GO(1, 1);
#undef RPS_GO
}

// Same contract as b200qc_rho_sb with the GEMM on tcgen05 int8 slices.  bplanes (sum_sb nslice * nsp * ceil(nsp / bn) * bn
// bytes at b_off[sb], ZERO-FILLED once by the caller: the rows past nsp of a last N tile are never written) and
// cscale (sum_sb nsp doubles) are per-call scratch.  bn = row tile of the sliced density: 128 selects the
// point-stationary kernel (rho_i8_ps_kernel; aplanes then hold 64-row tiles, b200qc_rho_i8_prepare row_tile = 64);
// 64, or 96 with nslice = 5, the row-tile-streaming kernel (rho_i8_kernel, aplanes in 128-row tiles).
extern "C" int b200qc_rho_sb_i8(const void *sbdesc, int nsb, int sbp, int max_nsp, int nslice, const int *idx,
                                const double *ao, const double *dm, int nao, const signed char *aplanes,
                                const int64_t *a_off, const double *rscale, signed char *bplanes,
                                const int64_t *b_off, double *cscale, int bn, double *rho, double *grad, void *stream) {
    QC_REQUIRE(sbp % I8_BM == 0, "superblock size must be a multiple of 128");
    QC_REQUIRE(nslice == 5 || nslice == 6, "nslice must be 5 or 6");
    QC_REQUIRE(bn == 128 || bn == 64 || (bn == 96 && nslice == 5), "density row tile: 128, 64, or 96 with 5 slices");
    QC_REQUIRE((int64_t)max_nsp * 6 * 4096 < (1LL << 31), "too many AOs per superblock for exact int32 accumulation");
    if (nsb == 0) return 0;
    const SBDesc *sbd = (const SBDesc *)sbdesc;
    cudaStream_t st = as_stream(stream);
    if (bn == 128) {
        if (nslice == 5)
            return rho_i8_ps_run<5>(sbd, nsb, sbp, max_nsp, idx, ao, dm, nao, aplanes, a_off, rscale, bplanes, b_off, cscale, rho, grad, st);
        return rho_i8_ps_run<6>(sbd, nsb, sbp, max_nsp, idx, ao, dm, nao, aplanes, a_off, rscale, bplanes, b_off, cscale, rho, grad, st);
    }
    if (nslice == 5 && bn == 96)
        return rho_i8_run<5, 96>(sbd, nsb, sbp, max_nsp, idx, ao, dm, nao, aplanes, a_off, rscale, bplanes, b_off, cscale, rho, grad, st);
    if (nslice == 5)
        return rho_i8_run<5, 64>(sbd, nsb, sbp, max_nsp, idx, ao, dm, nao, aplanes, a_off, rscale, bplanes, b_off, cscale, rho, grad, st);
    return rho_i8_run<6, 64>(sbd, nsb, sbp, max_nsp, idx, ao, dm, nao, aplanes, a_off, rscale, bplanes, b_off, cscale, rho, grad, st);
}
