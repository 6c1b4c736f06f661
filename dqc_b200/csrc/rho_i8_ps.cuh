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
// Epilogue (8 warps; TMEM lane = AO nu, warp = lane quarter x column half): the accumulators are drained to fp64
// registers (TMEM is released to the next M tile at once), multiplied with the fp64 AO values phi_c[g][nu] -- read
// COALESCED across the lanes (consecutive nu), unlike the row-per-thread reads of rho_i8_kernel -- and reduced over
// nu by a butterfly reduce-scatter across the warp (fixed order, no atomics); partial sums live in registers across
// the M tiles and meet in shared memory once per unit.
#pragma once
#include "rho_i8.cuh"

#define RPS_BN 64          // grid points per tile (MMA N)
#define RPS_NST 3          // A-operand (D_sb) ring depth
#define RPS_MAXSLOT 16     // B cache slots (one K step of the 64-point tile each)
#define RPS_STREAM 3       // slots that turn into a ring when the tile does not fit
#define RPS_CLUSTER 2      // default cluster size of the multicast A stream

__device__ __forceinline__ double ldcs_f64(const double *p) {
    double v;
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// Sum of v[i] over the 32 lanes for every i in [0, 32): lane l returns the total of element l.
// Five exchange rounds (16, 8, 4, 2, 1 elements), 31 shuffles of 64 bits; the order of the additions is fixed.
__device__ __forceinline__ double warp_reduce_scatter32(double (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; i++) {
            const double send = upper ? v[i] : v[i + off];
            const double keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
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
                 double *__restrict__ rho, double *__restrict__ grad, int nbc, int variant) {
    extern __shared__ __align__(1024) unsigned char i8_smem[];
    constexpr int A_STAGE = S * I8_A_PLANE, A_PART = A_STAGE / CL, B_ROWS = RPS_BN * 16, B_SLOT = S * I8_KT * RPS_BN;
    static_assert(A_STAGE % (16 * CL) == 0, "the A stage splits into 16-byte aligned parts");
    // instruction descriptor without N: D = S32, A = B = signed int8, both K-major, M = 128
    constexpr uint32_t IDESC0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_BM >> 4) << 24);
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
    constexpr int NCH = (NCOMP == 4) ? 4 : 1;     // reduce-scatter rounds per M tile: 4 x (8 points x 4 components) or 1 x 32 points
    __shared__ uint64_t afull[RPS_NST], aempty[RPS_NST], bfull[RPS_MAXSLOT], bempty[RPS_MAXSLOT], accum_full, accum_empty;
    __shared__ uint32_t tmem_base_smem;
    __shared__ double comb[3][2][NCH][32];        // partial sums of lane quarters 1..3, handed to quarter 0
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ptiles = sbp / RPS_BN;
    // a cluster walks groups of CL consecutive point tiles of one superblock; CTA r of the cluster owns tile r of the group
    const int crank = (CL > 1) ? (int)cluster_ctarank() : 0;
    const int gps = ptiles / CL, ngroups = nsb * gps;
    const int g0 = (int)blockIdx.x / CL, gstep = (int)gridDim.x / CL;

    if (tid == 0) {
        for (int i = 0; i < RPS_NST; i++) {
            mbar_init(&afull[i], 1);
            mbar_init(&aempty[i], CL);            // every CTA of the cluster releases a stage
        }
        for (int i = 0; i < RPS_MAXSLOT; i++) {
            mbar_init(&bfull[i], 1);
            mbar_init(&bempty[i], 1);
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
    if (CL > 1) cluster_sync_all();               // the peers' barriers exist before anything is multicast to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    const uint32_t abase = smem_u32(i8_smem);                 // A ring
    const uint32_t bbase = abase + RPS_NST * A_STAGE;         // B cache: nbc slots
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
                        const int slot = ait % RPS_NST;
                        mbar_wait(&aempty[slot], ((ait / RPS_NST) & 1) ^ 1);
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
                        const int slot = ait % RPS_NST;
                        mbar_wait(&afull[slot], (ait / RPS_NST) & 1);
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
        // ===== epilogue: TMEM lane quarter lg = warp % 4 (AO rows), column half = 32 of the 64 grid points =====
        const int lg = warp & 3, half = (warp - 4) >> 2;
        const int et = tid - 128;                                    // 0..255: (component, point row) of the L2 prefetch
        int nt = 0;
        for (int ug = g0; ug < ngroups; ug += gstep) {
            const int sb = ug / gps, pt = (ug - sb * gps) * CL + crank;
            const SBDesc d = sbd[sb];
            const int ntm = (d.nsp + I8_BM - 1) / I8_BM;
            const int64_t ld = d.nsp, cstride = (int64_t)sbp * ld;
            const int grow0 = pt * RPS_BN + half * 32;               // first grid row (inside the superblock) of this warp
            double acc[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) acc[ch] = 0.0;
            for (int mt = 0; mt < ntm; mt++, nt++) {
                const int nu = mt * I8_BM + lg * 32 + lane;          // AO row of D_sb == AO column of phi
                const bool live = nu < d.nsp;
                const double cs = live ? cscale[d.idx_off + nu] : 0.0;
                const double *col = ao + d.ao_off + (int64_t)grow0 * ld + (live ? nu : 0);
                // the fp64 AO values of the NEXT M tile (or of the first M tile of the next unit) start their way from
                // HBM into L2 now, one 1 KB row segment per thread, a whole MMA phase before they are read
                if (variant != 1 && variant != 3 && et < 64 * NCOMP) {
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
                mbar_wait(&accum_full, nt & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                double x[32];
                if (variant == 3) {
#pragma unroll
                    for (int j = 0; j < 32; j++) x[j] = 1.0;
                } else {
#pragma unroll
                    for (int ch = 0; ch < 4; ch++) {
                        double t8[8];
                        i8_recombine8<S>(tmem + ((uint32_t)(lg * 32) << 16), RPS_BN, half * 32 + ch * 8, t8);
#pragma unroll
                        for (int j = 0; j < 8; j++) x[ch * 8 + j] = t8[j] * cs;
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");      // all epilogue warps have drained TMEM
                if (warp == 4 && lane == 0)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&accum_empty)) : "memory");
                if (variant == 1 || variant == 3) {
#pragma unroll
                    for (int j = 0; j < 32; j++) acc[0] += x[j];
                    continue;
                }
                if (NCOMP == 4) {
                    // chunk ch = 8 points x 4 components; the loads of chunk ch + 1 are issued before chunk ch is reduced
                    double v[2][32];
                    auto load_chunk = [&](double (&t)[32], int ch) {
#pragma unroll
                        for (int c = 0; c < 4; c++)
#pragma unroll
                            for (int j = 0; j < 8; j++)
                                t[c * 8 + j] = live ? ldcs_f64(col + c * cstride + (int64_t)(ch * 8 + j) * ld) : 0.0;
                    };
                    load_chunk(v[0], 0);
#pragma unroll
                    for (int ch = 0; ch < 4; ch++) {
                        if (ch + 1 < 4) load_chunk(v[(ch + 1) & 1], ch + 1);
#pragma unroll
                        for (int c = 0; c < 4; c++)
#pragma unroll
                            for (int j = 0; j < 8; j++) v[ch & 1][c * 8 + j] *= x[ch * 8 + j];
                        acc[ch] += warp_reduce_scatter32(v[ch & 1], lane);    // lane l: component l / 8, point ch * 8 + l % 8
                    }
                } else {
                    double v[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = live ? ldcs_f64(col + (int64_t)j * ld) : 0.0;
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] *= x[j];
                    acc[0] += warp_reduce_scatter32(v, lane);         // lane l: point l
                }
            }
            // the four lane quarters of a column half meet in shared memory (the next write of `comb` is ordered
            // behind these reads by the bar.sync 1 of the next unit's first M tile)
            if (lg > 0) {
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) comb[lg - 1][half][ch][lane] = acc[ch];
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (lg == 0) {
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) {
                    const double t = ((acc[ch] + comb[0][half][ch][lane]) + comb[1][half][ch][lane]) + comb[2][half][ch][lane];
                    const int c = (NCOMP == 4) ? (lane >> 3) : 0;
                    const int grow = grow0 + ((NCOMP == 4) ? (ch * 8 + (lane & 7)) : lane);
                    const int64_t g = (int64_t)sb * sbp + grow;
                    const double sa_ = rscale[g];
                    if (c == 0) rho[g] = sa_ * t;
                    else grad[(int64_t)(c - 1) * ngrid_ld + g] = 2.0 * sa_ * t;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();               // no CTA leaves while a peer may still multicast into it
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// B-cache slots that fit beside the A ring in the 227 KB of shared memory (static barriers / comb: < 7 KB)
template <int S>
static constexpr int rps_slots() {
    return (227 * 1024 - 7 * 1024 - RPS_NST * S * I8_A_PLANE) / (S * I8_KT * RPS_BN) < RPS_MAXSLOT
               ? (227 * 1024 - 7 * 1024 - RPS_NST * S * I8_A_PLANE) / (S * I8_KT * RPS_BN)
               : RPS_MAXSLOT;
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
    for (;; nbc--) {
        QC_REQUIRE(nbc > RPS_STREAM, "rho_i8_ps_kernel does not fit in shared memory");
        smem = (size_t)RPS_NST * S * I8_A_PLANE + (size_t)nbc * S * I8_KT * RPS_BN;
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
                                ngl, rho, grad, nbc, g_i8_variant));
    else
        kern<<<grid, 384, smem, st>>>(sbd, nsb, sbp, ao, pplanes, p_off, dplanes, d_off, rscale, cscale, ngl, rho, grad, nbc,
                                      g_i8_variant);
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

template <int S>
static int rho_i8_ps_run(const SBDesc *sbd, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                         const double *dm, int nao, const signed char *pplanes, const int64_t *p_off,
                         const double *rscale, signed char *dplanes, const int64_t *d_off, double *cscale, double *rho,
                         double *grad, cudaStream_t st) {
    dim3 gg((unsigned)(max_nsp / 8), (unsigned)nsb);
    prof_begin(PROF_SB_GATHER, st);
    sb_gather_slice_dm_kernel<S, I8_BM><<<gg, 256, 0, st>>>(sbd, idx, dm, nao, d_off, dplanes, cscale);
    prof_end(st);
    QC_LAUNCHED(1);
    // B200QC_I8_MODE bits 12..16: number of B cache slots (experiments: what the cache does not take stays L1);
    // bits 17..18: cluster size of the multicast A stream (0 = default, 1 = no clusters, 2, 3 = clusters of 4)
    int nbc = (g_i8_mode >> 12) & 31;
    if (nbc < RPS_STREAM + 1 || nbc > rps_slots<S>()) nbc = rps_slots<S>();
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
