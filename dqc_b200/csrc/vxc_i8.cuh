// K4 on the 5th-generation tensor cores: M_sb = phi_sb^T vb_sb as an error-free sliced int8 GEMM
// (Ozaki scheme) on tcgen05.mma.kind::i8 with int32 accumulators in TMEM.
//
// tcgen05.mma has no f64 kind; the DMMA (mma.sync m8n8k4) form of this GEMM saturates the fp64 tensor
// pipe at ~30 TFLOP/s.  Here both fp64 operands are cut, per superblock and per column, into S int8
// slices of a block-fixed-point representation
//     x = 2^e * sum_s q_s 2^(-6 - 7 s),   q_s in [-64, 64],   e = exponent of the column maximum,
// every slice product sum_g q_s^A q_t^B is EXACT in int32 (|.| <= 512 * 4096 per product, <= 2^24 per
// anti-diagonal), products with s + t = d share the scale 2^(-12 - 7 d) and one TMEM accumulator, and the
// S accumulators are recombined exactly (int64) in the epilogue.  Truncation at s + t < S leaves a relative
// error of ~(S + 1) 2^(-7 S - 5) of (column max x column max x K): S = 6 reproduces the fp64 Vxc matrix
// to ~1e-12 (tools/ozaki_emul.py), S = 5 -- the default -- to 1e-10 at the full C60 size (bar: 1e-6), with
// S (S + 1) / 2 = 21 / 15 int8 MMAs per fp64 one.
//
// Kernel layout: CTA tile 128 (mu) x BN (nu), BN = 96 with S = 5 (480 of the 512 TMEM columns; 3-stage ring,
// 8 epilogue warps) or 64 (4-stage ring, 4 epilogue warps); K tile = 32 grid rows (one MMA K step).  tcgen05.mma
// reads both operands from shared memory for every instruction, so bytes per MMA -- not L2 or HBM -- bound the
// main loop: the wide tile with one slice less moves 42 % fewer of them.
// Operands are MN-major (the AO index is the contiguous one) in the no-swizzle canonical UMMA layout
// (core matrix = 16 bytes of MN x 8 rows of K).  The slicing kernels write the planes to global memory
// ALREADY in that tiled order, so a pipeline stage is two contiguous blocks and the producer is one
// thread issuing two 1-D bulk copies (cp.async.bulk + mbarrier complete_tx) -- no tensor maps, no
// per-thread address arithmetic.  Warp-specialised: warp 0 = producer, warp 1 = MMA issuer (one thread,
// S (S + 1) / 2 MMAs per stage, tcgen05.commit frees the stage), warps 4.. = epilogue (tcgen05.ld 32x32b,
// recombination, tile staged in shared memory, row-wise coalesced fp64 atomics into M[idx][idx]).
#pragma once
#include "sb_common.cuh"

#define I8_BM 128
#define I8_BN 64
#define I8_KT 32          // one MMA K step (32 int8) per pipeline stage
#define I8_THREADS 256
#define I8_STAGES 4       // BN = 64: 4 stages of operands in flight per SM (+ the epilogue staging tile)
#define I8_A_PLANE (I8_KT * I8_BM)   // 4096 bytes: [4 K groups][8 MN chunks][8 rows][16 bytes]
#define I8_B_PLANE (I8_KT * I8_BN)   // 2048 bytes: [4 K groups][4 MN chunks][8 rows][16 bytes]

// ---- slicing: X [rows][ld] fp64 per superblock -> S int8 planes in the TILED operand order + per-column scale
// W = 128 (A operand, M tiles, zero-padded to a whole tile by the caller's memset), out (per SB, bytes):
//     [tile = col / 128][k tile = row / 32][slice][(row % 32) / 8][(col % 128) / 16][row % 8][col % 16]
// W = 96 / 64 (B operand, N tiles): the slices sit INSIDE the 8-row K group,
//     [tile = col / W][k tile = row / 32][(row % 32) / 8][slice][(col % W) / 16][row % 8][col % 16]
// so that the slices t0 .. t0 + n - 1 of a tile form ONE MN-major operand of n W columns: tcgen05.mma re-reads both
// operands from shared memory for every instruction, and one MMA of N = n W (<= 256) against A slice s that updates
// the n adjacent accumulators s + t0 .. replaces n MMAs that would each re-read the A plane.
template <int S, int W>
struct BPlaneLayout {
    static constexpr int PLANE = I8_KT * W;
    static constexpr int KG = (W == I8_BM) ? (W / 16) * 128 : S * (W / 16) * 128;   // stride of an 8-row K group
    static constexpr int SL = (W == I8_BM) ? PLANE : (W / 16) * 128;                // stride of a slice
};
template <int S, int W>
__global__ void __launch_bounds__(256)
sb_slice_kernel(const SBDesc *__restrict__ sbd, const double *__restrict__ x, const int64_t *__restrict__ x_off,
                int x_is_ao, int sbp, const int64_t *__restrict__ p_off, signed char *__restrict__ planes,
                double *__restrict__ scales) {
    const SBDesc d = sbd[blockIdx.y];
    const int c0 = blockIdx.x * 64;
    if (c0 >= d.nsp) return;
    const int col = c0 + (threadIdx.x & 63), rg = threadIdx.x >> 6;   // 64 columns x 4 row groups
    const int64_t ld = d.nsp;
    const double *X = x + (x_is_ao ? d.ao_off : x_off[blockIdx.y]);
    __shared__ double smax[4][64];
    double m = 0.0;
    for (int r = rg; r < sbp; r += 4) m = fmax(m, fabs(X[(int64_t)r * ld + col]));
    smax[rg][threadIdx.x & 63] = m;
    __syncthreads();
    m = fmax(fmax(smax[0][threadIdx.x & 63], smax[1][threadIdx.x & 63]),
             fmax(smax[2][threadIdx.x & 63], smax[3][threadIdx.x & 63]));
    int e = 0;
    if (m > 0.0) frexp(m, &e);              // m = f 2^e, f in [0.5, 1)  =>  |x| / 2^e < 1
    const double inv = ldexp(64.0, -e);     // y = x 2^(6 - e), |y| < 64
    if (rg == 0) scales[d.idx_off + col] = ldexp(1.0, e);
    constexpr int PLANE = I8_KT * W;
    const int nk = sbp / I8_KT;
    // second pass: a thread owns 16 consecutive columns of one row -- 128 contiguous bytes in, one 16-byte store per
    // slice out (16 columns of a row are contiguous in the MN-major plane).  The 64 column scales sit in smem.
    __shared__ double sinv[64];
    if (rg == 0) sinv[threadIdx.x & 63] = inv;
    __syncthreads();
    const int cg = threadIdx.x & 3;                          // 16-column group inside the 64-column block
    const int cq = c0 + cg * 16;                             // first column of the group
    const int tile = cq / W, wc = cq % W;                    // (W is a multiple of 16: a group never straddles tiles)
    signed char *P = planes + p_off[blockIdx.y] + (int64_t)tile * nk * S * PLANE + (wc >> 4) * 128;
    for (int r = threadIdx.x >> 2; r < sbp; r += 64) {
        const double *src = X + (int64_t)r * ld + cq;
        unsigned int w[S][4];
#pragma unroll
        for (int s = 0; s < S; s++) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(src + j);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                double y = (h ? v.y : v.x) * sinv[cg * 16 + j + h];
#pragma unroll
                for (int s = 0; s < S; s++) {
                    w[s][(j + h) >> 2] |= ((unsigned int)slice_digit(y) & 0xffu) << (8 * ((j + h) & 3));
                }
            }
        }
        using L = BPlaneLayout<S, W>;
        signed char *Q = P + (int64_t)(r >> 5) * S * PLANE + ((r & 31) >> 3) * L::KG + (r & 7) * 16;
#pragma unroll
        for (int s = 0; s < S; s++) *reinterpret_cast<uint4 *>(Q + s * L::SL) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
}

// ---- fused K4a: vb = w (v phi + 2 g . grad phi) cut straight into the B-operand planes ----
// The slicer needs the column maxima of vb per superblock BEFORE it can cut, which is why round 1 wrote vb to HBM in
// fp64 (4.4 GB at C60) and read it back twice.  Here: bound, cut, verify, repair.
//  1. Any upper bound of the maximum works as the block exponent -- a bound that is 2^k too large costs k of the 7 S
//     mantissa bits -- and a cheap one follows from the triangle inequality, taken per group of 32 consecutive rows:
//         max_g |vb_g,nu| <= max_groups [ max_g |w v| * max_g |phi_nu| + sum_d max_g |2 w g_d| * max_g |d_d phi_nu| ]
//     with the per-group column maxima of phi / grad phi precomputed once (sb_colmax_kernel, fp32 rounded up: 1/64 of
//     the AO storage) and the coefficient maxima formed here.
//  2. One pass over the fp64 AO values forms vb and cuts it (vb never goes to HBM).  While cutting, the kernel records
//     the LARGEST scaled value it saw per column: that is the exact column maximum, for free.
//  3. A column whose bound turned out more than 2^VBS_LOOSE_BITS too large AND whose last digit would weigh more than
//     2^VBS_ABS_EXP gets its exact exponent written to `scales` and its 64-column block flagged;
//     vxc_vbslice_kernel<.., FIX = true> then re-cuts the flagged blocks only.  The bound has a long loose tail on a
//     molecular grid (tools/check_vb_bound.py, C60 / PBE: 38 % of the columns exact, 78 % within 4 bits, 10 % more than
//     20 bits off -- the core functions of the neighbouring atoms, large exactly where the Becke weight of this atom's
//     points vanishes), but almost all of those columns are tiny in absolute terms: 4.5 % of the blocks are re-cut,
//     Vxc agrees with the two-pass form to 5e-11 (2.6e-10 without any repair; the parity bar is 1e-6).  For a
//     potential that is uncorrelated from point to point (tests/test_gpu_xcpath.py) most blocks are re-cut.
// Memory access: a warp reads 8 consecutive rows, lane = 2 adjacent columns -> every load instruction is one contiguous
// 512-byte row segment (the round-2 first version, thread = 16 columns of a row, ran at 93 % L1 throughput with half
// of every sector wasted per instruction: 4.65 ms); the int8 bytes are staged in shared memory and leave as 16-byte
// pieces of whole 128-byte core matrices.
#define VBS_GROUP 32
#define VBS_LOOSE_BITS 4
#define VBS_ABS_EXP (-40)       // 2^-40 = 9e-13 (at -49: 27 % of the C60 blocks are re-cut)
#define VBS_PHASE 64            // rows per staging phase (two K tiles)
#define VBS_ROWB 80             // staged row: 64 bytes + 16 of padding (16-byte reads of 8 consecutive rows: no conflicts)

// colmax[((idx_off + col) * ngroups + group) * NCOMP + c] >= max over the group's rows of |ao_c[row][col]|
template <int NCOMP>
__global__ void __launch_bounds__(64 * NCOMP)
sb_colmax_kernel(const SBDesc *__restrict__ sbd, const double *__restrict__ ao, int sbp, float *__restrict__ colmax) {
    const SBDesc d = sbd[blockIdx.y];
    const int c0 = blockIdx.x * 64;
    if (c0 >= d.nsp) return;
    const int col = c0 + (threadIdx.x & 63), c = threadIdx.x >> 6;
    const int64_t ld = d.nsp;
    const int ngroups = sbp / VBS_GROUP;
    const double *X = ao + d.ao_off + (int64_t)c * sbp * ld + col;
    for (int grp = 0; grp < ngroups; grp++) {
        double m = 0.0;
#pragma unroll 8
        for (int r = 0; r < VBS_GROUP; r++) m = fmax(m, fabs(X[(int64_t)(grp * VBS_GROUP + r) * ld]));
        colmax[((int64_t)(d.idx_off + col) * ngroups + grp) * NCOMP + c] = __double2float_ru(m);
    }
}

// planes: the B operand of vxc_i8_gemm_kernel in the order sb_slice_kernel<S, W> writes (BPlaneLayout<S, W>).
// fixflag[sb * gridDim.x + column block]: written by the FIX = false pass (1 = re-cut this block), read by FIX = true.
template <int S, int W, int NCOMP, bool FIX>
__global__ void __launch_bounds__(256, 2)
vxc_vbslice_kernel(const SBDesc *__restrict__ sbd, const double *__restrict__ ao, int sbp, int64_t ngrid_ld,
                   const double *__restrict__ w, const double *__restrict__ vrho, const double *__restrict__ vgrad,
                   const float *__restrict__ colmax, const int64_t *__restrict__ p_off, signed char *__restrict__ planes,
                   double *__restrict__ scales, int *__restrict__ fixflag, int loose_bits) {
    extern __shared__ __align__(16) unsigned char vbs_raw[];
    // coef[NCOMP][sbp], wmax[ngroups][NCOMP] (doubles), then the staging tile [S][VBS_PHASE][VBS_ROWB] bytes
    __shared__ double sinv[64], bpart[4][64];
    __shared__ float obs[8][64];
    __shared__ int anyfix;
    const int sb = blockIdx.y;
    const SBDesc d = sbd[sb];
    const int c0 = blockIdx.x * 64;
    if (c0 >= d.nsp) return;
    if (FIX && fixflag[sb * gridDim.x + blockIdx.x] == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ngroups = sbp / VBS_GROUP;
    double *coef = reinterpret_cast<double *>(vbs_raw), *wmax = coef + NCOMP * sbp;
    unsigned char *stage = reinterpret_cast<unsigned char *>(wmax + NCOMP * ngroups);
    // 1. the row coefficients of this superblock; a warp handles whole 32-row groups and leaves their maxima
    for (int grp = warp; grp < ngroups; grp += 8) {
        const int r = grp * VBS_GROUP + lane;
        const int64_t g = (int64_t)sb * sbp + r;
        const double wg = w[g];
        double m[NCOMP];
        double cf = wg * vrho[g];
        coef[r] = cf;
        m[0] = fabs(cf);
        if (NCOMP == 4) {
#pragma unroll
            for (int dd = 0; dd < 3; dd++) {
                cf = 2.0 * wg * vgrad[(int64_t)dd * ngrid_ld + g];
                coef[(dd + 1) * sbp + r] = cf;
                m[dd + 1] = fabs(cf);
            }
        }
        if (!FIX) {
#pragma unroll
            for (int c = 0; c < NCOMP; c++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m[c] = fmax(m[c], __shfl_xor_sync(0xffffffffu, m[c], o));
                if (lane == 0) wmax[grp * NCOMP + c] = m[c];
            }
        }
    }
    if (tid == 0) anyfix = 0;
    __syncthreads();
    if (FIX) {
        // the exponents are final: exact for the columns the first pass found loose, the (tight) bound for the others
        if (tid < 64) sinv[tid] = 64.0 / scales[d.idx_off + c0 + tid];
    } else {
        // 2. bound of the column maxima -> block exponents of the 64 columns of this CTA (4 threads per column, each a
        //    quarter of the row groups)
        const int col = c0 + (tid & 63), part = tid >> 6;
        const float *cm = colmax + (int64_t)(d.idx_off + col) * ngroups * NCOMP;
        double b = 0.0;
        for (int grp = part; grp < ngroups; grp += 4) {
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < NCOMP; c++) t += wmax[grp * NCOMP + c] * (double)cm[grp * NCOMP + c];
            b = fmax(b, t);
        }
        bpart[part][tid & 63] = b;
        __syncthreads();
        if (tid < 64) {
            const double bb = fmax(fmax(bpart[0][tid], bpart[1][tid]), fmax(bpart[2][tid], bpart[3][tid])) * (1.0 + 1e-12);
            int e = 0;
            if (bb > 0.0) frexp(bb, &e);            // bb = f 2^e, f in [0.5, 1)  =>  |vb| / 2^e < 1
            sinv[tid] = ldexp(64.0, -e);
        }
    }
    __syncthreads();
    // 3. cut.  Warp = 8 consecutive rows (one K row group) of a 64-row phase, lane = columns c0 + 2 lane, + 1.
    using L = BPlaneLayout<S, W>;
    constexpr int PLANE = L::PLANE;
    const int nk = sbp / I8_KT;
    const int64_t ld = d.nsp, cs = (int64_t)sbp * ld;
    const double si0 = sinv[2 * lane], si1 = sinv[2 * lane + 1];
    const double *src0 = ao + d.ao_off + c0 + 2 * lane;
    float mo0 = 0.f, mo1 = 0.f;                     // largest |scaled value| seen: the exact column maxima
    for (int r0 = 0; r0 < sbp; r0 += VBS_PHASE) {
        // (two half-groups of 4 rows: 16 independent 16-byte loads in flight per thread at 128 registers)
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            const int rw = r0 + warp * 8 + hf * 4;
            double2 v[NCOMP][4];
#pragma unroll
            for (int c = 0; c < NCOMP; c++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    v[c][j] = *reinterpret_cast<const double2 *>(src0 + c * cs + (int64_t)(rw + j) * ld);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double y0 = coef[rw + j] * v[0][j].x, y1 = coef[rw + j] * v[0][j].y;
                if (NCOMP == 4) {
#pragma unroll
                    for (int c = 1; c < 4; c++) {
                        const double cf = coef[c * sbp + rw + j];
                        y0 += cf * v[c][j].x;
                        y1 += cf * v[c][j].y;
                    }
                }
                y0 *= si0;
                y1 *= si1;
                if (!FIX) {
                    mo0 = fmaxf(mo0, __double2float_ru(fabs(y0)));
                    mo1 = fmaxf(mo1, __double2float_ru(fabs(y1)));
                }
                unsigned char *q = stage + (warp * 8 + hf * 4 + j) * VBS_ROWB + 2 * lane;
#pragma unroll
                for (int sl = 0; sl < S; sl++) {
                    const unsigned int q0 = (unsigned int)slice_digit(y0) & 0xffu, q1 = (unsigned int)slice_digit(y1) & 0xffu;
                    *reinterpret_cast<unsigned short *>(q + sl * (VBS_PHASE * VBS_ROWB)) = (unsigned short)(q0 | (q1 << 8));
                }
            }
        }
        __syncthreads();
        // the phase leaves as 16-byte pieces; 8 consecutive threads write one 128-byte core matrix
        // piece = (((k tile of the phase * 4 + K row group) * S + slice) * 4 + 16-column chunk) * 8 + row % 8
        for (int pc = tid; pc < (VBS_PHASE / 8) * S * 4 * 8; pc += 256) {
            const int r8 = pc & 7, ch = (pc >> 3) & 3, t = pc >> 5;
            const int sl = t % S, kg = t / S;                        // kg = 8-row group inside the phase (0..7)
            const uint4 val = *reinterpret_cast<const uint4 *>(stage + sl * (VBS_PHASE * VBS_ROWB) + (kg * 8 + r8) * VBS_ROWB + ch * 16);
            const int cq = c0 + ch * 16, tile = cq / W, wc = cq % W;
            const int r = r0 + kg * 8 + r8;
            signed char *Q = planes + p_off[sb] + (int64_t)tile * nk * S * PLANE + (int64_t)(r >> 5) * S * PLANE +
                             ((r & 31) >> 3) * L::KG + sl * L::SL + (wc >> 4) * 128 + r8 * 16;
            *reinterpret_cast<uint4 *>(Q) = val;
        }
        __syncthreads();
    }
    if (FIX) return;
    // 4. verify the bound against the maxima actually seen; exact exponents for the loose columns
    obs[warp][2 * lane] = mo0;
    obs[warp][2 * lane + 1] = mo1;
    __syncthreads();
    if (tid < 64) {
        float mo = obs[0][tid];
#pragma unroll
        for (int q = 1; q < 8; q++) mo = fmaxf(mo, obs[q][tid]);
        double sc = 64.0 / sinv[tid];                                // 2^e of the bound
        // loose AND coarse in absolute terms: the last digit of a column cut with exponent e weighs 2^(e - 7 S); columns
        // whose bound is loose because the grid weights vanish exactly where the AO is large (the core functions of
        // the neighbouring atoms: thousands of them at C60, true maximum 2^-20 of the bound and below) stay as they are
        if (mo > 0.f && ((mo < ldexpf(64.f, -loose_bits) && sc > ldexp(1.0, VBS_ABS_EXP + 7 * S)) || mo >= 64.f)) {
            int e = 0;
            frexp((double)mo * (1.0 + 1e-6) / sinv[tid], &e);        // (mo is rounded up already; margin for the fp32 step)
            sc = ldexp(1.0, e);
            anyfix = 1;
        }
        scales[d.idx_off + c0 + tid] = sc;
    }
    __syncthreads();
    if (tid == 0) fixflag[sb * gridDim.x + blockIdx.x] = anyfix;
}

// ---- tcgen05 plumbing (PTX spellings as in the CUTLASS sm100 headers) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(a), "r"(phase) : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0));
}
// no-swizzle shared-memory matrix descriptor (version 1); byte offsets are stored without their 4 LSBs
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 16 columns x 32 lanes, no wait: issue several, then tmem_wait_ld() once
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Recombination of the S anti-diagonal accumulators of 16 columns: the int32 partial sums are merged exactly
// in int64 (sum_d acc_d 128^(S-1-d), |.| < 2^60), converted once and scaled -- 1 I2F per element instead of S.
template <int S, int ACC_STRIDE = 64>
__device__ __forceinline__ void i8_recombine16(uint32_t tmem_lane_base, int col0, double (&out)[16]) {
    uint32_t v[S][16];
#pragma unroll
    for (int dd = 0; dd < S; dd++) tmem_ld16_nowait(tmem_lane_base + dd * ACC_STRIDE + col0, v[dd]);
    tmem_wait_ld();
    const double sc = ldexp(1.0, -12 - 7 * (S - 1));
#pragma unroll
    for (int j = 0; j < 16; j++) {
        long long t = (long long)(int)v[0][j];
#pragma unroll
        for (int dd = 1; dd < S; dd++) t = (t << 7) + (long long)(int)v[dd][j];
        out[j] = __ll2double_rn(t) * sc;
    }
}

// 8-column variant (fewer live registers: kernels with more epilogue warps); acc_stride = TMEM columns between
// the accumulators of consecutive anti-diagonals
template <int S>
__device__ __forceinline__ void i8_recombine8(uint32_t tmem_lane_base, int acc_stride, int col0, double (&out)[8]) {
    uint32_t v[S][8];
#pragma unroll
    for (int dd = 0; dd < S; dd++) tmem_ld8_nowait(tmem_lane_base + dd * acc_stride + col0, v[dd]);
    tmem_wait_ld();
    const double sc = ldexp(1.0, -12 - 7 * (S - 1));
#pragma unroll
    for (int j = 0; j < 8; j++) {
        long long t = (long long)(int)v[0][j];
#pragma unroll
        for (int dd = 1; dd < S; dd++) t = (t << 7) + (long long)(int)v[dd][j];
        out[j] = __ll2double_rn(t) * sc;
    }
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- L2 cache hints and 2-CTA cluster multicast ----
// mode bits (b200qc_i8_mode): 1 = L2 evict_last hint on the re-used A planes, 2 = K4 in 2-CTA clusters with the A
// stage multicast to both CTAs (each CTA fetches half of it; the pair works on two N tiles of the same M tile),
// 4 = the same for K2 (the pair splits the N tiles of one 128-row block and adds its partial row sums atomically)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// the data AND the complete_tx land at the same CTA-relative offsets in every CTA of `mask`
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
static int g_i8_mode = 0;
extern "C" int b200qc_i8_mode(int flags) { g_i8_mode = flags; return 0; }

template <typename K, typename... Args>
static cudaError_t launch_cluster2(K kernel, int grid, int threads, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}


static int g_i8_variant = 0;   // timing experiments only (b200qc_i8_debug_variant)
extern "C" int b200qc_i8_debug_variant(int v) { g_i8_variant = v; return 0; }

// Persistent kernel: one CTA per SM walks tiles cta, cta + gridDim.x, ... of the flattened (superblock, M tile,
// N tile) list; the epilogue of tile i (fp64 recombination already staged in shared memory, atomics still to
// do) overlaps the main loop of tile i + 1.
// MC = 1: launched in 2-CTA clusters; `tile_off` / `ntiles` then describe PAIR units (superblock, M tile, pair of
// N tiles); CTA r of the cluster owns N tile 2 p + r (idle when the superblock has an odd number of N tiles and
// this is the missing one -- it still fetches and multicasts its half of every A stage).
// BN = N tile: 64, or 96 with S = 5 (5 x 96 = 480 TMEM columns).  tcgen05.mma reads both operands from shared memory
// for every instruction, so at N = 64 the tensor pipe waits on shared-memory bandwidth (21 x 6 KB read + 36 KB written
// per K step); N = 96 with S = 5 moves 42 % fewer bytes per unit of work and issues 29 % fewer MMAs.
// EW = epilogue warps: 4, or 8 (two per TMEM lane quarter, each draining half of the columns -- the drain is what
// keeps the MMA issuer waiting between tiles -- and half of the rows of the atomic phase); CTA = 128 + 32 EW threads.
template <int S, int MC, int BN, int EW>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
vxc_i8_gemm_kernel(const SBDesc *__restrict__ sbd, const int *__restrict__ tile_off, int nsb, int ntiles,
                   const int *__restrict__ idx, const signed char *__restrict__ aplanes,
                   const int64_t *__restrict__ a_off, const signed char *__restrict__ bplanes,
                   const int64_t *__restrict__ b_off, const double *__restrict__ ascale,
                   const double *__restrict__ bscale, int sbp, int nao, double *__restrict__ mat, int variant) {
    extern __shared__ __align__(1024) unsigned char i8_smem[];
    static_assert(S * BN <= 512, "accumulators exceed the tensor memory");
    constexpr int B_PLANE = I8_KT * BN, NSTAGE = (BN == 96) ? 3 : I8_STAGES, EPI_LD = BN + 1;
    constexpr int A_STAGE = S * I8_A_PLANE, B_STAGE = S * B_PLANE, STAGE = A_STAGE + B_STAGE;
    constexpr uint32_t LBO_A = (I8_BM / 16) * 128, LBO_B = BPlaneLayout<S, BN>::KG;   // stride between 8-row K groups
    constexpr int B_SL = BPlaneLayout<S, BN>::SL;                             // stride between the slices of a K group
    constexpr int NCAT = 256 / BN;                                            // B slices one MMA can take (N <= 256)
    // instruction descriptor without N: D = S32, A = B = signed int8, both MN-major, M = 128 (dense)
    constexpr uint32_t IDESC0 = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(I8_BM >> 4) << 24);
    __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], accum_full, accum_empty;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nk = sbp / I8_KT;
    double *tile = reinterpret_cast<double *>(i8_smem + NSTAGE * STAGE);   // epilogue staging, after the ring
    const int crank = MC ? (int)cluster_ctarank() : 0;
    const int u0 = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, ustep = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int A_HALF = A_STAGE / 2;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], MC ? 2 : 1);     // MC: both CTAs of the pair release a stage
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
    if (MC) cluster_sync_all();      // the peer's barriers are initialised before anything is multicast to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    const uint32_t sbase = smem_u32(i8_smem);

    // tile t -> (superblock, M tile, N tile); tile_off is the exclusive prefix of tiles (MC: pair units) per
    // superblock.  Returns false when this CTA has no N tile in the unit (MC, odd tile count).
    auto locate = [&](int t, int &sb, int &tm, int &tn) -> bool {
        int lo = 0, hi = nsb;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (tile_off[mid] <= t) lo = mid; else hi = mid;
        }
        sb = lo;
        const int ntn = (sbd[sb].nsp + BN - 1) / BN;
        const int r = t - tile_off[sb];
        if (MC) {
            const int ntp = (ntn + 1) >> 1;
            tm = r / ntp;
            tn = 2 * (r - tm * ntp) + crank;
            return tn < ntn;
        }
        tm = r / ntn;
        tn = r - tm * ntn;
        return true;
    };

    if (warp == 0) {
        // ===== producer: two contiguous blocks per stage =====
        if (lane == 0) {
            int it = 0;
            for (int t = u0; t < ntiles; t += ustep) {
                int sb, tm, tn;
                const bool active = locate(t, sb, tm, tn);
                const signed char *A = aplanes + a_off[sb] + (int64_t)tm * nk * A_STAGE;
                const signed char *B = bplanes + b_off[sb] + (int64_t)tn * nk * B_STAGE;
                for (int kt = 0; kt < nk; kt++, it++) {
                    const int slot = it % NSTAGE;
                    mbar_wait(&empty_bar[slot], ((it / NSTAGE) & 1) ^ 1);
                    if (MC) {
                        // this CTA's half of the A stage goes to both CTAs; the other half arrives from the peer
                        mbar_expect_tx(&full_bar[slot], A_STAGE + (active ? B_STAGE : 0));
                        bulk_g2s_mc(sbase + slot * STAGE + crank * A_HALF, A + (int64_t)kt * A_STAGE + crank * A_HALF, A_HALF,
                                    &full_bar[slot], (uint16_t)3);
                        if (active)
                            bulk_g2s(sbase + slot * STAGE + A_STAGE, B + (int64_t)kt * B_STAGE, B_STAGE, &full_bar[slot]);
                    } else {
                        mbar_expect_tx(&full_bar[slot], STAGE);
                        bulk_g2s(sbase + slot * STAGE, A + (int64_t)kt * A_STAGE, A_STAGE, &full_bar[slot]);
                        bulk_g2s(sbase + slot * STAGE + A_STAGE, B + (int64_t)kt * B_STAGE, B_STAGE, &full_bar[slot]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // MN-major, no swizzle: SBO = stride between 16-element MN chunks (128 B), LBO = stride between
            // 8-row K groups (A: 1024 B, B: 512 B)
            const uint64_t da0 = umma_desc(sbase, LBO_A, 128), db0 = umma_desc(sbase + A_STAGE, LBO_B, 128);
            int it = 0, nt = 0;
            for (int t = u0; t < ntiles; t += ustep) {
                int sb_, tm_, tn_;
                const bool active = MC ? locate(t, sb_, tm_, tn_) : true;
                if (active) {
                    mbar_wait(&accum_empty, (nt & 1) ^ 1);      // the epilogue has drained the previous tile
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int kt = 0; kt < nk; kt++, it++) {
                    const int slot = it % NSTAGE;
                    mbar_wait(&full_bar[slot], (it / NSTAGE) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = da0 + (uint64_t)((slot * STAGE) >> 4), db = db0 + (uint64_t)((slot * STAGE) >> 4);
                    if (variant != 2 && active) {      // (variant 2: timing experiment without the MMAs)
                    // A slice s2 against the B slices t0 .. t0 + n - 1 (one operand of n BN columns) -> the adjacent
                    // accumulators s2 + t0 .. s2 + t0 + n - 1
#pragma unroll
                    for (int s2 = 0; s2 < S; s2++)
#pragma unroll
                        for (int t0 = 0; t0 < S - s2; t0 += NCAT) {
                            const int n = (S - s2 - t0) < NCAT ? (S - s2 - t0) : NCAT;
                            umma_i8(tmem + (s2 + t0) * BN, da + (uint64_t)((s2 * I8_A_PLANE) >> 4),
                                    db + (uint64_t)((t0 * B_SL) >> 4), IDESC0 | ((uint32_t)((n * BN) >> 3) << 17),
                                    (kt > 0 || s2 > 0) ? 1u : 0u);
                        }
                    }
                    // frees the stage (in both CTAs of a pair) when the MMAs above retire
                    if (MC) umma_commit_mc(&empty_bar[slot], (uint16_t)3); else umma_commit(&empty_bar[slot]);
                }
                if (active) {
                    umma_commit(&accum_full);
                    nt++;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue warps: TMEM lanes 32 (warp % 4) ..; with EW = 8 warp w and w + 4 share a lane quarter and
        // split the columns of the tile =====
        constexpr int NCW = BN / (EW / 4);            // columns drained per warp
        const int lg = warp & 3, half = (warp - 4) >> 2;
        const int r = lg * 32 + lane;
        int nt = -1;
        for (int t = u0; t < ntiles; t += ustep) {
            int sb, tm, tn;
            if (!locate(t, sb, tm, tn)) continue;
            nt++;
            const SBDesc d = sbd[sb];
            const int m0 = tm * I8_BM, n0 = tn * BN;
            const int row = m0 + r;
            const double sa_ = row < d.nsp ? ascale[d.idx_off + row] : 0.0;
            mbar_wait(&accum_full, nt & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (variant == 1) {                        // timing experiment without the epilogue
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
                if (warp == 4 && lane == 0)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&accum_empty)) : "memory");
                continue;
            }
#pragma unroll
            for (int ch = 0; ch < NCW / 16; ch++) {
                double acc[16];
                const int c = half * NCW + ch * 16;
                i8_recombine16<S, BN>(tmem + ((uint32_t)(lg * 32) << 16), c, acc);
#pragma unroll
                for (int j = 0; j < 16; j++) tile[r * EPI_LD + c + j] = acc[j] * sa_;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");    // the epilogue warps: TMEM drained, tile staged
            if (warp == 4 && lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&accum_empty)) : "memory");
            const int *ix = idx + d.idx_off;
            int bcol[BN / 32];
            double bs[BN / 32];
#pragma unroll
            for (int cg = 0; cg < BN / 32; cg++) {      // the last N tile of a superblock may reach past nsp (zero planes)
                const int c = n0 + cg * 32 + lane;
                bcol[cg] = c < d.nsp ? ix[c] : nao;
                bs[cg] = c < d.nsp ? bscale[d.idx_off + c] : 0.0;
            }
            for (int rr = warp - 4; rr < I8_BM; rr += EW) {
                const int grow = m0 + rr;
                if (grow >= d.nsp) break;
                const int a = ix[grow];
                if (a >= nao) continue;
                double *dst = mat + (int64_t)a * nao;
#pragma unroll
                for (int cg = 0; cg < BN / 32; cg++)
                    if (bcol[cg] < nao) atomicAdd(dst + bcol[cg], tile[rr * EPI_LD + cg * 32 + lane] * bs[cg]);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");    // the staging tile is free again
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (MC) cluster_sync_all();      // no CTA leaves while its peer may still multicast into it
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// Slices the (static) AO values of every superblock once into the tiled A-operand order.
// aplanes: sum_sb ceil(nsp / 128) * 128 * sbp * nslice bytes, ZERO-FILLED by the caller (partial last M tile);
// a_off[sb]: byte offset of the SB's block; ascale: sum_sb nsp doubles.
extern "C" int b200qc_vxc_i8_prepare(const void *sbdesc, int nsb, int sbp, int max_nsp, int nslice, int ncomp,
                                     const double *ao, const int64_t *a_off, signed char *aplanes, double *ascale,
                                     float *colmax, void *stream) {
    QC_REQUIRE(nslice == 5 || nslice == 6, "nslice must be 5 or 6");
    QC_REQUIRE(sbp % I8_KT == 0, "superblock size must be a multiple of 32");
    QC_REQUIRE(ncomp == 1 || ncomp == 4, "ncomp must be 1 or 4");
    if (nsb == 0) return 0;
    dim3 grid((unsigned)(max_nsp / 64), (unsigned)nsb);
    const SBDesc *sbd = (const SBDesc *)sbdesc;
    if (colmax != nullptr) {    // static column maxima of phi (and grad phi) per 32-row group: the fused vb slicer's bound
        if (ncomp == 4) sb_colmax_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(sbd, ao, sbp, colmax);
        else sb_colmax_kernel<1><<<grid, 64, 0, as_stream(stream)>>>(sbd, ao, sbp, colmax);
        QC_LAUNCHED(1);
    }
    if (nslice == 5)
        sb_slice_kernel<5, I8_BM><<<grid, 256, 0, as_stream(stream)>>>(sbd, ao, nullptr, 1, sbp, a_off, aplanes, ascale);
    else
        sb_slice_kernel<6, I8_BM><<<grid, 256, 0, as_stream(stream)>>>(sbd, ao, nullptr, 1, sbp, a_off, aplanes, ascale);
    QC_LAUNCHED(1);
    return 0;
}

template <int S, int BN>
static int vxc_i8_run(const SBDesc *sbd, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                      const double *weights, const double *vrho, const double *vgrad, double *vb,
                      const int64_t *vb_off, const float *colmax, int *fixflag, const signed char *aplanes,
                      const int64_t *a_off, const double *ascale,
                      signed char *bplanes, const int64_t *b_off, double *bscale, const int *tile_off, int ntiles,
                      const int *ptile_off, int nptiles, int nao, double *mat, cudaStream_t st) {
    const int64_t ngl = (int64_t)nsb * sbp;
    dim3 gs((unsigned)(max_nsp / 64), (unsigned)nsb);
    if (colmax != nullptr) {
        // K4a fused: vb is cut into the int8 planes as it is formed (block exponents from the column-maximum bound),
        // then the blocks whose bound was loose are cut again with their exact exponents
        const size_t sm1 = sizeof(double) * (vgrad ? 4 : 1) * (sbp + sbp / VBS_GROUP) + (size_t)S * VBS_PHASE * VBS_ROWB;
        if (sm1 > 40 * 1024) {       // (the kernel also has ~5 KB of static shared memory)
            QC_CHECK(cudaFuncSetAttribute(vxc_vbslice_kernel<S, BN, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
            QC_CHECK(cudaFuncSetAttribute(vxc_vbslice_kernel<S, BN, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
            QC_CHECK(cudaFuncSetAttribute(vxc_vbslice_kernel<S, BN, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
            QC_CHECK(cudaFuncSetAttribute(vxc_vbslice_kernel<S, BN, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
        }
        // (B200QC_I8_MODE bits 20..22: looseness threshold of the repair pass, experiments; 0 = VBS_LOOSE_BITS)
        // (bit 23: never repair -- diagnostics of the bound, tools/check_vb_bound.py)
        const int loose = (g_i8_mode & (1 << 23)) ? 1000 : ((g_i8_mode >> 20) & 7) ? ((g_i8_mode >> 20) & 7) : VBS_LOOSE_BITS;
        prof_begin(PROF_VXC_VB, st);
        if (vgrad) {
            vxc_vbslice_kernel<S, BN, 4, false><<<gs, 256, sm1, st>>>(sbd, ao, sbp, ngl, weights, vrho, vgrad, colmax, b_off, bplanes, bscale, fixflag, loose);
            vxc_vbslice_kernel<S, BN, 4, true><<<gs, 256, sm1, st>>>(sbd, ao, sbp, ngl, weights, vrho, vgrad, colmax, b_off, bplanes, bscale, fixflag, loose);
        } else {
            vxc_vbslice_kernel<S, BN, 1, false><<<gs, 256, sm1, st>>>(sbd, ao, sbp, ngl, weights, vrho, vgrad, colmax, b_off, bplanes, bscale, fixflag, loose);
            vxc_vbslice_kernel<S, BN, 1, true><<<gs, 256, sm1, st>>>(sbd, ao, sbp, ngl, weights, vrho, vgrad, colmax, b_off, bplanes, bscale, fixflag, loose);
        }
        prof_end(st);
        QC_LAUNCHED(2);
    } else {
        // K4a unfused (exact column maxima): vb in fp64 (one streaming pass at the HBM roofline), then the slicer
        const int wpb = 8;
        const unsigned nb1 = (unsigned)((ngl + wpb - 1) / wpb);
        prof_begin(PROF_VXC_VB, st);
        if (vgrad)
            vxc_vb_sb_kernel<4><<<nb1, wpb * 32, 0, st>>>(sbd, ao, sbp, ngl, weights, vrho, vgrad, vb_off, vb);
        else
            vxc_vb_sb_kernel<1><<<nb1, wpb * 32, 0, st>>>(sbd, ao, sbp, ngl, weights, vrho, vgrad, vb_off, vb);
        prof_end(st);
        QC_LAUNCHED(1);
        prof_begin(PROF_I8_SLICE, st);
        sb_slice_kernel<S, BN><<<gs, 256, 0, st>>>(sbd, vb, vb_off, 0, sbp, b_off, bplanes, bscale);
        prof_end(st);
        QC_LAUNCHED(1);
    }
    constexpr int NSTAGE = (BN == 96) ? 3 : I8_STAGES;
    const size_t smem = (size_t)NSTAGE * S * (I8_A_PLANE + I8_KT * BN) + sizeof(double) * I8_BM * (BN + 1);
    prof_begin(PROF_VXC_GEMM, st);
    if ((g_i8_mode & 2) && ptile_off != nullptr && BN == 64) {
        QC_CHECK(cudaFuncSetAttribute(vxc_i8_gemm_kernel<S, 1, 64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QC_CHECK(launch_cluster2(vxc_i8_gemm_kernel<S, 1, 64, 4>, NUM_SMS, I8_THREADS, smem, st, sbd, ptile_off, nsb, nptiles, idx,
                                 aplanes, a_off, (const signed char *)bplanes, b_off, ascale, (const double *)bscale, sbp, nao,
                                 mat, g_i8_variant));
    } else {
        constexpr int EW = (BN == 96) ? 8 : 4;
        QC_CHECK(cudaFuncSetAttribute(vxc_i8_gemm_kernel<S, 0, BN, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        vxc_i8_gemm_kernel<S, 0, BN, EW><<<NUM_SMS, 128 + 32 * EW, smem, st>>>(sbd, tile_off, nsb, ntiles, idx, aplanes, a_off,
                                                                             bplanes, b_off, ascale, bscale, sbp, nao, mat,
                                                                             g_i8_variant);
    }
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

// Same contract as b200qc_vxc_sb, with the GEMM on tcgen05 int8 slices.  aplanes / a_off / ascale / colmax come from
// b200qc_vxc_i8_prepare; colmax != NULL selects the fused vb slicer (vb / vb_off are then unused and may be NULL),
// colmax == NULL the two-pass form with exact column maxima (vb: scratch of sum_sb sbp * nsp doubles); bplanes (nslice * sum_sb sbp * ceil(nsp / bn) * bn bytes at b_off[sb], ZERO-FILLED once by the
// caller: the columns past nsp of a last N tile are never written) and bscale (sum_sb nsp) are scratch;
// tile_off[sb] = exclusive prefix of ceil(nsp / 128) * ceil(nsp / bn) (device int32), ntiles = its total.
extern "C" int b200qc_vxc_sb_i8(const void *sbdesc, int nsb, int sbp, int max_nsp, int nslice, const int *idx,
                                const double *ao, const double *weights, const double *vrho, const double *vgrad,
                                int nao, const int64_t *vb_off, double *vb, const float *colmax, int *fixflag,
                                const signed char *aplanes, const int64_t *a_off, const double *ascale, signed char *bplanes,
                                const int64_t *b_off, double *bscale, int bn, const int *tile_off, int ntiles,
                                const int *ptile_off, int nptiles, double *mat, void *stream) {
    QC_REQUIRE(sbp % I8_KT == 0 && sbp % I8_BM == 0, "superblock size must be a multiple of 128");
    QC_REQUIRE(nslice == 5 || nslice == 6, "nslice must be 5 or 6");
    QC_REQUIRE(bn == 64 || (bn == 96 && nslice == 5), "N tile: 64, or 96 with 5 slices");
    QC_REQUIRE((int64_t)sbp * 6 * 4096 < (1LL << 31), "superblock too long for exact int32 accumulation");
    QC_REQUIRE(colmax != nullptr || (vb != nullptr && vb_off != nullptr), "either colmax or the vb scratch is needed");
    QC_REQUIRE(colmax == nullptr || sbp <= 1536, "fused vb slicer: superblock too long for its shared-memory table");
    QC_REQUIRE(colmax == nullptr || (fixflag != nullptr && sbp % VBS_PHASE == 0), "fused vb slicer: fixflag scratch missing");
    cudaStream_t st = as_stream(stream);
    QC_CHECK(cudaMemsetAsync(mat, 0, sizeof(double) * nao * nao, st));
    if (nsb == 0) return 0;
    const SBDesc *sbd = (const SBDesc *)sbdesc;
    if (nslice == 5 && bn == 96)
        return vxc_i8_run<5, 96>(sbd, nsb, sbp, max_nsp, idx, ao, weights, vrho, vgrad, vb, vb_off, colmax, fixflag, aplanes, a_off, ascale,
                                 bplanes, b_off, bscale, tile_off, ntiles, ptile_off, nptiles, nao, mat, st);
    if (nslice == 5)
        return vxc_i8_run<5, 64>(sbd, nsb, sbp, max_nsp, idx, ao, weights, vrho, vgrad, vb, vb_off, colmax, fixflag, aplanes, a_off, ascale,
                                 bplanes, b_off, bscale, tile_off, ntiles, ptile_off, nptiles, nao, mat, st);
    return vxc_i8_run<6, 64>(sbd, nsb, sbp, max_nsp, idx, ao, weights, vrho, vgrad, vb, vb_off, colmax, fixflag, aplanes, a_off, ascale,
                             bplanes, b_off, bscale, tile_off, ntiles, ptile_off, nptiles, nao, mat, st);
}
