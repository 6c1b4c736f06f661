// fp64 tensor-core tile engine shared by K2 (density on grid), K4 (Vxc integration) and the DF
// build GEMM.  tcgen05.mma has no f64 kind, and the parity bar of this path is 1e-6 elementwise /
// 1e-8 Ha on fp64 operands, so the tensor pipe is driven through the fp64 DMMA shape
// mma.sync.m8n8k4 (B200 keeps the full-rate fp64 tensor datapath).  Operand tiles are staged in
// shared memory by a 3-stage cp.async ring; paddings are chosen so that both the 16-byte async
// stores and the fragment loads are bank-conflict-free.
//
// CTA tile 128 (M) x 64 (N), K step 16, 256 threads = 8 warps arranged 4 (M) x 2 (N), each warp
// owns 32 x 32 = 4 x 4 DMMA tiles (32 accumulator registers per thread).
#pragma once
#include "common.cuh"

#define GM_BM 128
#define GM_BN 64
#define GM_BK 16
#define GM_STAGES 3
#define GM_THREADS 256
#define GM_A_STRIDE_K 20    // A stored [m][k], k contiguous:   20 mod 16 == 4
#define GM_A_STRIDE_M 132   // A stored [k][m], m contiguous:  132 mod 16 == 4
#define GM_B_STRIDE 68      // B stored [k][n]:                 68 mod 16 == 4
#define GM_A_TILE 2560      // max(128*20, 16*132) doubles
#define GM_B_TILE (GM_BK * GM_B_STRIDE)
#define GM_STAGE_DOUBLES (GM_A_TILE + GM_B_TILE)
#define GM_SMEM_BYTES (GM_STAGES * GM_STAGE_DOUBLES * 8)

__device__ __forceinline__ void dmma8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;  // src-size 0 => the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Accumulates acc += A(128 x K) * B(K x 64) for one CTA tile.
//   A_KCONTIG: element (m, k) at A[m * lda + k]   (rows >= m_valid are read as zero)
//   else     : element (m, k) at A[k * lda + m]   (columns >= m_valid are read as zero)
//   B: element (k, n) at B[k * ldb + n].  K = ktiles * 16; all K rows must be readable.
// acc[i][j][e] <-> row = warp_m*32 + i*8 + lane/4, col = warp_n*32 + j*8 + 2*(lane%4) + e.
template <bool A_KCONTIG>
__device__ __forceinline__ void gemm_tile_128x64(const double *__restrict__ A, int64_t lda, int m_valid,
                                                 const double *__restrict__ B, int64_t ldb, int ktiles,
                                                 double (&acc)[4][4][2], double *smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int lr = lane >> 2, lc = lane & 3;

    auto load_stage = [&](int kt, int slot) {
        double *As = smem + slot * GM_STAGE_DOUBLES;
        double *Bs = As + GM_A_TILE;
        const int64_t k0 = (int64_t)kt * GM_BK;
        if (A_KCONTIG) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = tid + i * GM_THREADS;
                const int row = c >> 3, ch = c & 7;
                const bool ok = row < m_valid;
                cp_async16(As + row * GM_A_STRIDE_K + ch * 2, ok ? A + (int64_t)row * lda + k0 + ch * 2 : A, ok);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = tid + i * GM_THREADS;
                const int kr = c >> 6, ch = c & 63;
                const bool ok = ch * 2 < m_valid;
                cp_async16(As + kr * GM_A_STRIDE_M + ch * 2, ok ? A + (k0 + kr) * lda + ch * 2 : A, ok);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int c = tid + i * GM_THREADS;
            const int kr = c >> 5, ch = c & 31;
            cp_async16(Bs + kr * GM_B_STRIDE + ch * 2, B + (k0 + kr) * ldb + ch * 2, true);
        }
    };

    __syncthreads();  // previous users of the ring are done
#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; s++) {
        if (s < ktiles) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < ktiles; kt++) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();
        const int nxt = kt + GM_STAGES - 1;
        if (nxt < ktiles) load_stage(nxt, nxt % GM_STAGES);
        cp_async_commit();
        const double *As = smem + (kt % GM_STAGES) * GM_STAGE_DOUBLES;
        const double *Bs = As + GM_A_TILE;
#pragma unroll
        for (int kk = 0; kk < GM_BK / 4; kk++) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
                a[i] = A_KCONTIG ? As[(wm + i * 8 + lr) * GM_A_STRIDE_K + kk * 4 + lc]
                                 : As[(kk * 4 + lc) * GM_A_STRIDE_M + wm + i * 8 + lr];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[(kk * 4 + lc) * GM_B_STRIDE + wn + j * 8 + lr];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Plain C = A * B (+ split over K into partial slabs), used by the DF build and tests.
//   A: (M, K) row-major (k contiguous), B: (K, N) row-major; M % 128 == 0 not required (guarded),
//   N % 64 == 0 and K % 16 == 0 required.
__global__ void __launch_bounds__(GM_THREADS, 2)
gemm_nn_kernel(const double *__restrict__ A, int64_t lda, const double *__restrict__ B, int64_t ldb,
               double *__restrict__ C, int64_t ldc, int M, int K) {
    extern __shared__ __align__(16) double gm_smem[];
    const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    gemm_tile_128x64<true>(A + (int64_t)m0 * lda, lda, M - m0, B + n0, ldb, K / GM_BK, acc, gm_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = m0 + wm + i * 8 + (lane >> 2);
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int col = n0 + wn + j * 8 + 2 * (lane & 3);
            *reinterpret_cast<double2 *>(C + (int64_t)row * ldc + col) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fp64 tensor-pipe peak: register-resident DMMA issue loop (no memory traffic), 8 independent
// accumulator chains per warp, 2 CTAs of 256 threads per SM.  The driver-written
// MEASURED_PEAKS.json holds HBM and bf16 numbers only, so the fp64 roofline denominator of K2/K4
// is measured by this kernel in the same process as the benchmark.
__global__ void __launch_bounds__(256, 2) dmma_peak_kernel(int iters, double *out) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma8x8x4(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;  // keeps the loop alive
}

// returns TFLOP/s (2 * 8 * 8 * 4 flop per DMMA) in *tflops
extern "C" int b200qc_peak_fp64_dmma(int iters, double *scratch, double *tflops, void *stream) {
    cudaStream_t st = as_stream(stream);
    const int nblk = NUM_SMS * 2;
    cudaEvent_t e0, e1;
    QC_CHECK(cudaEventCreate(&e0));
    QC_CHECK(cudaEventCreate(&e1));
    dmma_peak_kernel<<<nblk, 256, 0, st>>>(iters / 8 + 1, scratch);  // warm-up
    QC_CHECK(cudaEventRecord(e0, st));
    dmma_peak_kernel<<<nblk, 256, 0, st>>>(iters, scratch);
    QC_CHECK(cudaEventRecord(e1, st));
    QC_LAUNCHED(2);
    QC_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    QC_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 8 * 8 * 4 * 8.0 * iters * (256 / 32) * (double)nblk;
    *tflops = flop / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}
