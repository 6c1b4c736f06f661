// Measured issue-rate peak of tcgen05.mma.kind::i8 on this device: the denominator of the int8 rooflines that
// bench.py reports (MEASURED_PEAKS.json has no int8 entry).  One CTA per SM issues `iters` back-to-back
// M = 128, N = 256, K = 32 MMAs on resident shared-memory operands (no loads in the loop, two alternating
// 256-column accumulators) -- the int8 analogue of b200qc_peak_fp64_dmma.  At this shape an MMA reads 12 KB of
// shared memory in 128 tensor cycles, so the tensor pipe, not the operand port, sets the rate.
#pragma once
#include "vxc_i8.cuh"

__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters) {
    extern __shared__ __align__(1024) unsigned char i8_smem[];
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int A_BYTES = I8_KT * 128, B_BYTES = I8_KT * 256;
    for (int i = tid; i < (A_BYTES + B_BYTES) / 4; i += 128) reinterpret_cast<uint32_t *>(i8_smem)[i] = 0x01010101u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the MMA
    if (tid == 0) {
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    if (tid == 0) {
        // K-major, no swizzle (as in rho_i8.cuh): LBO = rows x 16 B between the two K chunks, SBO = 128 B
        const uint32_t sbase = smem_u32(i8_smem);
        const uint64_t da = umma_desc(sbase, 128 * 16, 128), db = umma_desc(sbase + A_BYTES, 256 * 16, 128);
        constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int it = 0; it < iters; it++) umma_i8(tmem + (it & 1) * 256, da, db, IDESC, it > 1 ? 1u : 0u);
        umma_commit(&done_bar);
        mbar_wait(&done_bar, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// *tops = int8 ops per second (2 per multiply-add) in units of 1e12, summed over all SMs
extern "C" int b200qc_peak_i8_mma(int iters, double *tops, void *stream) {
    QC_REQUIRE(iters >= 64 && tops != nullptr, "bad arguments");
    cudaStream_t st = as_stream(stream);
    const size_t smem = (size_t)I8_KT * (128 + 256);
    cudaEvent_t e0, e1;
    QC_CHECK(cudaEventCreate(&e0));
    QC_CHECK(cudaEventCreate(&e1));
    i8_peak_kernel<<<NUM_SMS, 128, smem, st>>>(iters / 8 + 64);   // warm-up
    QC_CHECK(cudaEventRecord(e0, st));
    i8_peak_kernel<<<NUM_SMS, 128, smem, st>>>(iters);
    QC_CHECK(cudaEventRecord(e1, st));
    QC_LAUNCHED(2);
    QC_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    QC_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *tops = 2.0 * 128 * 256 * 32 * (double)iters * NUM_SMS / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}
