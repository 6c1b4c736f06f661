// Prints which (TMEM lane, column) each register of tcgen05.ld.16x256b.x1 / .x2 returns, by writing
// value = lane * 1000 + column with tcgen05.st.32x32b first.  (No PTX manual offline: measured, then relied on.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(int *out) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tbase + ((uint32_t)(warp * 32) << 16);
    // 32x32b.x16 store: thread = lane, 16 columns
    uint32_t v[16];
    for (int j = 0; j < 16; j++) v[j] = (warp * 32 + lane) * 1000 + j;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(tmem), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t a[4], b[4], c[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(tmem));
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(tmem + (16u << 16)));
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]) : "r"(tmem));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    int *o = out + threadIdx.x * 16;
    for (int j = 0; j < 4; j++) o[j] = a[j];
    for (int j = 0; j < 4; j++) o[4 + j] = b[j];
    for (int j = 0; j < 8; j++) o[8 + j] = c[j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(32));
}
int main() {
    int *d, h[128 * 16];
    cudaMalloc(&d, sizeof(h));
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int t = 0; t < 128; t += (t < 32 ? 1 : 32)) {
        printf("thread %3d: x1@0:", t);
        for (int j = 0; j < 4; j++) printf(" (%d,%d)", h[t * 16 + j] / 1000, h[t * 16 + j] % 1000);
        printf(" | x1@16:");
        for (int j = 0; j < 4; j++) printf(" (%d,%d)", h[t * 16 + 4 + j] / 1000, h[t * 16 + 4 + j] % 1000);
        printf(" | x2@0:");
        for (int j = 0; j < 8; j++) printf(" (%d,%d)", h[t * 16 + 8 + j] / 1000, h[t * 16 + 8 + j] % 1000);
        printf("\n");
    }
    return 0;
}
