// Becke partition weights -- replaces the per-atom torch loop of
// dqc/grid/multiatoms_grid.py:173-273 (O(natom^2 ngrid) on the host in the reference).
// One thread per grid point; the point's distances to all atoms sit in a shared-memory column
// (stride = blockDim, conflict-free); pair data (R_ij, a_ij) are warp-uniform global reads.
#pragma once
#include "common.cuh"

#define BECKE_THREADS 128

__global__ void becke_pairs_kernel(const double *__restrict__ pos, int nat, double *__restrict__ rinv) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nat * nat) return;
    int i = idx / nat, j = idx % nat;
    // the reference adds the identity to the displacement of the diagonal (:187-190) -> |(1,1,1)|
    double e = (i == j) ? 1.0 : 0.0;
    double dx = pos[3 * j] - pos[3 * i] + e, dy = pos[3 * j + 1] - pos[3 * i + 1] + e,
           dz = pos[3 * j + 2] - pos[3 * i + 2] + e;
    rinv[idx] = sqrt(dx * dx + dy * dy + dz * dz);  // R_ij itself: the weight kernel divides, like the reference
}

__global__ void __launch_bounds__(BECKE_THREADS)
becke_weights_kernel(const double *__restrict__ xyz, const int *__restrict__ owner, int64_t ngrid,
                     const double *__restrict__ pos, int nat, const double *__restrict__ rinv,
                     const double *__restrict__ aij, double *__restrict__ w) {
    extern __shared__ double rg[];  // [nat][blockDim.x]
    const int tid = threadIdx.x;
    const int BT = blockDim.x;       // 128, or fewer points per CTA when many atoms must fit in shared memory
    const int64_t g = (int64_t)blockIdx.x * BT + tid;
    if (g >= ngrid) return;  // no block-wide barrier below
    const double x = xyz[3 * g], y = xyz[3 * g + 1], z = xyz[3 * g + 2];
    for (int k = 0; k < nat; k++) {
        const double dx = x - pos[3 * k], dy = y - pos[3 * k + 1], dz = z - pos[3 * k + 2];
        rg[k * BT + tid] = sqrt(dx * dx + dy * dy + dz * dz);
    }
    const int own = owner[g];
    const double sdiag = 0.5 * (1.0 + 1e-12) + 0.5;  // the i == j factor, kept for fidelity (:259)
    double psum = 0.0, pown = 0.0;
    for (int j = 0; j < nat; j++) {
        const double rj = rg[j * BT + tid];
        double P = sdiag;
        bool keep = true;
        for (int i = 0; i < nat; i++) {
            if (i == j) continue;
            double mu = (rj - rg[i * BT + tid]) / rinv[i * nat + j];
            if (aij) mu += aij[i * nat + j] * (1.0 - mu * mu);
            if (!(mu < 0.74)) {
                keep = false;
                break;
            }
            double f = mu;
#pragma unroll
            for (int it = 0; it < 3; it++) f = 0.5 * f * (3.0 - f * f);
            P *= 0.5 * (1.0 + 1e-12 - f);
        }
        if (keep) {
            psum += P;
            if (j == own) pown = P;
        }
    }
    w[g] = pown / psum;
}

extern "C" int b200qc_becke_weights(const double *xyz, const int *owner, int64_t ngrid,
                                    const double *atompos, int natom, const double *aij, double *w,
                                    void *stream) {
    QC_REQUIRE(natom >= 1, "no atoms");
    if (ngrid == 0) return 0;
    cudaStream_t st = as_stream(stream);
    double *rinv = nullptr;
    QC_CHECK(cudaMallocAsync(&rinv, sizeof(double) * natom * natom, st));
    becke_pairs_kernel<<<(natom * natom + 255) / 256, 256, 0, st>>>(atompos, natom, rinv);
    QC_LAUNCHED(1);
    int bt = BECKE_THREADS;          // points per CTA: halve until the distance column fits (<= 800 atoms)
    while (bt > 32 && sizeof(double) * natom * bt > 200 * 1024) bt >>= 1;
    const size_t smem = sizeof(double) * natom * bt;
    QC_REQUIRE(smem <= 200 * 1024, "too many atoms for the shared-memory distance column (> 800)");
    QC_CHECK(cudaFuncSetAttribute(becke_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nblk = (int)((ngrid + bt - 1) / bt);
    prof_begin(PROF_BECKE, st);
    becke_weights_kernel<<<nblk, bt, smem, st>>>(xyz, owner, ngrid, atompos, natom, rinv, aij, w);
    prof_end(st);
    QC_LAUNCHED(1);
    QC_CHECK(cudaFreeAsync(rinv, st));
    return 0;
}
