// K2 -- density and density gradient on the grid.
// Replaces HamiltonCGTO._dm2densinfo (dqc/hamilton/hcgto.py:371-443): per grid chunk
// `dmao = basis @ D`, rho = rowsum(dmao * basis), grad_d = 2 rowsum(dmao * grad_basis_d).
// Here: one CTA owns 128 grid points and sweeps the AO axis in 64-column tiles; each tile of
// X = phi D comes out of the fp64 tensor pipe (gemm_f64.cuh) and is contracted immediately, in
// registers, against the matching tile of phi / d phi -- X never touches memory, phi is read
// from HBM once per component, D (nao^2) lives in L2.
// Algorithmic work per launch: 2 ngrid nao^2 flop; ncomp ngrid nao 8 bytes.
#pragma once
#include "gemm_f64.cuh"

template <int NCOMP>
__global__ void __launch_bounds__(GM_THREADS, 2)
rho_kernel(const double *__restrict__ ao, int64_t ngrid_ld, int64_t ao_ld, const double *__restrict__ dm,
           double *__restrict__ rho, double *__restrict__ grad) {
    extern __shared__ __align__(16) double gm_smem[];
    __shared__ double red[2][GM_BM][NCOMP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int64_t m0 = (int64_t)blockIdx.x * GM_BM;
    const double *A = ao + m0 * ao_ld;  // component 0 = values
    double part[NCOMP][4];
#pragma unroll
    for (int c = 0; c < NCOMP; c++)
#pragma unroll
        for (int i = 0; i < 4; i++) part[c][i] = 0.0;

    const int ntile = (int)(ao_ld / GM_BN), ktiles = (int)(ao_ld / GM_BK);
    for (int nt = 0; nt < ntile; nt++) {
        const int n0 = nt * GM_BN;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        gemm_tile_128x64<true>(A, ao_ld, GM_BM, dm + n0, ao_ld, ktiles, acc, gm_smem);
#pragma unroll
        for (int c = 0; c < NCOMP; c++) {
            const double *comp = ao + ((int64_t)c * ngrid_ld + m0) * ao_ld + n0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double *row = comp + (int64_t)(wm + i * 8 + (lane >> 2)) * ao_ld + wn + 2 * (lane & 3);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double2 v = *reinterpret_cast<const double2 *>(row + j * 8);
                    part[c][i] += acc[i][j][0] * v.x + acc[i][j][1] * v.y;
                }
            }
        }
    }
    // quad reduction (the 4 lanes of a quad hold the same rows), then the two N-halves via smem
#pragma unroll
    for (int c = 0; c < NCOMP; c++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double v = part[c][i];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if ((lane & 3) == 0) red[warp & 1][wm + i * 8 + (lane >> 2)][c] = v;
        }
    __syncthreads();
    if (threadIdx.x < GM_BM) {
        const int r = threadIdx.x;
        rho[m0 + r] = red[0][r][0] + red[1][r][0];
        if (NCOMP == 4) {
#pragma unroll
            for (int d = 0; d < 3; d++) grad[(int64_t)d * ngrid_ld + m0 + r] = 2.0 * (red[0][r][d + 1] + red[1][r][d + 1]);
        }
    }
}

extern "C" int b200qc_rho(const double *ao, int64_t ngrid_ld, int64_t ao_ld, const double *dm, double *rho,
                          double *grad, void *stream) {
    QC_REQUIRE(ngrid_ld % GM_BM == 0 && ao_ld % GM_BN == 0, "ngrid_ld must be a multiple of 128 and ao_ld of 64");
    if (ngrid_ld == 0) return 0;
    const unsigned nblk = (unsigned)(ngrid_ld / GM_BM);
    prof_begin(PROF_RHO, as_stream(stream));
    if (grad) {
        QC_CHECK(cudaFuncSetAttribute(rho_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES));
        rho_kernel<4><<<nblk, GM_THREADS, GM_SMEM_BYTES, as_stream(stream)>>>(ao, ngrid_ld, ao_ld, dm, rho, grad);
    } else {
        QC_CHECK(cudaFuncSetAttribute(rho_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES));
        rho_kernel<1><<<nblk, GM_THREADS, GM_SMEM_BYTES, as_stream(stream)>>>(ao, ngrid_ld, ao_ld, dm, rho, grad);
    }
    prof_end(as_stream(stream));
    QC_LAUNCHED(1);
    return 0;
}
