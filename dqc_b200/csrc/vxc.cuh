// K4 -- Vxc integration.
// Replaces HamiltonCGTO._get_vxc_from_potinfo (dqc/hamilton/hcgto.py:445-495):
//   vb = v phi + sum_d (2 g_d) d_d phi ;  mat += (phi w)^T vb      per grid chunk.
// Two launches: (1) a streaming pass builds the weighted vb (HBM-bound: reads ncomp, writes 1
// AO-sized tensor); (2) mat = phi^T vb on the fp64 tensor pipe, split over the grid axis so that
// the nao^2/(128*64) output tiles times nsplit fill the 148 SMs, each split writing its own partial
// slab; (3) a fixed-order reduction of the slabs (deterministic -- no atomics).
// Algorithmic work: 2 ngrid nao^2 flop; (ncomp + 1) ngrid nao 8 bytes read + ngrid nao 8 written.
#pragma once
#include "gemm_f64.cuh"

template <int NCOMP>
__global__ void vxc_vb_kernel(const double *__restrict__ ao, int64_t ngrid_ld, int64_t ao_ld,
                              const double *__restrict__ w, const double *__restrict__ vrho,
                              const double *__restrict__ vgrad, double *__restrict__ vb) {
    // one warp per grid row, lanes along the AO axis (double2 per lane)
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= ngrid_ld) return;
    const int lane = threadIdx.x & 31;
    const double wg = w[g];
    const double c0 = wg * vrho[g];
    double c1 = 0, c2 = 0, c3 = 0;
    if (NCOMP == 4) {
        c1 = 2.0 * wg * vgrad[g];
        c2 = 2.0 * wg * vgrad[ngrid_ld + g];
        c3 = 2.0 * wg * vgrad[2 * ngrid_ld + g];
    }
    const double2 *p0 = reinterpret_cast<const double2 *>(ao + g * ao_ld);
    const double2 *p1 = reinterpret_cast<const double2 *>(ao + (ngrid_ld + g) * ao_ld);
    const double2 *p2 = reinterpret_cast<const double2 *>(ao + (2 * ngrid_ld + g) * ao_ld);
    const double2 *p3 = reinterpret_cast<const double2 *>(ao + (3 * ngrid_ld + g) * ao_ld);
    double2 *out = reinterpret_cast<double2 *>(vb + g * ao_ld);
    for (int c = lane; c < ao_ld / 2; c += 32) {
        double2 v = p0[c];
        double2 r = make_double2(c0 * v.x, c0 * v.y);
        if (NCOMP == 4) {
            v = p1[c]; r.x += c1 * v.x; r.y += c1 * v.y;
            v = p2[c]; r.x += c2 * v.x; r.y += c2 * v.y;
            v = p3[c]; r.x += c3 * v.x; r.y += c3 * v.y;
        }
        out[c] = r;
    }
}

// partial[split][mu][nu] = sum_{g in split} phi[g][mu] vb[g][nu]
__global__ void __launch_bounds__(GM_THREADS, 2)
vxc_gemm_kernel(const double *__restrict__ phi, const double *__restrict__ vb, int64_t ao_ld,
                int64_t rows_per_split, int64_t ngrid_ld, double *__restrict__ partial) {
    extern __shared__ __align__(16) double gm_smem[];
    const int n0 = blockIdx.x * GM_BN, m0 = blockIdx.y * GM_BM, split = blockIdx.z;
    const int64_t g0 = (int64_t)split * rows_per_split;
    const int64_t g1 = min(g0 + rows_per_split, ngrid_ld);
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    if (g1 > g0)
        gemm_tile_128x64<false>(phi + g0 * ao_ld + m0, ao_ld, (int)min((int64_t)GM_BM, ao_ld - m0),
                                vb + g0 * ao_ld + n0, ao_ld, (int)((g1 - g0) / GM_BK), acc, gm_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    double *P = partial + (int64_t)split * ao_ld * ao_ld;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = m0 + wm + i * 8 + (lane >> 2);
        if (row >= ao_ld) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int col = n0 + wn + j * 8 + 2 * (lane & 3);
            *reinterpret_cast<double2 *>(P + (int64_t)row * ao_ld + col) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

__global__ void slab_reduce_kernel(const double *__restrict__ partial, int nslab, int64_t n, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < nslab; k++) s += partial[(int64_t)k * n + i];
    out[i] = s;
}

static int vxc_nsplit(int64_t ngrid_ld, int64_t ao_ld) {
    const int64_t tiles = ((ao_ld + GM_BM - 1) / GM_BM) * (ao_ld / GM_BN);
    int64_t want = (4 * 2 * NUM_SMS + tiles - 1) / tiles;  // ~4 waves of 2 CTAs/SM
    const int64_t maxsplit = ngrid_ld / 1024 > 0 ? ngrid_ld / 1024 : 1;
    if (want > maxsplit) want = maxsplit;
    if (want < 1) want = 1;
    if (want > 256) want = 256;
    return (int)want;
}

extern "C" int64_t b200qc_vxc_worksize(int64_t ngrid_ld, int64_t ao_ld) {
    return ngrid_ld * ao_ld + (int64_t)vxc_nsplit(ngrid_ld, ao_ld) * ao_ld * ao_ld;
}

extern "C" int b200qc_vxc_mat(const double *ao, int64_t ngrid_ld, int64_t ao_ld, const double *weights,
                              const double *vrho, const double *vgrad, double *mat, double *work,
                              void *stream) {
    QC_REQUIRE(ngrid_ld % GM_BM == 0 && ao_ld % GM_BN == 0, "ngrid_ld must be a multiple of 128 and ao_ld of 64");
    cudaStream_t st = as_stream(stream);
    double *vb = work, *partial = work + ngrid_ld * ao_ld;
    const int wpb = 8;
    const unsigned nb1 = (unsigned)((ngrid_ld + wpb - 1) / wpb);
    prof_begin(PROF_VXC_VB, st);
    if (vgrad)
        vxc_vb_kernel<4><<<nb1, wpb * 32, 0, st>>>(ao, ngrid_ld, ao_ld, weights, vrho, vgrad, vb);
    else
        vxc_vb_kernel<1><<<nb1, wpb * 32, 0, st>>>(ao, ngrid_ld, ao_ld, weights, vrho, vgrad, vb);
    prof_end(st);
    QC_LAUNCHED(1);
    const int nsplit = vxc_nsplit(ngrid_ld, ao_ld);
    int64_t rows = (ngrid_ld + nsplit - 1) / nsplit;
    rows = (rows + GM_BK - 1) / GM_BK * GM_BK;
    dim3 grid((unsigned)(ao_ld / GM_BN), (unsigned)((ao_ld + GM_BM - 1) / GM_BM), (unsigned)nsplit);
    QC_CHECK(cudaFuncSetAttribute(vxc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES));
    prof_begin(PROF_VXC_GEMM, st);
    vxc_gemm_kernel<<<grid, GM_THREADS, GM_SMEM_BYTES, st>>>(ao, vb, ao_ld, rows, ngrid_ld, partial);
    prof_end(st);
    QC_LAUNCHED(1);
    const int64_t n = ao_ld * ao_ld;
    prof_begin(PROF_VXC_REDUCE, st);
    slab_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, nsplit, n, mat);
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}
