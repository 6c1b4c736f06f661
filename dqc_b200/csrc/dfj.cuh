// K9 -- density-fitted Coulomb matrix.  Replaces DFMol.get_elrep (dqc/df/dfmol.py:60-79):
//   temp_P = sum_ij D_ij (ij|P);  c = temp . inv_j2c;  J_ij = sum_P c_P (ij|P).
// (ij|P) is held once, packed over i >= j: B[pair][P], pair = i(i+1)/2 + j, row stride `ld`.
// Both contractions are GEMVs over B (0.25 flop/byte) => HBM-bound: 2 * npair * naux * 8 bytes per
// call.  Pass 1 streams the rows of B through CTAs that keep 8 columns per
// thread in registers (one contiguous slab of rows per CTA) and writes one partial row per CTA (fixed-order reduction afterwards: deterministic, no
// atomics); pass 2 is one warp per pair row with the fitted coefficients served from L1.  Pair rows that are
// negligible in every column (b200qc_dfj_rowmask) are read by neither pass.
#pragma once
#include "common.cuh"

#define DFJ_TPB 256
#define DFJ_SLOTS 4                       // double2 column slots per thread
#define DFJ_COLS (DFJ_TPB * DFJ_SLOTS * 2)  // columns per CTA in pass 1

__device__ __forceinline__ void pair_to_ij(int64_t r, int &i, int &j) {
    int64_t ii = (int64_t)((sqrt(8.0 * (double)r + 1.0) - 1.0) * 0.5);
    while (ii * (ii + 1) / 2 > r) ii--;
    while ((ii + 1) * (ii + 2) / 2 <= r) ii++;
    i = (int)ii;
    j = (int)(r - ii * (ii + 1) / 2);
}

// dvec[pair] = D_ij + D_ji (i > j) or D_ii
__global__ void dfj_gather_dm_kernel(const double *__restrict__ dm, int nao, int64_t npair, double *__restrict__ dvec) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= npair) return;
    int i, j;
    pair_to_ij(r, i, j);
    dvec[r] = i == j ? dm[(int64_t)i * nao + i] : dm[(int64_t)i * nao + j] + dm[(int64_t)j * nao + i];
}

// LIST: the CTA walks entries [i0, i1) of a list of pair rows (the rows b200qc_dfj_rowmask keeps, ascending) instead of
// the rows themselves; the next index is fetched one step ahead so that the row loads never wait for it.  (A mask test
// with a data-dependent `continue` in front of the loads cost more than the skipped rows saved.)
template <bool LIST>
__global__ void __launch_bounds__(DFJ_TPB)
dfj_pass1_kernel(const double *__restrict__ B, int64_t nrows, int64_t ld, const double *__restrict__ dvec,
                 int64_t rows_per_cta, double *__restrict__ partial, const int *__restrict__ rows) {
    const int64_t i0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t i1 = min(i0 + rows_per_cta, nrows);
    const int64_t c0 = (int64_t)blockIdx.y * DFJ_COLS + 2 * threadIdx.x;
    double2 acc[DFJ_SLOTS];
    bool ok[DFJ_SLOTS];
#pragma unroll
    for (int s = 0; s < DFJ_SLOTS; s++) {
        acc[s] = make_double2(0.0, 0.0);
        ok[s] = c0 + (int64_t)s * 2 * DFJ_TPB < ld;
    }
    if (LIST) {
        int64_t r = i0 < i1 ? rows[i0] : 0;
        for (int64_t i = i0; i < i1; i++) {
            const int64_t rn = i + 1 < i1 ? rows[i + 1] : 0;
            const double *row = B + r * ld + c0;
            const double d = __ldg(dvec + r);
#pragma unroll
            for (int s = 0; s < DFJ_SLOTS; s++)
                if (ok[s]) {
                    const double2 v = __ldcs(reinterpret_cast<const double2 *>(row + s * 2 * DFJ_TPB));
                    acc[s].x += d * v.x;
                    acc[s].y += d * v.y;
                }
            r = rn;
        }
    } else {
        const double *row = B + i0 * ld + c0;
        for (int64_t r = i0; r < i1; r++, row += ld) {
            const double d = __ldg(dvec + r);
#pragma unroll
            for (int s = 0; s < DFJ_SLOTS; s++)
                if (ok[s]) {
                    const double2 v = __ldcs(reinterpret_cast<const double2 *>(row + s * 2 * DFJ_TPB));
                    acc[s].x += d * v.x;
                    acc[s].y += d * v.y;
                }
        }
    }
    double *out = partial + (int64_t)blockIdx.x * ld + c0;
#pragma unroll
    for (int s = 0; s < DFJ_SLOTS; s++)
        if (ok[s]) *reinterpret_cast<double2 *>(out + s * 2 * DFJ_TPB) = acc[s];
}

// Narrow slabs (ld <= 512: the aux slice of one rank in an 8-GPU build is 428 columns): one double2 column slot per
// thread would leave a single load in flight per thread (0.6 of the HBM rate on the 8-GPU C60 build); here a thread
// keeps FOUR consecutive rows in flight for its column pair.
__global__ void __launch_bounds__(DFJ_TPB)
dfj_pass1_narrow_kernel(const double *__restrict__ B, int64_t nrows, int64_t ld, const double *__restrict__ dvec,
                        int64_t rows_per_cta, double *__restrict__ partial, const int *__restrict__ rows) {
    const int64_t i0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t i1 = min(i0 + rows_per_cta, nrows);
    const int64_t c0 = 2 * threadIdx.x;
    if (c0 >= ld) return;
    double2 acc[4];
#pragma unroll
    for (int u = 0; u < 4; u++) acc[u] = make_double2(0.0, 0.0);
    int64_t i = i0;
    for (; i + 4 <= i1; i += 4) {
        int64_t r[4];
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) r[u] = rows ? (int64_t)rows[i + u] : i + u;
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = __ldcs(reinterpret_cast<const double2 *>(B + r[u] * ld + c0));
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const double d = __ldg(dvec + r[u]);
            acc[u].x += d * v[u].x;
            acc[u].y += d * v[u].y;
        }
    }
    for (; i < i1; i++) {
        const int64_t r = rows ? (int64_t)rows[i] : i;
        const double d = __ldg(dvec + r);
        const double2 v = __ldcs(reinterpret_cast<const double2 *>(B + r * ld + c0));
        acc[0].x += d * v.x;
        acc[0].y += d * v.y;
    }
    // fixed order: the four row phases, then (dfj_reduce_kernel) the slabs
    *reinterpret_cast<double2 *>(partial + (int64_t)blockIdx.x * ld + c0) =
        make_double2((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y));
}

__global__ void dfj_reduce_kernel(const double *__restrict__ partial, int nslab, int64_t ld, int64_t n, double *__restrict__ t) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double s = 0.0;
    for (int k = 0; k < nslab; k++) s += partial[(int64_t)k * ld + c];
    t[c] = s;
}

// c[k] = sum_l t[l] M[l][k]   (temp @ inv_j2c, dfmol.py:72); block (32, 8): 32 columns x 8 row slices
__global__ void dfj_vecmat_kernel(const double *__restrict__ t, const double *__restrict__ M, int64_t n, double *__restrict__ c) {
    __shared__ double red[8][33];
    const int64_t k = (int64_t)blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    if (k < n)
        for (int64_t l = threadIdx.y; l < n; l += 8) s += t[l] * M[l * n + k];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && k < n) {
        double v = 0.0;
#pragma unroll
        for (int y = 0; y < 8; y++) v += red[y][threadIdx.x];
        c[k] = v;
    }
}

// J[i][j] = J[j][i] = sum_P B[pair][P] c[P]; one warp per row, two rows in flight
__global__ void __launch_bounds__(256)
dfj_pass2_kernel(const double *__restrict__ B, int64_t npair, int64_t naux, int64_t ld, const double *__restrict__ c,
                 int nao, double *__restrict__ vj, const unsigned char *__restrict__ mask) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t nhalf = ld / 2;
    for (int64_t r = warp; r < npair; r += nwarp) {
        const double2 *row = reinterpret_cast<const double2 *>(B + r * ld);
        double s0 = 0.0, s1 = 0.0;
        int64_t k = (mask && mask[r]) ? nhalf : lane;   // negligible pair row: J_ij = 0 without reading it
        for (; k + 32 < nhalf; k += 64) {
            const double2 v0 = __ldcs(row + k), v1 = __ldcs(row + k + 32);
            const double2 c0 = __ldg(reinterpret_cast<const double2 *>(c) + k);
            const double2 c1 = __ldg(reinterpret_cast<const double2 *>(c) + k + 32);
            s0 += v0.x * c0.x + v0.y * c0.y;
            s1 += v1.x * c1.x + v1.y * c1.y;
        }
        for (; k < nhalf; k += 32) {
            const double2 v0 = __ldcs(row + k);
            const double2 c0 = __ldg(reinterpret_cast<const double2 *>(c) + k);
            s0 += v0.x * c0.x + v0.y * c0.y;
        }
        double s = s0 + s1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            int i, j;
            pair_to_ij(r, i, j);
            vj[(int64_t)i * nao + j] = s;
            vj[(int64_t)j * nao + i] = s;
        }
    }
}

// mask[r] = 1 when every |(ij|P)| of pair row r (this rank's columns) is below thresh: both passes then skip the row.
// On C60/def2-SVP 22 % of the rows are below 1e-14 (pairs of tight functions on distant atoms), 35 % on the 113-atom
// system: the error they would add is of the order of the rounding of the sums they are left out of.
__global__ void __launch_bounds__(256)
dfj_rowmask_kernel(const double *__restrict__ B, int64_t npair, int64_t naux, int64_t ld, double thresh,
                   unsigned char *__restrict__ mask) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= npair) return;
    double m = 0.0;
    for (int64_t k = lane; k < naux; k += 32) m = fmax(m, fabs(B[r * ld + k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) mask[r] = m < thresh ? 1 : 0;
}

extern "C" int b200qc_dfj_rowmask(const double *j3c, int64_t nao, int64_t naux, int64_t ld, double thresh,
                                  unsigned char *mask, void *stream) {
    QC_REQUIRE(j3c && mask && ld >= naux, "bad arguments");
    const int64_t npair = nao * (nao + 1) / 2;
    dfj_rowmask_kernel<<<(unsigned)((npair + 7) / 8), 256, 0, as_stream(stream)>>>(j3c, npair, naux, ld, thresh, mask);
    QC_LAUNCHED(1);
    return 0;
}

__global__ void pack_tril_kernel(const double *__restrict__ full, int nao, int64_t naux, int64_t ld, int64_t npair,
                                 double *__restrict__ packed) {
    const int64_t r = blockIdx.x;
    if (r >= npair) return;
    int i, j;
    pair_to_ij(r, i, j);
    const double *src = full + ((int64_t)i * nao + j) * naux;
    double *dst = packed + r * ld;
    for (int64_t k = threadIdx.x; k < ld; k += blockDim.x) dst[k] = k < naux ? src[k] : 0.0;
}

static int dfj_nslab(int64_t npair, int64_t ld) {
    const int64_t ny = (ld + DFJ_COLS - 1) / DFJ_COLS;
    int64_t nx = (8 * NUM_SMS + ny - 1) / ny;
    if (nx > (npair + 31) / 32) nx = (npair + 31) / 32;
    if (nx < 1) nx = 1;
    return (int)nx;
}

// work layout (doubles): dvec[npair rounded up to even] | t[ld] | c[ld] | partial[nslab][ld]
// (every segment starts 16-byte aligned: pass 1 and pass 2 use double2 accesses)
static inline int64_t dfj_dvec_len(int64_t npair) { return (npair + 1) & ~(int64_t)1; }
extern "C" int64_t b200qc_dfj_worksize(int64_t nao, int64_t ld) {
    const int64_t npair = nao * (nao + 1) / 2;
    return dfj_dvec_len(npair) + 2 * ld + (int64_t)dfj_nslab(npair, ld) * ld;
}

// temp_P = sum_ij D_ij (ij|P) over this rank's columns (pass 1 only; multi-GPU DF splits the aux axis)
// pass 1 over a list of pair rows (ascending row indices, e.g. the rows b200qc_dfj_rowmask keeps); rows = NULL: all rows
extern "C" int b200qc_dfj_pass1_rows(const double *j3c, int64_t nao, int64_t naux, int64_t ld, const double *dm,
                                     double *temp, double *work, const int *rows, int64_t nrows, void *stream) {
    QC_REQUIRE(ld % 2 == 0 && ld >= naux, "ld must be even and >= naux");
    QC_REQUIRE(((uintptr_t)j3c | (uintptr_t)work) % 16 == 0, "j3c and work must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int64_t npair = nao * (nao + 1) / 2;
    if (!rows) nrows = npair;
    QC_REQUIRE(nrows >= 0 && nrows <= npair, "bad row list");
    double *dvec = work, *partial = work + dfj_dvec_len(npair) + 2 * ld;
    prof_begin(PROF_DFJ_SMALL, st);
    dfj_gather_dm_kernel<<<(unsigned)((npair + 255) / 256), 256, 0, st>>>(dm, (int)nao, npair, dvec);
    prof_end(st);
    QC_LAUNCHED(1);
    const int nslab = dfj_nslab(npair, ld);
    const int64_t per = (nrows + nslab - 1) / nslab;
    dim3 grid((unsigned)nslab, (unsigned)((ld + DFJ_COLS - 1) / DFJ_COLS));
    prof_begin(PROF_DFJ_PASS1, st);
    if (ld <= 2 * DFJ_TPB) dfj_pass1_narrow_kernel<<<grid, DFJ_TPB, 0, st>>>(j3c, nrows, ld, dvec, per, partial, rows);
    else if (rows) dfj_pass1_kernel<true><<<grid, DFJ_TPB, 0, st>>>(j3c, nrows, ld, dvec, per, partial, rows);
    else dfj_pass1_kernel<false><<<grid, DFJ_TPB, 0, st>>>(j3c, nrows, ld, dvec, per, partial, rows);
    prof_end(st);
    QC_LAUNCHED(1);
    prof_begin(PROF_DFJ_SMALL, st);
    dfj_reduce_kernel<<<(unsigned)((naux + 255) / 256), 256, 0, st>>>(partial, nslab, ld, naux, temp);
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

extern "C" int b200qc_dfj_pass1(const double *j3c, int64_t nao, int64_t naux, int64_t ld, const double *dm,
                                double *temp, double *work, void *stream) {
    return b200qc_dfj_pass1_rows(j3c, nao, naux, ld, dm, temp, work, nullptr, 0, stream);
}

// vj_ij = sum_P (ij|P) c_P over this rank's columns (c padded to ld with zeros by the caller)
extern "C" int b200qc_dfj_pass2_masked(const double *j3c, int64_t nao, int64_t naux, int64_t ld, const double *coef,
                                       double *vj, const unsigned char *mask, void *stream) {
    QC_REQUIRE(ld % 2 == 0 && ld >= naux, "ld must be even and >= naux");
    QC_REQUIRE(((uintptr_t)j3c | (uintptr_t)coef) % 16 == 0, "j3c and coef must be 16-byte aligned");
    const int64_t npair = nao * (nao + 1) / 2;
    prof_begin(PROF_DFJ_PASS2, as_stream(stream));
    dfj_pass2_kernel<<<NUM_SMS * 8, 256, 0, as_stream(stream)>>>(j3c, npair, naux, ld, coef, (int)nao, vj, mask);
    prof_end(as_stream(stream));
    QC_LAUNCHED(1);
    return 0;
}

extern "C" int b200qc_dfj_pass2(const double *j3c, int64_t nao, int64_t naux, int64_t ld, const double *coef,
                                double *vj, void *stream) {
    return b200qc_dfj_pass2_masked(j3c, nao, naux, ld, coef, vj, nullptr, stream);
}

extern "C" int b200qc_dfj(const double *j3c, int64_t nao, int64_t naux, int64_t ld, const double *inv_j2c,
                          const double *dm, double *vj, double *work, void *stream) {
    cudaStream_t st = as_stream(stream);
    const int64_t npair = nao * (nao + 1) / 2;
    double *t = work + dfj_dvec_len(npair), *c = t + ld;
    int rc = b200qc_dfj_pass1(j3c, nao, naux, ld, dm, t, work, stream);
    if (rc) return rc;
    QC_CHECK(cudaMemsetAsync(c, 0, sizeof(double) * ld, st));
    dfj_vecmat_kernel<<<(unsigned)((naux + 31) / 32), dim3(32, 8), 0, st>>>(t, inv_j2c, naux, c);
    QC_LAUNCHED(1);
    return b200qc_dfj_pass2(j3c, nao, naux, ld, c, vj, stream);
}

extern "C" int b200qc_pack_tril(const double *full, int64_t nao, int64_t naux, int64_t ld, double *packed,
                                void *stream) {
    QC_REQUIRE(ld >= naux, "ld must be >= naux");
    const int64_t npair = nao * (nao + 1) / 2;
    pack_tril_kernel<<<(unsigned)npair, 128, 0, as_stream(stream)>>>(full, (int)nao, naux, ld, npair, packed);
    QC_LAUNCHED(1);
    return 0;
}

// y[r] = sum_c A[r * ld + c] x[c]  -- the HBM-bound contraction of the stored-ERI regime
// (J = eri_j . vec(D), K = eri_k . vec(D), hcgto.py:209,234); warp per row, two loads in flight.
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const double *__restrict__ A, int64_t nrow, int64_t ncol, int64_t ld, const double *__restrict__ x,
                 double *__restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t nhalf = ncol / 2;
    for (int64_t r = warp; r < nrow; r += nwarp) {
        const double2 *row = reinterpret_cast<const double2 *>(A + r * ld);
        const double2 *xv = reinterpret_cast<const double2 *>(x);
        double s0 = 0.0, s1 = 0.0;
        int64_t k = lane;
        for (; k + 32 < nhalf; k += 64) {
            const double2 v0 = __ldcs(row + k), v1 = __ldcs(row + k + 32);
            const double2 c0 = __ldg(xv + k), c1 = __ldg(xv + k + 32);
            s0 += v0.x * c0.x + v0.y * c0.y;
            s1 += v1.x * c1.x + v1.y * c1.y;
        }
        for (; k < nhalf; k += 32) {
            const double2 v0 = __ldcs(row + k);
            const double2 c0 = __ldg(xv + k);
            s0 += v0.x * c0.x + v0.y * c0.y;
        }
        double s = s0 + s1;
        if ((ncol & 1) && lane == 0) s += A[r * ld + ncol - 1] * x[ncol - 1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[r] = s;
    }
}

extern "C" int b200qc_gemv(const double *A, int64_t nrow, int64_t ncol, int64_t ld, const double *x, double *y,
                           void *stream) {
    QC_REQUIRE(ld % 2 == 0 && ld >= ncol, "ld must be even and >= ncol");
    QC_REQUIRE(((uintptr_t)A | (uintptr_t)x) % 16 == 0, "A and x must be 16-byte aligned");
    if (nrow == 0) return 0;
    prof_begin(PROF_GEMV, as_stream(stream));
    gemv_rows_kernel<<<NUM_SMS * 8, 256, 0, as_stream(stream)>>>(A, nrow, ncol, ld, x, y);
    prof_end(as_stream(stream));
    QC_LAUNCHED(1);
    return 0;
}
