// K1 -- AO values (gradients, Laplacians) on grid points.
// Replaces GTOval_sph / GTOval_ip_sph / the GTOval_sph_deriv2 sum of eval_laplgto as called from
// dqc/hamilton/intor/gtoeval.py:196-260 (to_transpose=True layout: [comp][grid][ao], ao contiguous).
// DERIV = 0: phi; 1: phi, d/dx, d/dy, d/dz; 2: those four and the Laplacian (the meta-GGA branch, hcgto.py:183-186).
//
// HBM-write-bound: ncomp * ngrid * nao * 8 bytes written, 24 bytes/point read.  Layout of the work:
// one CTA owns 32 grid points (lane = point, so shell data are warp-uniform broadcasts and there is
// no divergence on l) and walks the AO axis in 64-column windows; the 8 warps split the shells of a
// window, park their values in a padded shared tile [comp][col][33] (conflict-free both ways) and
// the whole CTA then streams the tile out as 256-byte contiguous row segments.
#pragma once
#include "common.cuh"

#define AO_PTS 32
#define AO_WIN 64
#define AO_THREADS 256
#define AO_NCOMP(DERIV) ((DERIV) == 0 ? 1 : ((DERIV) == 1 ? 4 : 5))

// non-zero pattern of the cart -> real-spherical matrices of s, p and d shells (tables.cuh: p is the identity up to a
// constant; d rows xy, yz, z^2, xz, x^2 - y^2 over xx xy xz yy yz zz): the structural zeros are not multiplied
template <int L> __host__ __device__ constexpr bool ao_c2s_nz(int m, int c) {
    return L == 0 ? true
         : L == 1 ? m == c
         : L == 2 ? (m == 0 ? c == 1 : m == 1 ? c == 4 : m == 2 ? (c == 0 || c == 3 || c == 5) : m == 3 ? c == 2 : (c == 0 || c == 3))
                  : true;
}

// Point-major tile [comp][point][64 columns], no padding, XOR-swizzled so that BOTH phases are conflict-free: the
// evaluation writes one column for 32 points (lane = point; unswizzled, the 512-byte rows would put every lane on the
// same bank), the write-out reads 16-byte column pairs along a row.  Pair c (columns 2c, 2c + 1) of row p sits at pair
// c ^ (p & 7), its two halves swapped when bit 3 of p is set: 16 consecutive points -> 16 different 8-byte slots.
#define AO_TS 64
__device__ __forceinline__ int ao_pm_index(int comp, int p, int col) {
    return (comp * AO_PTS + p) * AO_TS + ((((col >> 1) ^ (p & 7)) << 1) | ((col & 1) ^ ((p >> 3) & 1)));
}

// PM = false: tile[comp][col][33] (column-major, conflict-free both ways); PM = true: the swizzled point-major tile
template <int L, int DERIV, bool PM = false>
__device__ __forceinline__ void ao_shell_to_tile(const ShellRec &sh, const double *__restrict__ env,
                                                 double x, double y, double z, int col0, int lane,
                                                 double *tile) {
    constexpr int NC = NCART(L), NS = 2 * L + 1;
    const double r2 = x * x + y * y + z * z;
    // rad = sum c e^(-a r^2);  grad rad = r drad, drad = sum -2 a c e^(-a r^2);  lapl rad = d2rad = sum (4 a^2 r^2 - 6 a) c e^(-a r^2)
    double rad = 0.0, drad = 0.0, d2rad = 0.0;
    for (int p = 0; p < sh.nprim; p++) {
        const double a = env[sh.ptr_exp + p];
        // a primitive below e^-50 = 2e-22 on all 32 points of the warp is skipped (tight core functions away from their
        // nucleus: most of the exps of a contracted s shell); the branch is warp-uniform
        if (!__any_sync(0xffffffffu, a * r2 < 50.0)) continue;
        const double e = env[sh.ptr_coef + p] * exp(-a * r2);
        rad += e;
        if (DERIV) drad -= 2.0 * a * e;
        if (DERIV == 2) d2rad += (4.0 * a * a * r2 - 6.0 * a) * e;
    }
    double px[L + 2], py[L + 2], pz[L + 2];
    px[0] = py[0] = pz[0] = 1.0;
#pragma unroll
    for (int k = 1; k <= L + 1; k++) {
        px[k] = px[k - 1] * x;
        py[k] = py[k - 1] * y;
        pz[k] = pz[k - 1] * z;
    }
    double cv[AO_NCOMP(DERIV)][NC];
    {
        int c = 0;
#pragma unroll
        for (int a = L; a >= 0; a--)
#pragma unroll
            for (int b = L - a; b >= 0; b--) {
                const int g = L - a - b;
                const double mono = px[a] * py[b] * pz[g];
                cv[0][c] = mono * rad;
                if (DERIV) {
                    const double dmx = (a ? a * px[a > 0 ? a - 1 : 0] : 0.0) * py[b] * pz[g];
                    const double dmy = px[a] * (b ? b * py[b > 0 ? b - 1 : 0] : 0.0) * pz[g];
                    const double dmz = px[a] * py[b] * (g ? g * pz[g > 0 ? g - 1 : 0] : 0.0);
                    cv[1][c] = dmx * rad + mono * x * drad;
                    cv[2][c] = dmy * rad + mono * y * drad;
                    cv[3][c] = dmz * rad + mono * z * drad;
                }
                if (DERIV == 2) {
                    // lapl (m R) = (lapl m) R + 2 grad m . grad R + m lapl R,  grad m . r = L m
                    const double lm = (a > 1 ? a * (a - 1) * px[a > 1 ? a - 2 : 0] : 0.0) * py[b] * pz[g] +
                                      px[a] * (b > 1 ? b * (b - 1) * py[b > 1 ? b - 2 : 0] : 0.0) * pz[g] +
                                      px[a] * py[b] * (g > 1 ? g * (g - 1) * pz[g > 1 ? g - 2 : 0] : 0.0);
                    cv[4][c] = lm * rad + mono * (2.0 * L * drad + d2rad);
                }
                c++;
            }
    }
    const double *M = c2s_ptr(L);
#pragma unroll
    for (int m = 0; m < NS; m++) {
        const int col = col0 + m;
        if (col < 0 || col >= AO_WIN) continue;
#pragma unroll
        for (int comp = 0; comp < AO_NCOMP(DERIV); comp++) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < NC; c++)
                if (ao_c2s_nz<L>(m, c)) s += M[m * NC + c] * cv[comp][c];
            tile[PM ? ao_pm_index(comp, lane, col) : (comp * AO_WIN + col) * 33 + lane] = s;
        }
    }
}

template <int DERIV>
__global__ void __launch_bounds__(AO_THREADS)
ao_eval_kernel(const ShellRec *__restrict__ shells, const double *__restrict__ env,
               const int *__restrict__ ao_loc, int sh0, int sh1, const double *__restrict__ coords,
               int64_t ngrid, double *__restrict__ ao, int64_t ngrid_ld, int64_t ao_ld) {
    extern __shared__ double tile[];
    constexpr int NCOMP = AO_NCOMP(DERIV);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * AO_PTS + lane;
    const bool live = g < ngrid;
    double gx = 0, gy = 0, gz = 0;
    if (live) {
        gx = coords[3 * g];
        gy = coords[3 * g + 1];
        gz = coords[3 * g + 2];
    }
    const int ao_base = ao_loc[sh0];
    const int nao = ao_loc[sh1] - ao_base;
    int s_lo = sh0;  // first shell that may touch the current window
    for (int c0 = 0; c0 < nao; c0 += AO_WIN) {
        // advance s_lo to the first shell whose range ends after c0; find s_hi = first shell starting >= c0+WIN
        while (s_lo < sh1 && ao_loc[s_lo + 1] - ao_base <= c0) s_lo++;
        int s_hi = s_lo;
        while (s_hi < sh1 && ao_loc[s_hi] - ao_base < c0 + AO_WIN) s_hi++;
        for (int s = s_lo + warp; s < s_hi; s += AO_THREADS / 32) {
            const ShellRec sh = shells[s];
            const double x = gx - sh.x, y = gy - sh.y, z = gz - sh.z;
            const int col0 = sh.ao_off - ao_base - c0;
            switch (sh.l) {
                case 0: ao_shell_to_tile<0, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                case 1: ao_shell_to_tile<1, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                case 2: ao_shell_to_tile<2, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                case 3: ao_shell_to_tile<3, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                default: ao_shell_to_tile<4, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
            }
        }
        __syncthreads();
        // stream the tile out: warp w writes points 4w..4w+3, lanes run along the AO axis
        const int ncol = min(AO_WIN, nao - c0);
#pragma unroll
        for (int comp = 0; comp < NCOMP; comp++)
#pragma unroll
            for (int pp = 0; pp < 4; pp++) {
                const int p = warp * 4 + pp;
                const int64_t gp = (int64_t)blockIdx.x * AO_PTS + p;
                if (gp >= ngrid) continue;
                double *row = ao + ((int64_t)comp * ngrid_ld + gp) * ao_ld + c0;
#pragma unroll
                for (int cc = 0; cc < AO_WIN; cc += 32) {
                    const int col = cc + lane;
                    if (col < ncol) row[col] = tile[(comp * AO_WIN + col) * 33 + p];
                }
            }
        __syncthreads();
    }
}

extern "C" int b200qc_eval_gto(const b200qc_basis *basis, int sh0, int sh1, int deriv,
                               const double *coords, int64_t ngrid, double *ao, int64_t ngrid_ld,
                               int64_t ao_ld, void *stream) {
    if (qc_require_basis_device(basis)) return 2;
    QC_REQUIRE(basis != nullptr, "null basis");
    QC_REQUIRE(0 <= sh0 && sh0 < sh1 && sh1 <= basis->nbas, "bad shell range");
    QC_REQUIRE(deriv >= 0 && deriv <= 2, "deriv must be 0, 1 or 2");
    QC_REQUIRE(ngrid_ld >= ngrid && ao_ld >= basis->h_ao_loc[sh1] - basis->h_ao_loc[sh0], "leading dims too small");
    if (ngrid == 0) return 0;
    const int nblk = (int)((ngrid + AO_PTS - 1) / AO_PTS);
    const size_t smem = sizeof(double) * AO_NCOMP(deriv) * AO_WIN * 33;
    prof_begin(PROF_AO_EVAL, as_stream(stream));
    if (deriv == 2) {
        QC_CHECK(cudaFuncSetAttribute(ao_eval_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ao_eval_kernel<2><<<nblk, AO_THREADS, smem, as_stream(stream)>>>(
            basis->d_shells, basis->d_env, basis->d_ao_loc, sh0, sh1, coords, ngrid, ao, ngrid_ld, ao_ld);
    } else if (deriv == 1) {
        QC_CHECK(cudaFuncSetAttribute(ao_eval_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ao_eval_kernel<1><<<nblk, AO_THREADS, smem, as_stream(stream)>>>(
            basis->d_shells, basis->d_env, basis->d_ao_loc, sh0, sh1, coords, ngrid, ao, ngrid_ld, ao_ld);
    } else {
        ao_eval_kernel<0><<<nblk, AO_THREADS, smem, as_stream(stream)>>>(
            basis->d_shells, basis->d_env, basis->d_ao_loc, sh0, sh1, coords, ngrid, ao, ngrid_ld, ao_ld);
    }
    prof_end(as_stream(stream));
    QC_LAUNCHED(1);
    return 0;
}
