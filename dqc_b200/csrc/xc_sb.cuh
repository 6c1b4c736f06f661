// Block-sparse XC grid path ("superblocks").
//
// The reference evaluates every AO on every grid point (non0tab = 1, gtoeval.py:211) and runs the
// density / Vxc contractions dense (hcgto.py:399-418, 461-481).  On a molecular grid most AOs are
// numerically zero on most points: on the C60/def2-SVP sg3 grid 54 % of the AOs exceed 1e-12 on an
// average block, and the contractions scale with the SQUARE of that fraction.  Here the grid is cut
// into superblocks (SB) of SBP consecutive points (the grid is atom-major, radial-major, so an SB is
// spatially compact); each SB keeps only the shells whose envelope can exceed `eps` somewhere on it:
//   * AO values are stored compacted per SB:  ao[sb] = [ncomp][SBP][nsp]  (nsp = kept AOs padded to 64)
//   * rho:   D is gathered to D_sb[nsp][nsp] once per build, X = phi_sb D_sb on the fp64 tensor pipe
//            and contracted against phi / grad phi in registers (same tile engine as rho.cuh)
//   * Vxc:   M_sb = phi_sb^T vb_sb per SB (K = SBP), tiles added into M[idx][idx] with fp64 atomics
// eps bounds the dropped |phi| (and |grad phi|): the results differ from the dense ones by O(eps),
// i.e. ~1e-11 at the default 1e-12 -- five orders below the 1e-6 parity bar.  eps = 0 keeps everything.
#pragma once
#include <cstdlib>
#include "gemm_f64.cuh"
#include "sb_common.cuh"


// ---- screening: flags[sb][shell] = 1 if the shell's envelope exceeds eps on some point of the SB ----
__global__ void __launch_bounds__(256)
ao_screen_kernel(const ShellRec *__restrict__ shells, const double *__restrict__ env, int sh0, int nshell,
                 const double *__restrict__ coords, int64_t ngrid, int sbp, double eps, int deriv,
                 unsigned char *__restrict__ flags) {
    const int sb = blockIdx.x;
    const int64_t g0 = (int64_t)sb * sbp;
    __shared__ int hit;
    for (int s = 0; s < nshell; s++) {
        if (threadIdx.x == 0) hit = 0;
        __syncthreads();
        const ShellRec sh = shells[sh0 + s];
        const double ang = sqrt((2 * sh.l + 1) / (4.0 * 3.14159265358979323846));
        bool mine = false;
        for (int p = threadIdx.x; p < sbp && !mine; p += blockDim.x) {
            const int64_t g = g0 + p;
            if (g >= ngrid) break;
            const double x = coords[3 * g] - sh.x, y = coords[3 * g + 1] - sh.y, z = coords[3 * g + 2] - sh.z;
            const double r2 = x * x + y * y + z * z, r = sqrt(r2);
            double rl = 1.0;
            for (int k = 0; k < sh.l; k++) rl *= r;
            double v = 0.0, dv = 0.0, lv = 0.0;
            for (int q = 0; q < sh.nprim; q++) {
                const double a = env[sh.ptr_exp + q];
                const double e = fabs(env[sh.ptr_coef + q]) * exp(-a * r2);
                v += e;
                dv += e * (2.0 * a * r * rl + (sh.l > 0 ? sh.l * rl / fmax(r, 1e-300) : 0.0));
                lv += e * (4.0 * a * a * r2 + 2.0 * a * (2 * sh.l + 3));   // |lapl (r^l Y_lm e^(-a r^2))| / (r^l |Y_lm|)
            }
            v *= rl * ang;
            dv *= ang * 2.0 * (sh.l + 1);   // generous bound on |grad (r^l Y_lm)| / r^(l-1)
            lv *= rl * ang;
            if (v > eps || (deriv && dv > eps) || (deriv == 2 && lv > eps)) mine = true;
        }
        if (mine) hit = 1;
        __syncthreads();
        if (threadIdx.x == 0) flags[(int64_t)sb * nshell + s] = (unsigned char)(eps <= 0.0 ? 1 : hit);
        __syncthreads();
    }
}

extern "C" int b200qc_ao_screen(const b200qc_basis *basis, int sh0, int sh1, const double *coords, int64_t ngrid,
                                int sbp, double eps, int deriv, unsigned char *flags, void *stream) {
    if (qc_require_basis_device(basis)) return 2;
    QC_REQUIRE(basis && 0 <= sh0 && sh0 < sh1 && sh1 <= basis->nbas, "bad shell range");
    QC_REQUIRE(sbp > 0 && sbp % GM_BM == 0, "superblock size must be a multiple of 128");
    if (ngrid == 0) return 0;
    const int nsb = (int)((ngrid + sbp - 1) / sbp);
    ao_screen_kernel<<<nsb, 256, 0, as_stream(stream)>>>(basis->d_shells, basis->d_env, sh0, sh1 - sh0, coords, ngrid,
                                                         sbp, eps, deriv, flags);
    QC_LAUNCHED(1);
    return 0;
}

// ---- compact AO evaluation: same per-shell arithmetic as ao_eval.cuh, columns = kept shells of the SB ----
// LMAX = highest angular momentum of the basis (2: s, p, d only -- the f / g code paths of LMAX = 4 cost 228 registers per
// thread, one CTA per SM; without them two to three CTAs fit)
template <int DERIV, int LMAX>
__global__ void __launch_bounds__(AO_THREADS, LMAX <= 2 ? 2 : 1)
ao_eval_sb_kernel(const ShellRec *__restrict__ shells, const double *__restrict__ env, const SBDesc *__restrict__ sbd,
                  const int *__restrict__ shell_ids, const int *__restrict__ shell_col,
                  const double *__restrict__ coords, int64_t ngrid, int sbp, double *__restrict__ ao) {
    extern __shared__ double tile[];
    constexpr int NCOMP = AO_NCOMP(DERIV);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunks = sbp / AO_PTS;
    const int sb = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const SBDesc d = sbd[sb];
    const int row0 = chunk * AO_PTS;
    const int64_t g = (int64_t)sb * sbp + row0 + lane;
    const bool live = g < ngrid;
    double gx = 0, gy = 0, gz = 0;
    if (live) {
        gx = coords[3 * g]; gy = coords[3 * g + 1]; gz = coords[3 * g + 2];
    }
    const int *ids = shell_ids + d.shell_off;
    const int *cols = shell_col + d.shell_off;   // compact first column of each kept shell
    int s_lo = 0;
    for (int c0 = 0; c0 < d.nsp; c0 += AO_WIN) {
        while (s_lo < d.nshell && cols[s_lo] + 2 * shells[ids[s_lo]].l + 1 <= c0) s_lo++;
        int s_hi = s_lo;
        while (s_hi < d.nshell && cols[s_hi] < c0 + AO_WIN) s_hi++;
        if (s_hi == s_lo) continue;   // nothing but padding in this window (buffer is pre-zeroed)
        for (int s = s_lo + warp; s < s_hi; s += AO_THREADS / 32) {
            const ShellRec sh = shells[ids[s]];
            const double x = gx - sh.x, y = gy - sh.y, z = gz - sh.z;
            const int col0 = cols[s] - c0;
            if (LMAX <= 2) {
                switch (sh.l) {
                    case 0: ao_shell_to_tile<0, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                    case 1: ao_shell_to_tile<1, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                    default: ao_shell_to_tile<2, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                }
            } else {
                switch (sh.l) {
                    case 0: ao_shell_to_tile<0, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                    case 1: ao_shell_to_tile<1, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                    case 2: ao_shell_to_tile<2, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                    case 3: ao_shell_to_tile<3, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                    default: ao_shell_to_tile<4, DERIV>(sh, env, x, y, z, col0, lane, tile); break;
                }
            }
        }
        __syncthreads();
        // columns of this window that belong to kept shells: [cols[s_lo], end of shell s_hi-1)
        const int cbeg = max(c0, cols[s_lo]);
        const int cend = min(c0 + AO_WIN, cols[s_hi - 1] + 2 * shells[ids[s_hi - 1]].l + 1);
#pragma unroll
        for (int comp = 0; comp < NCOMP; comp++)
#pragma unroll
            for (int pp = 0; pp < 4; pp++) {
                const int p = warp * 4 + pp;
                const int64_t gp = (int64_t)sb * sbp + row0 + p;
                if (gp >= ngrid) continue;
                double *row = ao + d.ao_off + ((int64_t)comp * sbp + row0 + p) * d.nsp;
#pragma unroll
                for (int cc = 0; cc < AO_WIN; cc += 32) {
                    const int col = c0 + cc + lane;
                    if (col >= cbeg && col < cend) row[col] = tile[(comp * AO_WIN + cc + lane) * 33 + p];
                }
            }
        __syncthreads();
    }
}

// ---- K1, staged form (default) ----
// A CTA owns AO_CH consecutive 32-point chunks of one superblock.  It first stages the superblock's kept-shell list in
// shared memory -- centre, l, first compact column and the primitive exponents / coefficients (offsets by a CTA-wide
// prefix sum) -- so that the evaluation loop has no dependent global loads (the first form walked ids -> shell record
// -> env for every (warp, shell): 12 % warp occupancy at 228 registers, long-scoreboard bound, 0.14 of the HBM write
// rate).  Per 64-column window the eight warps split the shells (lane = point), park the values in a POINT-major tile
// (XOR-swizzled: conflict-free for the column writes and the row reads) and the CTA streams it out with 128-bit loads /
// stores: one 512-byte row segment per warp instruction.
#define AO_CH 4
struct AOStage {
    double x, y, z;
    int l, nprim, col, poff;
};

template <int DERIV, int LMAX>
__global__ void __launch_bounds__(AO_THREADS, LMAX <= 2 ? 2 : 1)
ao_eval_sb2_kernel(const ShellRec *__restrict__ shells, const double *__restrict__ env, const SBDesc *__restrict__ sbd,
                   const int *__restrict__ shell_ids, const int *__restrict__ shell_col,
                   const double *__restrict__ coords, int64_t ngrid, int sbp, double *__restrict__ ao, int max_shell,
                   int max_prim) {
    extern __shared__ __align__(16) double tile[];
    constexpr int NCOMP = AO_NCOMP(DERIV);
    __shared__ int wsum[AO_THREADS / 32];
    AOStage *stg = reinterpret_cast<AOStage *>(tile + NCOMP * AO_PTS * AO_TS);
    double *pdat = reinterpret_cast<double *>(stg + max_shell);   // [exponents: max_prim][coefficients: max_prim]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int groups = sbp / (AO_PTS * AO_CH);
    const int sb = blockIdx.x / groups, grp = blockIdx.x % groups;
    if ((int64_t)sb * sbp + (int64_t)grp * AO_CH * AO_PTS >= ngrid) return;
    const SBDesc d = sbd[sb];
    const int *ids = shell_ids + d.shell_off;
    const int *cols = shell_col + d.shell_off;   // compact first column of each kept shell
    // ---- stage the kept shells ----
    int run = 0;
    for (int base = 0; base < d.nshell; base += AO_THREADS) {
        const int s = base + threadIdx.x;
        ShellRec sh = {};
        int np = 0;
        if (s < d.nshell) {
            sh = shells[ids[s]];
            np = sh.nprim;
        }
        int inc = np;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w2 = 0; w2 < AO_THREADS / 32; w2++) {
            if (w2 < warp) woff += wsum[w2];
            tot += wsum[w2];
        }
        if (s < d.nshell) {
            AOStage a;
            a.x = sh.x; a.y = sh.y; a.z = sh.z;
            a.l = sh.l; a.nprim = np; a.col = cols[s];
            a.poff = run + woff + inc - np;
            stg[s] = a;
            for (int p = 0; p < np; p++) {
                pdat[a.poff + p] = env[sh.ptr_exp + p];
                pdat[max_prim + a.poff + p] = env[sh.ptr_coef + p];
            }
        }
        run += tot;
        __syncthreads();
    }
    for (int ch = 0; ch < AO_CH; ch++) {
        const int row0 = (grp * AO_CH + ch) * AO_PTS;
        if ((int64_t)sb * sbp + row0 >= ngrid) break;
        const int64_t g = (int64_t)sb * sbp + row0 + lane;
        double gx = 0, gy = 0, gz = 0;
        if (g < ngrid) {
            gx = coords[3 * g]; gy = coords[3 * g + 1]; gz = coords[3 * g + 2];
        }
        int s_lo = 0;
        for (int c0 = 0; c0 < d.nsp; c0 += AO_WIN) {
            while (s_lo < d.nshell && stg[s_lo].col + 2 * stg[s_lo].l + 1 <= c0) s_lo++;
            int s_hi = s_lo;
            while (s_hi < d.nshell && stg[s_hi].col < c0 + AO_WIN) s_hi++;
            if (s_hi == s_lo) continue;   // nothing but padding in this window (buffer is pre-zeroed)
            // kept shells fill the columns contiguously: only the last window has a tail of padding columns
            const int cend = min(AO_WIN, stg[s_hi - 1].col + 2 * stg[s_hi - 1].l + 1 - c0);
            if (cend < AO_WIN) {
                const int nz = AO_WIN - cend;
                for (int e = threadIdx.x; e < NCOMP * AO_PTS * nz; e += AO_THREADS) {
                    const int r = e / nz;
                    tile[ao_pm_index(r / AO_PTS, r % AO_PTS, cend + e % nz)] = 0.0;
                }
            }
            for (int s = s_lo + warp; s < s_hi; s += AO_THREADS / 32) {
                const AOStage a = stg[s];
                ShellRec sh;
                sh.l = a.l; sh.nprim = a.nprim;
                sh.ptr_exp = a.poff; sh.ptr_coef = max_prim + a.poff;
                const double x = gx - a.x, y = gy - a.y, z = gz - a.z;
                const int col0 = a.col - c0;
                if (LMAX <= 2) {
                    switch (a.l) {
                        case 0: ao_shell_to_tile<0, DERIV, true>(sh, pdat, x, y, z, col0, lane, tile); break;
                        case 1: ao_shell_to_tile<1, DERIV, true>(sh, pdat, x, y, z, col0, lane, tile); break;
                        default: ao_shell_to_tile<2, DERIV, true>(sh, pdat, x, y, z, col0, lane, tile); break;
                    }
                } else {
                    switch (a.l) {
                        case 0: ao_shell_to_tile<0, DERIV, true>(sh, pdat, x, y, z, col0, lane, tile); break;
                        case 1: ao_shell_to_tile<1, DERIV, true>(sh, pdat, x, y, z, col0, lane, tile); break;
                        case 2: ao_shell_to_tile<2, DERIV, true>(sh, pdat, x, y, z, col0, lane, tile); break;
                        case 3: ao_shell_to_tile<3, DERIV, true>(sh, pdat, x, y, z, col0, lane, tile); break;
                        default: ao_shell_to_tile<4, DERIV, true>(sh, pdat, x, y, z, col0, lane, tile); break;
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int comp = 0; comp < NCOMP; comp++)
#pragma unroll
                for (int pp = 0; pp < 4; pp++) {
                    const int p = warp * 4 + pp;
                    if ((int64_t)sb * sbp + row0 + p >= ngrid) continue;
                    double2 v = *reinterpret_cast<const double2 *>(tile + (comp * AO_PTS + p) * AO_TS +
                                                                   2 * (lane ^ (p & 7)));
                    if (p & 8) {
                        const double t = v.x;
                        v.x = v.y;
                        v.y = t;
                    }
                    double *row = ao + d.ao_off + ((int64_t)comp * sbp + row0 + p) * d.nsp + c0;
                    *reinterpret_cast<double2 *>(row + 2 * lane) = v;
                }
            __syncthreads();
        }
    }
}

extern "C" int b200qc_eval_gto_sb(const b200qc_basis *basis, int deriv, const double *coords, int64_t ngrid, int sbp,
                                  int nsb, const void *sbdesc, const int *shell_ids, const int *shell_col, double *ao,
                                  void *stream) {
    if (qc_require_basis_device(basis)) return 2;
    QC_REQUIRE(basis && deriv >= 0 && deriv <= 2, "bad arguments");
    QC_REQUIRE(sbp % AO_PTS == 0, "superblock size must be a multiple of 32");
    if (ngrid == 0 || nsb == 0) return 0;
    const size_t smem = sizeof(double) * AO_NCOMP(deriv) * AO_WIN * 33;
    const unsigned nblk = (unsigned)nsb * (unsigned)(sbp / AO_PTS);
    cudaStream_t st = as_stream(stream);
    int lmax = 0;
    for (const ShellRec &sh : basis->h_shells) lmax = std::max(lmax, sh.l);
    // staged form: shared-memory room for the longest kept-shell list and its primitives (one-off host look at the lists)
    int max_shell = 0, max_prim = 0;
    {
        std::vector<SBDesc> hd(nsb);
        QC_CHECK(cudaMemcpyAsync(hd.data(), sbdesc, sizeof(SBDesc) * nsb, cudaMemcpyDeviceToHost, st));
        QC_CHECK(cudaStreamSynchronize(st));
        int64_t nids = 0;
        for (const SBDesc &d : hd) nids = std::max<int64_t>(nids, (int64_t)d.shell_off + d.nshell);
        std::vector<int> hids(nids);
        if (nids) QC_CHECK(cudaMemcpyAsync(hids.data(), shell_ids, sizeof(int) * nids, cudaMemcpyDeviceToHost, st));
        QC_CHECK(cudaStreamSynchronize(st));
        for (const SBDesc &d : hd) {
            int np = 0;
            for (int s = 0; s < d.nshell; s++) {
                const int id = hids[d.shell_off + s];
                QC_REQUIRE(id >= 0 && id < basis->nbas, "kept-shell list holds a shell outside the basis");
                np += basis->h_shells[id].nprim;
            }
            max_shell = std::max(max_shell, d.nshell);
            max_prim = std::max(max_prim, np);
        }
        max_prim += max_prim & 1;
    }
    const size_t smem2 = sizeof(double) * AO_NCOMP(deriv) * AO_PTS * AO_TS + sizeof(AOStage) * max_shell +
                         sizeof(double) * 2 * max_prim;
    const bool staged = smem2 <= 200 * 1024 && sbp % (AO_PTS * AO_CH) == 0 && !getenv("B200QC_K1_UNSTAGED");
    const unsigned nblk2 = (unsigned)nsb * (unsigned)(sbp / (AO_PTS * AO_CH));
    prof_begin(PROF_AO_EVAL, st);
#define AO_SB_LAUNCH(DERIV, LMAX)                                                                                        \
    do {                                                                                                                 \
        if (staged) {                                                                                                    \
            QC_CHECK(cudaFuncSetAttribute(ao_eval_sb2_kernel<DERIV, LMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          (int)smem2));                                                                  \
            ao_eval_sb2_kernel<DERIV, LMAX><<<nblk2, AO_THREADS, smem2, st>>>(                                           \
                basis->d_shells, basis->d_env, (const SBDesc *)sbdesc, shell_ids, shell_col, coords, ngrid, sbp, ao,     \
                max_shell, max_prim);                                                                                    \
        } else {                                                                                                         \
            QC_CHECK(cudaFuncSetAttribute(ao_eval_sb_kernel<DERIV, LMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                          (int)smem));                                                                   \
            ao_eval_sb_kernel<DERIV, LMAX><<<nblk, AO_THREADS, smem, st>>>(basis->d_shells, basis->d_env,                \
                                                                          (const SBDesc *)sbdesc, shell_ids, shell_col,  \
                                                                          coords, ngrid, sbp, ao);                       \
        }                                                                                                                \
    } while (0)
    if (lmax <= 2) {
        if (deriv == 2) AO_SB_LAUNCH(2, 2); else if (deriv == 1) AO_SB_LAUNCH(1, 2); else AO_SB_LAUNCH(0, 2);
    } else {
        if (deriv == 2) AO_SB_LAUNCH(2, 4); else if (deriv == 1) AO_SB_LAUNCH(1, 4); else AO_SB_LAUNCH(0, 4);
    }
#undef AO_SB_LAUNCH
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

// ---- rho ----
// D_sb[i][j] = D[idx_i][idx_j] (0 for padding); D is (nao, nao) row-major
__global__ void sb_gather_dm_kernel(const SBDesc *__restrict__ sbd, const int *__restrict__ idx,
                                    const double *__restrict__ dm, int nao, double *__restrict__ dsb) {
    const SBDesc d = sbd[blockIdx.y];
    const int *ix = idx + d.idx_off;
    const int64_t n2 = (int64_t)d.nsp * d.nsp;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d.nsp), j = (int)(e - (int64_t)i * d.nsp);
        const int a = ix[i], b = ix[j];
        dsb[d.d_off + e] = (a < nao && b < nao) ? dm[(int64_t)a * nao + b] : 0.0;
    }
}

template <int NCOMP>
__global__ void __launch_bounds__(GM_THREADS, 2)
rho_sb_kernel(const SBDesc *__restrict__ sbd, const double *__restrict__ ao, const double *__restrict__ dsb, int sbp,
              int64_t ngrid_ld, double *__restrict__ rho, double *__restrict__ grad, double *__restrict__ lapl) {
    // NCOMP = 5 (meta-GGA storage): component 4 is the AO Laplacian, lapl[g] = sum_nu X_g,nu lapl phi_g,nu
    extern __shared__ __align__(16) double gm_smem[];
    __shared__ double red[2][GM_BM][NCOMP];
    const int tiles = sbp / GM_BM;
    const int sb = blockIdx.x / tiles, tile = blockIdx.x % tiles;
    const SBDesc d = sbd[sb];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int64_t ld = d.nsp;
    const double *base = ao + d.ao_off + (int64_t)tile * GM_BM * ld;   // component 0, this tile's rows
    const double *D = dsb + d.d_off;
    double part[NCOMP][4];
#pragma unroll
    for (int c = 0; c < NCOMP; c++)
#pragma unroll
        for (int i = 0; i < 4; i++) part[c][i] = 0.0;
    const int ntile = d.nsp / GM_BN, ktiles = d.nsp / GM_BK;
    for (int nt = 0; nt < ntile; nt++) {
        const int n0 = nt * GM_BN;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        gemm_tile_128x64<true>(base, ld, GM_BM, D + n0, ld, ktiles, acc, gm_smem);
#pragma unroll
        for (int c = 0; c < NCOMP; c++) {
            const double *comp = base + (int64_t)c * sbp * ld + n0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double *row = comp + (int64_t)(wm + i * 8 + (lane >> 2)) * ld + wn + 2 * (lane & 3);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double2 v = *reinterpret_cast<const double2 *>(row + j * 8);
                    part[c][i] += acc[i][j][0] * v.x + acc[i][j][1] * v.y;
                }
            }
        }
    }
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
        const int64_t g = (int64_t)sb * sbp + (int64_t)tile * GM_BM + r;
        rho[g] = red[0][r][0] + red[1][r][0];
        if (NCOMP >= 4) {
#pragma unroll
            for (int dd = 0; dd < 3; dd++) grad[(int64_t)dd * ngrid_ld + g] = 2.0 * (red[0][r][dd + 1] + red[1][r][dd + 1]);
        }
        if (NCOMP == 5) lapl[g] = red[0][r][NCOMP - 1] + red[1][r][NCOMP - 1];
    }
}

// gg[g] = sum_d sum_mu,nu d_d phi_g,mu D_mu,nu d_d phi_g,nu (hcgto.py:427-429): three GEMMs X_d = (d_d phi) D_sb with the
// row dots against the same component, accumulated over d.  lapl rho = 2 (lapl part + gg), tau = gg / 2.
__global__ void __launch_bounds__(GM_THREADS, 2)
rho_sb_gg_kernel(const SBDesc *__restrict__ sbd, const double *__restrict__ ao, const double *__restrict__ dsb, int sbp,
                 double *__restrict__ gg) {
    extern __shared__ __align__(16) double gm_smem[];
    __shared__ double red[2][GM_BM];
    const int tiles = sbp / GM_BM;
    const int sb = blockIdx.x / tiles, tile = blockIdx.x % tiles;
    const SBDesc d = sbd[sb];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int64_t ld = d.nsp;
    const double *D = dsb + d.d_off;
    double part[4] = {0.0, 0.0, 0.0, 0.0};
    const int ntile = d.nsp / GM_BN, ktiles = d.nsp / GM_BK;
    for (int dd = 1; dd <= 3; dd++) {
        const double *base = ao + d.ao_off + ((int64_t)dd * sbp + (int64_t)tile * GM_BM) * ld;   // component dd, this tile's rows
        for (int nt = 0; nt < ntile; nt++) {
            const int n0 = nt * GM_BN;
            double acc[4][4][2];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
            gemm_tile_128x64<true>(base, ld, GM_BM, D + n0, ld, ktiles, acc, gm_smem);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double *row = base + n0 + (int64_t)(wm + i * 8 + (lane >> 2)) * ld + wn + 2 * (lane & 3);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double2 v = *reinterpret_cast<const double2 *>(row + j * 8);
                    part[i] += acc[i][j][0] * v.x + acc[i][j][1] * v.y;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double v = part[i];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if ((lane & 3) == 0) red[warp & 1][wm + i * 8 + (lane >> 2)] = v;
    }
    __syncthreads();
    if (threadIdx.x < GM_BM) {
        const int r = threadIdx.x;
        gg[(int64_t)sb * sbp + (int64_t)tile * GM_BM + r] = red[0][r] + red[1][r];
    }
}

// dm: (nao, nao) symmetric AO-basis density; rho (ngrid_ld = nsb * sbp), grad (3, ngrid_ld) or NULL;
// dsb: scratch of sum_sb nsp^2 doubles
extern "C" int b200qc_rho_sb(const void *sbdesc, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                             const double *dm, int nao, double *dsb, double *rho, double *grad, void *stream) {
    QC_REQUIRE(sbp % GM_BM == 0, "superblock size must be a multiple of 128");
    if (nsb == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const SBDesc *sbd = (const SBDesc *)sbdesc;
    const int64_t ngl = (int64_t)nsb * sbp;
    prof_begin(PROF_SB_GATHER, st);
    dim3 gg((unsigned)(((int64_t)max_nsp * max_nsp + 1023) / 1024 > 64 ? 64 : ((int64_t)max_nsp * max_nsp + 1023) / 1024),
            (unsigned)nsb);
    sb_gather_dm_kernel<<<gg, 256, 0, st>>>(sbd, idx, dm, nao, dsb);
    prof_end(st);
    QC_LAUNCHED(1);
    const unsigned nblk = (unsigned)nsb * (unsigned)(sbp / GM_BM);
    prof_begin(PROF_RHO, st);
    if (grad) {
        QC_CHECK(cudaFuncSetAttribute(rho_sb_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES));
        rho_sb_kernel<4><<<nblk, GM_THREADS, GM_SMEM_BYTES, st>>>(sbd, ao, dsb, sbp, ngl, rho, grad, nullptr);
    } else {
        QC_CHECK(cudaFuncSetAttribute(rho_sb_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES));
        rho_sb_kernel<1><<<nblk, GM_THREADS, GM_SMEM_BYTES, st>>>(sbd, ao, dsb, sbp, ngl, rho, grad, nullptr);
    }
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

// Meta-GGA densities (hcgto.py:399-438) on the 5-component superblock storage [phi, dx, dy, dz, lapl]:
// rho, grad (3, ngrid_ld), lapl = lapl rho = 2 (sum X lapl phi + gg), kin = tau = gg / 2.  fp64 DMMA engine.
__global__ void mgga_finish_kernel(int64_t n, double *__restrict__ lapl, double *__restrict__ kin) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double gg = kin[i];
    lapl[i] = 2.0 * (lapl[i] + gg);
    kin[i] = 0.5 * gg;
}

extern "C" int b200qc_rho_sb_mgga(const void *sbdesc, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                                  const double *dm, int nao, double *dsb, double *rho, double *grad, double *lapl,
                                  double *kin, void *stream) {
    QC_REQUIRE(sbp % GM_BM == 0, "superblock size must be a multiple of 128");
    QC_REQUIRE(rho && grad && lapl && kin, "all four outputs are needed");
    if (nsb == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const SBDesc *sbd = (const SBDesc *)sbdesc;
    const int64_t ngl = (int64_t)nsb * sbp;
    prof_begin(PROF_SB_GATHER, st);
    dim3 gg((unsigned)(((int64_t)max_nsp * max_nsp + 1023) / 1024 > 64 ? 64 : ((int64_t)max_nsp * max_nsp + 1023) / 1024),
            (unsigned)nsb);
    sb_gather_dm_kernel<<<gg, 256, 0, st>>>(sbd, idx, dm, nao, dsb);
    prof_end(st);
    QC_LAUNCHED(1);
    const unsigned nblk = (unsigned)nsb * (unsigned)(sbp / GM_BM);
    prof_begin(PROF_RHO, st);
    QC_CHECK(cudaFuncSetAttribute(rho_sb_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES));
    rho_sb_kernel<5><<<nblk, GM_THREADS, GM_SMEM_BYTES, st>>>(sbd, ao, dsb, sbp, ngl, rho, grad, lapl);
    QC_CHECK(cudaFuncSetAttribute(rho_sb_gg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES));
    rho_sb_gg_kernel<<<nblk, GM_THREADS, GM_SMEM_BYTES, st>>>(sbd, ao, dsb, sbp, kin);
    mgga_finish_kernel<<<(unsigned)((ngl + 255) / 256), 256, 0, st>>>(ngl, lapl, kin);
    prof_end(st);
    QC_LAUNCHED(3);
    return 0;
}

// ---- Vxc ---- (vxc_vb_sb_kernel lives in sb_common.cuh)
// M[idx_i][idx_j] += sum_{g in sb} phi[g][i] vb[g][j]; grid (max tiles, nsb)
__global__ void __launch_bounds__(GM_THREADS, 2)
vxc_sb_gemm_kernel(const SBDesc *__restrict__ sbd, const int *__restrict__ idx, const double *__restrict__ ao,
                   const double *__restrict__ vb, const int64_t *__restrict__ vb_off, int sbp, int nao,
                   double *__restrict__ mat, int acomp) {
    // acomp: component of the AO storage on the left (0 = phi; 1..3 = d_d phi for the meta-GGA tau term)
    extern __shared__ __align__(16) double gm_smem[];
    const SBDesc d = sbd[blockIdx.y];
    const int ntn = d.nsp / GM_BN, ntm = (d.nsp + GM_BM - 1) / GM_BM;
    if ((int)blockIdx.x >= ntn * ntm) return;
    const int tm = blockIdx.x / ntn, tn = blockIdx.x % ntn;
    const int m0 = tm * GM_BM, n0 = tn * GM_BN;
    const int64_t ld = d.nsp;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    gemm_tile_128x64<false>(ao + d.ao_off + (int64_t)acomp * sbp * ld + m0, ld, min(GM_BM, d.nsp - m0), vb + vb_off[blockIdx.y] + n0, ld,
                            sbp / GM_BK, acc, gm_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int *ix = idx + d.idx_off;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = m0 + wm + i * 8 + (lane >> 2);
        if (row >= d.nsp) continue;
        const int a = ix[row];
        if (a >= nao) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int col = n0 + wn + j * 8 + 2 * (lane & 3);
            const int b0 = ix[col], b1 = ix[col + 1];
            if (b0 < nao) atomicAdd(mat + (int64_t)a * nao + b0, acc[i][j][0]);
            if (b1 < nao) atomicAdd(mat + (int64_t)a * nao + b1, acc[i][j][1]);
        }
    }
}

// mat (nao, nao) is OVERWRITTEN with sum_g w phi^T (vrho phi + 2 vgrad . grad phi).
// vb: scratch of sum_sb sbp * nsp doubles; vb_off[sb]: its per-SB offsets (device, int64)
extern "C" int b200qc_vxc_sb(const void *sbdesc, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                             const double *weights, const double *vrho, const double *vgrad, int nao,
                             const int64_t *vb_off, double *vb, double *mat, void *stream) {
    QC_REQUIRE(sbp % GM_BM == 0, "superblock size must be a multiple of 128");
    cudaStream_t st = as_stream(stream);
    QC_CHECK(cudaMemsetAsync(mat, 0, sizeof(double) * nao * nao, st));
    if (nsb == 0) return 0;
    const SBDesc *sbd = (const SBDesc *)sbdesc;
    const int64_t ngl = (int64_t)nsb * sbp;
    const int wpb = 8;
    const unsigned nb1 = (unsigned)((ngl + wpb - 1) / wpb);
    prof_begin(PROF_VXC_VB, st);
    if (vgrad)
        vxc_vb_sb_kernel<4><<<nb1, wpb * 32, 0, st>>>(sbd, ao, sbp, ngl, weights, vrho, vgrad, vb_off, vb);
    else
        vxc_vb_sb_kernel<1><<<nb1, wpb * 32, 0, st>>>(sbd, ao, sbp, ngl, weights, vrho, vgrad, vb_off, vb);
    prof_end(st);
    QC_LAUNCHED(1);
    const int maxtiles = (max_nsp / GM_BN) * ((max_nsp + GM_BM - 1) / GM_BM);
    dim3 grid((unsigned)maxtiles, (unsigned)nsb);
    QC_CHECK(cudaFuncSetAttribute(vxc_sb_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES));
    prof_begin(PROF_VXC_GEMM, st);
    vxc_sb_gemm_kernel<<<grid, GM_THREADS, GM_SMEM_BYTES, st>>>(sbd, idx, ao, vb, vb_off, sbp, nao, mat, 0);
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

// Meta-GGA Vxc (hcgto.py:463-489) on the 5-component storage:
//   mat = sum_g w phi^T (vrho phi + 2 vgrad . grad phi + 2 vlapl lapl phi) + sum_d (d_d phi)^T w (2 vlapl + vkin / 2) (d_d phi)
// (the caller symmetrises, like the reference).  vlapl / vkin are de/d(lapl rho) and de/dtau.  fp64 DMMA engine.
__global__ void sb_scale_rows_kernel(const SBDesc *__restrict__ sbd, const double *__restrict__ ao, int sbp, int64_t ngrid_ld,
                                     int comp, const double *__restrict__ w, const double *__restrict__ vlapl,
                                     const double *__restrict__ vkin, const int64_t *__restrict__ vb_off,
                                     double *__restrict__ vb) {
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per grid row
    if (g >= ngrid_ld) return;
    const int sb = (int)(g / sbp), r = (int)(g - (int64_t)sb * sbp);
    const SBDesc d = sbd[sb];
    const int lane = threadIdx.x & 31;
    const double cf = w[g] * (2.0 * vlapl[g] + 0.5 * vkin[g]);
    const int64_t ld = d.nsp;
    const double2 *p = reinterpret_cast<const double2 *>(ao + d.ao_off + ((int64_t)comp * sbp + r) * ld);
    double2 *out = reinterpret_cast<double2 *>(vb + vb_off[sb] + (int64_t)r * ld);
    for (int c = lane; c < ld / 2; c += 32) {
        const double2 v = p[c];
        out[c] = make_double2(cf * v.x, cf * v.y);
    }
}

extern "C" int b200qc_vxc_sb_mgga(const void *sbdesc, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                                  const double *weights, const double *vrho, const double *vgrad, const double *vlapl,
                                  const double *vkin, int nao, const int64_t *vb_off, double *vb, double *mat,
                                  void *stream) {
    QC_REQUIRE(sbp % GM_BM == 0, "superblock size must be a multiple of 128");
    QC_REQUIRE(vrho && vgrad && vlapl && vkin, "all four potentials are needed");
    cudaStream_t st = as_stream(stream);
    QC_CHECK(cudaMemsetAsync(mat, 0, sizeof(double) * nao * nao, st));
    if (nsb == 0) return 0;
    const SBDesc *sbd = (const SBDesc *)sbdesc;
    const int64_t ngl = (int64_t)nsb * sbp;
    const int wpb = 8;
    const unsigned nb1 = (unsigned)((ngl + wpb - 1) / wpb);
    const int maxtiles = (max_nsp / GM_BN) * ((max_nsp + GM_BM - 1) / GM_BM);
    dim3 grid((unsigned)maxtiles, (unsigned)nsb);
    QC_CHECK(cudaFuncSetAttribute(vxc_sb_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES));
    prof_begin(PROF_VXC_VB, st);
    vxc_vb_sb_kernel<5><<<nb1, wpb * 32, 0, st>>>(sbd, ao, sbp, ngl, weights, vrho, vgrad, vb_off, vb, vlapl);
    prof_end(st);
    prof_begin(PROF_VXC_GEMM, st);
    vxc_sb_gemm_kernel<<<grid, GM_THREADS, GM_SMEM_BYTES, st>>>(sbd, idx, ao, vb, vb_off, sbp, nao, mat, 0);
    prof_end(st);
    QC_LAUNCHED(2);
    for (int dd = 1; dd <= 3; dd++) {
        prof_begin(PROF_VXC_VB, st);
        sb_scale_rows_kernel<<<nb1, wpb * 32, 0, st>>>(sbd, ao, sbp, ngl, dd, weights, vlapl, vkin, vb_off, vb);
        prof_end(st);
        prof_begin(PROF_VXC_GEMM, st);
        vxc_sb_gemm_kernel<<<grid, GM_THREADS, GM_SMEM_BYTES, st>>>(sbd, idx, ao, vb, vb_off, sbp, nao, mat, dd);
        prof_end(st);
        QC_LAUNCHED(2);
    }
    return 0;
}
