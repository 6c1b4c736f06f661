// Rys-quadrature integral engine: one-electron (overlap, kinetic, nuclear attraction / rinv) and
// two-electron integrals over 2, 3 or 4 centres -- replaces the libcint drivers the reference
// reaches through dqclibs (GTOint2c, GTOnr3c_drv, GTOnr2e_fill_drv; dqc/hamilton/intor/
// molintor.py:624-688) and feeds the direct J/K digestion (jk.cuh).
//
// Work decomposition.  A launch handles ONE angular-momentum class (li lj | lk ll) so that all
// control flow is warp-uniform.  A task = one contracted shell quartet (bra pair x ket pair from two
// explicit pair lists; absent centres of the 3c / 2c / 1e cases are dummy s shells with zero
// exponent and unit coefficient).  G = 8, 16 or 32 lanes cooperate on a task (picked from the
// number of cartesian components of the class), so a warp carries 32/G tasks:
//   1. every lane walks the primitive quartets of its task (pair quantities recomputed per lane,
//      two exps -- cheaper than a trip through memory);
//   2. lanes 0..2n-1 evaluate the n Rys roots and weights (Chebyshev table, rys.cuh) -> shared;
//   3. lanes (root, xyz) run the 2-D vertical recurrence + both horizontal transfers into a
//      per-(root, xyz) table in shared memory (stride padded odd: conflict-free);
//   4. every lane owns <= NACC cartesian components and accumulates sum_r Ix Iy Iz in registers;
//      the packed table offsets of each component sit in a CTA-wide shared array.
// After the primitive loop the block is moved through shared memory for the cartesian -> real
// spherical transform (one index at a time, ping-pong) and handed to a sink (dense store, packed
// (ij|P) store, Schwarz bound, or the J/K digestion).
#pragma once
#include "common.cuh"
#include "rys.cuh"

#define INT_MODE_ERI 0
#define INT_MODE_OVLP 1
#define INT_MODE_KIN 2
#define INT_MODE_NUC 3

#define SINK_DENSE 0      // out[(i)*s0 + (j)*s1 + (k)*s2 + (l)*s3]
#define SINK_PACK3C 1     // out[pair(i >= j) * s2 + k], s2 = naux
#define SINK_SCHWARZ 2    // out[ish * s0 + jsh] = sqrt(max |(ab|ab)|)  (bra == ket)
#define SINK_JK 3         // digestion into J / K (jk.cuh)
#define SINK_STORE_JK 4   // bra i >= j, ket k >= l: the 4 images into out[i][j][k][l] and out2[i][k][j][l]

#define INT_THREADS 128
#define INT_MAX_COMP 1312  // 32 lanes x 41 components

struct IntClass {
    int mode, sink;
    int l[4], present[4], nc[4], ns[4];
    int ncomp, nroots;
    int nij, nkl, ljt;      // table extents: i+j <= nij-1, k+l <= nkl-1, j index runs to ljt
    int fsize, region;      // final table size and padded per-(root, xyz) region (doubles)
    int tkin;               // offset of the kinetic table inside a region (KIN only)
    int gstride;            // doubles of shared memory per lane group
    int c2s_off[4];         // offsets of the 4 c2s matrices in the CTA-shared area
    int head;               // doubles of the CTA-shared head (c2s matrices + component offsets)
    int ncharge;
    int cart;               // raw cartesian output: the "c2s" matrices are identities (ns == nc)
};

struct IntArgs {
    const ShellRec *shells;
    const double *env;
    const int2 *bra, *ket;   // pair lists (second = -1: absent centre)
    int64_t nbra, nket;
    const double *charges;   // NUC: (ncharge, 4) = x, y, z, q   (value = sum_C q_C <a| 1/|r-C| |b>)
    double *out, *out2;
    int64_t s0, s1, s2, s3;  // dense strides / sink parameters
    int off[4];              // AO offset of each slot's slice start (subtracted before striding)
    // J/K sink
    const double *dm;
    double *vj, *vk;
    int nset, nao;
    const double *qbra, *qket;  // Schwarz bounds aligned with the pair lists (JK sink) or NULL
    double thresh;
    int same_lists;             // bra and ket lists are the same list: use only ket <= bra
};

__device__ __forceinline__ void int_build_table(const IntClass &K, double *tab, double w0, double c00,
                                                double c01, double b10, double b01, double b00, double ab,
                                                double cd) {
    const int nij = K.nij, nkl = K.nkl;
    double *W = tab + K.fsize;
    double *line = W + nij * nkl;
    W[0] = w0;
    if (nij > 1) W[nkl] = c00 * w0;
    for (int n = 1; n < nij - 1; n++) W[(n + 1) * nkl] = c00 * W[n * nkl] + n * b10 * W[(n - 1) * nkl];
    for (int m = 0; m < nkl - 1; m++)
        for (int n = 0; n < nij; n++) {
            double v = c01 * W[n * nkl + m];
            if (m > 0) v += m * b01 * W[n * nkl + m - 1];
            if (n > 0) v += n * b00 * W[(n - 1) * nkl + m];
            W[n * nkl + m + 1] = v;
        }
    const int li = K.l[0], ljt = K.ljt, lk = K.l[2], ll = K.l[3];
    for (int j = 0; j <= ljt; j++) {
        if (j > 0)
            for (int n = 0; n < nij - j; n++)
                for (int m = 0; m < nkl; m++) W[n * nkl + m] = W[(n + 1) * nkl + m] + ab * W[n * nkl + m];
        for (int i = 0; i <= li; i++) {
            for (int m = 0; m < nkl; m++) line[m] = W[i * nkl + m];
            for (int l = 0; l <= ll; l++) {
                if (l > 0)
                    for (int m = 0; m < nkl - l; m++) line[m] = line[m + 1] + cd * line[m];
                for (int k = 0; k <= lk; k++) tab[((i * (ljt + 1) + j) * (lk + 1) + k) * (ll + 1) + l] = line[k];
            }
        }
    }
}

struct SlotData {
    double x, y, z;
    int np, pe, pc, ao;
};

__device__ __forceinline__ SlotData int_slot(const ShellRec *shells, int sh) {
    SlotData s;
    const ShellRec r = shells[sh];
    s.x = r.x; s.y = r.y; s.z = r.z;
    s.np = r.nprim; s.pe = r.ptr_exp; s.pc = r.ptr_coef; s.ao = r.ao_off;
    return s;
}

// Computes the real-spherical block of one task.  Returns a pointer (inside gsm) to the block laid
// out [ns0][ns1][ns2][ns3].  All 32 lanes of the warp must call this together (valid == false for
// idle groups).
template <int G, int NACC>
__device__ __forceinline__ double *int_compute_block(const IntClass &K, const IntArgs &A, bool valid, int ish,
                                                      int jsh, int ksh, int lsh, double *gsm, const double *head,
                                                      const int *coff, const int *cofft, int lg) {
    double acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; a++) acc[a] = 0.0;
    double *rw = gsm;
    double *tables = gsm + 16;
    const int nr = K.nroots;
    SlotData si = {0.0, 0.0, 0.0, 1, -1, -1, 0};
    SlotData sj = si, sk = si, sl = si;
    if (valid) {
        si = int_slot(A.shells, ish);
        if (jsh >= 0) sj = int_slot(A.shells, jsh); else { sj = si; sj.np = 1; sj.pe = -1; }
        if (K.mode == INT_MODE_ERI) {
            sk = int_slot(A.shells, ksh);
            if (lsh >= 0) sl = int_slot(A.shells, lsh); else { sl = sk; sl.np = 1; sl.pe = -1; }
        }
    }
    const double abx = si.x - sj.x, aby = si.y - sj.y, abz = si.z - sj.z;
    const double ab2 = abx * abx + aby * aby + abz * abz;
    double cdx = 0, cdy = 0, cdz = 0, cd2 = 0;
    int nketp = 1;
    if (K.mode == INT_MODE_ERI) {
        cdx = sk.x - sl.x; cdy = sk.y - sl.y; cdz = sk.z - sl.z;
        cd2 = cdx * cdx + cdy * cdy + cdz * cdz;
        nketp = sk.np * sl.np;
    } else if (K.mode == INT_MODE_NUC) {
        nketp = K.ncharge;
    }
    const int total = valid ? si.np * sj.np * nketp : 0;
    const int tmax = __reduce_max_sync(0xffffffffu, total);
    const double PI = 3.14159265358979323846;

    for (int n = 0; n < tmax; n++) {
        bool act = n < total;
        double p = 1, q = 1, px = 0, py = 0, pz = 0, pqx = 0, pqy = 0, pqz = 0, qcx = 0, qcy = 0, qcz = 0;
        double pref = 0, a0 = 0, a1 = 0, x = 0, aj = 0;
        if (act) {
            const int ib = n / nketp, kq = n - ib * nketp;
            const int ip = ib / sj.np, jp = ib - ip * sj.np;
            const double ai = A.env[si.pe + ip];
            double cc = A.env[si.pc + ip];
            if (sj.pe >= 0) { aj = A.env[sj.pe + jp]; cc *= A.env[sj.pc + jp]; }
            p = ai + aj;
            const double ea = ai * aj / p * ab2;
            px = (ai * si.x + aj * sj.x) / p; py = (ai * si.y + aj * sj.y) / p; pz = (ai * si.z + aj * sj.z) / p;
            if (K.mode == INT_MODE_ERI) {
                const int kp = kq / sl.np, lp = kq - kp * sl.np;
                const double ak = A.env[sk.pe + kp];
                double al = 0;
                cc *= A.env[sk.pc + kp];
                if (sl.pe >= 0) { al = A.env[sl.pe + lp]; cc *= A.env[sl.pc + lp]; }
                q = ak + al;
                const double ec = ak * al / q * cd2;
                if (ea + ec > 80.0) act = false;
                const double qx = (ak * sk.x + al * sl.x) / q, qy = (ak * sk.y + al * sl.y) / q,
                             qz = (ak * sk.z + al * sl.z) / q;
                pqx = px - qx; pqy = py - qy; pqz = pz - qz;
                qcx = qx - sk.x; qcy = qy - sk.y; qcz = qz - sk.z;
                const double pq = p + q;
                x = p * q / pq * (pqx * pqx + pqy * pqy + pqz * pqz);
                pref = cc * 34.98683665524972497 /* 2 pi^2.5 */ / (p * q * sqrt(pq)) * exp(-ea - ec);
                a0 = q / pq; a1 = p / pq;
            } else if (K.mode == INT_MODE_NUC) {
                const double *ch = A.charges + 4 * kq;
                pqx = px - ch[0]; pqy = py - ch[1]; pqz = pz - ch[2];
                x = p * (pqx * pqx + pqy * pqy + pqz * pqz);
                pref = cc * ch[3] * 2.0 * PI / p * exp(-ea);
                a0 = 1.0; a1 = 0.0;
                if (ea > 80.0) act = false;
            } else {
                const double s = PI / p;
                pref = cc * s * sqrt(s) * exp(-ea);
                if (ea > 80.0) act = false;
            }
        }
        // ---- Rys roots / weights ----
        if (K.mode == INT_MODE_ERI || K.mode == INT_MODE_NUC) {
            if (act)
                for (int f = lg; f < 2 * nr; f += G) rw[f] = rys_eval(nr, f, x);
        } else if (lg == 0) {
            rw[0] = 0.0;
            rw[1] = 1.0;
        }
        __syncwarp();
        // ---- 2-D tables ----
        if (act) {
            const double pax = px - si.x, pay = py - si.y, paz = pz - si.z;
            for (int t = lg; t < 3 * nr; t += G) {
                const int r = t / 3, d = t - 3 * r;
                const double u = rw[r];
                const double pa = d == 0 ? pax : (d == 1 ? pay : paz);
                const double pq_ = d == 0 ? pqx : (d == 1 ? pqy : pqz);
                const double qc = d == 0 ? qcx : (d == 1 ? qcy : qcz);
                const double ab = d == 0 ? abx : (d == 1 ? aby : abz);
                const double cd = d == 0 ? cdx : (d == 1 ? cdy : cdz);
                const double w0 = d == 2 ? rw[nr + r] * pref : 1.0;
                double *tab = tables + t * K.region;
                int_build_table(K, tab, w0, pa - a0 * u * pq_, qc + a1 * u * pq_, (1.0 - a0 * u) / (2 * p),
                                (1.0 - a1 * u) / (2 * q), u / (2 * (p + q)), ab, cd);
                if (K.mode == INT_MODE_KIN) {
                    // T(i,j) = -1/2 [ j(j-1) S(i,j-2) - 2b(2j+1) S(i,j) + 4b^2 S(i,j+2) ]
                    const int li = K.l[0], lj = K.l[1], ljt = K.ljt;
                    double *T = tab + K.tkin;
                    for (int i = 0; i <= li; i++)
                        for (int j = 0; j <= lj; j++) {
                            double v = -2.0 * aj * (2 * j + 1) * tab[i * (ljt + 1) + j] +
                                       4.0 * aj * aj * tab[i * (ljt + 1) + j + 2];
                            if (j >= 2) v += j * (j - 1) * tab[i * (ljt + 1) + j - 2];
                            T[i * (lj + 1) + j] = -0.5 * v;
                        }
                }
            }
        }
        __syncwarp();
        // ---- contraction over roots ----
        if (act) {
            if (K.mode != INT_MODE_KIN) {
#pragma unroll
                for (int a = 0; a < NACC; a++) {
                    const int c = lg + G * a;
                    if (c < K.ncomp) {
                        const int o = coff[c];
                        const double *tx = tables + (o & 1023), *ty = tables + K.region + ((o >> 10) & 1023),
                                     *tz = tables + 2 * K.region + (o >> 20);
                        double s = 0.0;
                        for (int r = 0; r < nr; r++) s += tx[3 * r * K.region] * ty[3 * r * K.region] * tz[3 * r * K.region];
                        acc[a] += s;
                    }
                }
            } else {
#pragma unroll
                for (int a = 0; a < NACC; a++) {
                    const int c = lg + G * a;
                    if (c < K.ncomp) {
                        const int o = coff[c], ot = cofft[c];
                        const double *t0 = tables, *t1 = tables + K.region, *t2 = tables + 2 * K.region;
                        const double sx = t0[o & 1023], sy = t1[(o >> 10) & 1023], sz = t2[o >> 20];
                        const double kx = t0[K.tkin + (ot & 1023)], ky = t1[K.tkin + ((ot >> 10) & 1023)],
                                     kz = t2[K.tkin + (ot >> 20)];
                        acc[a] += kx * sy * sz + sx * ky * sz + sx * sy * kz;
                    }
                }
            }
        }
        __syncwarp();
    }

    // ---- cartesian -> real spherical, one index at a time (last index first) ----
    double *b0 = gsm, *b1 = gsm + K.ncomp;
#pragma unroll
    for (int a = 0; a < NACC; a++) {
        const int c = lg + G * a;
        if (c < K.ncomp) b0[c] = acc[a];
    }
    __syncwarp();
    int cur[4] = {K.nc[0], K.nc[1], K.nc[2], K.nc[3]};
    for (int s = 3; s >= 0; s--) {
        if (!K.present[s]) continue;
        const int nc = K.nc[s], ns = K.ns[s];
        int inner = 1, outer = 1;
        for (int q2 = s + 1; q2 < 4; q2++) inner *= cur[q2];
        for (int q2 = 0; q2 < s; q2++) outer *= cur[q2];
        const double *M = head + K.c2s_off[s];
        const int nout = outer * ns * inner;
        for (int e = lg; e < nout; e += G) {
            const int o = e / (ns * inner), rem = e - o * ns * inner;
            const int m = rem / inner, in_ = rem - m * inner;
            const double *src = b0 + (int64_t)o * nc * inner + in_;
            double v = 0.0;
            for (int c = 0; c < nc; c++) v += M[m * nc + c] * src[c * inner];
            b1[e] = v;
        }
        __syncwarp();
        double *t = b0; b0 = b1; b1 = t;
        cur[s] = ns;
    }
    return b0;
}

// fills the CTA-shared head: c2s matrices and packed component offsets
__device__ __forceinline__ void int_fill_head(const IntClass &K, double *head, int *coff, int *cofft) {
    for (int s = 0; s < 4; s++) {
        if (!K.present[s]) continue;
        const double *M = c2s_ptr(K.l[s]);
        for (int e = threadIdx.x; e < K.ns[s] * K.nc[s]; e += blockDim.x)
            head[K.c2s_off[s] + e] = K.cart ? ((e / K.nc[s] == e % K.nc[s]) ? 1.0 : 0.0) : M[e];
    }
    const int s1 = K.ljt + 1, s2 = K.l[2] + 1, s3 = K.l[3] + 1;
    for (int c = threadIdx.x; c < K.ncomp; c += blockDim.x) {
        int r = c;
        const int cl = r % K.nc[3]; r /= K.nc[3];
        const int ck = r % K.nc[2]; r /= K.nc[2];
        const int cj = r % K.nc[1];
        const int ci = r / K.nc[1];
        int o[3], ot[3];
        for (int d = 0; d < 3; d++) {
            const int pi = c_cart_pow[K.l[0]][ci][d], pj = K.present[1] ? c_cart_pow[K.l[1]][cj][d] : 0;
            const int pk = K.present[2] ? c_cart_pow[K.l[2]][ck][d] : 0;
            const int pl = K.present[3] ? c_cart_pow[K.l[3]][cl][d] : 0;
            o[d] = ((pi * s1 + pj) * s2 + pk) * s3 + pl;
            ot[d] = pi * (K.l[1] + 1) + pj;
        }
        coff[c] = o[0] | (o[1] << 10) | (o[2] << 20);
        cofft[c] = ot[0] | (ot[1] << 10) | (ot[2] << 20);
    }
}

template <int G, int NACC>
__global__ void __launch_bounds__(INT_THREADS)
int_dense_kernel(const IntClass K, const IntArgs A) {
    extern __shared__ __align__(16) double int_smem[];
    double *head = int_smem;
    int *coff = reinterpret_cast<int *>(int_smem + K.head - K.ncomp);  // 2 * ncomp ints = ncomp doubles
    int *cofft = coff + K.ncomp;
    int_fill_head(K, head, coff, cofft);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int lg = lane % G;
    const int grp = threadIdx.x / G;
    constexpr int GPC = INT_THREADS / G;
    double *gsm = int_smem + K.head + (int64_t)grp * K.gstride;
    const int64_t task = (int64_t)blockIdx.x * GPC + grp;
    const int64_t ntask = A.nbra * A.nket;   // SCHWARZ: nket == 1 and the ket pair is the bra pair
    bool valid = task < ntask;
    int ish = 0, jsh = -1, ksh = -1, lsh = -1;
    if (valid) {
        const int64_t ib = task / A.nket, ik = task - ib * A.nket;
        const int2 b = A.bra[ib];
        ish = b.x; jsh = b.y;
        if (K.sink == SINK_SCHWARZ) {
            ksh = ish; lsh = jsh;
        } else if (K.mode == INT_MODE_ERI) {
            const int2 k2 = A.ket[ik];
            ksh = k2.x; lsh = k2.y;
        }
    }
    const double *blk = int_compute_block<G, NACC>(K, A, valid, ish, jsh, ksh, lsh, gsm, head, coff, cofft, lg);
    if (!valid) return;
    const int n0 = K.ns[0], n1 = K.ns[1], n2 = K.ns[2], n3 = K.ns[3];
    const int ntot = n0 * n1 * n2 * n3;
    if (K.sink == SINK_SCHWARZ) {
        double m = 0.0;
        for (int e = lg; e < n0 * n1; e += G) {
            const int a = e / n1, b = e - a * n1;
            m = fmax(m, fabs(blk[((a * n1 + b) * n2 + a) * n3 + b]));
        }
        for (int o = G / 2; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lg == 0) {
            const double v = sqrt(m);
            A.out[(int64_t)ish * A.s0 + jsh] = v;
            A.out[(int64_t)jsh * A.s0 + ish] = v;
        }
        return;
    }
    const int ai = A.shells[ish].ao_off - A.off[0];
    const int aj = jsh >= 0 ? A.shells[jsh].ao_off - A.off[1] : 0;
    const int ak = ksh >= 0 ? A.shells[ksh].ao_off - A.off[2] : 0;
    const int al = lsh >= 0 ? A.shells[lsh].ao_off - A.off[3] : 0;
    if (K.sink == SINK_DENSE) {
        for (int e = lg; e < ntot; e += G) {
            int r = e;
            const int d = r % n3; r /= n3;
            const int c = r % n2; r /= n2;
            const int b = r % n1;
            const int a = r / n1;
            A.out[(ai + a) * A.s0 + (aj + b) * A.s1 + (ak + c) * A.s2 + (al + d) * A.s3] = blk[e];
        }
    } else if (K.sink == SINK_STORE_JK) {
        // stored-ERI regime: out = (ij|kl) as [i][j][k][l] (J layout), out2 = (ij|kl) as [i][k][j][l] (K layout)
        const int64_t q1 = A.s2, q2 = q1 * q1, q3 = q2 * q1;   // strides of a dense nao^4 tensor
        for (int e = lg; e < ntot; e += G) {
            int r = e;
            const int d = r % n3; r /= n3;
            const int c = r % n2; r /= n2;
            const int b = r % n1;
            const int a = r / n1;
            const int64_t I = ai + a, J = aj + b, Kk = ak + c, L = al + d;
            const double v = blk[e];
            A.out[I * q3 + J * q2 + Kk * q1 + L] = v;
            A.out[J * q3 + I * q2 + Kk * q1 + L] = v;
            A.out[I * q3 + J * q2 + L * q1 + Kk] = v;
            A.out[J * q3 + I * q2 + L * q1 + Kk] = v;
            A.out2[I * q3 + Kk * q2 + J * q1 + L] = v;
            A.out2[J * q3 + Kk * q2 + I * q1 + L] = v;
            A.out2[I * q3 + L * q2 + J * q1 + Kk] = v;
            A.out2[J * q3 + L * q2 + I * q1 + Kk] = v;
        }
    } else {  // SINK_PACK3C: i >= j only, k contiguous
        for (int e = lg; e < ntot; e += G) {
            int r = e;
            const int c = r % n2; r /= n2;
            const int b = r % n1;
            const int a = r / n1;
            const int64_t I = ai + a, J = aj + b;
            if (I >= J) A.out[(I * (I + 1) / 2 + J) * A.s2 + ak + c] = blk[e];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side

static int int_make_class(IntClass &K, int mode, int sink, const int l[4], const int present[4], int ncharge,
                          int cart = 0) {
    K.mode = mode;
    K.sink = sink;
    K.cart = cart;
    K.ncomp = 1;
    int L = 0;
    for (int s = 0; s < 4; s++) {
        K.l[s] = present[s] ? l[s] : 0;
        K.present[s] = present[s];
        K.nc[s] = NCART(K.l[s]);
        K.ns[s] = cart ? K.nc[s] : 2 * K.l[s] + 1;
        K.ncomp *= K.nc[s];
        L += K.l[s];
    }
    K.ljt = K.l[1] + (mode == INT_MODE_KIN ? 2 : 0);
    K.nroots = (mode == INT_MODE_ERI || mode == INT_MODE_NUC) ? L / 2 + 1 : 1;
    K.nij = K.l[0] + K.ljt + 1;
    K.nkl = K.l[2] + K.l[3] + 1;
    K.fsize = (K.l[0] + 1) * (K.ljt + 1) * (K.l[2] + 1) * (K.l[3] + 1);
    K.tkin = K.fsize + K.nij * K.nkl + K.nkl;
    K.region = K.tkin + (mode == INT_MODE_KIN ? (K.l[0] + 1) * (K.l[1] + 1) : 0);
    if (K.region % 2 == 0) K.region++;
    const int tab = 3 * K.nroots * K.region;
    const int c2sbuf = 2 * K.ncomp;
    K.gstride = 16 + (tab > c2sbuf ? tab : c2sbuf);
    if (K.gstride % 2) K.gstride++;
    int off = 0;
    for (int s = 0; s < 4; s++) {
        K.c2s_off[s] = off;
        if (present[s]) off += K.ns[s] * K.nc[s];
    }
    if (off % 2) off++;
    K.head = off + K.ncomp;  // + 2 * ncomp ints
    if (K.head % 2) K.head++;
    K.ncharge = ncharge;
    if (K.fsize > 1023) return 1;   // packed offsets are 10 bits
    if (K.nroots > RYS_NMAX) return 1;
    if (K.ncomp > INT_MAX_COMP) return 1;
    return 0;
}

template <int G, int NACC>
static int int_launch_t(const IntClass &K, const IntArgs &A, cudaStream_t st) {
    constexpr int GPC = INT_THREADS / G;
    const size_t smem = sizeof(double) * ((size_t)K.head + (size_t)GPC * K.gstride);
    QC_REQUIRE(smem <= 220 * 1024, "integral class needs too much shared memory");
    QC_CHECK(cudaFuncSetAttribute(int_dense_kernel<G, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntask = A.nbra * A.nket;
    const int64_t nblk = (ntask + GPC - 1) / GPC;
    QC_REQUIRE(nblk < 2147483647LL, "too many integral tasks in one launch");
    prof_begin(PROF_INTS, st);
    int_dense_kernel<G, NACC><<<(unsigned)nblk, INT_THREADS, smem, st>>>(K, A);
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

static int int_launch(const IntClass &K, const IntArgs &A, cudaStream_t st) {
    if (A.nbra == 0 || A.nket == 0) return 0;
    if (K.ncomp <= 16) return int_launch_t<8, 2>(K, A, st);
    if (K.ncomp <= 64) return int_launch_t<8, 8>(K, A, st);
    if (K.ncomp <= 256) return int_launch_t<16, 16>(K, A, st);
    if (K.ncomp <= 512) return int_launch_t<32, 16>(K, A, st);
    return int_launch_t<32, 41>(K, A, st);
}

// Device pair lists per angular-momentum class, built on the host from shell slices.
struct PairLists {
    std::vector<int2> h;                  // all pairs, grouped by class
    std::vector<int> cls_off, cls_la, cls_lb;  // class c = pairs [cls_off[c], cls_off[c+1])
    int2 *d = nullptr;
    int upload(cudaStream_t st) {
        if (h.empty()) return 0;
        if (cudaMallocAsync(&d, sizeof(int2) * h.size(), st) != cudaSuccess) return 1;
        if (cudaMemcpyAsync(d, h.data(), sizeof(int2) * h.size(), cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
        return 0;
    }
    void release(cudaStream_t st) {
        if (d) cudaFreeAsync(d, st);
        d = nullptr;
    }
};

// pairs (a in [a0,a1)) x (b in [b0,b1)) grouped by (la, lb); b0 < 0: single-centre list (b = -1);
// tri: keep only a >= b.
static void make_pair_lists(const b200qc_basis *B, int a0, int a1, int b0, int b1, bool tri, PairLists &P) {
    P.h.clear(); P.cls_off.clear(); P.cls_la.clear(); P.cls_lb.clear();
    for (int la = 0; la <= B200QC_LMAX; la++)
        for (int lb = 0; lb <= (b0 < 0 ? 0 : B200QC_LMAX); lb++) {
            const size_t start = P.h.size();
            for (int a = a0; a < a1; a++) {
                if (B->h_shells[a].l != la) continue;
                if (b0 < 0) {
                    P.h.push_back(make_int2(a, -1));
                    continue;
                }
                for (int b = b0; b < b1; b++) {
                    if (B->h_shells[b].l != lb) continue;
                    if (tri && b > a) continue;
                    P.h.push_back(make_int2(a, b));
                }
            }
            if (P.h.size() > start) {
                P.cls_off.push_back((int)start);
                P.cls_la.push_back(la);
                P.cls_lb.push_back(lb);
            }
        }
    P.cls_off.push_back((int)P.h.size());
}

static int int_require_ready(const b200qc_basis *basis) {
    QC_REQUIRE(basis != nullptr, "null basis");
    const int dev = qc_current_device();
    QC_REQUIRE(dev >= 0 && dev == basis->device,
               "the current CUDA device is not the one this basis was uploaded on (wrap the call in torch.cuda.device)");
    QC_REQUIRE(g_rys_ready[dev], "Rys table not uploaded on this device (b200qc_rys_upload)");
    return 0;
}

// Runs every (bra class) x (ket class) launch of a two-electron integral.
static int int_run_eri(const b200qc_basis *B, PairLists &bra, PairLists &ket, bool bra_pair, bool ket_pair, int sink,
                       IntArgs A, cudaStream_t st) {
    if (bra.upload(st) || ket.upload(st)) {
        b200qc_set_error("pair list upload failed");
        return 1;
    }
    A.shells = B->d_shells;
    A.env = B->d_env;
    int rc = 0;
    for (size_t cb = 0; cb + 1 < bra.cls_off.size() && !rc; cb++)
        for (size_t ck = 0; ck + 1 < ket.cls_off.size() && !rc; ck++) {
            IntClass K;
            const int l[4] = {bra.cls_la[cb], bra.cls_lb[cb], ket.cls_la[ck], ket.cls_lb[ck]};
            const int pr[4] = {1, bra_pair ? 1 : 0, 1, ket_pair ? 1 : 0};
            if (int_make_class(K, INT_MODE_ERI, sink, l, pr, 0, B->cart)) {
                b200qc_set_error("angular-momentum class outside the supported range (<= 1312 cartesian components)");
                rc = 2;
                break;
            }
            IntArgs a = A;
            a.bra = bra.d + bra.cls_off[cb];
            a.nbra = bra.cls_off[cb + 1] - bra.cls_off[cb];
            a.ket = ket.d + ket.cls_off[ck];
            a.nket = ket.cls_off[ck + 1] - ket.cls_off[ck];
            rc = int_launch(K, a, st);
        }
    bra.release(st);
    ket.release(st);
    return rc;
}

extern "C" int b200qc_int1e(const b200qc_basis *basis, int kind, const int *sl, const double *h_rinv_orig,
                            double *out, void *stream) {
    if (int_require_ready(basis)) return 2;
    QC_REQUIRE(kind >= 0 && kind <= 3, "kind must be 0..3");
    QC_REQUIRE(0 <= sl[0] && sl[0] <= sl[1] && sl[1] <= basis->nbas && 0 <= sl[2] && sl[2] <= sl[3] &&
                   sl[3] <= basis->nbas, "bad shell slice");
    cudaStream_t st = as_stream(stream);
    PairLists bra;
    make_pair_lists(basis, sl[0], sl[1], sl[2], sl[3], false, bra);
    if (bra.h.empty()) return 0;
    // point charges: -Z at every nucleus (int1e_nuc) or +1 at rinv_orig (int1e_rinv)
    std::vector<double> ch;
    if (kind == 2) {
        for (int a = 0; a < basis->natm; a++) {
            const int pc = basis->h_atm[a * ATM_SLOTS + 1];
            ch.insert(ch.end(), {basis->h_env[pc], basis->h_env[pc + 1], basis->h_env[pc + 2],
                                 -(double)basis->h_atm[a * ATM_SLOTS]});
        }
    } else if (kind == 3) {
        QC_REQUIRE(h_rinv_orig != nullptr, "rinv origin missing");
        ch.insert(ch.end(), {h_rinv_orig[0], h_rinv_orig[1], h_rinv_orig[2], 1.0});
    }
    double *d_ch = nullptr;
    if (!ch.empty()) {
        QC_CHECK(cudaMallocAsync(&d_ch, sizeof(double) * ch.size(), st));
        QC_CHECK(cudaMemcpyAsync(d_ch, ch.data(), sizeof(double) * ch.size(), cudaMemcpyHostToDevice, st));
    }
    if (bra.upload(st)) {
        b200qc_set_error("pair list upload failed");
        return 1;
    }
    const int mode = kind == 0 ? INT_MODE_OVLP : (kind == 1 ? INT_MODE_KIN : INT_MODE_NUC);
    IntArgs A = {};
    A.shells = basis->d_shells;
    A.env = basis->d_env;
    A.charges = d_ch;
    A.out = out;
    const int nj = basis->h_ao_loc[sl[3]] - basis->h_ao_loc[sl[2]];
    A.s0 = nj; A.s1 = 1; A.s2 = 0; A.s3 = 0;
    A.off[0] = basis->h_ao_loc[sl[0]];
    A.off[1] = basis->h_ao_loc[sl[2]];
    int2 dummy = make_int2(-1, -1);
    (void)dummy;
    int rc = 0;
    for (size_t cb = 0; cb + 1 < bra.cls_off.size() && !rc; cb++) {
        IntClass K;
        const int l[4] = {bra.cls_la[cb], bra.cls_lb[cb], 0, 0};
        const int pr[4] = {1, 1, 0, 0};
        if (int_make_class(K, mode, SINK_DENSE, l, pr, (int)ch.size() / 4, basis->cart)) {
            b200qc_set_error("unsupported angular momentum class");
            rc = 2;
            break;
        }
        IntArgs a = A;
        a.bra = bra.d + bra.cls_off[cb];
        a.nbra = bra.cls_off[cb + 1] - bra.cls_off[cb];
        a.ket = nullptr;
        a.nket = 1;
        rc = int_launch(K, a, st);
    }
    bra.release(st);
    if (d_ch) cudaFreeAsync(d_ch, st);
    // the host vectors behind the async copies must outlive them
    QC_CHECK(cudaStreamSynchronize(st));
    return rc;
}

extern "C" int b200qc_int2c2e(const b200qc_basis *basis, const int *sl, double *out, void *stream) {
    if (int_require_ready(basis)) return 2;
    cudaStream_t st = as_stream(stream);
    PairLists bra, ket;
    make_pair_lists(basis, sl[0], sl[1], -1, -1, false, bra);
    make_pair_lists(basis, sl[2], sl[3], -1, -1, false, ket);
    IntArgs A = {};
    A.out = out;
    A.s0 = 0; A.s1 = 0; A.s2 = 1; A.s3 = 0;
    A.s0 = basis->h_ao_loc[sl[3]] - basis->h_ao_loc[sl[2]];
    A.off[0] = basis->h_ao_loc[sl[0]];
    A.off[2] = basis->h_ao_loc[sl[2]];
    int rc = int_run_eri(basis, bra, ket, false, false, SINK_DENSE, A, st);
    QC_CHECK(cudaStreamSynchronize(st));
    return rc;
}

static int int3c2e_impl(const b200qc_basis *basis, const int *sl, double *out, bool packed, int64_t ld,
                        cudaStream_t st) {
    if (int_require_ready(basis)) return 2;
    PairLists bra, ket;
    if (packed) QC_REQUIRE(sl[0] == sl[2] && sl[1] == sl[3], "packed (ij|P) needs identical i and j slices");
    make_pair_lists(basis, sl[0], sl[1], sl[2], sl[3], packed, bra);
    make_pair_lists(basis, sl[4], sl[5], -1, -1, false, ket);
    IntArgs A = {};
    A.out = out;
    const int64_t nj = basis->h_ao_loc[sl[3]] - basis->h_ao_loc[sl[2]];
    const int64_t nk = basis->h_ao_loc[sl[5]] - basis->h_ao_loc[sl[4]];
    if (packed) QC_REQUIRE(ld >= nk, "ld must be >= naux");
    A.s0 = nj * nk; A.s1 = nk; A.s2 = packed ? ld : 1; A.s3 = 0;
    A.off[0] = basis->h_ao_loc[sl[0]];
    A.off[1] = basis->h_ao_loc[sl[2]];
    A.off[2] = basis->h_ao_loc[sl[4]];
    int rc = int_run_eri(basis, bra, ket, true, false, packed ? SINK_PACK3C : SINK_DENSE, A, st);
    QC_CHECK(cudaStreamSynchronize(st));
    return rc;
}

extern "C" int b200qc_int3c2e(const b200qc_basis *basis, const int *sl, double *out, void *stream) {
    return int3c2e_impl(basis, sl, out, false, 0, as_stream(stream));
}
extern "C" int b200qc_int3c2e_packed(const b200qc_basis *basis, const int *sl, double *out, int64_t ld,
                                     void *stream) {
    return int3c2e_impl(basis, sl, out, true, ld, as_stream(stream));
}

extern "C" int b200qc_int2e(const b200qc_basis *basis, const int *sl, double *out, void *stream) {
    if (int_require_ready(basis)) return 2;
    cudaStream_t st = as_stream(stream);
    PairLists bra, ket;
    make_pair_lists(basis, sl[0], sl[1], sl[2], sl[3], false, bra);
    make_pair_lists(basis, sl[4], sl[5], sl[6], sl[7], false, ket);
    IntArgs A = {};
    A.out = out;
    const int64_t nj = basis->h_ao_loc[sl[3]] - basis->h_ao_loc[sl[2]];
    const int64_t nk = basis->h_ao_loc[sl[5]] - basis->h_ao_loc[sl[4]];
    const int64_t nl = basis->h_ao_loc[sl[7]] - basis->h_ao_loc[sl[6]];
    A.s0 = nj * nk * nl; A.s1 = nk * nl; A.s2 = nl; A.s3 = 1;
    A.off[0] = basis->h_ao_loc[sl[0]];
    A.off[1] = basis->h_ao_loc[sl[2]];
    A.off[2] = basis->h_ao_loc[sl[4]];
    A.off[3] = basis->h_ao_loc[sl[6]];
    int rc = int_run_eri(basis, bra, ket, true, true, SINK_DENSE, A, st);
    QC_CHECK(cudaStreamSynchronize(st));
    return rc;
}

// Raw cartesian output for every integral entry point of this basis handle (flag != 0), or back to real spherical AOs.
// The caller's ao_loc (b200qc_basis_upload) must count (l + 1)(l + 2) / 2 functions per shell while the flag is set.
// Derivative integrals are linear combinations of such blocks over shells with l + 1 (coefficients c_p 2 a_p) and
// l - 1: d/dx [x^a y^b z^c e^(-a r^2)] = a x^(a-1) ... - 2 a x^(a+1) ...  (dqc_b200/hamilton/intor/deriv.py; the
// reference gets them from libcint's int1e_ip* / int2e_ip1, molintor.py:178-578).
extern "C" int b200qc_basis_set_cartesian(b200qc_basis *basis, int flag) {
    QC_REQUIRE(basis != nullptr, "null basis");
    if (flag) {
        for (int i = 0; i < basis->nbas; i++)
            QC_REQUIRE(basis->h_ao_loc[i + 1] - basis->h_ao_loc[i] == NCART(basis->h_shells[i].l),
                       "ao_loc of a cartesian basis must count (l + 1)(l + 2) / 2 functions per shell");
    }
    basis->cart = flag ? 1 : 0;
    return 0;
}

// Real-spherical transform of the kernels: out[(2 l + 1)][(l + 1)(l + 2) / 2], row m = -l..l (p: x, y, z), column =
// cartesian component in libcint order, including the angular normalisation.
extern "C" int b200qc_c2s_matrix(int l, double *h_out) {
    QC_REQUIRE(l >= 0 && l <= B200QC_LMAX && h_out != nullptr, "bad arguments");
    h_fill_c2s(l, h_out);
    return 0;
}

// Stored-ERI regime (small molecules): both dense layouts of (ij|kl) over shells [sh0, sh1) filled from
// the quartets with i >= j, k >= l (4-fold symmetry).  eri_j[i][j][k][l] = eri_k[i][k][j][l] = (ij|kl).
extern "C" int b200qc_eri_store(const b200qc_basis *basis, int sh0, int sh1, double *eri_j, double *eri_k,
                                void *stream) {
    if (int_require_ready(basis)) return 2;
    QC_REQUIRE(0 <= sh0 && sh0 < sh1 && sh1 <= basis->nbas && eri_j && eri_k, "bad arguments");
    QC_REQUIRE(!basis->cart, "the stored-ERI regime works on spherical AOs");
    cudaStream_t st = as_stream(stream);
    PairLists bra, ket;
    make_pair_lists(basis, sh0, sh1, sh0, sh1, true, bra);
    make_pair_lists(basis, sh0, sh1, sh0, sh1, true, ket);
    IntArgs A = {};
    A.out = eri_j;
    A.out2 = eri_k;
    A.s2 = basis->h_ao_loc[sh1] - basis->h_ao_loc[sh0];
    for (int q = 0; q < 4; q++) A.off[q] = basis->h_ao_loc[sh0];
    int rc = int_run_eri(basis, bra, ket, true, true, SINK_STORE_JK, A, st);
    QC_CHECK(cudaStreamSynchronize(st));
    return rc;
}
