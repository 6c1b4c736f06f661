// K8 -- direct Coulomb / exchange build: (ij|kl) blocks from the Rys engine (ints.cuh) are digested
// into J and K as they are produced, never stored.  Replaces the dense-ERI contractions
//   J_kl = sum_ij D_ij (ij|kl)            dqc/hamilton/hcgto.py:204-211
//   K_jk = sum_il D_il (ij|kl)            dqc/hamilton/hcgto.py:224-241
// (the reference materialises nao^4 doubles: 12.7 TB for Taxol/def2-SVP).
//
// A plan (built once per geometry) holds the shell pairs that survive the Schwarz test, in canonical orientation
// l(i) >= l(j), grouped into classes (l_i, l_j, bucket of the primitive-pair count) and sorted inside a class by their
// Schwarz bound Q_ij = sqrt(max |(ij|ij)|); per pair a geometry record and its primitive-pair data (JKPair / JKPrim,
// common.cuh); per class pair (bra class >= ket class) the work items: (bra pair, chunk of ket pairs with
// Q_bra Q_ket >= thresh and ket <= bra in the canonical order).  Unique quartets only (8-fold symmetry); D must be
// symmetric (the caller symmetrises, which leaves the reference's symmetrised J and K unchanged).
// Two engines digest the items:
//   * classes of s, p, d shells (<= 81 primitive pairs per shell pair): the register-resident quartet engine of
//     jk_reg.cuh (third translation unit; chunks of 128 kets, one lane per contracted quartet);
//   * everything else (f, g shells): jk_kernel below -- a lane group per quartet, 2-D tables in shared memory
//     (int_compute_block of ints.cuh), chunks of 16 kets, J_ij partial sums of the bra tile in shared memory.
// The class-pair launches of a build go round-robin onto four side streams (b200qc_jkplan_run).  fp64 atomics for the
// ket-side J and the four K tiles (sums of O(1e4) terms: reproducible to ~1e-14, far inside the 1e-6 parity bar).
#pragma once
#include "ints.cuh"
#include <algorithm>
#include <cstdlib>

#define JK_CHUNK 16
#define JK_MAXSET 2
#define JK_NSTREAM 4
static cudaStream_t g_jk_streams[64][JK_NSTREAM];
static cudaEvent_t g_jk_join[64][JK_NSTREAM], g_jk_fork[64];
static bool g_jk_streams_ready[64] = {};

struct JKClassPair {
    IntClass K;
    int bra_off, nbra, ket_off, nket;   // ranges inside the plan's pair array
    int64_t *d_work_off = nullptr;       // (nbra + 1) prefix of chunk counts
    int *d_nket_of_bra = nullptr;        // admissible kets per bra
    int64_t nitems = 0;
    int same;
    int reg = 0;                         // 1: register-resident quartet engine (jk_reg.cuh), work items in d_items
    int2 *d_items = nullptr;             // (bra, first ket) per chunk of JKR_CHUNK kets
};

struct b200qc_jkplan {
    const b200qc_basis *basis;
    int sh0, sh1, ao0, nao;
    int2 *d_pairs = nullptr;
    double *d_q = nullptr;
    JKPair *d_jkpairs = nullptr;         // the same pairs as d_pairs with their geometry (register engine)
    JKPrim *d_prims = nullptr;           // primitive-pair data, pp_off of a pair points in here
    double *d_scratch_j = nullptr;       // (nao, nao): J sink of K-only runs of the register engine
    std::vector<JKClassPair> cps;
    int64_t nquartets = 0, nquartets_reg = 0;
    double flops_int = 0.0;              // integral evaluation (roots, 2-D tables, sum over roots), one pass per quartet
    double ncomp_total = 0.0;            // sum over quartets of the cartesian block size (digestion: 2 flops per use)
};

// fp64 operations of ONE primitive quartet of a class as the Rys scheme needs them (each class evaluated in one pass):
// 2 nr Clenshaw sums of the root table, per (root, xyz) the vertical recurrence and both horizontal transfers, then
// nr (2 mul + 1 add) per cartesian component.  The roofline numerator of the J/K kernels (bench.py).
static double jk_prim_flops(const int l[4]) {
    const int nr = (l[0] + l[1] + l[2] + l[3]) / 2 + 1;
    const int nij = l[0] + l[1] + 1, nkl = l[2] + l[3] + 1;
    double tab = 0.0;
    tab += 3.0 * (nij - 1);                              // n recurrence at m = 0
    tab += 5.0 * nij * (nkl - 1);                        // m recurrence
    for (int j = 1; j <= l[1]; j++) tab += 2.0 * (nij - j) * nkl;
    for (int ll = 1; ll <= l[3]; ll++) tab += 2.0 * (nkl - ll) * (l[0] + 1) * (l[1] + 1);
    const double ncomp = (double)NCART(l[0]) * NCART(l[1]) * NCART(l[2]) * NCART(l[3]);
    return 40.0 + 2.0 * nr * (2.0 * 13 + 2) + nr * (3.0 * tab + 12.0) + 3.0 * nr * ncomp;
}

// Primitive pairs whose Gaussian-product factor exp(-a_i a_j |AB|^2 / (a_i + a_j)) is below e^-50 = 2e-22 are left out
// of the register engine's pair data (the shared-memory engine skips primitive quartets beyond e^-80).
#define JKR_EACUT 50.0
static const int g_jkr_buckets[] = {1, 2, 3, 4, 6, 9, 16, 36, JKR_MAXPP};
#define JKR_NBUCKET ((int)(sizeof(g_jkr_buckets) / sizeof(int)))
// bucket of a primitive-pair count: lanes of a warp walk kets of one bucket, so their trip counts are close
static int jkr_bucket(int npp) {
    for (int b = 0; b < JKR_NBUCKET; b++)
        if (npp <= g_jkr_buckets[b]) return b;
    return JKR_NBUCKET;   // too many primitive pairs to stage: shared-memory engine
}

template <int G, int NACC>
__global__ void __launch_bounds__(INT_THREADS)
jk_kernel(const IntClass K, const IntArgs A, const int64_t *__restrict__ work_off, const int *__restrict__ nket_of_bra,
          int64_t nitems, int item0, int item_stride) {
    extern __shared__ __align__(16) double int_smem[];
    double *head = int_smem;
    int *coff = reinterpret_cast<int *>(int_smem + K.head - K.ncomp);
    int *cofft = coff + K.ncomp;
    int_fill_head(K, head, coff, cofft);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int lg = lane % G;
    const int grp = threadIdx.x / G;
    constexpr int GPC = INT_THREADS / G;
    // per group: [engine scratch gstride][dtile 6 * 25 * nset][jbra 25 * nset]
    const int n0 = K.ns[0], n1 = K.ns[1], n2 = K.ns[2], n3 = K.ns[3];
    const int nset = A.nset;
    const int gtot = K.gstride + 7 * 25 * JK_MAXSET;
    double *gsm = int_smem + K.head + (int64_t)grp * gtot;
    double *dt = gsm + K.gstride;              // D tiles: kl, ij, jl, il, jk, ik  (each <= 25 per set)
    double *jb = dt + 6 * 25 * JK_MAXSET;      // J_ij partial of the bra tile
    const int64_t slot = (int64_t)blockIdx.x * GPC + grp;
    const int64_t item = item0 + slot * item_stride;
    const bool have = item < nitems;
    int ib = 0, k0 = 0, k1 = 0;
    if (have) {
        // binary search: work_off[ib] <= item < work_off[ib + 1]
        int lo = 0, hi = (int)A.nbra;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (work_off[mid] <= item) lo = mid; else hi = mid;
        }
        ib = lo;
        k0 = (int)(item - work_off[ib]) * JK_CHUNK;
        k1 = min(k0 + JK_CHUNK, nket_of_bra[ib]);
    }
    const int nk = have ? k1 - k0 : 0;
    const int nkmax = __reduce_max_sync(0xffffffffu, nk);
    int ish = 0, jsh = 0, ai = 0, aj = 0;
    if (have) {
        const int2 b = A.bra[ib];
        ish = b.x; jsh = b.y;
        ai = A.shells[ish].ao_off - A.off[0];
        aj = A.shells[jsh].ao_off - A.off[0];
        for (int e = lg; e < 25 * JK_MAXSET; e += G) jb[e] = 0.0;
    }
    const int nao = A.nao;
    const int64_t nn = (int64_t)nao * nao;
    for (int kk = 0; kk < nkmax; kk++) {
        const bool valid = kk < nk;
        int ksh = 0, lsh = 0;
        if (valid) {
            const int2 k2 = A.ket[k0 + kk];
            ksh = k2.x; lsh = k2.y;
        }
        const double *blk = int_compute_block<G, NACC>(K, A, valid, ish, jsh, ksh, lsh, gsm, head, coff, cofft, lg);
        if (valid) {
            const int ak = A.shells[ksh].ao_off - A.off[0], al = A.shells[lsh].ao_off - A.off[0];
            double f = 1.0;
            if (ish == jsh) f *= 0.5;
            if (ksh == lsh) f *= 0.5;
            if (ish == ksh && jsh == lsh) f *= 0.5;
            // stage the six density tiles
            for (int s = 0; s < nset; s++) {
                const double *D = A.dm + s * nn;
                double *t = dt + s * 150;
                for (int e = lg; e < n2 * n3; e += G) t[e] = D[(int64_t)(ak + e / n3) * nao + al + e % n3];
                for (int e = lg; e < n0 * n1; e += G) t[25 + e] = D[(int64_t)(ai + e / n1) * nao + aj + e % n1];
                for (int e = lg; e < n1 * n3; e += G) t[50 + e] = D[(int64_t)(aj + e / n3) * nao + al + e % n3];
                for (int e = lg; e < n0 * n3; e += G) t[75 + e] = D[(int64_t)(ai + e / n3) * nao + al + e % n3];
                for (int e = lg; e < n1 * n2; e += G) t[100 + e] = D[(int64_t)(aj + e / n2) * nao + ak + e % n2];
                for (int e = lg; e < n0 * n2; e += G) t[125 + e] = D[(int64_t)(ai + e / n2) * nao + ak + e % n2];
            }
        }
        __syncwarp();
        if (valid) {
            const int ak = A.shells[ksh].ao_off - A.off[0], al = A.shells[lsh].ao_off - A.off[0];
            double f = 1.0;
            if (ish == jsh) f *= 0.5;
            if (ksh == lsh) f *= 0.5;
            if (ish == ksh && jsh == lsh) f *= 0.5;
            for (int s = 0; s < nset; s++) {
                const double *t = dt + s * 150;
                if (A.vj) {
                    double *J = A.vj + s * nn;
                    // J_ij += 2f sum_kl B D_kl   (kept in shared memory across the chunk)
                    for (int e = lg; e < n0 * n1; e += G) {
                        const double *b = blk + (int64_t)e * n2 * n3;
                        double v = 0.0;
                        for (int c = 0; c < n2 * n3; c++) v += b[c] * t[c];
                        jb[s * 25 + e] += 2.0 * f * v;
                    }
                    // J_kl += 2f sum_ij B D_ij
                    for (int e = lg; e < n2 * n3; e += G) {
                        double v = 0.0;
                        for (int c = 0; c < n0 * n1; c++) v += blk[(int64_t)c * n2 * n3 + e] * t[25 + c];
                        atomicAdd(J + (int64_t)(ak + e / n3) * nao + al + e % n3, 2.0 * f * v);
                    }
                }
                if (A.vk) {
                    double *Kx = A.vk + s * nn;
                    // K_ik += f sum_jl B D_jl
                    for (int e = lg; e < n0 * n2; e += G) {
                        const int a = e / n2, c = e % n2;
                        double v = 0.0;
                        for (int b = 0; b < n1; b++)
                            for (int d = 0; d < n3; d++) v += blk[((a * n1 + b) * n2 + c) * n3 + d] * t[50 + b * n3 + d];
                        atomicAdd(Kx + (int64_t)(ai + a) * nao + ak + c, f * v);
                    }
                    // K_jk += f sum_il B D_il
                    for (int e = lg; e < n1 * n2; e += G) {
                        const int b = e / n2, c = e % n2;
                        double v = 0.0;
                        for (int a = 0; a < n0; a++)
                            for (int d = 0; d < n3; d++) v += blk[((a * n1 + b) * n2 + c) * n3 + d] * t[75 + a * n3 + d];
                        atomicAdd(Kx + (int64_t)(aj + b) * nao + ak + c, f * v);
                    }
                    // K_il += f sum_jk B D_jk
                    for (int e = lg; e < n0 * n3; e += G) {
                        const int a = e / n3, d = e % n3;
                        double v = 0.0;
                        for (int b = 0; b < n1; b++)
                            for (int c = 0; c < n2; c++) v += blk[((a * n1 + b) * n2 + c) * n3 + d] * t[100 + b * n2 + c];
                        atomicAdd(Kx + (int64_t)(ai + a) * nao + al + d, f * v);
                    }
                    // K_jl += f sum_ik B D_ik
                    for (int e = lg; e < n1 * n3; e += G) {
                        const int b = e / n3, d = e % n3;
                        double v = 0.0;
                        for (int a = 0; a < n0; a++)
                            for (int c = 0; c < n2; c++) v += blk[((a * n1 + b) * n2 + c) * n3 + d] * t[125 + a * n2 + c];
                        atomicAdd(Kx + (int64_t)(aj + b) * nao + al + d, f * v);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (have && A.vj)
        for (int s = 0; s < nset; s++)
            for (int e = lg; e < n0 * n1; e += G)
                atomicAdd(A.vj + s * nn + (int64_t)(ai + e / n1) * nao + aj + e % n1, jb[s * 25 + e]);
}

// out = acc + acc^T
__global__ void jk_symm_kernel(double *__restrict__ m, int nao, int nset) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nn = (int64_t)nao * nao;
    if (idx >= nn * nset) return;
    const int s = (int)(idx / nn);
    const int64_t r = idx - s * nn;
    const int i = (int)(r / nao), j = (int)(r - (int64_t)i * nao);
    if (i > j) return;
    double *M = m + s * nn;
    const double v = M[(int64_t)i * nao + j] + M[(int64_t)j * nao + i];
    M[(int64_t)i * nao + j] = v;
    M[(int64_t)j * nao + i] = v;
}

template <int G, int NACC>
static int jk_launch_t(const JKClassPair &cp, const IntArgs &A, int rank, int world, cudaStream_t st) {
    constexpr int GPC = INT_THREADS / G;
    const int gtot = cp.K.gstride + 7 * 25 * JK_MAXSET;
    const size_t smem = sizeof(double) * ((size_t)cp.K.head + (size_t)GPC * gtot);
    QC_REQUIRE(smem <= 220 * 1024, "J/K class needs too much shared memory");
    QC_CHECK(cudaFuncSetAttribute(jk_kernel<G, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t mine = (cp.nitems - rank + world - 1) / world;   // items rank, rank + world, ...
    if (mine <= 0) return 0;
    const int64_t nblk = (mine + GPC - 1) / GPC;
    QC_REQUIRE(nblk < 2147483647LL, "too many J/K work items in one launch");
    prof_begin(PROF_JK, st);
    jk_kernel<G, NACC><<<(unsigned)nblk, INT_THREADS, smem, st>>>(cp.K, A, cp.d_work_off, cp.d_nket_of_bra, cp.nitems,
                                                                  rank, world);
    prof_end(st);
    QC_LAUNCHED(1);
    return 0;
}

extern "C" int b200qc_jkplan_free(b200qc_jkplan *p) {
    if (!p) return 0;
    cudaFree(p->d_pairs);
    cudaFree(p->d_q);
    cudaFree(p->d_jkpairs);
    cudaFree(p->d_prims);
    cudaFree(p->d_scratch_j);
    for (auto &cp : p->cps) {
        cudaFree(cp.d_work_off);
        cudaFree(cp.d_nket_of_bra);
        cudaFree(cp.d_items);
    }
    delete p;
    return 0;
}

extern "C" int b200qc_jkplan_create(const b200qc_basis *basis, int sh0, int sh1, double thresh, b200qc_jkplan **out,
                                    void *stream) {
    if (int_require_ready(basis)) return 2;
    QC_REQUIRE(!basis->cart, "the direct J/K engine works on spherical AOs");
    QC_REQUIRE(0 <= sh0 && sh0 < sh1 && sh1 <= basis->nbas && out, "bad arguments");
    cudaStream_t st = as_stream(stream);
    const int nb = basis->nbas;
    // 1. Schwarz bounds of every pair i >= j of the slice
    PairLists tri;
    make_pair_lists(basis, sh0, sh1, sh0, sh1, true, tri);
    double *d_q = nullptr;
    QC_CHECK(cudaMalloc(&d_q, sizeof(double) * nb * nb));
    QC_CHECK(cudaMemsetAsync(d_q, 0, sizeof(double) * nb * nb, st));
    if (tri.upload(st)) {
        b200qc_set_error("pair list upload failed");
        return 1;
    }
    for (size_t c = 0; c + 1 < tri.cls_off.size(); c++) {
        IntClass K;
        const int l[4] = {tri.cls_la[c], tri.cls_lb[c], tri.cls_la[c], tri.cls_lb[c]};
        const int pr[4] = {1, 1, 1, 1};
        QC_REQUIRE(!int_make_class(K, INT_MODE_ERI, SINK_SCHWARZ, l, pr, 0), "class outside the supported range");
        IntArgs A = {};
        A.shells = basis->d_shells;
        A.env = basis->d_env;
        A.bra = tri.d + tri.cls_off[c];
        A.nbra = tri.cls_off[c + 1] - tri.cls_off[c];
        A.ket = A.bra;
        A.nket = 1;
        A.out = d_q;
        A.s0 = nb;
        int rc = int_launch(K, A, st);
        if (rc) return rc;
    }
    std::vector<double> q((size_t)nb * nb);
    QC_CHECK(cudaMemcpyAsync(q.data(), d_q, sizeof(double) * nb * nb, cudaMemcpyDeviceToHost, st));
    QC_CHECK(cudaStreamSynchronize(st));
    tri.release(st);
    cudaFree(d_q);
    double qmax = 0.0;
    for (const int2 &p : tri.h) qmax = std::max(qmax, q[(size_t)p.x * nb + p.y]);

    // 2. surviving pairs in canonical orientation l(i) >= l(j), with their primitive-pair data; classes =
    //    (l_i, l_j, bucket of the primitive-pair count), inside a class sorted by Q descending
    auto *plan = new b200qc_jkplan();
    plan->basis = basis;
    plan->sh0 = sh0; plan->sh1 = sh1;
    plan->ao0 = basis->h_ao_loc[sh0];
    plan->nao = basis->h_ao_loc[sh1] - plan->ao0;
    struct PairTmp {
        int2 p;
        double q;
        int key;                 // (l_i * 8 + l_j) * 16 + bucket
        std::vector<JKPrim> prims;
    };
    std::vector<PairTmp> all;
    static const double c2s_const[2] = {0.282094791773878143, 0.488602511902919921};   // s, p (tables.cuh)
    const bool reg_ok = !getenv("B200QC_JK_NOREG");
    // small systems (a few thousand shell pairs) are bound by the number of launches, not by the spread of the primitive
    // loops inside a warp: one bucket per (l_i, l_j) there -- 21 class pairs instead of ~400 for benzene / cc-pVDZ
    // (B200QC_JK_ONE_BUCKET = 0 / 1 overrides the size test: the tests run both forms on the same molecule)
    const char *ob = getenv("B200QC_JK_ONE_BUCKET");
    const bool one_bucket = ob ? ob[0] == '1' : tri.h.size() < 4096;
    for (size_t e = 0; e < tri.h.size(); e++) {
        int2 p = tri.h[e];
        const double qq = q[(size_t)p.x * nb + p.y];
        if (qq * qmax < thresh) continue;
        if (basis->h_shells[p.x].l < basis->h_shells[p.y].l) std::swap(p.x, p.y);
        const ShellRec &si = basis->h_shells[p.x], &sj = basis->h_shells[p.y];
        PairTmp t;
        t.p = p;
        t.q = qq;
        const double abx = si.x - sj.x, aby = si.y - sj.y, abz = si.z - sj.z;
        const double ab2 = abx * abx + aby * aby + abz * abz;
        const double sc = (si.l <= 1 ? c2s_const[si.l] : 1.0) * (sj.l <= 1 ? c2s_const[sj.l] : 1.0);
        for (int ip = 0; ip < si.nprim; ip++)
            for (int jp = 0; jp < sj.nprim; jp++) {
                const double ai = basis->h_env[si.ptr_exp + ip], aj = basis->h_env[sj.ptr_exp + jp];
                const double pp = ai + aj, ea = ai * aj / pp * ab2;
                if (ea > JKR_EACUT) continue;
                JKPrim r;
                r.p = pp;
                r.hp = 0.5 / pp;
                r.px = (ai * si.x + aj * sj.x) / pp;
                r.py = (ai * si.y + aj * sj.y) / pp;
                r.pz = (ai * si.z + aj * sj.z) / pp;
                r.c = sc * basis->h_env[si.ptr_coef + ip] * basis->h_env[sj.ptr_coef + jp] * std::exp(-ea) / pp;
                t.prims.push_back(r);
            }
        int bucket = reg_ok ? jkr_bucket((int)t.prims.size()) : JKR_NBUCKET;
        if (one_bucket && bucket < JKR_NBUCKET) bucket = 0;
        t.key = (si.l * 8 + sj.l) * 16 + bucket;
        all.push_back(std::move(t));
    }
    std::stable_sort(all.begin(), all.end(), [](const PairTmp &a, const PairTmp &b) {
        return a.key != b.key ? a.key < b.key : a.q > b.q;
    });
    std::vector<int2> pairs;
    std::vector<double> qs;
    std::vector<JKPair> jkpairs;
    std::vector<JKPrim> prims;
    std::vector<int> coff, cla, clb, cbucket;
    for (size_t e = 0; e < all.size(); e++) {
        const PairTmp &t = all[e];
        if (e == 0 || t.key != all[e - 1].key) {
            coff.push_back((int)e);
            cla.push_back(t.key / 128);
            clb.push_back((t.key / 16) % 8);
            cbucket.push_back(t.key % 16);
        }
        const ShellRec &si = basis->h_shells[t.p.x], &sj = basis->h_shells[t.p.y];
        JKPair r;
        r.ax = si.x; r.ay = si.y; r.az = si.z;
        r.abx = si.x - sj.x; r.aby = si.y - sj.y; r.abz = si.z - sj.z;
        r.ish = t.p.x; r.jsh = t.p.y;
        r.ao_i = si.ao_off - plan->ao0; r.ao_j = sj.ao_off - plan->ao0;
        r.npp = (int)t.prims.size();
        r.pp_off = (int)prims.size();
        prims.insert(prims.end(), t.prims.begin(), t.prims.end());
        jkpairs.push_back(r);
        pairs.push_back(t.p);
        qs.push_back(t.q);
    }
    coff.push_back((int)pairs.size());
    if (pairs.empty()) {
        *out = plan;
        return 0;
    }
    if (prims.empty()) prims.push_back(JKPrim{1.0, 0.5, 0.0, 0.0, 0.0, 0.0});
    QC_CHECK(cudaMalloc(&plan->d_pairs, sizeof(int2) * pairs.size()));
    QC_CHECK(cudaMemcpy(plan->d_pairs, pairs.data(), sizeof(int2) * pairs.size(), cudaMemcpyHostToDevice));
    QC_CHECK(cudaMalloc(&plan->d_jkpairs, sizeof(JKPair) * jkpairs.size()));
    QC_CHECK(cudaMemcpy(plan->d_jkpairs, jkpairs.data(), sizeof(JKPair) * jkpairs.size(), cudaMemcpyHostToDevice));
    QC_CHECK(cudaMalloc(&plan->d_scratch_j, sizeof(double) * (size_t)plan->nao * plan->nao));
    QC_CHECK(cudaMemset(plan->d_scratch_j, 0, sizeof(double) * (size_t)plan->nao * plan->nao));
    QC_CHECK(cudaMalloc(&plan->d_prims, sizeof(JKPrim) * prims.size()));
    QC_CHECK(cudaMemcpy(plan->d_prims, prims.data(), sizeof(JKPrim) * prims.size(), cudaMemcpyHostToDevice));
    // 3. class pairs (bra class >= ket class) with their work items
    const RysTable &rt = g_rys_fine[basis->device];   // the register engine reads the refined root table
    const size_t ncls = cla.size();
    for (size_t cb = 0; cb < ncls; cb++)
        for (size_t ck = 0; ck <= cb; ck++) {
            JKClassPair cp;
            const int l[4] = {cla[cb], clb[cb], cla[ck], clb[ck]};
            const int pr[4] = {1, 1, 1, 1};
            cp.reg = cbucket[cb] < JKR_NBUCKET && cbucket[ck] < JKR_NBUCKET && jkr_supported(l) &&
                     rt.nint == 128 && rt.deg == 9;
            if (int_make_class(cp.K, INT_MODE_ERI, SINK_JK, l, pr, 0)) {
                b200qc_jkplan_free(plan);
                b200qc_set_error("J/K class outside the supported range");
                return 2;
            }
            cp.bra_off = coff[cb]; cp.nbra = coff[cb + 1] - coff[cb];
            cp.ket_off = coff[ck]; cp.nket = coff[ck + 1] - coff[ck];
            cp.same = cb == ck;
            const int chunk = cp.reg ? JKR_CHUNK : JK_CHUNK;
            std::vector<int64_t> woff(cp.nbra + 1, 0);
            std::vector<int> nkb(cp.nbra, 0);
            std::vector<int2> items;
            const double *qk = qs.data() + cp.ket_off;
            int64_t nq = 0;
            for (int b = 0; b < cp.nbra; b++) {
                const double qb = qs[cp.bra_off + b];
                // kets sorted descending: count those with qb * qk >= thresh
                int lo = 0, hi = cp.nket;
                while (lo < hi) {
                    const int mid = (lo + hi) / 2;
                    if (qb * qk[mid] >= thresh) lo = mid + 1; else hi = mid;
                }
                int n = lo;
                if (cp.same) n = std::min(n, b + 1);
                nkb[b] = n;
                woff[b + 1] = woff[b] + (n + chunk - 1) / chunk;
                if (cp.reg)
                    for (int k0 = 0; k0 < n; k0 += chunk) items.push_back(make_int2(b, k0));
                nq += n;
            }
            plan->nquartets += nq;
            if (cp.reg) plan->nquartets_reg += nq;
            {   // operation counts: sum_b npp_b * (sum of npp over its admissible kets)
                std::vector<double> pre(cp.nket + 1, 0.0);
                for (int k = 0; k < cp.nket; k++) pre[k + 1] = pre[k] + jkpairs[cp.ket_off + k].npp;
                double nprimq = 0.0;
                for (int b = 0; b < cp.nbra; b++) nprimq += (double)jkpairs[cp.bra_off + b].npp * pre[nkb[b]];
                plan->flops_int += nprimq * jk_prim_flops(l);
                plan->ncomp_total += (double)nq * NCART(l[0]) * NCART(l[1]) * NCART(l[2]) * NCART(l[3]);
            }
            cp.nitems = woff[cp.nbra];
            if (cp.nitems == 0) continue;
            QC_CHECK(cudaMalloc(&cp.d_nket_of_bra, sizeof(int) * nkb.size()));
            QC_CHECK(cudaMemcpy(cp.d_nket_of_bra, nkb.data(), sizeof(int) * nkb.size(), cudaMemcpyHostToDevice));
            if (cp.reg) {
                QC_CHECK(cudaMalloc(&cp.d_items, sizeof(int2) * items.size()));
                QC_CHECK(cudaMemcpy(cp.d_items, items.data(), sizeof(int2) * items.size(), cudaMemcpyHostToDevice));
            } else {
                QC_CHECK(cudaMalloc(&cp.d_work_off, sizeof(int64_t) * woff.size()));
                QC_CHECK(cudaMemcpy(cp.d_work_off, woff.data(), sizeof(int64_t) * woff.size(), cudaMemcpyHostToDevice));
            }
            plan->cps.push_back(cp);
        }
    *out = plan;
    return 0;
}

extern "C" int64_t b200qc_jkplan_nquartets(const b200qc_jkplan *p) { return p ? p->nquartets : 0; }
// fp64 operations of one build (integrals once per quartet + digestion: 2 tile contractions for J, 4 for K)
extern "C" double b200qc_jkplan_flops(const b200qc_jkplan *p, int with_j, int with_k) {
    return p ? p->flops_int + 2.0 * p->ncomp_total * (2.0 * (with_j != 0) + 4.0 * (with_k != 0)) : 0.0;
}
// quartets that go through the register-resident engine (jk_reg.cuh); the rest use the shared-memory engine
extern "C" int64_t b200qc_jkplan_nquartets_reg(const b200qc_jkplan *p) { return p ? p->nquartets_reg : 0; }

// dm: (nset, nao, nao) symmetric; vj / vk: (nset, nao, nao) or NULL.  rank / world: this process
// digests work items rank, rank + world, ... (partial J / K; the caller all-reduces).
extern "C" int b200qc_jkplan_run(const b200qc_jkplan *plan, const double *dm, int nset, double *vj, double *vk,
                                 int rank, int world, void *stream) {
    QC_REQUIRE(plan && dm && nset >= 1 && nset <= JK_MAXSET, "bad arguments (nset must be 1 or 2)");
    QC_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
    cudaStream_t st = as_stream(stream);
    const int64_t nn = (int64_t)plan->nao * plan->nao;
    if (vj) QC_CHECK(cudaMemsetAsync(vj, 0, sizeof(double) * nn * nset, st));
    if (vk) QC_CHECK(cudaMemsetAsync(vk, 0, sizeof(double) * nn * nset, st));
    const RysTable &rt = g_rys_host[plan->basis->device];
    // The class-pair kernels are independent (they only add into J / K): they go round-robin onto JK_NSTREAM side
    // streams forked from the caller's stream, so the tail of one persistent grid overlaps the start of the next
    // (hundreds of launches per build; sharded builds shrink every one of them).  The profiler then brackets the whole
    // region once on the caller's stream instead of every launch.
    const int dev = plan->basis->device;
    const bool multi = !getenv("B200QC_JK_ONE_STREAM");
    cudaStream_t js[JK_NSTREAM];
    if (multi) {
        if (!g_jk_streams_ready[dev]) {
            for (int i = 0; i < JK_NSTREAM; i++) {
                QC_CHECK(cudaStreamCreateWithFlags(&g_jk_streams[dev][i], cudaStreamNonBlocking));
                QC_CHECK(cudaEventCreateWithFlags(&g_jk_join[dev][i], cudaEventDisableTiming));
            }
            QC_CHECK(cudaEventCreateWithFlags(&g_jk_fork[dev], cudaEventDisableTiming));
            g_jk_streams_ready[dev] = true;
        }
        prof_begin(PROF_JK, st);
        QC_CHECK(cudaEventRecord(g_jk_fork[dev], st));
        for (int i = 0; i < JK_NSTREAM; i++) {
            js[i] = g_jk_streams[dev][i];
            QC_CHECK(cudaStreamWaitEvent(js[i], g_jk_fork[dev], 0));
        }
    } else {
        for (int i = 0; i < JK_NSTREAM; i++) js[i] = st;
    }
    const bool prof_was_on = g_prof_on;
    if (multi) g_prof_on = false;
    int rc_all = 0;
    size_t icp = 0;
    for (const JKClassPair &cp : plan->cps) {
        cudaStream_t st = js[icp++ % JK_NSTREAM];   // shadows the caller's stream inside the loop
        if (cp.reg) {
            JKRArgs R = {};
            for (int s = 0; s < 4; s++) R.l[s] = cp.K.l[s];
            R.bra = plan->d_jkpairs + cp.bra_off;
            R.ket = plan->d_jkpairs + cp.ket_off;
            R.prims = plan->d_prims;
            R.items = cp.d_items;
            R.nket_of_bra = cp.d_nket_of_bra;
            R.nitems = cp.nitems;
            R.item0 = rank; R.item_stride = world;
            R.nao = plan->nao;
            const int nr = cp.K.nroots;
            const RysTable &rf = g_rys_fine[plan->basis->device];
            R.rys_coef = rf.coef[nr - 1];
            R.rys_nint = rf.nint; R.rys_deg = rf.deg;
            R.rys_h = rf.h; R.rys_xmax = rt.xmax;
            for (int r = 0; r < nr; r++) {
                R.herm_u[r] = rt.herm[nr - 1][0][r];
                R.herm_w[r] = rt.herm[nr - 1][1][r];
            }
            h_fill_c2s(2, R.c2s_d);
            for (int s = 0; s < nset; s++) {   // one density per launch: the block stays in registers
                R.dm = dm + s * nn;
                R.vj = vj ? vj + s * nn : plan->d_scratch_j;   // K only: the J+K kernels with a scratch J
                R.vk = vk ? vk + s * nn : nullptr;
                rc_all = jkr_launch(R, st);
                if (rc_all) break;
            }
            if (rc_all) break;
            continue;
        }
        IntArgs A = {};
        A.shells = plan->basis->d_shells;
        A.env = plan->basis->d_env;
        A.bra = plan->d_pairs + cp.bra_off;
        A.nbra = cp.nbra;
        A.ket = plan->d_pairs + cp.ket_off;
        A.nket = cp.nket;
        A.off[0] = plan->ao0;
        A.dm = dm; A.vj = vj; A.vk = vk;
        A.nset = nset; A.nao = plan->nao;
        int rc;
        const int nc = cp.K.ncomp;
        if (nc <= 16) rc = jk_launch_t<8, 2>(cp, A, rank, world, st);
        else if (nc <= 64) rc = jk_launch_t<8, 8>(cp, A, rank, world, st);
        else if (nc <= 256) rc = jk_launch_t<16, 16>(cp, A, rank, world, st);
        else if (nc <= 512) rc = jk_launch_t<32, 16>(cp, A, rank, world, st);
        else rc = jk_launch_t<32, 41>(cp, A, rank, world, st);
        if (rc) { rc_all = rc; break; }
    }
    if (multi) {
        g_prof_on = prof_was_on;
        for (int i = 0; i < JK_NSTREAM; i++) {
            QC_CHECK(cudaEventRecord(g_jk_join[dev][i], js[i]));
            QC_CHECK(cudaStreamWaitEvent(st, g_jk_join[dev][i], 0));
        }
        prof_end(st);
    }
    if (rc_all) return rc_all;
    const int64_t tot = nn * nset;
    if (vj) { jk_symm_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(vj, plan->nao, nset); QC_LAUNCHED(1); }
    if (vk) { jk_symm_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(vk, plan->nao, nset); QC_LAUNCHED(1); }
    return 0;
}

extern "C" int b200qc_jk_direct(const b200qc_basis *basis, int sh0, int sh1, const double *dm, int nset, double *vj,
                                double *vk, void *stream) {
    b200qc_jkplan *plan = nullptr;
    int rc = b200qc_jkplan_create(basis, sh0, sh1, 1e-13, &plan, stream);
    if (rc) return rc;
    rc = b200qc_jkplan_run(plan, dm, nset, vj, vk, 0, 1, stream);
    cudaStreamSynchronize(as_stream(stream));
    b200qc_jkplan_free(plan);
    return rc;
}
