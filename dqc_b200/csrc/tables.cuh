// Host-side construction of the constant tables + basis upload + error plumbing.
#pragma once
#include "common.cuh"
#include <cmath>
#include <mutex>

static thread_local std::string g_last_error;
void b200qc_set_error(const std::string &msg) { g_last_error = msg; }

extern "C" const char *b200qc_last_error(void) { return g_last_error.c_str(); }
extern "C" int b200qc_version(void) { return 100; }
extern "C" int64_t b200qc_launch_count(void) { return g_launch_count; }

extern "C" int b200qc_profile(int on) {
    for (ProfRec &r : g_prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    g_prof_on = on != 0;
    return 0;
}
extern "C" int b200qc_profile_nkernels(void) { return PROF_N; }
extern "C" const char *b200qc_profile_name(int id) { return (id >= 0 && id < PROF_N) ? g_prof_names[id] : ""; }
// ms_total / counts: PROF_N entries each; synchronises the device, then clears the records
extern "C" int b200qc_profile_read(double *ms_total, int64_t *counts) {
    QC_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < PROF_N; i++) {
        ms_total[i] = 0.0;
        counts[i] = 0;
    }
    for (ProfRec &r : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            ms_total[r.id] += ms;
            counts[r.id]++;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    return 0;
}

namespace {
double h_binom(int n, int k) {
    if (k < 0 || k > n) return 0.0;
    double r = 1.0;
    for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}
double h_fact(int n) {
    double r = 1.0;
    for (int i = 2; i <= n; i++) r *= i;
    return r;
}
// libcint cartesian order: lx descending, then ly descending
int h_cart_index(int l, int lx, int ly) {
    int idx = 0;
    for (int x = l; x > lx; x--) idx += l - x + 1;
    return idx + (l - lx - ly);
}
// Real solid harmonics r^l Y_lm as polynomials in x, y, z (Helgaker/Jorgensen/Olsen eq. 6.4.47),
// m = -l..l; s and p follow libcint's special cases (p ordered x, y, z).
void h_fill_c2s(int l, double *M) {
    int nc = NCART(l), ns = 2 * l + 1;
    for (int i = 0; i < ns * nc; i++) M[i] = 0.0;
    if (l == 0) {
        M[0] = 0.282094791773878143;
        return;
    }
    if (l == 1) {
        M[0] = M[4] = M[8] = 0.488602511902919921;
        return;
    }
    const double pi = 3.14159265358979323846;
    double pref = std::sqrt((2 * l + 1) / (4 * pi));
    for (int m = -l; m <= l; m++) {
        int am = m < 0 ? -m : m;
        double N = std::sqrt(2.0 * h_fact(l + am) * h_fact(l - am) / (m == 0 ? 2.0 : 1.0)) /
                   (std::pow(2.0, am) * h_fact(l));
        int vm2 = m < 0 ? 1 : 0;
        for (int t = 0; t <= (l - am) / 2; t++)
            for (int u = 0; u <= t; u++)
                for (int v2 = vm2; v2 <= am; v2 += 2) {
                    int sp = t + (v2 - vm2) / 2;
                    double C = ((sp & 1) ? -1.0 : 1.0) * std::pow(0.25, t) * h_binom(l, t) *
                               h_binom(l - t, am + t) * h_binom(t, u) * h_binom(am, v2);
                    int ex = 2 * t + am - 2 * u - v2, ey = 2 * u + v2;
                    if (ex < 0) continue;
                    M[(m + l) * nc + h_cart_index(l, ex, ey)] += pref * N * C;
                }
    }
}
// __constant__ symbols live once PER DEVICE: the tables are uploaded on first use on every device a basis is
// created on (a second Hamiltonian on another GPU of the same process would otherwise read zero-filled tables).
#define QC_MAX_DEVICES 64
std::mutex g_tables_mutex;
bool g_tables_ready[QC_MAX_DEVICES] = {};
}  // namespace

static int qc_current_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= QC_MAX_DEVICES) return -1;
    return dev;
}

static int qc_init_tables() {
    const int dev = qc_current_device();
    if (dev < 0) {
        b200qc_set_error("no current CUDA device");
        return 1;
    }
    std::lock_guard<std::mutex> lock(g_tables_mutex);
    if (g_tables_ready[dev]) return 0;
    int status = 0;
    {
        C2STables t;
        h_fill_c2s(0, t.s0);
        h_fill_c2s(1, t.s1);
        h_fill_c2s(2, t.s2);
        h_fill_c2s(3, t.s3);
        h_fill_c2s(4, t.s4);
        signed char pw[B200QC_LMAX + 1][15][3] = {};
        for (int l = 0; l <= B200QC_LMAX; l++) {
            int k = 0;
            for (int x = l; x >= 0; x--)
                for (int y = l - x; y >= 0; y--) {
                    pw[l][k][0] = (signed char)x;
                    pw[l][k][1] = (signed char)y;
                    pw[l][k][2] = (signed char)(l - x - y);
                    k++;
                }
        }
        cudaError_t e = cudaMemcpyToSymbol(c_c2s, &t, sizeof(t));
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_cart_pow, pw, sizeof(pw));
        if (e != cudaSuccess) {
            b200qc_set_error(std::string("constant table upload failed: ") + cudaGetErrorString(e));
            status = 1;
        }
    }
    if (status == 0) g_tables_ready[dev] = true;
    return status;
}

// Kernels that read the per-device constant tables refuse a basis that lives on another device.
static int qc_require_basis_device(const b200qc_basis *b) {
    QC_REQUIRE(b != nullptr, "null basis");
    QC_REQUIRE(qc_current_device() == b->device,
               "the current CUDA device is not the one this basis was uploaded on (wrap the call in torch.cuda.device)");
    return 0;
}

extern "C" int b200qc_basis_upload(const int *h_atm, int natm, const int *h_bas, int nbas,
                                   const double *h_env, int nenv, const int *h_ao_loc,
                                   b200qc_basis **out) {
    QC_REQUIRE(out != nullptr && natm > 0 && nbas > 0, "bad arguments");
    if (qc_init_tables()) return 1;
    auto *b = new b200qc_basis();
    b->device = qc_current_device();
    b->natm = natm;
    b->nbas = nbas;
    b->nenv = nenv;
    b->h_atm.assign(h_atm, h_atm + natm * ATM_SLOTS);
    b->h_bas.assign(h_bas, h_bas + nbas * BAS_SLOTS);
    b->h_env.assign(h_env, h_env + nenv);
    b->h_ao_loc.assign(h_ao_loc, h_ao_loc + nbas + 1);
    b->h_shells.resize(nbas);
    for (int i = 0; i < nbas; i++) {
        const int *bs = h_bas + i * BAS_SLOTS;
        ShellRec &s = b->h_shells[i];
        int pc = h_atm[bs[0] * ATM_SLOTS + 1];
        s.x = h_env[pc];
        s.y = h_env[pc + 1];
        s.z = h_env[pc + 2];
        s.l = bs[1];
        s.nprim = bs[2];
        s.ptr_exp = bs[5];
        s.ptr_coef = bs[6];
        s.ao_off = h_ao_loc[i];
        s.atom = bs[0];
        if (s.l > B200QC_LMAX) {
            delete b;
            b200qc_set_error("angular momentum above g is not supported");
            return 2;
        }
        if (bs[3] != 1) {
            delete b;
            b200qc_set_error("only nctr = 1 shells (the reference never emits anything else)");
            return 2;
        }
    }
    QC_CHECK(cudaMalloc(&b->d_atm, sizeof(int) * natm * ATM_SLOTS));
    QC_CHECK(cudaMalloc(&b->d_bas, sizeof(int) * nbas * BAS_SLOTS));
    QC_CHECK(cudaMalloc(&b->d_ao_loc, sizeof(int) * (nbas + 1)));
    QC_CHECK(cudaMalloc(&b->d_env, sizeof(double) * nenv));
    QC_CHECK(cudaMalloc(&b->d_shells, sizeof(ShellRec) * nbas));
    QC_CHECK(cudaMemcpy(b->d_atm, h_atm, sizeof(int) * natm * ATM_SLOTS, cudaMemcpyHostToDevice));
    QC_CHECK(cudaMemcpy(b->d_bas, h_bas, sizeof(int) * nbas * BAS_SLOTS, cudaMemcpyHostToDevice));
    QC_CHECK(cudaMemcpy(b->d_ao_loc, h_ao_loc, sizeof(int) * (nbas + 1), cudaMemcpyHostToDevice));
    QC_CHECK(cudaMemcpy(b->d_env, h_env, sizeof(double) * nenv, cudaMemcpyHostToDevice));
    QC_CHECK(cudaMemcpy(b->d_shells, b->h_shells.data(), sizeof(ShellRec) * nbas, cudaMemcpyHostToDevice));
    *out = b;
    return 0;
}

extern "C" int b200qc_basis_free(b200qc_basis *b) {
    if (!b) return 0;
    cudaFree(b->d_atm);
    cudaFree(b->d_bas);
    cudaFree(b->d_ao_loc);
    cudaFree(b->d_env);
    cudaFree(b->d_shells);
    delete b;
    return 0;
}

// Finer form of the root table for the register-resident J/K engine (jk_reg.cuh): every interval of the base table is
// split into `nsub` and re-expanded in Chebyshev polynomials truncated to `ncoef` coefficients.  A degree-13 expansion
// on intervals of width 1 and a degree-9 expansion on intervals of width 1/2 have the same truncation error (3e-16 ..
// 6e-16 relative: the base table's own), but the engine's Clenshaw sums -- more than half of the work of an (ss|ss) ..
// (pp|ps) primitive quartet -- are 10 steps instead of 14.  Pure host arithmetic: in (nint, nf, deg + 1), out
// (nint * nsub, nf, ncoef).
extern "C" int b200qc_rys_refine(const double *h_coef, int nint, int nf, int deg, int nsub, int ncoef, double *h_out) {
    QC_REQUIRE(h_coef && h_out && nint > 0 && nf > 0 && deg >= 1 && deg < 32 && nsub >= 1 && ncoef >= 1 && ncoef <= deg + 1,
               "bad arguments");
    const int N = deg + 1;
    const double PI = 3.14159265358979323846;
    std::vector<double> node(N), val(N);
    for (int k = 0; k < N; k++) node[k] = std::cos(PI * (k + 0.5) / N);
    for (int it = 0; it < nint; it++)
        for (int f = 0; f < nf; f++) {
            const double *c = h_coef + ((size_t)it * nf + f) * N;
            for (int sub = 0; sub < nsub; sub++) {
                const double a = -1.0 + 2.0 * sub / nsub, b = a + 2.0 / nsub;
                for (int k = 0; k < N; k++) {   // base polynomial at the Chebyshev nodes of the sub-interval (Clenshaw)
                    const double t = a + (node[k] + 1.0) * (b - a) * 0.5;
                    double b1 = 0.0, b2 = 0.0;
                    for (int j = N - 1; j >= 1; j--) {
                        const double b0 = c[j] + 2.0 * t * b1 - b2;
                        b2 = b1;
                        b1 = b0;
                    }
                    val[k] = c[0] + t * b1 - b2;
                }
                double *o = h_out + (((size_t)it * nsub + sub) * nf + f) * ncoef;
                for (int m = 0; m < ncoef; m++) {   // discrete Chebyshev transform (exact for degree <= deg)
                    double sacc = 0.0;
                    for (int k = 0; k < N; k++) sacc += val[k] * std::cos(PI * m * (k + 0.5) / N);
                    o[m] = sacc * (m == 0 ? 1.0 : 2.0) / N;
                }
            }
        }
    return 0;
}
#define QC_RYS_FINE_NSUB 2
#define QC_RYS_FINE_NCOEF 10

static std::vector<double *> g_rys_dev[QC_MAX_DEVICES];
static RysTable g_rys_fine[QC_MAX_DEVICES];   // the refined table (device pointers; h, nint, deg of the fine form)
static RysTable g_rys_host[QC_MAX_DEVICES];   // host copy of each device's table descriptor (jk.cuh hands it to jk_reg.cuh)   // per device, like the __constant__ RysTable that points at them

extern "C" int b200qc_rys_upload(int nmax, double h, int deg, double xmax,
                                 const double *const *h_coef, const double *const *h_herm) {
    QC_REQUIRE(nmax >= 1 && nmax <= RYS_NMAX, "nmax out of range");
    const int dev = qc_current_device();
    QC_REQUIRE(dev >= 0, "no current CUDA device");
    for (double *p : g_rys_dev[dev]) cudaFree(p);
    g_rys_dev[dev].clear();
    g_rys_ready[dev] = false;
    RysTable t = {};
    t.nmax = nmax;
    t.deg = deg;
    t.h = h;
    t.xmax = xmax;
    t.nint = (int)std::llround(xmax / h);
    for (int n = 1; n <= nmax; n++) {
        size_t cnt = (size_t)t.nint * 2 * n * (deg + 1);
        double *d = nullptr;
        QC_CHECK(cudaMalloc(&d, cnt * sizeof(double)));
        QC_CHECK(cudaMemcpy(d, h_coef[n - 1], cnt * sizeof(double), cudaMemcpyHostToDevice));
        g_rys_dev[dev].push_back(d);
        t.coef[n - 1] = d;
        for (int r = 0; r < n; r++) {
            t.herm[n - 1][0][r] = h_herm[n - 1][r];
            t.herm[n - 1][1][r] = h_herm[n - 1][n + r];
        }
    }
    QC_CHECK(cudaMemcpyToSymbol(c_rys, &t, sizeof(t)));
    g_rys_host[dev] = t;
    // the refined form for jk_reg.cuh
    RysTable tf = t;
    tf.h = h / QC_RYS_FINE_NSUB;
    tf.deg = QC_RYS_FINE_NCOEF - 1;
    tf.nint = t.nint * QC_RYS_FINE_NSUB;
    for (int n = 1; n <= nmax; n++) {
        std::vector<double> fine((size_t)tf.nint * 2 * n * QC_RYS_FINE_NCOEF);
        if (b200qc_rys_refine(h_coef[n - 1], t.nint, 2 * n, deg, QC_RYS_FINE_NSUB, QC_RYS_FINE_NCOEF, fine.data())) return 2;
        double *d = nullptr;
        QC_CHECK(cudaMalloc(&d, fine.size() * sizeof(double)));
        QC_CHECK(cudaMemcpy(d, fine.data(), fine.size() * sizeof(double), cudaMemcpyHostToDevice));
        g_rys_dev[dev].push_back(d);
        tf.coef[n - 1] = d;
    }
    g_rys_fine[dev] = tf;
    g_rys_ready[dev] = true;
    return 0;
}
