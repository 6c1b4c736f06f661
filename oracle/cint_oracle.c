/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (dqc_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * CPU restatement of the integral / AO-evaluation arithmetic that the reference obtains from its
 * un-vendored native dependency `dqclibs` (libcint + PySCF libcgto; pinned only as
 * "dqclibs>=0.1.0", reference setup.py:54-63).  The reference's call sites this file restates:
 *   - GTOint2c   with int1e_{ovlp,kin,nuc,rinv}_sph, int2c2e_sph   dqc/hamilton/intor/molintor.py:624-644
 *   - GTOnr3c_drv with int3c2e_sph (fill s1)                        dqc/hamilton/intor/molintor.py:646-665
 *   - GTOnr2e_fill_drv with int2e_sph                               dqc/hamilton/intor/molintor.py:667-688
 *   - GTOval_sph / GTOval_ip_sph                                    dqc/hamilton/intor/gtoeval.py:196-239
 * on the libcint (atm, bas, env) layout built by dqc/hamilton/intor/lcintwrap.py:36-117.
 *
 * Algorithm: McMurchie-Davidson (Hermite Gaussians + Boys function) -- deliberately a different
 * algorithm from the Rys-quadrature CUDA kernels it checks.  libcint conventions restated from the
 * published library behaviour: contracted coefficients in env already carry the radial
 * normalisation (dqc/utils/datastruct.py:34-61); s and p shells carry the extra constants
 * 1/sqrt(4 pi) and sqrt(3/(4 pi)); l >= 2 use real solid harmonics r^l Y_lm ordered m = -l..l,
 * p is ordered (x, y, z); cartesians are ordered lexicographically (xx, xy, xz, yy, yz, zz ...).
 *
 * Output axis order is the one the reference hands to its callers AFTER its swapaxes:
 *   2-centre out[i][j]; 3-centre out[i][j][P]; 4-centre out[i][j][k][l] = (ij|kl).
 *
 * PARITY PIN: integral values are pinned through the reference's golden RHF/3-21G energies
 * (dqc/test/test_hf.py:18-51) and the H2 density points (dqc/test/test_hamilton.py:115-142);
 * individual integrals have no in-tree golden numbers (they are compared to live PySCF there).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ATM_SLOTS 6
#define BAS_SLOTS 8
#define PTR_RINV_ORIG 4
#define LMAX 6
#define NCART(l) (((l) + 1) * ((l) + 2) / 2)

typedef struct {
    int l, nprim;
    const double *ex, *co, *R;
} Shell;

static void get_shell(Shell *s, int ish, const int *atm, const int *bas, const double *env) {
    const int *b = bas + ish * BAS_SLOTS;
    s->l = b[1];
    s->nprim = b[2];
    s->ex = env + b[5];
    s->co = env + b[6];
    s->R = env + atm[b[0] * ATM_SLOTS + 1];
}

/* ------------------------------------------------------------------ Boys function F_n(x) */
static void boys(int nmax, double x, double *F) {
    if (x < 35.0) {
        /* convergent series for the top order, then stable downward recursion */
        double ex = exp(-x);
        double term = 1.0 / (2 * nmax + 1), sum = term;
        for (int k = 1; k < 400; k++) {
            term *= 2 * x / (2 * nmax + 2 * k + 1);
            sum += term;
            if (term < 1e-17 * sum) break;
        }
        F[nmax] = ex * sum;
        for (int n = nmax; n > 0; n--) F[n - 1] = (2 * x * F[n] + ex) / (2 * n - 1);
    } else {
        /* asymptotic F_0 and upward recursion (stable for large x) */
        double ex = exp(-x);
        F[0] = 0.5 * sqrt(M_PI / x) * erf(sqrt(x));
        for (int n = 0; n < nmax; n++) F[n + 1] = ((2 * n + 1) * F[n] - ex) / (2 * x);
    }
}

/* --------------------------------------------- Hermite expansion coefficients (1 dimension)
 * E[i][j][t], i<=la, j<=lb, t<=i+j, WITHOUT the exp(-mu XAB^2) factor.  b may be 0 (one-centre). */
static void hermite_E(int la, int lb, double a, double b, double XAB, double *E) {
    int nt = la + lb + 1;
    double p = a + b;
    double XPA = -b / p * XAB, XPB = a / p * XAB, o2p = 0.5 / p;
#define EE(i, j, t) E[((i) * (lb + 1) + (j)) * nt + (t)]
    memset(E, 0, sizeof(double) * (la + 1) * (lb + 1) * nt);
    EE(0, 0, 0) = 1.0;
    for (int i = 0; i < la; i++)
        for (int t = 0; t <= i + 1; t++) {
            double v = XPA * (t <= i ? EE(i, 0, t) : 0.0);
            if (t > 0) v += o2p * EE(i, 0, t - 1);
            if (t + 1 <= i) v += (t + 1) * EE(i, 0, t + 1);
            EE(i + 1, 0, t) = v;
        }
    for (int i = 0; i <= la; i++)
        for (int j = 0; j < lb; j++)
            for (int t = 0; t <= i + j + 1; t++) {
                double v = XPB * (t <= i + j ? EE(i, j, t) : 0.0);
                if (t > 0) v += o2p * EE(i, j, t - 1);
                if (t + 1 <= i + j) v += (t + 1) * EE(i, j, t + 1);
                EE(i, j + 1, t) = v;
            }
#undef EE
}

/* --------------------------------------------- Hermite Coulomb integrals R_{tuv}, t+u+v <= L
 * R[(t*(L+1)+u)*(L+1)+v]; work must hold (L+1)^4 doubles. */
static void hermite_R(int L, double alpha, const double *PC, double *R, double *work) {
    int n1 = L + 1;
    double F[4 * LMAX + 2];
    double x = alpha * (PC[0] * PC[0] + PC[1] * PC[1] + PC[2] * PC[2]);
    boys(L, x, F);
#define W(n, t, u, v) work[(((n) * n1 + (t)) * n1 + (u)) * n1 + (v)]
    double m2a = 1.0;
    for (int n = 0; n <= L; n++) {
        W(n, 0, 0, 0) = m2a * F[n];
        m2a *= -2.0 * alpha;
    }
    for (int tot = 1; tot <= L; tot++)
        for (int t = 0; t <= tot; t++)
            for (int u = 0; u <= tot - t; u++) {
                int v = tot - t - u;
                for (int n = 0; n <= L - tot; n++) {
                    double val;
                    if (t > 0) {
                        val = PC[0] * W(n + 1, t - 1, u, v);
                        if (t > 1) val += (t - 1) * W(n + 1, t - 2, u, v);
                    } else if (u > 0) {
                        val = PC[1] * W(n + 1, t, u - 1, v);
                        if (u > 1) val += (u - 1) * W(n + 1, t, u - 2, v);
                    } else {
                        val = PC[2] * W(n + 1, t, u, v - 1);
                        if (v > 1) val += (v - 1) * W(n + 1, t, u, v - 2);
                    }
                    W(n, t, u, v) = val;
                }
            }
    for (int t = 0; t <= L; t++)
        for (int u = 0; u <= L - t; u++)
            for (int v = 0; v <= L - t - u; v++) R[(t * n1 + u) * n1 + v] = W(0, t, u, v);
#undef W
}

/* --------------------------------------------- cartesian -> real spherical (libcint convention)
 * c2s[l] is (2l+1) x ncart, row-major; built once by orc_init() from the closed formula for real
 * solid harmonics (Helgaker, Jorgensen, Olsen, eq. 6.4.47) times sqrt((2l+1)/(4 pi)); p kept as
 * (x,y,z). */
static double *c2s_tab[LMAX + 1];
static int c2s_ready = 0;

static double binom(int n, int k) {
    if (k < 0 || k > n) return 0.0;
    double r = 1.0;
    for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}
static double fact(int n) {
    double r = 1.0;
    for (int i = 2; i <= n; i++) r *= i;
    return r;
}
static int cart_index(int l, int lx, int ly) {
    /* position of (lx, ly, lz) in libcint's ordering: lx descending, then ly descending */
    int idx = 0;
    for (int x = l; x > lx; x--) idx += l - x + 1;
    return idx + (l - lx - ly);
}

void orc_init(void) {
    if (c2s_ready) return;
    for (int l = 0; l <= LMAX; l++) {
        int nc = NCART(l), ns = 2 * l + 1;
        double *M = (double *)calloc((size_t)ns * nc, sizeof(double));
        if (l == 0) {
            M[0] = 0.282094791773878143;
        } else if (l == 1) {
            M[0 * 3 + 0] = M[1 * 3 + 1] = M[2 * 3 + 2] = 0.488602511902919921;
        } else {
            double pref = sqrt((2 * l + 1) / (4 * M_PI));
            for (int m = -l; m <= l; m++) {
                int am = abs(m);
                double N = sqrt(2.0 * fact(l + am) * fact(l - am) / (m == 0 ? 2.0 : 1.0)) /
                           (pow(2.0, am) * fact(l));
                int vm2 = (m < 0) ? 1 : 0; /* 2*v_m */
                for (int t = 0; t <= (l - am) / 2; t++)
                    for (int u = 0; u <= t; u++)
                        for (int v2 = vm2; v2 <= am; v2 += 2) { /* v2 = 2v */
                            /* C = (-1)^(t + v - v_m) (1/4)^t C(l,t) C(l-t,|m|+t) C(t,u) C(|m|,2v) */
                            int sgnpow = t + (v2 - vm2) / 2;
                            double C = ((sgnpow & 1) ? -1.0 : 1.0) * pow(0.25, t) * binom(l, t) *
                                       binom(l - t, am + t) * binom(t, u) * binom(am, v2);
                            int ex = 2 * t + am - 2 * u - v2;
                            int ey = 2 * u + v2;
                            if (ex < 0) continue;
                            M[(m + l) * nc + cart_index(l, ex, ey)] += pref * N * C;
                        }
            }
        }
        c2s_tab[l] = M;
    }
    c2s_ready = 1;
}

const double *orc_c2s(int l) {
    orc_init();
    return c2s_tab[l];
}

static void cart_powers(int l, int *lx, int *ly, int *lz) {
    int k = 0;
    for (int x = l; x >= 0; x--)
        for (int y = l - x; y >= 0; y--) {
            lx[k] = x;
            ly[k] = y;
            lz[k] = l - x - y;
            k++;
        }
}

/* transform index `axis` (of length ncart(l)) of a dense tensor to spherical.
 * in: [pre][ncart][post] -> out: [pre][nsph][post] */
static void c2s_axis(const double *in, double *out, int pre, int l, int post) {
    int nc = NCART(l), ns = 2 * l + 1;
    const double *M = c2s_tab[l];
    for (int a = 0; a < pre; a++)
        for (int m = 0; m < ns; m++) {
            double *o = out + ((size_t)a * ns + m) * post;
            for (int q = 0; q < post; q++) o[q] = 0.0;
            for (int c = 0; c < nc; c++) {
                double w = M[m * nc + c];
                if (w == 0.0) continue;
                const double *src = in + ((size_t)a * nc + c) * post;
                for (int q = 0; q < post; q++) o[q] += w * src[q];
            }
        }
}

/* ------------------------------------------------------------------ one-electron integrals */
/* kind: 0 overlap, 1 kinetic, 2 nuclear attraction (sum_A -Z_A/|r-R_A|), 3 rinv (1/|r-env[4:7]|) */
static void shellpair_1e(int kind, const Shell *A, const Shell *B, const int *atm, int natm,
                         const double *env, double *cart /* ncartA*ncartB */) {
    int la = A->l, lb = B->l;
    int nca = NCART(la), ncb = NCART(lb);
    int ax[NCART(LMAX)], ay[NCART(LMAX)], az[NCART(LMAX)], bx[NCART(LMAX)], by[NCART(LMAX)], bz[NCART(LMAX)];
    cart_powers(la, ax, ay, az);
    cart_powers(lb, bx, by, bz);
    int lb2 = lb + 2; /* kinetic needs j+2 */
    int nt = la + lb2 + 1;
    double *E[3];
    for (int d = 0; d < 3; d++) E[d] = (double *)malloc(sizeof(double) * (la + 1) * (lb2 + 1) * nt);
    int L = la + lb;
    double *R = (double *)malloc(sizeof(double) * (L + 1) * (L + 1) * (L + 1));
    double *work = (double *)malloc(sizeof(double) * (L + 1) * (L + 1) * (L + 1) * (L + 1));
    memset(cart, 0, sizeof(double) * nca * ncb);
    double AB[3] = {A->R[0] - B->R[0], A->R[1] - B->R[1], A->R[2] - B->R[2]};
    double AB2 = AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2];
#define EX(d, i, j, t) E[d][((i) * (lb2 + 1) + (j)) * nt + (t)]
    for (int ip = 0; ip < A->nprim; ip++)
        for (int jp = 0; jp < B->nprim; jp++) {
            double a = A->ex[ip], b = B->ex[jp], p = a + b;
            double cc = A->co[ip] * B->co[jp] * exp(-a * b / p * AB2);
            for (int d = 0; d < 3; d++) hermite_E(la, lb2, a, b, AB[d], E[d]);
            double P[3] = {(a * A->R[0] + b * B->R[0]) / p, (a * A->R[1] + b * B->R[1]) / p,
                           (a * A->R[2] + b * B->R[2]) / p};
            if (kind <= 1) {
                double s1 = sqrt(M_PI / p);
                for (int ia = 0; ia < nca; ia++)
                    for (int ib = 0; ib < ncb; ib++) {
                        int i[3] = {ax[ia], ay[ia], az[ia]}, j[3] = {bx[ib], by[ib], bz[ib]};
                        double S[3], D2[3];
                        for (int d = 0; d < 3; d++) {
                            S[d] = EX(d, i[d], j[d], 0) * s1;
                            /* d^2/dx^2 acting on the ket primitive */
                            double v = 4 * b * b * EX(d, i[d], j[d] + 2, 0) - 2 * b * (2 * j[d] + 1) * EX(d, i[d], j[d], 0);
                            if (j[d] >= 2) v += j[d] * (j[d] - 1) * EX(d, i[d], j[d] - 2, 0);
                            D2[d] = v * s1;
                        }
                        double val = (kind == 0) ? S[0] * S[1] * S[2]
                                                 : -0.5 * (D2[0] * S[1] * S[2] + S[0] * D2[1] * S[2] + S[0] * S[1] * D2[2]);
                        cart[ia * ncb + ib] += cc * val;
                    }
            } else {
                int ncen = (kind == 2) ? natm : 1;
                for (int ic = 0; ic < ncen; ic++) {
                    const double *C = (kind == 2) ? env + atm[ic * ATM_SLOTS + 1] : env + PTR_RINV_ORIG;
                    double q = (kind == 2) ? -(double)atm[ic * ATM_SLOTS + 0] : 1.0;
                    double PC[3] = {P[0] - C[0], P[1] - C[1], P[2] - C[2]};
                    hermite_R(L, p, PC, R, work);
                    double pref = cc * q * 2 * M_PI / p;
                    for (int ia = 0; ia < nca; ia++)
                        for (int ib = 0; ib < ncb; ib++) {
                            double s = 0.0;
                            for (int t = 0; t <= ax[ia] + bx[ib]; t++)
                                for (int u = 0; u <= ay[ia] + by[ib]; u++)
                                    for (int v = 0; v <= az[ia] + bz[ib]; v++)
                                        s += EX(0, ax[ia], bx[ib], t) * EX(1, ay[ia], by[ib], u) *
                                             EX(2, az[ia], bz[ib], v) * R[(t * (L + 1) + u) * (L + 1) + v];
                            cart[ia * ncb + ib] += pref * s;
                        }
                }
            }
        }
#undef EX
    for (int d = 0; d < 3; d++) free(E[d]);
    free(R);
    free(work);
}

/* out[(i - i0)][(j - j0)] over shells [ish0,ish1) x [jsh0,jsh1) */
void orc_int1e(int kind, double *out, const int *shls_slice, const int *ao_loc, const int *atm,
               int natm, const int *bas, int nbas, const double *env) {
    orc_init();
    int ish0 = shls_slice[0], ish1 = shls_slice[1], jsh0 = shls_slice[2], jsh1 = shls_slice[3];
    int ni = ao_loc[ish1] - ao_loc[ish0], nj = ao_loc[jsh1] - ao_loc[jsh0];
    (void)ni;
    (void)nbas;
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int ish = ish0; ish < ish1; ish++)
        for (int jsh = jsh0; jsh < jsh1; jsh++) {
            Shell A, B;
            get_shell(&A, ish, atm, bas, env);
            get_shell(&B, jsh, atm, bas, env);
            int nca = NCART(A.l), ncb = NCART(B.l), nsa = 2 * A.l + 1, nsb = 2 * B.l + 1;
            double *cart = (double *)malloc(sizeof(double) * nca * ncb);
            double *t1 = (double *)malloc(sizeof(double) * nsa * ncb);
            double *t2 = (double *)malloc(sizeof(double) * nsa * nsb);
            shellpair_1e(kind, &A, &B, atm, natm, env, cart);
            c2s_axis(cart, t1, 1, A.l, ncb);
            c2s_axis(t1, t2, nsa, B.l, 1);
            int i0 = ao_loc[ish] - ao_loc[ish0], j0 = ao_loc[jsh] - ao_loc[jsh0];
            for (int i = 0; i < nsa; i++)
                for (int j = 0; j < nsb; j++) out[(size_t)(i0 + i) * nj + j0 + j] = t2[i * nsb + j];
            free(cart);
            free(t1);
            free(t2);
        }
}

/* ------------------------------------------------------------------ two-electron integrals
 * Generic (AB|CD); a missing centre is a Shell with l=0, one primitive, exponent 0, coef 1 placed
 * on its partner's centre.  Result: cartesian block [nca][ncb][ncc][ncd]. */
static const double ZERO_EX = 0.0, ONE_CO = 1.0;
/* libcint applies the s/p solid-harmonic constants only to the shells that exist (int2c2e: two
 * CINTcommon_fac_sp factors, int3c2e: three); quartet_sph() pushes the dummy s shells through
 * the same c2s table, so that 1/sqrt(4 pi) is taken out again per dummy. */
static const double UNDO_DUMMY_S = 1.0 / 0.282094791773878143;

static void dummy_shell(Shell *s, const Shell *partner) {
    s->l = 0;
    s->nprim = 1;
    s->ex = &ZERO_EX;
    s->co = &ONE_CO;
    s->R = partner->R;
}

static void quartet_cart(const Shell *A, const Shell *B, const Shell *C, const Shell *D, double *out) {
    int la = A->l, lb = B->l, lc = C->l, ld = D->l;
    int nca = NCART(la), ncb = NCART(lb), ncc = NCART(lc), ncd = NCART(ld);
    int Lab = la + lb, Lcd = lc + ld, L = Lab + Lcd;
    int n1 = L + 1;
    int ax[NCART(LMAX)], ay[NCART(LMAX)], az[NCART(LMAX)], bx[NCART(LMAX)], by[NCART(LMAX)], bz[NCART(LMAX)];
    int cx[NCART(LMAX)], cy[NCART(LMAX)], cz[NCART(LMAX)], dx[NCART(LMAX)], dy[NCART(LMAX)], dz[NCART(LMAX)];
    cart_powers(la, ax, ay, az);
    cart_powers(lb, bx, by, bz);
    cart_powers(lc, cx, cy, cz);
    cart_powers(ld, dx, dy, dz);
    size_t ntot = (size_t)nca * ncb * ncc * ncd;
    memset(out, 0, sizeof(double) * ntot);
    double *Eab[3], *Ecd[3];
    for (int d = 0; d < 3; d++) {
        Eab[d] = (double *)malloc(sizeof(double) * (la + 1) * (lb + 1) * (Lab + 1));
        Ecd[d] = (double *)malloc(sizeof(double) * (lc + 1) * (ld + 1) * (Lcd + 1));
    }
    double *R = (double *)malloc(sizeof(double) * n1 * n1 * n1);
    double *work = (double *)malloc(sizeof(double) * n1 * n1 * n1 * n1);
    /* half-contracted: G[cd comp][t][u][v] = sum_{tau nu phi} (-1)^(..) Ecd R[t+tau][u+nu][v+phi] */
    int nab1 = Lab + 1;
    double *G = (double *)malloc(sizeof(double) * ncc * ncd * nab1 * nab1 * nab1);
    double AB[3] = {A->R[0] - B->R[0], A->R[1] - B->R[1], A->R[2] - B->R[2]};
    double CD[3] = {C->R[0] - D->R[0], C->R[1] - D->R[1], C->R[2] - D->R[2]};
    double AB2 = AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2];
    double CD2 = CD[0] * CD[0] + CD[1] * CD[1] + CD[2] * CD[2];
#define EA(d, i, j, t) Eab[d][((i) * (lb + 1) + (j)) * (Lab + 1) + (t)]
#define EC(d, i, j, t) Ecd[d][((i) * (ld + 1) + (j)) * (Lcd + 1) + (t)]
    for (int ip = 0; ip < A->nprim; ip++)
        for (int jp = 0; jp < B->nprim; jp++) {
            double a = A->ex[ip], b = B->ex[jp], p = a + b;
            double cab = A->co[ip] * B->co[jp] * exp(-a * b / p * AB2);
            for (int d = 0; d < 3; d++) hermite_E(la, lb, a, b, AB[d], Eab[d]);
            double P[3] = {(a * A->R[0] + b * B->R[0]) / p, (a * A->R[1] + b * B->R[1]) / p,
                           (a * A->R[2] + b * B->R[2]) / p};
            for (int kp = 0; kp < C->nprim; kp++)
                for (int lp = 0; lp < D->nprim; lp++) {
                    double c = C->ex[kp], dd = D->ex[lp], q = c + dd;
                    double ccd = C->co[kp] * D->co[lp] * exp(-c * dd / q * CD2);
                    for (int d = 0; d < 3; d++) hermite_E(lc, ld, c, dd, CD[d], Ecd[d]);
                    double Q[3] = {(c * C->R[0] + dd * D->R[0]) / q, (c * C->R[1] + dd * D->R[1]) / q,
                                   (c * C->R[2] + dd * D->R[2]) / q};
                    double PQ[3] = {P[0] - Q[0], P[1] - Q[1], P[2] - Q[2]};
                    double alpha = p * q / (p + q);
                    hermite_R(L, alpha, PQ, R, work);
                    double pref = cab * ccd * 2 * pow(M_PI, 2.5) / (p * q * sqrt(p + q));
                    for (int ic = 0; ic < ncc; ic++)
                        for (int id = 0; id < ncd; id++) {
                            double *g = G + (size_t)(ic * ncd + id) * nab1 * nab1 * nab1;
                            int Tx = cx[ic] + dx[id], Ty = cy[ic] + dy[id], Tz = cz[ic] + dz[id];
                            for (int t = 0; t <= Lab; t++)
                                for (int u = 0; u <= Lab - t; u++)
                                    for (int v = 0; v <= Lab - t - u; v++) {
                                        double s = 0.0;
                                        for (int ta = 0; ta <= Tx; ta++)
                                            for (int nu = 0; nu <= Ty; nu++)
                                                for (int ph = 0; ph <= Tz; ph++) {
                                                    double e = EC(0, cx[ic], dx[id], ta) * EC(1, cy[ic], dy[id], nu) *
                                                               EC(2, cz[ic], dz[id], ph);
                                                    if ((ta + nu + ph) & 1) e = -e;
                                                    s += e * R[((t + ta) * n1 + (u + nu)) * n1 + (v + ph)];
                                                }
                                        g[(t * nab1 + u) * nab1 + v] = s;
                                    }
                        }
                    for (int ia = 0; ia < nca; ia++)
                        for (int ib = 0; ib < ncb; ib++) {
                            int Tx = ax[ia] + bx[ib], Ty = ay[ia] + by[ib], Tz = az[ia] + bz[ib];
                            for (int icd = 0; icd < ncc * ncd; icd++) {
                                const double *g = G + (size_t)icd * nab1 * nab1 * nab1;
                                double s = 0.0;
                                for (int t = 0; t <= Tx; t++)
                                    for (int u = 0; u <= Ty; u++)
                                        for (int v = 0; v <= Tz; v++)
                                            s += EA(0, ax[ia], bx[ib], t) * EA(1, ay[ia], by[ib], u) *
                                                 EA(2, az[ia], bz[ib], v) * g[(t * nab1 + u) * nab1 + v];
                                out[(size_t)(ia * ncb + ib) * ncc * ncd + icd] += pref * s;
                            }
                        }
                }
        }
#undef EA
#undef EC
    for (int d = 0; d < 3; d++) {
        free(Eab[d]);
        free(Ecd[d]);
    }
    free(R);
    free(work);
    free(G);
}

/* spherical block of (AB|CD): sph[nsa][nsb][nsc][nsd]; returns via out buffer (caller sized) */
static void quartet_sph(const Shell *A, const Shell *B, const Shell *C, const Shell *D, double *sph) {
    int nca = NCART(A->l), ncb = NCART(B->l), ncc = NCART(C->l), ncd = NCART(D->l);
    int nsa = 2 * A->l + 1, nsb = 2 * B->l + 1, nsc = 2 * C->l + 1;
    size_t nmax = (size_t)nca * ncb * ncc * ncd;
    double *t0 = (double *)malloc(sizeof(double) * nmax);
    double *t1 = (double *)malloc(sizeof(double) * nmax);
    quartet_cart(A, B, C, D, t0);
    c2s_axis(t0, t1, 1, A->l, ncb * ncc * ncd);
    c2s_axis(t1, t0, nsa, B->l, ncc * ncd);
    c2s_axis(t0, t1, nsa * nsb, C->l, ncd);
    c2s_axis(t1, sph, nsa * nsb * nsc, D->l, 1);
    free(t0);
    free(t1);
}

/* (P|Q): out[P][Q] over shells [s0,s1) x [s2,s3) */
void orc_int2c2e(double *out, const int *shls_slice, const int *ao_loc, const int *atm, int natm,
                 const int *bas, int nbas, const double *env) {
    orc_init();
    (void)natm;
    (void)nbas;
    int i0s = shls_slice[0], i1s = shls_slice[1], j0s = shls_slice[2], j1s = shls_slice[3];
    int nj = ao_loc[j1s] - ao_loc[j0s];
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int ish = i0s; ish < i1s; ish++)
        for (int jsh = j0s; jsh < j1s; jsh++) {
            Shell A, B, C, D;
            get_shell(&A, ish, atm, bas, env);
            get_shell(&C, jsh, atm, bas, env);
            dummy_shell(&B, &A);
            dummy_shell(&D, &C);
            int nsa = 2 * A.l + 1, nsc = 2 * C.l + 1;
            double *sph = (double *)malloc(sizeof(double) * NCART(A.l) * NCART(C.l));
            quartet_sph(&A, &B, &C, &D, sph);
            int i0 = ao_loc[ish] - ao_loc[i0s], j0 = ao_loc[jsh] - ao_loc[j0s];
            for (int i = 0; i < nsa; i++)
                for (int j = 0; j < nsc; j++)
                    out[(size_t)(i0 + i) * nj + j0 + j] = sph[i * nsc + j] * UNDO_DUMMY_S * UNDO_DUMMY_S;
            free(sph);
        }
}

/* (ij|P): out[i][j][P] over shells [s0,s1) x [s2,s3) x [s4,s5) */
void orc_int3c2e(double *out, const int *shls_slice, const int *ao_loc, const int *atm, int natm,
                 const int *bas, int nbas, const double *env) {
    orc_init();
    (void)natm;
    (void)nbas;
    int i0s = shls_slice[0], i1s = shls_slice[1], j0s = shls_slice[2], j1s = shls_slice[3];
    int k0s = shls_slice[4], k1s = shls_slice[5];
    size_t nj = ao_loc[j1s] - ao_loc[j0s], nk = ao_loc[k1s] - ao_loc[k0s];
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int ish = i0s; ish < i1s; ish++)
        for (int jsh = j0s; jsh < j1s; jsh++) {
            Shell A, B, C, D;
            get_shell(&A, ish, atm, bas, env);
            get_shell(&B, jsh, atm, bas, env);
            int nsa = 2 * A.l + 1, nsb = 2 * B.l + 1;
            int i0 = ao_loc[ish] - ao_loc[i0s], j0 = ao_loc[jsh] - ao_loc[j0s];
            for (int ksh = k0s; ksh < k1s; ksh++) {
                get_shell(&C, ksh, atm, bas, env);
                dummy_shell(&D, &C);
                int nsc = 2 * C.l + 1;
                double *sph = (double *)malloc(sizeof(double) * NCART(A.l) * NCART(B.l) * NCART(C.l));
                quartet_sph(&A, &B, &C, &D, sph);
                int k0 = ao_loc[ksh] - ao_loc[k0s];
                for (int i = 0; i < nsa; i++)
                    for (int j = 0; j < nsb; j++)
                        for (int k = 0; k < nsc; k++)
                            out[((size_t)(i0 + i) * nj + j0 + j) * nk + k0 + k] =
                                sph[(i * nsb + j) * nsc + k] * UNDO_DUMMY_S;
                free(sph);
            }
        }
}

/* (ij|kl): out[i][j][k][l], all four ranges from shls_slice[8]; no screening, like the reference
 * (prescreen = NULL, molintor.py:676).  Uses ij<->ji, kl<->lk, (ij)<->(kl) symmetry only when the
 * four ranges are identical. */
void orc_int2e(double *out, const int *shls_slice, const int *ao_loc, const int *atm, int natm,
               const int *bas, int nbas, const double *env) {
    orc_init();
    (void)natm;
    (void)nbas;
    int s0[4], s1[4];
    size_t n[4];
    for (int q = 0; q < 4; q++) {
        s0[q] = shls_slice[2 * q];
        s1[q] = shls_slice[2 * q + 1];
        n[q] = ao_loc[s1[q]] - ao_loc[s0[q]];
    }
    int same = 1;
    for (int q = 1; q < 4; q++) same = same && s0[q] == s0[0] && s1[q] == s1[0];
    int nsh = s1[0] - s0[0];
    if (same) {
        int npair = nsh * (nsh + 1) / 2;
#pragma omp parallel for schedule(dynamic)
        for (int ij = 0; ij < npair; ij++) {
            int i = (int)((sqrt(8.0 * ij + 1) - 1) / 2);
            while (i * (i + 1) / 2 > ij) i--;
            while ((i + 1) * (i + 2) / 2 <= ij) i++;
            int j = ij - i * (i + 1) / 2;
            for (int kl = 0; kl <= ij; kl++) {
                int k = (int)((sqrt(8.0 * kl + 1) - 1) / 2);
                while (k * (k + 1) / 2 > kl) k--;
                while ((k + 1) * (k + 2) / 2 <= kl) k++;
                int l = kl - k * (k + 1) / 2;
                Shell A, B, C, D;
                get_shell(&A, s0[0] + i, atm, bas, env);
                get_shell(&B, s0[0] + j, atm, bas, env);
                get_shell(&C, s0[0] + k, atm, bas, env);
                get_shell(&D, s0[0] + l, atm, bas, env);
                int na = 2 * A.l + 1, nb = 2 * B.l + 1, nc = 2 * C.l + 1, nd = 2 * D.l + 1;
                double *sph = (double *)malloc(sizeof(double) * NCART(A.l) * NCART(B.l) * NCART(C.l) * NCART(D.l));
                quartet_sph(&A, &B, &C, &D, sph);
                size_t N = n[0];
                int oi = ao_loc[s0[0] + i] - ao_loc[s0[0]], oj = ao_loc[s0[0] + j] - ao_loc[s0[0]];
                int ok = ao_loc[s0[0] + k] - ao_loc[s0[0]], ol = ao_loc[s0[0] + l] - ao_loc[s0[0]];
                for (int a = 0; a < na; a++)
                    for (int b = 0; b < nb; b++)
                        for (int c = 0; c < nc; c++)
                            for (int d = 0; d < nd; d++) {
                                double v = sph[((a * nb + b) * nc + c) * nd + d];
                                size_t I = oi + a, J = oj + b, K = ok + c, Lq = ol + d;
                                out[((I * N + J) * N + K) * N + Lq] = v;
                                out[((J * N + I) * N + K) * N + Lq] = v;
                                out[((I * N + J) * N + Lq) * N + K] = v;
                                out[((J * N + I) * N + Lq) * N + K] = v;
                                out[((K * N + Lq) * N + I) * N + J] = v;
                                out[((Lq * N + K) * N + I) * N + J] = v;
                                out[((K * N + Lq) * N + J) * N + I] = v;
                                out[((Lq * N + K) * N + J) * N + I] = v;
                            }
                free(sph);
            }
        }
        return;
    }
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int ish = s0[0]; ish < s1[0]; ish++)
        for (int jsh = s0[1]; jsh < s1[1]; jsh++)
            for (int ksh = s0[2]; ksh < s1[2]; ksh++)
                for (int lsh = s0[3]; lsh < s1[3]; lsh++) {
                    Shell A, B, C, D;
                    get_shell(&A, ish, atm, bas, env);
                    get_shell(&B, jsh, atm, bas, env);
                    get_shell(&C, ksh, atm, bas, env);
                    get_shell(&D, lsh, atm, bas, env);
                    int na = 2 * A.l + 1, nb = 2 * B.l + 1, nc = 2 * C.l + 1, nd = 2 * D.l + 1;
                    double *sph = (double *)malloc(sizeof(double) * NCART(A.l) * NCART(B.l) * NCART(C.l) * NCART(D.l));
                    quartet_sph(&A, &B, &C, &D, sph);
                    size_t oi = ao_loc[ish] - ao_loc[s0[0]], oj = ao_loc[jsh] - ao_loc[s0[1]];
                    size_t ok = ao_loc[ksh] - ao_loc[s0[2]], ol = ao_loc[lsh] - ao_loc[s0[3]];
                    for (int a = 0; a < na; a++)
                        for (int b = 0; b < nb; b++)
                            for (int c = 0; c < nc; c++)
                                for (int d = 0; d < nd; d++)
                                    out[(((oi + a) * n[1] + oj + b) * n[2] + ok + c) * n[3] + ol + d] =
                                        sph[((a * nb + b) * nc + c) * nd + d];
                    free(sph);
                }
}

/* ------------------------------------------------------------------ AO values on grid points
 * deriv = 2: out[g][mu] = Laplacian of phi_mu (eval_laplgto, to_transpose=True: the reference sums the xx, yy, zz
 *            components of GTOval_sph_deriv2, dqc/hamilton/intor/gtoeval.py:66-73)
 * deriv = 0: out[g][mu]              (what eval_gto(..., to_transpose=True) returns)
 * deriv = 1: out[comp][g][mu], comp = 0 value? NO: reference keeps them separate; here
 *            out[0..2][g][mu] = d/dx, d/dy, d/dz phi_mu (eval_gradgto, to_transpose=True).
 * coords is (ngrid, 3) C-order. */
void orc_eval_gto(int deriv, int ngrid, const double *coords, double *out, const int *shls_slice,
                  const int *ao_loc, const int *atm, int natm, const int *bas, int nbas,
                  const double *env) {
    orc_init();
    (void)natm;
    (void)nbas;
    int sh0 = shls_slice[0], sh1 = shls_slice[1];
    size_t nao = ao_loc[sh1] - ao_loc[sh0];
    int ncomp = deriv == 1 ? 3 : 1;
#pragma omp parallel for schedule(static)
    for (int g = 0; g < ngrid; g++) {
        double cartv[3][NCART(LMAX)];
        int lx[NCART(LMAX)], ly[NCART(LMAX)], lz[NCART(LMAX)];
        for (int ish = sh0; ish < sh1; ish++) {
            Shell S;
            get_shell(&S, ish, atm, bas, env);
            int l = S.l, nc = NCART(l), ns = 2 * l + 1;
            double x = coords[3 * g] - S.R[0], y = coords[3 * g + 1] - S.R[1], z = coords[3 * g + 2] - S.R[2];
            double r2 = x * x + y * y + z * z;
            double rad = 0.0, drad = 0.0, d2rad = 0.0; /* sum c e^{-a r2}; sum -2 a c e^{-a r2}; Laplacian of the radial sum */
            for (int p = 0; p < S.nprim; p++) {
                double e = S.co[p] * exp(-S.ex[p] * r2);
                rad += e;
                drad += -2.0 * S.ex[p] * e;
                d2rad += (4.0 * S.ex[p] * S.ex[p] * r2 - 6.0 * S.ex[p]) * e;
            }
            cart_powers(l, lx, ly, lz);
            double px[LMAX + 2], py[LMAX + 2], pz[LMAX + 2];
            px[0] = py[0] = pz[0] = 1.0;
            for (int k = 1; k <= l + 1; k++) {
                px[k] = px[k - 1] * x;
                py[k] = py[k - 1] * y;
                pz[k] = pz[k - 1] * z;
            }
            for (int c = 0; c < nc; c++) {
                int a = lx[c], b = ly[c], cz = lz[c];
                double mono = px[a] * py[b] * pz[cz];
                if (!deriv) {
                    cartv[0][c] = mono * rad;
                } else if (deriv == 2) {
                    /* lapl (m R) = (lapl m) R + 2 grad m . grad R + m lapl R, grad R = r drad, grad m . r = l m */
                    double lm = (a > 1 ? a * (a - 1) * px[a - 2] : 0.0) * py[b] * pz[cz] +
                                px[a] * (b > 1 ? b * (b - 1) * py[b - 2] : 0.0) * pz[cz] +
                                px[a] * py[b] * (cz > 1 ? cz * (cz - 1) * pz[cz - 2] : 0.0);
                    cartv[0][c] = lm * rad + mono * (2.0 * l * drad + d2rad);
                } else {
                    double dmx = (a ? a * px[a - 1] : 0.0) * py[b] * pz[cz];
                    double dmy = px[a] * (b ? b * py[b - 1] : 0.0) * pz[cz];
                    double dmz = px[a] * py[b] * (cz ? cz * pz[cz - 1] : 0.0);
                    cartv[0][c] = dmx * rad + mono * x * drad;
                    cartv[1][c] = dmy * rad + mono * y * drad;
                    cartv[2][c] = dmz * rad + mono * z * drad;
                }
            }
            const double *M = c2s_tab[l];
            size_t mu0 = ao_loc[ish] - ao_loc[sh0];
            for (int comp = 0; comp < ncomp; comp++)
                for (int m = 0; m < ns; m++) {
                    double s = 0.0;
                    for (int c = 0; c < nc; c++) s += M[m * nc + c] * cartv[comp][c];
                    out[((size_t)comp * ngrid + g) * nao + mu0 + m] = s;
                }
        }
    }
}
