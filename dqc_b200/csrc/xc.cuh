// K3 -- pointwise LDA/GGA exchange-correlation (value + first derivatives).
// Replaces pylibxc's LibXCFunctional.compute as driven by dqc/xc/libxc_wrapper.py:380-413 with the
// input/output massaging of dqc/xc/libxc.py:124-242 fused in (sigma = |grad rho|^2 is formed here,
// the energy comes back per unit volume, the gradient potential as 2 vsigma grad rho).
// Functional forms are the published ones (Slater; Perdew-Wang 92 with libxc's `pw` and `pw_mod`
// parameter sets; PBE exchange and correlation); derivatives are hand-derived and checked in
// tests against torch autograd of the energy (the reference's own default route).
#pragma once
#include "common.cuh"

#define XC_LDA_X 1
#define XC_LDA_C_PW 2
#define XC_LDA_C_PW_MOD 3
#define XC_GGA_X_PBE 101
#define XC_GGA_C_PBE 102
#define XC_LDA_C_VWN 4       // VWN5 (libxc lda_c_vwn)
#define XC_LDA_C_VWN_RPA 5   // VWN-RPA (libxc lda_c_vwn_rpa, the LDA part of libxc's B3LYP)
#define XC_GGA_X_B88 103
#define XC_GGA_C_LYP 104
#define XC_MGGA_X_SCAN 201  // SCAN exchange (libxc mgga_x_scan); needs tau, through b200qc_xc_mgga_unpol only
#define XC_MAX_TERMS 8
#define XC_RHO_CUT 1e-15   // densities at or below this contribute exactly zero (libxc-style threshold)

struct XCTerms {
    int n;
    int id[XC_MAX_TERMS];
    double coef[XC_MAX_TERMS];
};

namespace xc {
constexpr double PI = 3.14159265358979323846;
constexpr double KAPPA = 0.8040;
constexpr double BETA = 0.06672455060314922;
constexpr double MU = BETA * PI * PI / 3.0;
constexpr double GAMMA = 0.031090690869654895034;  // (1 - ln 2) / pi^2
constexpr double CX = -0.73855876638202240588;     // -3/4 (3/pi)^(1/3)
constexpr double FZ_DEN = 0.51984209978974632953;  // 2^(4/3) - 2

struct PWParams {
    double a[3], alpha1[3], b1[3], b2[3], b3[3], b4[3], fz20;
};
__device__ __forceinline__ PWParams pw_params(bool mod) {
    PWParams p;
    if (mod) {
        p.a[0] = 0.0310907; p.a[1] = 0.01554535; p.a[2] = 0.0168869;
        p.fz20 = 1.709920934161365617563962776245;
    } else {
        p.a[0] = 0.031091; p.a[1] = 0.015545; p.a[2] = 0.016887;
        p.fz20 = 1.709921;
    }
    p.alpha1[0] = 0.21370; p.alpha1[1] = 0.20548; p.alpha1[2] = 0.11125;
    p.b1[0] = 7.5957; p.b1[1] = 14.1189; p.b1[2] = 10.357;
    p.b2[0] = 3.5876; p.b2[1] = 6.1977; p.b2[2] = 3.6231;
    p.b3[0] = 1.6382; p.b3[1] = 3.3662; p.b3[2] = 0.88026;
    p.b4[0] = 0.49294; p.b4[1] = 0.62517; p.b4[2] = 0.49671;
    return p;
}
// G(rs) = -2a(1+alpha1 rs) ln(1 + 1/Q1) and dG/drs
__device__ __forceinline__ void pw_g(const PWParams &p, int k, double rs, double &g, double &dg) {
    const double srs = sqrt(rs);
    const double a = p.a[k];
    const double Q1 = 2 * a * (p.b1[k] * srs + p.b2[k] * rs + p.b3[k] * rs * srs + p.b4[k] * rs * rs);
    const double dQ1 = 2 * a * (0.5 * p.b1[k] / srs + p.b2[k] + 1.5 * p.b3[k] * srs + 2 * p.b4[k] * rs);
    const double lg = log1p(1.0 / Q1);
    const double Q0 = -2 * a * (1 + p.alpha1[k] * rs);
    g = Q0 * lg;
    dg = -2 * a * p.alpha1[k] * lg - Q0 * dQ1 / (Q1 * (Q1 + 1.0));
}
// PW92 eps_c(rs, zeta) with d/drs and d/dzeta
__device__ __forceinline__ void pw_eps(bool mod, double rs, double zeta, double &eps, double &deps_drs,
                                       double &deps_dz) {
    const PWParams p = pw_params(mod);
    double g0, d0;
    pw_g(p, 0, rs, g0, d0);
    if (zeta == 0.0) {
        eps = g0;
        deps_drs = d0;
        deps_dz = 0.0;
        return;
    }
    double g1, d1, g2, d2;
    pw_g(p, 1, rs, g1, d1);
    pw_g(p, 2, rs, g2, d2);
    const double opz = fmax(1.0 + zeta, 1e-15), omz = fmax(1.0 - zeta, 1e-15);
    const double f = (opz * cbrt(opz) + omz * cbrt(omz) - 2.0) / FZ_DEN;
    const double df = (4.0 / 3.0) * (cbrt(opz) - cbrt(omz)) / FZ_DEN;
    const double z3 = zeta * zeta * zeta, z4 = z3 * zeta;
    const double w = g1 - g0 + g2 / p.fz20;
    eps = g0 + z4 * f * w - f * g2 / p.fz20;
    deps_drs = d0 + z4 * f * (d1 - d0 + d2 / p.fz20) - f * d2 / p.fz20;
    deps_dz = (4 * z3 * f + z4 * df) * w - df * g2 / p.fz20;
}

// ---- forward-mode duals: functionals written once as an energy expression, derivatives by the chain rule ----
// (B88, LYP, VWN: the libxc forms are restated from the original papers; the kernel's derivatives are checked
// against torch autograd of the oracle's expressions, the reference's own default route, base_xc.py:39-125.)
template <int N>
struct Dual {
    double v, d[N];
};
#define DUAL_OP __device__ __forceinline__
template <int N> DUAL_OP Dual<N> dconst(double c) { Dual<N> r; r.v = c;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = 0.0; return r; }
template <int N> DUAL_OP Dual<N> dvar(double c, int k, double seed = 1.0) { Dual<N> r = dconst<N>(c); r.d[k] = seed; return r; }
template <int N> DUAL_OP Dual<N> operator+(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> DUAL_OP Dual<N> operator-(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> DUAL_OP Dual<N> operator*(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> DUAL_OP Dual<N> operator/(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; const double ib = 1.0 / b.v; r.v = a.v * ib;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
template <int N> DUAL_OP Dual<N> operator+(const Dual<N> &a, double c) { Dual<N> r = a; r.v += c; return r; }
template <int N> DUAL_OP Dual<N> operator+(double c, const Dual<N> &a) { return a + c; }
template <int N> DUAL_OP Dual<N> operator-(const Dual<N> &a, double c) { Dual<N> r = a; r.v -= c; return r; }
template <int N> DUAL_OP Dual<N> operator-(double c, const Dual<N> &a) { Dual<N> r; r.v = c - a.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = -a.d[i]; return r; }
template <int N> DUAL_OP Dual<N> operator-(const Dual<N> &a) { return 0.0 - a; }
template <int N> DUAL_OP Dual<N> operator*(const Dual<N> &a, double c) { Dual<N> r; r.v = a.v * c;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = a.d[i] * c; return r; }
template <int N> DUAL_OP Dual<N> operator*(double c, const Dual<N> &a) { return a * c; }
template <int N> DUAL_OP Dual<N> operator/(const Dual<N> &a, double c) { return a * (1.0 / c); }
template <int N> DUAL_OP Dual<N> operator/(double c, const Dual<N> &a) { return dconst<N>(c) / a; }
// f(a) with derivative df
template <int N> DUAL_OP Dual<N> dchain(const Dual<N> &a, double f, double df) { Dual<N> r; r.v = f;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = df * a.d[i]; return r; }
template <int N> DUAL_OP Dual<N> dsqrt(const Dual<N> &a) { const double f = sqrt(a.v); return dchain(a, f, a.v > 0.0 ? 0.5 / f : 0.0); }
template <int N> DUAL_OP Dual<N> dcbrt(const Dual<N> &a) { const double f = cbrt(a.v); return dchain(a, f, a.v != 0.0 ? f / (3.0 * a.v) : 0.0); }
template <int N> DUAL_OP Dual<N> dlog(const Dual<N> &a) { return dchain(a, log(a.v), 1.0 / a.v); }
template <int N> DUAL_OP Dual<N> dexp(const Dual<N> &a) { const double f = exp(a.v); return dchain(a, f, f); }
template <int N> DUAL_OP Dual<N> datan(const Dual<N> &a) { return dchain(a, atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
template <int N> DUAL_OP Dual<N> dasinh(const Dual<N> &a) { return dchain(a, asinh(a.v), 1.0 / sqrt(1.0 + a.v * a.v)); }

// Becke 88 exchange of ONE spin channel: e_s = -rho_s^(4/3) [ (3/2)(3/4pi)^(1/3) + beta x^2 / (1 + 6 beta x asinh x) ],
// x = |grad rho_s| / rho_s^(4/3), beta = 0.0042 (Becke, PRA 38, 3098 (1988))
template <int N> DUAL_OP Dual<N> b88_spin(const Dual<N> &rs, const Dual<N> &sss) {
    const double BETA88 = 0.0042, CX88 = 0.93052573634910002;   // (3/2) (3 / 4 pi)^(1/3)
    const Dual<N> c = dcbrt(rs);
    const Dual<N> r43 = rs * c;
    const Dual<N> x = dsqrt(sss) / r43;
    return -(r43 * (CX88 + BETA88 * x * x / (1.0 + 6.0 * BETA88 * x * dasinh(x))));
}
// Lee-Yang-Parr correlation in the gradient-only form of Miehlich, Savin, Stoll, Preuss, CPL 157, 200 (1989)
template <int N> DUAL_OP Dual<N> lyp_energy(const Dual<N> &ra, const Dual<N> &rb, const Dual<N> &saa, const Dual<N> &sab,
                                            const Dual<N> &sbb) {
    const double A = 0.04918, B = 0.132, C = 0.2533, Dd = 0.349;
    const double CF = 2.8712340001881918;                        // (3/10) (3 pi^2)^(2/3)
    const Dual<N> rho = ra + rb;
    const Dual<N> c13 = dcbrt(rho);
    const Dual<N> rm13 = 1.0 / c13;
    const Dual<N> den = 1.0 + Dd * rm13;
    const Dual<N> r2 = rho * rho;
    const Dual<N> omega = dexp(-(C * rm13)) / den / (r2 * rho * c13 * c13);         // rho^(-11/3)
    const Dual<N> delta = C * rm13 + Dd * rm13 / den;
    const Dual<N> ca = dcbrt(ra), cb = dcbrt(rb);
    const Dual<N> ra83 = ra * ra * ca * ca, rb83 = rb * rb * cb * cb;
    const Dual<N> stot = saa + 2.0 * sab + sbb;
    const Dual<N> rab = ra * rb;
    const Dual<N> t1 = 12.699208415745595 * CF * (ra83 + rb83);                       // 2^(11/3) CF (...)
    const Dual<N> t2 = (47.0 / 18.0 - (7.0 / 18.0) * delta) * stot;
    const Dual<N> t3 = (2.5 - delta / 18.0) * (saa + sbb);
    const Dual<N> t4 = ((delta - 11.0) / 9.0) * ((ra * saa + rb * sbb) / rho);
    const Dual<N> inner = rab * (t1 + t2 - t3 - t4) - (2.0 / 3.0) * r2 * stot + ((2.0 / 3.0) * r2 - ra * ra) * sbb +
                          ((2.0 / 3.0) * r2 - rb * rb) * saa;
    return -(A * 4.0 * rab / (den * rho)) - (A * B) * omega * inner;
}
// Vosko-Wilk-Nusair: eps(x = sqrt(rs)) of one parameter set (VWN, Can. J. Phys. 58, 1200 (1980), eq. 4.4)
template <int N> DUAL_OP Dual<N> vwn_aux(const Dual<N> &x, double A, double x0, double b, double c) {
    const double Q = sqrt(4.0 * c - b * b), X0 = x0 * x0 + b * x0 + c;
    const Dual<N> X = x * x + b * x + c;
    const Dual<N> at = datan(Q / (2.0 * x + b));
    const Dual<N> xm = x - x0;
    return A * (dlog(x * x / X) + (2.0 * b / Q) * at -
                (b * x0 / X0) * (dlog(xm * xm / X) + (2.0 * (b + 2.0 * x0) / Q) * at));
}
// energy per volume; rpa = false: VWN5 with the spin-stiffness interpolation, true: the RPA parameter sets with
// the plain f(zeta) interpolation between para- and ferromagnetic (libxc lda_c_vwn / lda_c_vwn_rpa)
template <int N> DUAL_OP Dual<N> vwn_energy(bool rpa, const Dual<N> &ra, const Dual<N> &rb) {
    const Dual<N> rho = ra + rb;
    const Dual<N> x = dsqrt(0.62035049089940009 / dcbrt(rho));                       // sqrt(rs), rs = (3 / 4 pi rho)^(1/3)
    Dual<N> zeta = (ra - rb) / rho;
    const Dual<N> opz = 1.0 + zeta, omz = 1.0 - zeta;
    const Dual<N> fz = (opz * dcbrt(opz) + omz * dcbrt(omz) - 2.0) / FZ_DEN;
    Dual<N> eps;
    if (rpa) {
        const Dual<N> ep = vwn_aux(x, 0.0310907, -0.409286, 13.0720, 42.7198);
        const Dual<N> ef = vwn_aux(x, 0.01554535, -0.743294, 20.1231, 101.578);
        eps = ep + (ef - ep) * fz;
    } else {
        const Dual<N> ep = vwn_aux(x, 0.0310907, -0.10498, 3.72744, 12.9352);
        const Dual<N> ef = vwn_aux(x, 0.01554535, -0.32500, 7.06042, 18.0578);
        const Dual<N> ac = vwn_aux(x, -1.0 / (6.0 * PI * PI), -0.0047584, 1.13107, 13.0045);
        const Dual<N> z2 = zeta * zeta, z4 = z2 * z2;
        eps = ep + ac * fz * (1.0 - z4) / 1.7099209341613657 + (ef - ep) * fz * z4;
    }
    return rho * eps;
}
// unpolarised front ends of the dual-number functionals: (rho, sigma) -> e, de/drho, de/dsigma
__device__ __forceinline__ void dual_unpol(int id, double rho, double sigma, double &e, double &vr, double &vs) {
    typedef Dual<2> T;
    // spin-resolved inputs of the closed-shell point: rho_s = rho / 2, sigma_ss' = sigma / 4
    const T ra = dvar<2>(0.5 * rho, 0, 0.5), sq = dvar<2>(0.25 * sigma, 1, 0.25);
    T r;
    switch (id) {
        case XC_GGA_X_B88: r = 2.0 * b88_spin(ra, sq); break;
        case XC_GGA_C_LYP: r = lyp_energy(ra, ra, sq, sq, sq); break;
        case XC_LDA_C_VWN: r = vwn_energy(false, ra, ra); break;
        default: r = vwn_energy(true, ra, ra); break;
    }
    e = r.v; vr = r.d[0]; vs = r.d[1];
}

// ---- unpolarised: (rho, sigma) -> e (per volume), de/drho, de/dsigma ----
__device__ __forceinline__ void lda_x_unpol(double rho, double &e, double &vr) {
    const double r13 = cbrt(rho);
    e = CX * rho * r13;
    vr = (4.0 / 3.0) * CX * r13;
}
__device__ __forceinline__ void lda_c_pw_unpol(bool mod, double rho, double &e, double &vr) {
    const double rs = cbrt(3.0 / (4 * PI * rho));
    double eps, drs, dz;
    pw_eps(mod, rs, 0.0, eps, drs, dz);
    e = rho * eps;
    vr = eps - (rs / 3.0) * drs;
}
__device__ __forceinline__ void gga_x_pbe_unpol(double rho, double sigma, double &e, double &vr, double &vs) {
    const double r13 = cbrt(rho);
    const double elda = CX * rho * r13;
    const double c = 4.0 * 9.5707800006273050 ;  // 4 (3 pi^2)^(2/3)
    const double r83 = rho * rho * r13 * r13;     // rho^(8/3)
    const double s2 = sigma / (c * r83);
    const double den = 1.0 + MU * s2 / KAPPA;
    const double F = 1.0 + KAPPA - KAPPA / den;
    const double dF = MU / (den * den);  // dF/ds2
    e = elda * F;
    vr = (4.0 / 3.0) * (elda / rho) * F - elda * dF * (8.0 / 3.0) * s2 / rho;
    vs = elda * dF / (c * r83);
}
// PBE correlation, general spin polarisation.  Outputs e, de/drho_up, de/drho_dn, de/dsigma_total.
__device__ __forceinline__ void gga_c_pbe_core(double rho, double zeta, double sigma, double &e, double &vu,
                                               double &vd, double &vs) {
    const double rs = cbrt(3.0 / (4 * PI * rho));
    double eps, deps_drs, deps_dz;
    pw_eps(true, rs, zeta, eps, deps_drs, deps_dz);
    double phi = 1.0, dphi = 0.0;
    if (zeta != 0.0) {
        const double opz = fmax(1.0 + zeta, 1e-15), omz = fmax(1.0 - zeta, 1e-15);
        const double a = cbrt(opz), b = cbrt(omz);
        phi = 0.5 * (a * a + b * b);
        dphi = (1.0 / 3.0) * (1.0 / a - 1.0 / b);
    }
    const double phi3 = phi * phi * phi;
    const double r13 = cbrt(rho);
    // t^2 = sigma pi / (16 (3 pi^2)^(1/3) phi^2 rho^(7/3))
    const double ct = PI / (16.0 * 3.0936677262801355);
    const double r73 = rho * rho * r13;
    const double y = ct * sigma / (phi * phi * r73);
    const double gp3 = GAMMA * phi3;
    const double q = -eps / gp3;
    const double E = expm1(q);
    const double bg = BETA / GAMMA;
    const double A = bg / E;
    const double Ay = A * y;
    const double N = 1.0 + Ay, D = 1.0 + Ay + Ay * Ay;
    const double P = bg * y * N / D;
    const double lg = log1p(P);
    const double H = gp3 * lg;
    const double P_y = bg * (1.0 + 2.0 * Ay) / (D * D);
    const double P_A = -bg * A * y * y * y * (2.0 + Ay) / (D * D);
    const double H_y = gp3 * P_y / (1.0 + P);
    const double H_A = gp3 * P_A / (1.0 + P);
    const double A_q = -A * (E + 1.0) / E;
    const double H_eps = H_A * A_q * (-1.0 / gp3);
    const double H_phi = 3.0 * GAMMA * phi * phi * lg + H_A * A_q * (3.0 * eps / (gp3 * phi)) + H_y * (-2.0 * y / phi);
    e = rho * (eps + H);
    // d/drho at fixed zeta, then the zeta part
    const double deps_drho = -(rs / (3.0 * rho)) * deps_drs;
    const double common = (eps + H) + rho * (deps_drho * (1.0 + H_eps) + H_y * (-(7.0 / 3.0) * y / rho));
    const double dz = rho * (deps_dz * (1.0 + H_eps) + H_phi * dphi);  // rho * d(eps+H)/dzeta
    vu = common + dz * (1.0 - zeta) / rho;
    vd = common - dz * (1.0 + zeta) / rho;
    vs = rho * H_y * ct / (phi * phi * r73);
}
// SCAN exchange, unpolarised (Sun, Ruzsinszky, Perdew, PRL 115, 036402 (2015); the closed form the reference's own
// test checks libxc against, dqc/test/test_xc.py:427-455): e = e_x^LDA(rho) F_x(s, alpha),
//   s = |grad rho| / (2 rho kF), alpha = (tau - tau_W) / tau_unif, tau_W = sigma / (8 rho), tau_unif = 0.3 kF^2 rho.
// Derivatives by forward-mode duals in (rho, sigma, tau); the functional does not depend on lapl rho.
__device__ __forceinline__ void mgga_x_scan_unpol(double rho, double sigma, double tau, double &e, double &vr,
                                                  double &vs, double &vt) {
    typedef Dual<3> T;
    const T r = dvar<3>(rho, 0), sg = dvar<3>(sigma, 1), t = dvar<3>(tau, 2);
    const double a1 = 4.9479, c1x = 0.667, c2x = 0.8, dx = 1.24, mu_ak = 10.0 / 81.0, b3 = 0.5, k1 = 0.065, h0 = 1.174;
    const double b2 = sqrt(5913.0 / 405000.0), b1 = 511.0 / 13500.0 / (2.0 * b2);
    const double b4 = mu_ak * mu_ak / k1 - 1606.0 / 18225.0 - b1 * b1;
    const T kf = dcbrt(r * (3.0 * 9.869604401089358)), kf2 = kf * kf;
    const T s2 = sg / (4.0 * r * r * kf2);
    const T alpha = (t - sg / (8.0 * r)) / (0.3 * kf2 * r);
    const T oma = 1.0 - alpha;
    const T w1 = b1 * s2 + b2 * oma * dexp(-b3 * (oma * oma));
    const T x = mu_ak * s2 * (1.0 + (b4 / mu_ak) * s2 * dexp(s2 * (-fabs(b4) / mu_ak))) + w1 * w1;
    const T h1 = (1.0 + k1) - k1 / (1.0 + x / k1);
    T gs = dconst<3>(1.0);                           // 1 - exp(-a1 / sqrt(s)) -> 1 (all derivatives 0) as s -> 0
    if (s2.v > 1e-40) gs = 1.0 - dexp(-a1 / dsqrt(dsqrt(s2)));
    T fa = dconst<3>(0.0);                           // both branches and all their derivatives vanish at alpha = 1
    if (oma.v > 1e-12) fa = dexp(-c1x * alpha / oma);
    else if (oma.v < -1e-12) fa = -dx * dexp(c2x / oma);
    const T fx = (h1 + fa * (h0 - h1)) * gs;
    const T r13 = dcbrt(r);
    const T res = (-0.7385587663820224 * r * r13) * fx;     // -(3/4) (3/pi)^(1/3) rho^(4/3)
    e = res.v; vr = res.d[0]; vs = res.d[1]; vt = res.d[2];
}
}  // namespace xc

// rho (n), grad (3, ld) -> edens (n), vrho (n), vgrad (3, ld)
__global__ void xc_unpol_kernel(XCTerms terms, int64_t n, int64_t ld, const double *__restrict__ rho,
                                const double *__restrict__ grad, double *__restrict__ edens,
                                double *__restrict__ vrho, double *__restrict__ vgrad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r = rho[i];
    double gx = 0, gy = 0, gz = 0, sigma = 0;
    if (grad) {
        gx = grad[i];
        gy = grad[ld + i];
        gz = grad[2 * ld + i];
        sigma = gx * gx + gy * gy + gz * gz;
    }
    double e = 0, vr = 0, vs = 0;
    if (r > XC_RHO_CUT) {
        for (int k = 0; k < terms.n; k++) {
            double ek = 0, vrk = 0, vsk = 0;
            switch (terms.id[k]) {
                case XC_LDA_X: xc::lda_x_unpol(r, ek, vrk); break;
                case XC_LDA_C_PW: xc::lda_c_pw_unpol(false, r, ek, vrk); break;
                case XC_LDA_C_PW_MOD: xc::lda_c_pw_unpol(true, r, ek, vrk); break;
                case XC_GGA_X_PBE: xc::gga_x_pbe_unpol(r, sigma, ek, vrk, vsk); break;
                case XC_GGA_C_PBE: {
                    double vd;
                    xc::gga_c_pbe_core(r, 0.0, sigma, ek, vrk, vd, vsk);
                    break;
                }
                default: xc::dual_unpol(terms.id[k], r, sigma, ek, vrk, vsk); break;
            }
            e += terms.coef[k] * ek;
            vr += terms.coef[k] * vrk;
            vs += terms.coef[k] * vsk;
        }
    }
    if (edens) edens[i] = e;
    if (vrho) vrho[i] = vr;
    if (vgrad) {
        vgrad[i] = 2.0 * vs * gx;
        vgrad[ld + i] = 2.0 * vs * gy;
        vgrad[2 * ld + i] = 2.0 * vs * gz;
    }
}

// spin-polarised: rho (2, ld), grad (2, 3, ld) -> edens (n), vrho (2, ld), vgrad (2, 3, ld)
// vgrad_u = 2 vs_uu grad_u + vs_ud grad_d, vgrad_d = 2 vs_dd grad_d + vs_ud grad_u (libxc.py:212-215)
__global__ void xc_pol_kernel(XCTerms terms, int64_t n, int64_t ld, const double *__restrict__ rho,
                              const double *__restrict__ grad, double *__restrict__ edens,
                              double *__restrict__ vrho, double *__restrict__ vgrad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ru = rho[i], rd = rho[ld + i];
    double gu[3] = {0, 0, 0}, gd[3] = {0, 0, 0};
    if (grad) {
        for (int d = 0; d < 3; d++) {
            gu[d] = grad[d * ld + i];
            gd[d] = grad[(3 + d) * ld + i];
        }
    }
    const double suu = gu[0] * gu[0] + gu[1] * gu[1] + gu[2] * gu[2];
    const double sdd = gd[0] * gd[0] + gd[1] * gd[1] + gd[2] * gd[2];
    const double sud = gu[0] * gd[0] + gu[1] * gd[1] + gu[2] * gd[2];
    const double rt = ru + rd;
    double e = 0, vu = 0, vd = 0, vsuu = 0, vsud = 0, vsdd = 0;
    for (int k = 0; k < terms.n; k++) {
        double ek = 0, vuk = 0, vdk = 0, suuk = 0, sudk = 0, sddk = 0;
        const int id = terms.id[k];
        if (id == XC_LDA_X || id == XC_GGA_X_PBE || id == XC_GGA_X_B88) {
            // spin scaling: E[ru, rd] = (E[2 ru] + E[2 rd]) / 2
            for (int s = 0; s < 2; s++) {
                const double r2 = 2.0 * (s ? rd : ru);
                if (!(r2 > XC_RHO_CUT)) continue;
                double es, vr, vs = 0;
                if (id == XC_LDA_X) xc::lda_x_unpol(r2, es, vr);
                else if (id == XC_GGA_X_B88) xc::dual_unpol(id, r2, 4.0 * (s ? sdd : suu), es, vr, vs);
                else xc::gga_x_pbe_unpol(r2, 4.0 * (s ? sdd : suu), es, vr, vs);
                ek += 0.5 * es;
                // d/dr_s [0.5 E(2 r_s, 4 sigma_ss)] = vr ; d/dsigma_ss = 2 vs
                if (s) { vdk = vr; sddk = 2.0 * vs; } else { vuk = vr; suuk = 2.0 * vs; }
            }
        } else if (rt > XC_RHO_CUT) {
            const double zeta = fmin(fmax((ru - rd) / rt, -1.0), 1.0);
            if (id == XC_LDA_C_PW || id == XC_LDA_C_PW_MOD) {
                const double rs = cbrt(3.0 / (4 * xc::PI * rt));
                double eps, drs, dz;
                xc::pw_eps(id == XC_LDA_C_PW_MOD, rs, zeta, eps, drs, dz);
                ek = rt * eps;
                const double common = eps - (rs / 3.0) * drs;
                vuk = common + dz * (1.0 - zeta);
                vdk = common - dz * (1.0 + zeta);
            } else if (id == XC_GGA_C_PBE) {
                double vs;
                xc::gga_c_pbe_core(rt, zeta, suu + 2.0 * sud + sdd, ek, vuk, vdk, vs);
                suuk = vs; sudk = 2.0 * vs; sddk = vs;
            } else if (id == XC_GGA_C_LYP) {
                typedef xc::Dual<5> T;
                const T r = xc::lyp_energy(xc::dvar<5>(ru, 0), xc::dvar<5>(rd, 1), xc::dvar<5>(suu, 2),
                                           xc::dvar<5>(sud, 3), xc::dvar<5>(sdd, 4));
                ek = r.v; vuk = r.d[0]; vdk = r.d[1]; suuk = r.d[2]; sudk = r.d[3]; sddk = r.d[4];
            } else if (id == XC_LDA_C_VWN || id == XC_LDA_C_VWN_RPA) {
                typedef xc::Dual<2> T;
                const T r = xc::vwn_energy(id == XC_LDA_C_VWN_RPA, xc::dvar<2>(ru, 0), xc::dvar<2>(rd, 1));
                ek = r.v; vuk = r.d[0]; vdk = r.d[1];
            }
        }
        const double c = terms.coef[k];
        e += c * ek; vu += c * vuk; vd += c * vdk;
        vsuu += c * suuk; vsud += c * sudk; vsdd += c * sddk;
    }
    if (edens) edens[i] = e;
    if (vrho) {
        vrho[i] = vu;
        vrho[ld + i] = vd;
    }
    if (vgrad) {
        for (int d = 0; d < 3; d++) {
            vgrad[d * ld + i] = 2.0 * vsuu * gu[d] + vsud * gd[d];
            vgrad[(3 + d) * ld + i] = 2.0 * vsdd * gd[d] + vsud * gu[d];
        }
    }
}

// meta-GGA: rho (n), grad (3, ld), tau (n) -> edens, vrho, vgrad = 2 (de/dsigma) grad rho, vlapl (= 0: no functional here
// depends on lapl rho), vtau = de/dtau.  LDA / GGA terms of the sum take the same formulas as xc_unpol_kernel.
__global__ void xc_mgga_unpol_kernel(XCTerms terms, int64_t n, int64_t ld, const double *__restrict__ rho,
                                     const double *__restrict__ grad, const double *__restrict__ tau,
                                     double *__restrict__ edens, double *__restrict__ vrho, double *__restrict__ vgrad,
                                     double *__restrict__ vlapl, double *__restrict__ vtau) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r = rho[i], tk = tau[i];
    const double gx = grad[i], gy = grad[ld + i], gz = grad[2 * ld + i];
    const double sigma = gx * gx + gy * gy + gz * gz;
    double e = 0, vr = 0, vs = 0, vt = 0;
    if (r > XC_RHO_CUT) {
        for (int k = 0; k < terms.n; k++) {
            double ek = 0, vrk = 0, vsk = 0, vtk = 0;
            switch (terms.id[k]) {
                case XC_LDA_X: xc::lda_x_unpol(r, ek, vrk); break;
                case XC_LDA_C_PW: xc::lda_c_pw_unpol(false, r, ek, vrk); break;
                case XC_LDA_C_PW_MOD: xc::lda_c_pw_unpol(true, r, ek, vrk); break;
                case XC_GGA_X_PBE: xc::gga_x_pbe_unpol(r, sigma, ek, vrk, vsk); break;
                case XC_GGA_C_PBE: {
                    double vd;
                    xc::gga_c_pbe_core(r, 0.0, sigma, ek, vrk, vd, vsk);
                    break;
                }
                case XC_MGGA_X_SCAN: xc::mgga_x_scan_unpol(r, sigma, tk, ek, vrk, vsk, vtk); break;
                default: xc::dual_unpol(terms.id[k], r, sigma, ek, vrk, vsk); break;
            }
            e += terms.coef[k] * ek;
            vr += terms.coef[k] * vrk;
            vs += terms.coef[k] * vsk;
            vt += terms.coef[k] * vtk;
        }
    }
    if (edens) edens[i] = e;
    if (vrho) vrho[i] = vr;
    if (vgrad) {
        vgrad[i] = 2.0 * vs * gx;
        vgrad[ld + i] = 2.0 * vs * gy;
        vgrad[2 * ld + i] = 2.0 * vs * gz;
    }
    if (vlapl) vlapl[i] = 0.0;
    if (vtau) vtau[i] = vt;
}

static int xc_pack_terms(int nterm, const int *ids, const double *coefs, XCTerms &t, bool &gga, bool allow_mgga = false) {
    QC_REQUIRE(nterm >= 1 && nterm <= XC_MAX_TERMS, "1..8 functional terms supported");
    t.n = nterm;
    gga = false;
    for (int k = 0; k < nterm; k++) {
        const int id = ids[k];
        QC_REQUIRE(id == XC_LDA_X || id == XC_LDA_C_PW || id == XC_LDA_C_PW_MOD || id == XC_GGA_X_PBE ||
                       id == XC_GGA_C_PBE || id == XC_LDA_C_VWN || id == XC_LDA_C_VWN_RPA || id == XC_GGA_X_B88 ||
                       id == XC_GGA_C_LYP || (allow_mgga && id == XC_MGGA_X_SCAN), "unknown functional id");
        gga = gga || id >= 100;
        t.id[k] = id;
        t.coef[k] = coefs[k];
    }
    return 0;
}

extern "C" int b200qc_xc_unpol(int nterm, const int *h_func_ids, const double *h_coefs, int64_t n,
                               int64_t ld, const double *rho, const double *grad, double *edens,
                               double *vrho, double *vgrad, void *stream) {
    XCTerms t;
    bool gga;
    if (int rc = xc_pack_terms(nterm, h_func_ids, h_coefs, t, gga)) return rc;
    QC_REQUIRE(!gga || grad != nullptr, "GGA functional needs the density gradient");
    if (n == 0) return 0;
    prof_begin(PROF_XC, as_stream(stream));
    xc_unpol_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(t, n, ld, rho, grad, edens, vrho,
                                                                               gga ? vgrad : nullptr);
    prof_end(as_stream(stream));
    QC_LAUNCHED(1);
    if (!gga && vgrad) QC_CHECK(cudaMemsetAsync(vgrad, 0, sizeof(double) * 3 * ld, as_stream(stream)));
    return 0;
}

extern "C" int b200qc_xc_pol(int nterm, const int *h_func_ids, const double *h_coefs, int64_t n,
                             int64_t ld, const double *rho, const double *grad, double *edens,
                             double *vrho, double *vgrad, void *stream) {
    XCTerms t;
    bool gga;
    if (int rc = xc_pack_terms(nterm, h_func_ids, h_coefs, t, gga)) return rc;
    QC_REQUIRE(!gga || grad != nullptr, "GGA functional needs the density gradient");
    if (n == 0) return 0;
    prof_begin(PROF_XC, as_stream(stream));
    xc_pol_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(t, n, ld, rho, gga ? grad : nullptr,
                                                                             edens, vrho, gga ? vgrad : nullptr);
    prof_end(as_stream(stream));
    QC_LAUNCHED(1);
    if (!gga && vgrad) QC_CHECK(cudaMemsetAsync(vgrad, 0, sizeof(double) * 6 * ld, as_stream(stream)));
    return 0;
}

// Meta-GGA functionals (family 4), unpolarised; the spin-polarised exchange follows from the spin-scaling relation on the
// host side (dqc_b200/xc/b200xc.py).  lapl is accepted for interface symmetry with libxc and not read.
extern "C" int b200qc_xc_mgga_unpol(int nterm, const int *h_func_ids, const double *h_coefs, int64_t n, int64_t ld,
                                    const double *rho, const double *grad, const double *lapl, const double *tau,
                                    double *edens, double *vrho, double *vgrad, double *vlapl, double *vtau,
                                    void *stream) {
    (void)lapl;
    XCTerms t;
    bool gga;
    if (int rc = xc_pack_terms(nterm, h_func_ids, h_coefs, t, gga, true)) return rc;
    QC_REQUIRE(rho && grad && tau, "meta-GGA functionals need rho, grad rho and tau");
    if (n == 0) return 0;
    prof_begin(PROF_XC, as_stream(stream));
    xc_mgga_unpol_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(t, n, ld, rho, grad, tau, edens, vrho,
                                                                                    vgrad, vlapl, vtau);
    prof_end(as_stream(stream));
    QC_LAUNCHED(1);
    return 0;
}
