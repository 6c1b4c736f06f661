"""ORACLE (test infrastructure only): XC energy densities in plain torch fp64, potentials by
autograd -- the reference's own default route (dqc/xc/base_xc.py:39-125, exercised against libxc
in dqc/test/test_xc.py:327-388).  The arithmetic the reference gets from libxc 6 (pylibxc2, absent
here) is restated from the published functional forms:
  lda_x       Slater exchange                      -- pinned by dqc/test/test_xc.py:390-391,416-417
  lda_c_pw    Perdew-Wang 92 (original parameters) -- pinned (rtol 1e-5) by test_xc.py:393-414
  gga_x_pbe   PBE exchange                         -- pinned (rtol 1e-5) by test_xc.py:419-425
  gga_c_pbe   PBE correlation on PW92-modified     -- PARITY UNPINNED in-tree (libxc formula restated)
  mgga_x_scan SCAN exchange (Sun et al., PRL 115, 036402) -- pinned by the closed form the reference checks libxc
              against (test_xc.py:427-455, unpolarised and by spin scaling polarised: :268-273)
Conventions follow dqc/xc/libxc.py:124-242 and libxc_wrapper.py:380-413: the energy returned is
per unit VOLUME (zk * rho); the potential bundle is value = de/drho, grad = 2 (de/dsigma) grad rho
(unpolarised) -- which is exactly d e / d(grad rho), what autograd gives.
All formulas are written spin-resolved; the unpolarised case calls them with rho/2, grad/2.
"""
import math
import torch

PI = math.pi
_PW = {  # a, alpha1, beta1..beta4 for (eps0, eps1, -alpha_c); fz20
    "pw": dict(a=(0.031091, 0.015545, 0.016887), fz20=1.709921),
    "pw_mod": dict(a=(0.0310907, 0.01554535, 0.0168869), fz20=1.709920934161365617563962776245),
}
_PW_ALPHA1 = (0.21370, 0.20548, 0.11125)
_PW_BETA = ((7.5957, 3.5876, 1.6382, 0.49294), (14.1189, 6.1977, 3.3662, 0.62517),
            (10.357, 3.6231, 0.88026, 0.49671))
PBE_KAPPA = 0.8040
PBE_BETA = 0.06672455060314922
PBE_MU = PBE_BETA * PI * PI / 3.0
PBE_GAMMA = (1.0 - math.log(2.0)) / (PI * PI)
FAMILY = {"lda_x": 1, "lda_c_pw": 1, "lda_c_pw_mod": 1, "gga_x_pbe": 2, "gga_c_pbe": 2,
          "lda_c_vwn": 1, "lda_c_vwn_rpa": 1, "gga_x_b88": 2, "gga_c_lyp": 2, "mgga_x_scan": 4}
# B88 / LYP / VWN: PARITY UNPINNED in-tree (the reference reaches them only through libxc, absent here, and its
# tests hold no numbers for them).  Restated from the original papers; cross-checked against literature atomic
# energies in tests/test_oracle_golden.py (flagged there as external, not from the reference).
_VWN5 = ((0.0310907, -0.10498, 3.72744, 12.9352), (0.01554535, -0.32500, 7.06042, 18.0578),
         (-1.0 / (6 * math.pi ** 2), -0.0047584, 1.13107, 13.0045))
_VWN_RPA = ((0.0310907, -0.409286, 13.0720, 42.7198), (0.01554535, -0.743294, 20.1231, 101.578))


def _vwn_aux(x, A, x0, b, c):
    """VWN, Can. J. Phys. 58, 1200 (1980), eq. 4.4; x = sqrt(rs)."""
    Q = math.sqrt(4 * c - b * b)
    X = x * x + b * x + c
    X0 = x0 * x0 + b * x0 + c
    at = torch.atan(Q / (2 * x + b))
    return A * (torch.log(x * x / X) + 2 * b / Q * at
                - b * x0 / X0 * (torch.log((x - x0) ** 2 / X) + 2 * (b + 2 * x0) / Q * at))


def _lda_x_unpol(rho):
    return -0.75 * (3.0 / PI) ** (1.0 / 3) * rho ** (4.0 / 3)


def _pw_g(rs, k, a):
    b1, b2, b3, b4 = _PW_BETA[k]
    den = 2 * a * (b1 * torch.sqrt(rs) + b2 * rs + b3 * rs ** 1.5 + b4 * rs * rs)
    return -2 * a * (1 + _PW_ALPHA1[k] * rs) * torch.log1p(1.0 / den)


def _pw_eps(rs, zeta, variant):
    a, fz20 = _PW[variant]["a"], _PW[variant]["fz20"]
    g0, g1, g2 = (_pw_g(rs, k, a[k]) for k in range(3))
    fz = ((1 + zeta) ** (4.0 / 3) + (1 - zeta) ** (4.0 / 3) - 2) / (2 ** (4.0 / 3) - 2)
    z4 = zeta ** 4
    return g0 + z4 * fz * (g1 - g0 + g2 / fz20) - fz * g2 / fz20


def edens_pol(name, ru, rd, gu=None, gd=None):
    """Energy per unit volume from spin densities ru, rd (n,) and gradients gu, gd (3, n)."""
    rho = ru + rd
    if name == "lda_x":
        return 0.5 * (_lda_x_unpol(2 * ru) + _lda_x_unpol(2 * rd))
    if name in ("lda_c_pw", "lda_c_pw_mod"):
        rs = (3.0 / (4 * PI * rho)) ** (1.0 / 3)
        return rho * _pw_eps(rs, (ru - rd) / rho, "pw" if name == "lda_c_pw" else "pw_mod")
    if name == "gga_x_pbe":
        def one(r2, g2):  # r2 = 2 rho_sigma, g2 = grad of it
            kf = (3 * PI * PI * r2) ** (1.0 / 3)
            s2 = (g2 * g2).sum(0) / (2 * kf * r2) ** 2
            return _lda_x_unpol(r2) * (1 + PBE_KAPPA - PBE_KAPPA / (1 + PBE_MU * s2 / PBE_KAPPA))
        return 0.5 * (one(2 * ru, 2 * gu) + one(2 * rd, 2 * gd))
    if name == "gga_c_pbe":
        g = gu + gd
        sigma = (g * g).sum(0)
        zeta = (ru - rd) / rho
        rs = (3.0 / (4 * PI * rho)) ** (1.0 / 3)
        eps = _pw_eps(rs, zeta, "pw_mod")
        phi = 0.5 * ((1 + zeta) ** (2.0 / 3) + (1 - zeta) ** (2.0 / 3))
        kf = (3 * PI * PI * rho) ** (1.0 / 3)
        ks = torch.sqrt(4 * kf / PI)
        t2 = sigma / (2 * phi * ks * rho) ** 2
        gp3 = PBE_GAMMA * phi ** 3
        A = PBE_BETA / PBE_GAMMA / torch.expm1(-eps / gp3)
        At2 = A * t2
        H = gp3 * torch.log1p(PBE_BETA / PBE_GAMMA * t2 * (1 + At2) / (1 + At2 + At2 * At2))
        return rho * (eps + H)
    if name in ("lda_c_vwn", "lda_c_vwn_rpa"):
        x = torch.sqrt((3.0 / (4 * PI * rho)) ** (1.0 / 3))
        zeta = (ru - rd) / rho
        fz = ((1 + zeta) ** (4.0 / 3) + (1 - zeta) ** (4.0 / 3) - 2) / (2 ** (4.0 / 3) - 2)
        if name == "lda_c_vwn_rpa":     # libxc lda_c_vwn_rpa: plain f(zeta) interpolation of the RPA fits
            ep, ef = (_vwn_aux(x, *p) for p in _VWN_RPA)
            return rho * (ep + (ef - ep) * fz)
        ep, ef, ac = (_vwn_aux(x, *p) for p in _VWN5)   # libxc lda_c_vwn = VWN5
        fpp0 = 4.0 / (9.0 * (2 ** (1.0 / 3) - 1))
        z4 = zeta ** 4
        return rho * (ep + ac * fz * (1 - z4) / fpp0 + (ef - ep) * fz * z4)
    if name == "gga_x_b88":
        def one(r, g):   # one spin channel (Becke, PRA 38, 3098 (1988)), beta = 0.0042
            r43 = r ** (4.0 / 3)
            x = torch.sqrt((g * g).sum(0)) / r43
            return -r43 * (1.5 * (3.0 / (4 * PI)) ** (1.0 / 3) + 0.0042 * x * x / (1 + 6 * 0.0042 * x * torch.asinh(x)))
        return one(ru, gu) + one(rd, gd)
    if name == "gga_c_lyp":
        # Miehlich, Savin, Stoll, Preuss, CPL 157, 200 (1989), eq. 2 (LYP without the Laplacian)
        a, b, c, d = 0.04918, 0.132, 0.2533, 0.349
        cf = 0.3 * (3 * PI * PI) ** (2.0 / 3)
        saa, sbb, sab = (gu * gu).sum(0), (gd * gd).sum(0), (gu * gd).sum(0)
        stot = saa + 2 * sab + sbb
        rm13 = rho ** (-1.0 / 3)
        den = 1 + d * rm13
        omega = torch.exp(-c * rm13) / den * rho ** (-11.0 / 3)
        delta = c * rm13 + d * rm13 / den
        inner = ru * rd * (2 ** (11.0 / 3) * cf * (ru ** (8.0 / 3) + rd ** (8.0 / 3))
                           + (47.0 / 18 - 7 * delta / 18) * stot - (2.5 - delta / 18) * (saa + sbb)
                           - (delta - 11) / 9 * (ru / rho * saa + rd / rho * sbb)) \
            - 2.0 / 3 * rho * rho * stot + (2.0 / 3 * rho * rho - ru * ru) * sbb + (2.0 / 3 * rho * rho - rd * rd) * saa
        return -a * 4 / den * ru * rd / rho - a * b * omega * inner
    raise KeyError(name)


def scan_x_unpol(rho, grad, tau):
    """SCAN exchange energy per volume of an unpolarised density (Sun, Ruzsinszky, Perdew 2015, eqs. 5-9 and the
    supplementary parameters): e = e_x^LDA(rho) F_x(s, alpha), s = |grad rho| / (2 rho kF),
    alpha = (tau - |grad rho|^2 / (8 rho)) / (0.3 kF^2 rho).  lapl rho does not enter."""
    a1, c1x, c2x, dx, k1, h0, b3 = 4.9479, 0.667, 0.8, 1.24, 0.065, 1.174, 0.5
    mu_ak = 10.0 / 81.0
    b2 = math.sqrt(5913.0 / 405000.0)
    b1 = 511.0 / 13500.0 / (2.0 * b2)
    b4 = mu_ak ** 2 / k1 - 1606.0 / 18225.0 - b1 ** 2
    # densities at or below 1e-15 contribute exactly zero (libxc's dens_threshold; the CUDA kernel's XC_RHO_CUT): on a
    # molecular grid rho underflows far from the nuclei and alpha would be 0 / 0
    live = rho > 1e-15
    rho_in = rho
    rho = torch.where(live, rho, torch.ones_like(rho))
    grad = torch.where(live, grad, torch.zeros_like(grad))
    tau = torch.where(live, tau, torch.ones_like(tau))
    sigma = (grad * grad).sum(0)
    kf = (3.0 * PI * PI * rho) ** (1.0 / 3.0)
    s2 = sigma / (4.0 * rho * rho * kf * kf)
    alpha = (tau - sigma / (8.0 * rho)) / (0.3 * kf * kf * rho)
    oma = 1.0 - alpha
    x = mu_ak * s2 * (1.0 + b4 * s2 / mu_ak * torch.exp(-abs(b4) * s2 / mu_ak)) \
        + (b1 * s2 + b2 * oma * torch.exp(-b3 * oma * oma)) ** 2
    h1 = 1.0 + k1 - k1 / (1.0 + x / k1)
    s2c = s2.clamp_min(1e-40)                                  # s -> 0: g -> 1 with vanishing derivatives
    gs = 1.0 - torch.exp(-a1 / s2c ** 0.25)
    pos, neg = oma > 1e-12, oma < -1e-12
    # (each branch sees harmless arguments where it is masked out: an inf there would turn into NaN gradients)
    op = torch.where(pos, oma, torch.ones_like(oma))
    on = torch.where(neg, oma, -torch.ones_like(oma))
    fa = torch.where(pos, torch.exp(-c1x * (1.0 - op) / op), torch.zeros_like(oma)) \
        - torch.where(neg, dx * torch.exp(c2x / on), torch.zeros_like(oma))
    fx = (h1 + fa * (h0 - h1)) * gs
    e = -0.75 * (3.0 / PI) ** (1.0 / 3.0) * rho ** (4.0 / 3.0) * fx
    return torch.where(live, e, 0.0 * rho_in)


def edens_unpol(name, rho, grad=None):
    half_g = None if grad is None else 0.5 * grad
    return edens_pol(name, 0.5 * rho, 0.5 * rho, half_g, half_g)


def parse(xcstr):
    """"a*f1 + f2" -> [(coef, name), ...] (the reference builds the same sum with eval,
    dqc/api/getxc.py:38-59)."""
    terms = []
    for part in xcstr.replace("-", "+-").split("+"):
        part = part.strip()
        if not part:
            continue
        coef, name = 1.0, part
        if "*" in part:
            a, b = [x.strip() for x in part.split("*")]
            try:
                coef, name = float(a), b
            except ValueError:
                coef, name = float(b), a
        terms.append((coef, name))
    return terms


def family(xcstr):
    return max(FAMILY[n] for _, n in parse(xcstr))


def eval_unpol(xcstr, rho, grad=None):
    """Returns (edens (n,), vrho (n,), vgrad (3, n) or None) for the unpolarised case."""
    fam = family(xcstr)
    rho = rho.detach().clone().requires_grad_(True)
    g = None if fam == 1 else grad.detach().clone().requires_grad_(True)
    e = sum(c * edens_unpol(n, rho, g if FAMILY[n] == 2 else None) for c, n in parse(xcstr))
    inputs = (rho,) if g is None else (rho, g)
    grads = torch.autograd.grad(e.sum(), inputs, allow_unused=True)
    vg = None
    if g is not None:
        vg = grads[1] if grads[1] is not None else torch.zeros_like(g)
    return e.detach(), grads[0].detach(), (vg.detach() if vg is not None else None)


def eval_pol(xcstr, ru, rd, gu=None, gd=None):
    """Returns (edens, (vrho_u, vrho_d), (vgrad_u, vgrad_d) or None)."""
    fam = family(xcstr)
    ru = ru.detach().clone().requires_grad_(True)
    rd = rd.detach().clone().requires_grad_(True)
    if fam == 2:
        gu = gu.detach().clone().requires_grad_(True)
        gd = gd.detach().clone().requires_grad_(True)
    e = sum(c * (edens_pol(n, ru, rd, gu, gd) if FAMILY[n] == 2 else edens_pol(n, ru, rd))
            for c, n in parse(xcstr))
    inputs = (ru, rd) if fam == 1 else (ru, rd, gu, gd)
    grads = torch.autograd.grad(e.sum(), inputs, allow_unused=True)
    grads = [x if x is not None else torch.zeros_like(i) for x, i in zip(grads, inputs)]
    if fam == 1:
        return e.detach(), (grads[0], grads[1]), None
    return e.detach(), (grads[0], grads[1]), (grads[2], grads[3])


def eval_unpol_mgga(xcstr, rho, grad, lapl, kin):
    """Family 4, unpolarised: (edens, vrho, vgrad (3, n), vlapl, vkin) -- potentials by autograd with respect to
    (rho, grad rho, lapl rho, tau) like the reference's default route."""
    rho = rho.detach().clone().requires_grad_(True)
    g = grad.detach().clone().requires_grad_(True)
    lp = lapl.detach().clone().requires_grad_(True)
    k = kin.detach().clone().requires_grad_(True)
    e = 0.0
    for c, n in parse(xcstr):
        if FAMILY[n] == 4:
            assert n == "mgga_x_scan"
            e = e + c * scan_x_unpol(rho, g, k)
        else:
            e = e + c * edens_unpol(n, rho, g if FAMILY[n] == 2 else None)
    grads = torch.autograd.grad(e.sum(), (rho, g, lp, k), allow_unused=True)
    grads = [x if x is not None else torch.zeros_like(i) for x, i in zip(grads, (rho, g, lp, k))]
    return (e.detach(),) + tuple(x.detach() for x in grads)


def eval_pol_mgga(xcstr, ru, rd, gu, gd, lu, ld, ku, kd):
    """Family 4, polarised exchange by spin scaling e[ru, rd] = (e[2 ru] + e[2 rd]) / 2 (test_xc.py:272-273):
    returns (edens, (v_u bundle), (v_d bundle)), each bundle = (vrho, vgrad, vlapl, vkin)."""
    for _, n in parse(xcstr):
        assert FAMILY[n] == 4 and "_x_" in n, "spin scaling holds for exchange functionals only"
    eu = eval_unpol_mgga(xcstr, 2 * ru, 2 * gu, 2 * lu, 2 * ku)
    ed = eval_unpol_mgga(xcstr, 2 * rd, 2 * gd, 2 * ld, 2 * kd)
    return 0.5 * (eu[0] + ed[0]), eu[1:], ed[1:]
