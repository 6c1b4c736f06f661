"""Pins the CPU oracle (oracle/) against the golden numbers the reference's own tests hold for the
Fock-build path.  Every constant below is copied from the cited reference test, nothing else.
These are CPU tests (-m "not gpu")."""
import json
import os
import numpy as np
import pytest
import torch
from tests import util
from oracle import fock_ref, scf_ref, xc_ref, cint

dtype = torch.float64
GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.json")))


def _diatomic(atomzs, dist, basis="3-21g"):
    pos = (np.array([[-0.5, 0.0, 0.0], [0.5, 0.0, 0.0]]) * dist).tolist()
    w, p = util.make_wrapper(atomzs, pos, basis)
    return w, p.numpy()


# dqc/test/test_hf.py:17-31 (RHF/3-21G from PySCF, rtol 1e-7 at :42-45)
RHF = [(c["atomzs"], c["dist"], c["energy"]) for c in GOLDEN["rhf_321g"]["cases"]]


@pytest.mark.parametrize("atomzs,dist,etrue", RHF)
def test_rhf_321g_golden(atomzs, dist, etrue):
    w, pos = _diatomic(atomzs, dist)
    h = fock_ref.RefHamilton(w).build_eri()
    e, _ = scf_ref.run_scf(h, atomzs, pos, sum(atomzs))
    assert abs(e - etrue) <= 1e-7 * abs(etrue)


# dqc/test/test_hf.py:141-153 (UHF atoms, rtol 1e-7 at :177-189)
UHF_ATOMS = [(c["z"], c["spin"], c["energy"]) for c in GOLDEN["uhf_atoms_321g"]["cases"]]


@pytest.mark.parametrize("z,spin,etrue", UHF_ATOMS)
def test_uhf_atoms_golden(z, spin, etrue):
    w, pos = util.make_wrapper([z], [[0.0, 0.0, 0.0]], "3-21g")
    h = fock_ref.RefHamilton(w).build_eri()
    e, _ = scf_ref.run_scf(h, [z], pos.numpy(), z, spin=spin)
    assert abs(e - etrue) <= 1e-7 * abs(etrue)


def test_uhf_no_golden():
    # dqc/test/test_hf.py:154-161,191-206: NO, dist 2.0, spin 1, -128.477807 Ha, rtol 1e-8
    w, pos = _diatomic([7, 8], 2.0)
    h = fock_ref.RefHamilton(w).build_eri()
    e, _ = scf_ref.run_scf(h, [7, 8], pos, 15, spin=1)
    assert abs(e - GOLDEN["uhf_no_321g"]["energy"]) <= 1e-8 * 128.477807 * 10  # constant is printed to 9 digits


def test_h2_density_points_golden():
    # dqc/test/test_hamilton.py:95-142: converged HF density of H2/3-21G ("H 0 0 0.8; H 0 0 -0.8")
    # on 5 points of the z axis, values from PySCF quoted there (default allclose tolerances)
    w, pos = util.make_wrapper([1, 1], [[0.0, 0.0, 0.8], [0.0, 0.0, -0.8]], "3-21g")
    h = fock_ref.RefHamilton(w).build_eri()
    _, dm = scf_ref.run_scf(h, [1, 1], pos.numpy(), 2)
    xyz = np.array(GOLDEN["h2_density_points"]["xyz"])
    dens = h.aodm2dens(dm, xyz).numpy()
    true = np.array(GOLDEN["h2_density_points"]["density"])
    assert np.allclose(dens, true, rtol=1e-5, atol=1e-8)


def test_vext_constant_is_identity():
    # dqc/test/test_hamilton.py:144-155: <mu| w |nu> integrated on the grid reproduces w * S
    from dqc_b200.grid.radial_grid import RadialGrid, DE2Transformation
    from dqc_b200.grid.lebedev_grid import LebedevGrid
    w, pos = util.make_wrapper([1], [[0.0, 0.0, 0.0]], "3-21g")
    atm, bas, env = w.atm_bas_env
    g = LebedevGrid(RadialGrid(99, "uniform", DE2Transformation(alpha=2.7, rmin=1e-7, rmax=15 * 1.5)), prec=41)
    ao = cint.eval_gto(atm, bas, env, g.get_rgrid().numpy(), 0)
    S = (ao * g.get_dvolume().numpy()[:, None]).T @ ao
    assert np.allclose(S, cint.int1e("ovlp", atm, bas, env), atol=4e-5)


# ---- XC analytic forms of dqc/test/test_xc.py:390-425 -------------------------------------
def test_lda_x_formula():
    rho = torch.logspace(-3, 2, 50, dtype=dtype)
    e, v, _ = xc_ref.eval_unpol("lda_x", rho)
    assert torch.allclose(e, -0.75 * (3 / np.pi) ** (1. / 3) * rho ** (4. / 3))   # test_xc.py:390-391
    assert torch.allclose(v, -(3 / np.pi) ** (1. / 3) * rho ** (1. / 3))          # test_xc.py:416-417


def test_pw92_formula():
    # test_xc.py:393-414 (ldac_e_true): PW92 with the original parameters, unpolarised and polarised
    def ldac_e_true(rhou, rhod):
        rho = rhou + rhod
        zeta = (rhou - rhod) / rho
        rs = (3. / (4 * np.pi * rho)) ** (1. / 3)

        def G(rs, A, a1, b1, b2, b3, b4, p=1):
            return -2 * A * (1 + a1 * rs) * torch.log(
                1 + 1. / (2 * A * (b1 * rs ** 0.5 + b2 * rs + b3 * rs ** 1.5 + b4 * rs ** (p + 1))))
        ec0 = G(rs, 0.031091, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294)
        ec1 = G(rs, 0.015545, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517)
        mac = G(rs, 0.016887, 0.11125, 10.357, 3.6231, 0.88026, 0.49671)
        fz = ((1 + zeta) ** (4. / 3) + (1 - zeta) ** (4. / 3) - 2) / (2 ** (4. / 3) - 2)
        fz20 = 1.709921
        return rho * (ec0 - mac * fz * (1 - zeta ** 4) / fz20 + (ec1 - ec0) * fz * zeta ** 4)
    ru = torch.logspace(-3, 1.5, 40, dtype=dtype)
    rd = ru.flip(0) * 0.7
    assert torch.allclose(xc_ref.edens_pol("lda_c_pw", ru, rd), ldac_e_true(ru, rd), rtol=1e-10)
    assert torch.allclose(xc_ref.edens_unpol("lda_c_pw", ru), ldac_e_true(ru / 2, ru / 2), rtol=1e-10)


def test_pbe_x_formula():
    # test_xc.py:419-425 (pbe_e_true)
    rho = torch.logspace(-2, 1.5, 30, dtype=dtype)
    g = torch.stack([0.3 * rho, -0.2 * rho ** 1.2, 0.5 * rho ** 0.9])
    kf = (3 * np.pi ** 2 * rho) ** (1. / 3)
    s = torch.sqrt((g * g).sum(0)) / (2 * rho * kf)
    kappa, mu = 0.804, 0.21951
    fx = 1 + kappa - kappa / (1 + mu * s * s / kappa)
    etrue = -0.75 * (3 / np.pi) ** (1. / 3) * rho ** (4. / 3) * fx
    assert torch.allclose(xc_ref.edens_unpol("gga_x_pbe", rho, g), etrue, rtol=1e-5)


# ---- B88 / LYP / VWN: no numbers in the reference's tests (it reaches them only through libxc).  The checks
# below are EXTERNAL cross-checks (literature values / a pinned functional), flagged as such in DESIGN.md.
def _radial(n=20000, rmax=40.0):
    r = torch.linspace(1e-6, rmax, n, dtype=dtype)
    w = torch.full((n,), float(r[1] - r[0]), dtype=dtype)
    w[0] *= 0.5
    w[-1] *= 0.5
    return r, 4 * np.pi * r * r * w


def test_b88_hydrogen_atom_literature():
    # Becke, PRA 38, 3098 (1988), table I: exchange energy of the H atom, exact -0.3125, LSD -0.2680, B88 -0.3098
    r, w = _radial()
    rho = torch.exp(-2 * r) / np.pi                 # fully polarised: one spin channel
    g = torch.zeros(3, r.shape[0], dtype=dtype)
    g[0] = -2 * rho
    e_b88 = 0.5 * float((w * xc_ref.edens_pol("gga_x_b88", rho, rho, g, g)).sum())   # one of two equal channels
    e_lsd = 0.5 * float((w * xc_ref.edens_pol("lda_x", rho, rho)).sum())
    assert abs(e_lsd - (-0.2680)) < 1e-4
    assert abs(e_b88 - (-0.3098)) < 1e-4


def test_lyp_one_electron_zero_and_helium_literature():
    r, w = _radial()
    # one-electron densities have no LYP correlation (Lee, Yang, Parr, PRB 37, 785 (1988))
    rho = torch.exp(-2 * r) / np.pi
    g = torch.zeros(3, r.shape[0], dtype=dtype)
    g[0] = -2 * rho
    tiny = torch.full_like(rho, 1e-300)
    assert abs(float((w * xc_ref.edens_pol("gga_c_lyp", rho, tiny, g, torch.zeros_like(g))).sum())) < 1e-12
    # He: E_c(LYP) = -0.0437 on the HF density (same paper, table I); the hydrogenic zeta = 27/16 density used
    # here is close to it (-0.0439)
    zt = 27.0 / 16
    rho = 2 * zt ** 3 / np.pi * torch.exp(-2 * zt * r)
    g[0] = -2 * zt * rho
    assert abs(float((w * xc_ref.edens_unpol("gga_c_lyp", rho, g)).sum()) - (-0.0437)) < 5e-4


def test_vwn_against_pinned_pw92_and_rpa_values():
    # VWN5 and PW92 are fits of the same Ceperley-Alder data: para- and ferromagnetic energies agree to < 6e-4 Ha
    for rs in (0.5, 1.0, 2.0, 5.0, 10.0, 50.0):
        rho = torch.tensor([3 / (4 * np.pi * rs ** 3)], dtype=dtype)
        zero = torch.tensor([1e-300], dtype=dtype)
        for args in ((rho / 2, rho / 2), (rho, zero)):
            e_vwn = float(xc_ref.edens_pol("lda_c_vwn", *args) / rho)
            e_pw = float(xc_ref.edens_pol("lda_c_pw", *args) / rho)
            assert abs(e_vwn - e_pw) < 6e-4
    # RPA correlation energy per electron of the unpolarised gas (VWN 1980, table 5 region): rs = 1 -> -0.079 Ha
    rho = torch.tensor([3 / (4 * np.pi)], dtype=dtype)
    assert abs(float(xc_ref.edens_unpol("lda_c_vwn_rpa", rho) / rho) - (-0.0793)) < 1e-3
    # both interpolations reduce to the paramagnetic fit at zeta = 0
    assert float(xc_ref.edens_pol("lda_c_vwn", rho / 2, rho / 2)) == pytest.approx(float(xc_ref.edens_unpol("lda_c_vwn", rho)))


@pytest.mark.parametrize("name", ["lda_x", "lda_c_pw", "lda_c_pw_mod", "gga_x_pbe", "gga_c_pbe",
                                  "lda_c_vwn", "lda_c_vwn_rpa", "gga_x_b88", "gga_c_lyp"])
def test_potential_is_derivative_of_energy(name):
    # v = de/drho checked by central finite differences (fp64)
    rho = torch.logspace(-2, 1, 12, dtype=dtype)
    g = torch.stack([0.4 * rho, 0.1 * rho, -0.3 * rho])
    fam = xc_ref.FAMILY[name]
    _, v, vg = xc_ref.eval_unpol(name, rho, g if fam == 2 else None)
    h = 1e-6 * rho
    ep = xc_ref.edens_unpol(name, rho + h, g if fam == 2 else None)
    em = xc_ref.edens_unpol(name, rho - h, g if fam == 2 else None)
    assert torch.allclose(v, (ep - em) / (2 * h), rtol=1e-6, atol=1e-9)


# ---- KS energies of dqc/test/test_ks.py:40-111 (H2, 6-311++G**, PySCF values, atol 1.3e-3 there because the grids
# differ): pins the oracle's whole XC chain -- atomic grids, Becke weights, AO values, rho / grad rho, functional,
# Vxc integration, SCF -- the same numbers the CUDA path reproduces in tests/test_gpu_hamilton.py
@pytest.mark.parametrize("xc,etrue", [("lda_x", -0.979143262), ("gga_x_pbe", -1.068217310366847)])
def test_rks_h2_golden_oracle(xc, etrue):
    from dqc_b200.grid.factory import get_predefined_grid
    from oracle import becke_ref
    pos = np.array([[-0.5, 0.0, 0.0], [0.5, 0.0, 0.0]])
    w, _ = util.make_wrapper([1, 1], pos.tolist(), "6-311++G**")
    one = get_predefined_grid("sg2", [1], torch.zeros(1, 3, dtype=dtype), device=torch.device("cpu"))
    n1 = one.get_rgrid().shape[0]
    pts = np.concatenate([one.get_rgrid().numpy() + p for p in pos])
    dvol = np.concatenate([one.get_dvolume().numpy()] * 2) * becke_ref.becke_weights(pts, np.repeat([0, 1], n1), pos)
    h = fock_ref.RefHamilton(w).build_eri()
    h.setup_grid(pts, dvol, xc)
    e, _ = scf_ref.run_scf(h, [1, 1], pos, 2, method="ks")
    assert abs(e - etrue) < 1.3e-3


@pytest.mark.parametrize("name", ["gga_x_pbe", "gga_c_pbe", "gga_x_b88", "gga_c_lyp"])
def test_gradient_potential_is_derivative_of_energy(name):
    # the gradient part of the potential, d e / d(grad rho) = 2 (de/dsigma) grad rho, by central
    # finite differences of the energy density (unpolarised and one spin channel of the polarised form)
    rho = torch.logspace(-2, 1, 10, dtype=dtype)
    g = torch.stack([0.4 * rho, 0.1 * rho ** 1.1, -0.3 * rho ** 0.9])
    _, _, vg = xc_ref.eval_unpol(name, rho, g)
    for d in range(3):
        h = 1e-6 * (g[d].abs() + 1e-3)
        gp, gm = g.clone(), g.clone()
        gp[d] += h
        gm[d] -= h
        fd = (xc_ref.edens_unpol(name, rho, gp) - xc_ref.edens_unpol(name, rho, gm)) / (2 * h)
        assert torch.allclose(vg[d], fd, rtol=1e-6, atol=1e-9)
    ru, rd = 0.7 * rho, 0.3 * rho
    gu, gd = 0.6 * g, 0.4 * g.flip(0)
    _, (vu, vd), (vgu, vgd) = xc_ref.eval_pol(name, ru, rd, gu, gd)
    h = 1e-6 * ru
    fd = (xc_ref.edens_pol(name, ru + h, rd, gu, gd) - xc_ref.edens_pol(name, ru - h, rd, gu, gd)) / (2 * h)
    assert torch.allclose(vu, fd, rtol=1e-6, atol=1e-9)
    h = 1e-6 * (gd[1].abs() + 1e-3)
    gp, gm = gd.clone(), gd.clone()
    gp[1] += h
    gm[1] -= h
    fd = (xc_ref.edens_pol(name, ru, rd, gu, gp) - xc_ref.edens_pol(name, ru, rd, gu, gm)) / (2 * h)
    assert torch.allclose(vgd[1], fd, rtol=1e-6, atol=1e-9)


# ---- embedded basis tables against EXTERNAL literature energies (not reference-held; flagged in DESIGN.md) ----
@pytest.mark.parametrize("case", GOLDEN["external_literature_rhf"]["cases"], ids=lambda c: c["name"])
def test_literature_rhf_energies_verify_basis_tables(case):
    """The sto-3g (H, O) and cc-pvdz (H, O) tables under dqc_b200/data/basis reproduce published RHF energies of
    water; def2-svp has no energy known offline to better than 1e-3 and stays 'typed, unverified by energy'."""
    pos = np.array(case["pos"]) * (1.0 if case["unit"] == "bohr" else 1.0 / 0.52917721092)
    w, p = util.make_wrapper(case["atomzs"], pos.tolist(), case["basis"])
    if case["enuc"] is not None:
        assert abs(fock_ref.nuclei_energy(case["atomzs"], pos) - case["enuc"]) < 1e-10
    h = fock_ref.RefHamilton(w).build_eri()
    e, _ = scf_ref.run_scf(h, case["atomzs"], pos, sum(case["atomzs"]))
    assert abs(e - case["energy"]) < case["atol"]


def test_pbe_c_formula_independent_restatement():
    """gga_c_pbe against the formula of Perdew, Burke, Ernzerhof, PRL 77, 3865 (1996), eqs. 3, 7, 8, written out
    here independently of oracle/xc_ref.py (EXTERNAL cross-check: the reference reaches PBE-c only through libxc).
    libxc's gga_c_pbe uses beta = 0.06672455060314922, gamma = (1 - ln 2) / pi^2 and the PW92 'mod' parameters."""
    rho = torch.logspace(-3, 1.5, 40, dtype=dtype)
    g = torch.stack([0.3 * rho ** 1.1, -0.2 * rho ** 1.2, 0.5 * rho ** 0.9])
    gnorm = torch.sqrt((g * g).sum(0))
    rs = (3. / (4 * np.pi * rho)) ** (1. / 3)
    A_, a1, b1, b2, b3, b4 = 0.0310907, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294
    ec = -2 * A_ * (1 + a1 * rs) * torch.log(1 + 1. / (2 * A_ * (b1 * rs ** 0.5 + b2 * rs + b3 * rs ** 1.5 + b4 * rs ** 2)))
    beta, gamma = 0.06672455060314922, (1 - np.log(2.0)) / np.pi ** 2
    kf = (3 * np.pi ** 2 * rho) ** (1. / 3)
    ks = torch.sqrt(4 * kf / np.pi)
    t = gnorm / (2 * ks * rho)                      # phi = 1 for the unpolarised gas
    A = beta / gamma / (torch.exp(-ec / gamma) - 1)
    H = gamma * torch.log(1 + beta / gamma * t ** 2 * (1 + A * t ** 2) / (1 + A * t ** 2 + A ** 2 * t ** 4))
    assert torch.allclose(xc_ref.edens_unpol("gga_c_pbe", rho, g), rho * (ec + H), rtol=1e-9)


# ---- meta-GGA: SCAN exchange pinned by the closed form the reference checks libxc against
def _scan_e_true(rho, gradn, lapl, tau):
    # dqc/test/test_xc.py:427-455 (scan_e_true), restated here independently of oracle/xc_ref.py
    kf = (3 * np.pi * np.pi * rho) ** (1. / 3)
    norm_gradn = torch.sqrt((gradn * gradn).sum(0))
    s = norm_gradn / (2 * rho * kf)
    tau_w = norm_gradn ** 2 / (8 * rho)
    tau_unif = 0.3 * kf ** 2 * rho
    alpha = (tau - tau_w) / tau_unif
    s2 = s * s
    a1, c1x, c2x, dx = 4.9479, 0.667, 0.8, 1.24
    mu_ak = 10. / 81
    b2 = (5913 / 405000.) ** 0.5
    b1 = 511 / 13500 / (2 * b2)
    b3, k1 = 0.5, 0.065
    b4 = mu_ak ** 2 / k1 - 1606 / 18225 - b1 ** 2
    x = mu_ak * s2 * (1 + (b4 * s2 / mu_ak) * torch.exp(-abs(b4) * s2 / mu_ak)) + \
        (b1 * s2 + b2 * (1 - alpha) * torch.exp(-b3 * (1 - alpha) ** 2)) ** 2
    h1 = 1 + k1 * (1 - k1 / (k1 + x))
    h0 = 1.174
    gs = 1 - torch.exp(-a1 / torch.sqrt(s))
    theta_1ma = ((1 - alpha) > 0) * 1.0
    theta_am1 = ((alpha - 1) > 0) * 1.0
    fa = torch.exp(-c1x * alpha / (1 - alpha)) * theta_1ma - dx * torch.exp(c2x / (1 - alpha)) * theta_am1
    return -0.75 * (3 / np.pi) ** (1. / 3) * rho ** (4. / 3) * (h1 + fa * (h0 - h1)) * gs


def _mgga_points(n=60, seed=3):
    g = torch.Generator().manual_seed(seed)
    rho = 10 ** (torch.rand(n, dtype=dtype, generator=g) * 4 - 3)
    grad = torch.randn(3, n, dtype=dtype, generator=g) * rho ** (4. / 3) * 1.5
    tau_w = (grad * grad).sum(0) / (8 * rho)
    tau_unif = 0.3 * (3 * np.pi ** 2 * rho) ** (2. / 3) * rho
    alpha = torch.cat([torch.rand(n // 2, dtype=dtype, generator=g) * 0.95,          # both branches of f(alpha)
                       1.05 + torch.rand(n - n // 2, dtype=dtype, generator=g) * 3])
    tau = tau_w + alpha * tau_unif
    lapl = torch.randn(n, dtype=dtype, generator=g)
    return rho, grad, lapl, tau


def test_scan_x_formula():
    # unpolarised and, by spin scaling, polarised (test_xc.py:268-273)
    rho, grad, lapl, tau = _mgga_points()
    e = xc_ref.eval_unpol_mgga("mgga_x_scan", rho, grad, lapl, tau)[0]
    assert torch.allclose(e, _scan_e_true(rho, grad, lapl, tau), rtol=1e-12)
    ru, rd, gu, gd = 0.7 * rho, 0.3 * rho, 0.6 * grad, 0.4 * grad
    ku, kd = 0.65 * tau, 0.35 * tau
    ep = xc_ref.eval_pol_mgga("mgga_x_scan", ru, rd, gu, gd, 0.5 * lapl, 0.5 * lapl, ku, kd)[0]
    etrue = 0.5 * (_scan_e_true(2 * ru, 2 * gu, lapl, 2 * ku) + _scan_e_true(2 * rd, 2 * gd, lapl, 2 * kd))
    ok = torch.isfinite(etrue)       # (the reference's expression is inf * 0 where alpha is within ~1e-3 above 1)
    assert int(ok.sum()) >= ok.numel() - 2 and torch.allclose(ep[ok], etrue[ok], rtol=1e-9)


def test_scan_potentials_are_derivatives_of_the_energy():
    rho, grad, lapl, tau = _mgga_points(40, seed=5)
    e, vr, vg, vl, vk = xc_ref.eval_unpol_mgga("mgga_x_scan", rho, grad, lapl, tau)
    f = lambda r, g, k: xc_ref.scan_x_unpol(r, g, k)
    h = 1e-6
    assert torch.allclose(vr, (f(rho * (1 + h), grad, tau) - f(rho * (1 - h), grad, tau)) / (2 * h * rho), rtol=2e-6, atol=1e-9)
    assert torch.allclose(vk, (f(rho, grad, tau * (1 + h)) - f(rho, grad, tau * (1 - h))) / (2 * h * tau), rtol=2e-6, atol=1e-9)
    for d in range(3):
        dg = torch.zeros_like(grad)
        dg[d] = h * grad[d].abs().clamp_min(1e-3)
        fd = (f(rho, grad + dg, tau) - f(rho, grad - dg, tau)) / (2 * dg[d])
        assert torch.allclose(vg[d], fd, rtol=5e-6, atol=1e-7)       # (atol: round-off of the difference quotient)
    assert float(vl.abs().max()) == 0.0              # SCAN does not depend on lapl rho
