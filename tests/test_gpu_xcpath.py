"""GPU parity of the XC grid path (K1 AO eval, Becke weights, K2 density, K3 functional, K4 Vxc)
against the CPU oracle, all through the C-ABI (dqc_b200._lib -> libb200qc.so)."""
import numpy as np
import pytest
import torch
from tests import util

pytestmark = pytest.mark.gpu


def _pad_dm(dm, ld, dev):
    out = torch.zeros(ld, ld, dtype=torch.float64, device=dev)
    out[:dm.shape[0], :dm.shape[1]] = dm.to(dev)
    return out


@pytest.mark.parametrize("which", ["h2o-def2svp", "highl", "c-etbjfit"])
@pytest.mark.parametrize("ngrid", [1, 77, 1000])
def test_ao_eval_matches_oracle(cuda, which, ngrid):
    from dqc_b200 import _lib
    from oracle import cint
    if which == "h2o-def2svp":
        w, _ = util.make_wrapper(*util.H2O, "def2-svp")
    elif which == "highl":
        w, _ = util.highl_wrapper()
    else:
        w, _ = util.make_wrapper([6], [[0.1, 0.2, -0.3]], "etb-jfit")
    atm, bas, env = w.atm_bas_env
    pts = util.random_points(ngrid, seed=ngrid)
    db = w.device_basis(cuda)
    for deriv in (0, 1, 2):
        ao = _lib.eval_gto(db, 0, len(w), torch.tensor(pts, device=cuda), deriv)
        torch.cuda.synchronize()
        nao = w.nao()
        ref0 = cint.eval_gto(atm, bas, env, pts, 0)
        got = ao[:, :ngrid, :nao].cpu().numpy()
        assert np.abs(got[0] - ref0).max() < 1e-12
        if deriv:
            ref1 = cint.eval_gto(atm, bas, env, pts, 1)
            assert np.abs(got[1:4] - ref1).max() < 1e-11
        if deriv == 2:          # Laplacian of the AOs (eval_laplgto, the meta-GGA storage)
            ref2 = cint.eval_gto(atm, bas, env, pts, 2)
            assert np.abs(got[4] - ref2).max() < 1e-10 * max(1.0, np.abs(ref2).max())
        # padding stays zero
        assert float(ao[:, ngrid:, :].abs().max()) == 0.0 if ao.shape[1] > ngrid else True
        assert float(ao[:, :, nao:].abs().max()) == 0.0 if ao.shape[2] > nao else True


@pytest.mark.parametrize("adjust,radii", [("becke", True), ("treutler", True), ("becke", False)])
def test_becke_weights_match_oracle(cuda, adjust, radii):
    from dqc_b200 import _lib
    from dqc_b200.utils.periodictable import atom_expected_radii
    from oracle import becke_ref
    zs, pos = util.CH4ISH
    pos = np.array(pos)
    rng = np.random.RandomState(3)
    owner = np.repeat(np.arange(len(zs)), 700)
    # points scattered around their owner, some far away (so that the 0.74 cut-off triggers)
    xyz = pos[owner] + rng.normal(size=(len(owner), 3)) * rng.choice([0.3, 1.5, 4.0], size=(len(owner), 1))
    rad = np.array([atom_expected_radii[z] for z in zs]) if radii else None
    ref = becke_ref.becke_weights(xyz, owner, pos, rad, adjust)
    aij = None
    if radii:
        r = torch.tensor(rad if adjust == "becke" else rad ** 0.5, device=cuda)
        u = (r - r.unsqueeze(1)) / (r + r.unsqueeze(1))
        aij = torch.clamp(u / (u * u - 1), min=-0.45, max=0.45).contiguous()
    got = _lib.becke_weights(torch.tensor(xyz, device=cuda), torch.tensor(owner, device=cuda),
                             torch.tensor(pos, device=cuda), aij).cpu().numpy()
    assert (ref == 0).sum() > 0  # the cut-off is exercised
    assert np.abs(got - ref).max() < 1e-13


XCS = ["lda_x", "lda_c_pw", "lda_c_pw_mod", "gga_x_pbe", "gga_c_pbe", "lda_x + lda_c_pw",
       "gga_x_pbe + gga_c_pbe", "0.7*gga_x_pbe + 0.3*lda_x + gga_c_pbe",
       "lda_c_vwn", "lda_c_vwn_rpa", "gga_x_b88", "gga_c_lyp",
       "0.08*lda_x + 0.72*gga_x_b88 + 0.19*lda_c_vwn_rpa + 0.81*gga_c_lyp"]     # the semi-local part of B3LYP


@pytest.mark.parametrize("xcstr", XCS)
def test_xc_unpol_matches_autograd_oracle(cuda, xcstr):
    from dqc_b200 import _lib
    from oracle import xc_ref
    g = torch.Generator().manual_seed(1)
    n = 4097
    rho = 10 ** (torch.rand(n, dtype=torch.float64, generator=g) * 9 - 7)   # 1e-7 .. 1e2
    grad = torch.randn(3, n, dtype=torch.float64, generator=g) * rho ** (4.0 / 3) * 2
    fam = xc_ref.family(xcstr)
    e_ref, vr_ref, vg_ref = xc_ref.eval_unpol(xcstr, rho, grad if fam == 2 else None)
    e, vr, vg = _lib.xc_unpol(xc_ref.parse(xcstr), rho.to(cuda), grad.to(cuda).contiguous() if fam == 2 else None)
    rel = lambda a, b: float(((a.cpu() - b).abs() / (b.abs() + 1e-12 * b.abs().max())).max())
    assert rel(e, e_ref) < 1e-11
    assert rel(vr, vr_ref) < 1e-10
    if fam == 2:
        assert float((vg.cpu() - vg_ref).abs().max() / vg_ref.abs().max()) < 1e-11
        assert rel(vg, vg_ref) < 1e-8


@pytest.mark.parametrize("xcstr", XCS)
def test_xc_pol_matches_autograd_oracle(cuda, xcstr):
    from dqc_b200 import _lib
    from oracle import xc_ref
    g = torch.Generator().manual_seed(2)
    n = 3001
    rho = 10 ** (torch.rand(2, n, dtype=torch.float64, generator=g) * 8 - 6)
    grad = torch.randn(2, 3, n, dtype=torch.float64, generator=g) * (rho ** (4.0 / 3)).unsqueeze(1) * 2
    fam = xc_ref.family(xcstr)
    args = (rho[0], rho[1]) + ((grad[0], grad[1]) if fam == 2 else ())
    e_ref, (vu, vd), vg_ref = xc_ref.eval_pol(xcstr, *args)
    e, vr, vg = _lib.xc_pol(xc_ref.parse(xcstr), rho.to(cuda).contiguous(),
                            grad.to(cuda).contiguous() if fam == 2 else None)
    rel = lambda a, b: float(((a.cpu() - b).abs() / (b.abs() + 1e-12 * b.abs().max())).max())
    # LYP is a difference of large terms: both implementations carry ~1e-10 of cancellation noise at
    # strongly polarised, steep-gradient points
    loose = 100.0 if "lyp" in xcstr else 1.0
    assert rel(e, e_ref) < 1e-11 * loose
    assert rel(vr[0], vu) < 1e-9 * loose and rel(vr[1], vd) < 1e-9 * loose
    if fam == 2:
        for s in range(2):
            assert float((vg[s].cpu() - vg_ref[s]).abs().max() / vg_ref[s].abs().max()) < 1e-10


def test_xc_pol_reduces_to_unpol(cuda):
    from dqc_b200 import _lib
    g = torch.Generator().manual_seed(5)
    n = 999
    rho = 10 ** (torch.rand(n, dtype=torch.float64, generator=g) * 6 - 4)
    grad = torch.randn(3, n, dtype=torch.float64, generator=g) * rho
    terms = [(1.0, "gga_x_pbe"), (1.0, "gga_c_pbe"), (0.5, "lda_c_pw")]
    e, vr, vg = _lib.xc_unpol(terms, rho.to(cuda), grad.to(cuda))
    rp = torch.stack([rho / 2, rho / 2]).to(cuda)
    gp = torch.stack([grad / 2, grad / 2]).to(cuda)
    e2, vr2, vg2 = _lib.xc_pol(terms, rp, gp)
    assert torch.allclose(e, e2, rtol=1e-12, atol=0)
    assert torch.allclose(vr, vr2[0], rtol=1e-11, atol=1e-300)
    assert torch.allclose(vg, vg2[0], rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("which,xcstr", [("h2o-def2svp", "lda_x + lda_c_pw"), ("h2o-def2svp", "gga_x_pbe + gga_c_pbe"),
                                         ("highl", "gga_x_pbe")])
def test_rho_and_vxc_match_oracle(cuda, which, xcstr):
    """K2 + K3 + K4 against oracle/fock_ref.py on a few thousand points (identity orthogonaliser so
    that the kernels themselves are compared; the basis change is plain torch on both sides)."""
    from dqc_b200 import _lib
    from oracle import fock_ref, xc_ref
    w, _ = util.make_wrapper(*util.H2O, "def2-svp") if which == "h2o-def2svp" else util.highl_wrapper()
    ngrid = 3000
    pts = util.random_points(ngrid, seed=11, span=2.5)
    wts = np.random.RandomState(5).uniform(0.0, 0.3, ngrid)
    h = fock_ref.RefHamilton(w, orthozer=False)
    h.setup_grid(pts, wts, xcstr)
    nao = w.nao()
    dm = util.seeded_dm(nao, max(1, nao // 4), seed=3)
    rho_ref, grad_ref = h.dm2densinfo(dm)
    fam = xc_ref.family(xcstr)

    db = w.device_basis(cuda)
    ao = _lib.eval_gto(db, 0, len(w), torch.tensor(pts, device=cuda), 1 if fam == 2 else 0)
    ld = ao.shape[2]
    rho, grad = _lib.rho(ao, _pad_dm(dm, ld, cuda), fam == 2)
    assert float((rho[:ngrid].cpu() - rho_ref).abs().max()) < 1e-11 * max(1.0, float(rho_ref.abs().max()))
    if fam == 2:
        assert float((grad[:, :ngrid].cpu() - grad_ref).abs().max()) < 1e-10 * max(1.0, float(grad_ref.abs().max()))
    # potentials from the oracle densities (so K4 is tested in isolation), then the full chain
    _, vrho_ref, vgrad_ref = xc_ref.eval_unpol(xcstr, rho_ref, grad_ref)
    mat_ref = h.vxc_from_potinfo(vrho_ref, vgrad_ref)   # includes (M + M^T)/2
    ngl = ao.shape[1]
    wpad = torch.zeros(ngl, dtype=torch.float64, device=cuda)
    wpad[:ngrid] = torch.tensor(wts, device=cuda)
    vr = torch.zeros(ngl, dtype=torch.float64, device=cuda)
    vr[:ngrid] = vrho_ref.to(cuda)
    vg = None
    if fam == 2:
        vg = torch.zeros(3, ngl, dtype=torch.float64, device=cuda)
        vg[:, :ngrid] = vgrad_ref.to(cuda)
    mat = _lib.vxc_mat(ao, wpad, vr, vg)[:nao, :nao]
    mat = 0.5 * (mat + mat.T)
    assert float((mat.cpu() - mat_ref).abs().max()) < 1e-10
    # full chain on the device
    e, vr2, vg2 = _lib.xc_unpol(xc_ref.parse(xcstr), rho, grad)
    mat2 = _lib.vxc_mat(ao, wpad, vr2, vg2)[:nao, :nao]
    mat2 = 0.5 * (mat2 + mat2.T)
    assert float((mat2.cpu() - mat_ref).abs().max()) < 1e-9
    exc = float((e * wpad).sum())
    exc_ref = float((xc_ref.eval_unpol(xcstr, rho_ref, grad_ref)[0] * torch.tensor(wts)).sum())
    assert abs(exc - exc_ref) < 1e-10


def test_gemm_engine_large_random(cuda):
    """K2/K4 at a size with many tiles and odd padding, against torch fp64 matmul on the device
    (plain library GEMM used only as a checker here)."""
    from dqc_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(0)
    ngl, ld, nao = 128 * 37, 64 * 5, 300
    ao = torch.zeros(4, ngl, ld, dtype=torch.float64)
    ao[:, :, :nao] = torch.randn(4, ngl, nao, dtype=torch.float64, generator=g)
    ao = ao.to(cuda)
    dm = torch.zeros(ld, ld, dtype=torch.float64)
    d0 = torch.randn(nao, nao, dtype=torch.float64, generator=g)
    dm[:nao, :nao] = d0 + d0.T
    dm = dm.to(cuda)
    rho, grad = _lib.rho(ao, dm, True)
    x = ao[0] @ dm
    assert torch.allclose(rho, (x * ao[0]).sum(-1), rtol=1e-12, atol=1e-9)
    for d in range(3):
        assert torch.allclose(grad[d], 2 * (x * ao[d + 1]).sum(-1), rtol=1e-12, atol=1e-9)
    w = torch.rand(ngl, dtype=torch.float64, device=cuda)
    vr = torch.randn(ngl, dtype=torch.float64, device=cuda)
    vg = torch.randn(3, ngl, dtype=torch.float64, device=cuda)
    mat = _lib.vxc_mat(ao, w, vr, vg)
    vb = vr[:, None] * ao[0] + 2 * (vg[0][:, None] * ao[1] + vg[1][:, None] * ao[2] + vg[2][:, None] * ao[3])
    ref = (ao[0] * w[:, None]).T @ vb
    assert float((mat - ref).abs().max() / ref.abs().max()) < 1e-13
    mat1 = _lib.vxc_mat(ao[:1].contiguous(), w, vr, None)
    ref1 = (ao[0] * w[:, None]).T @ (vr[:, None] * ao[0])
    assert float((mat1 - ref1).abs().max() / ref1.abs().max()) < 1e-13


@pytest.mark.parametrize("eps,sbp", [(0.0, 1024), (1e-12, 1024), (1e-12, 128), (1e-8, 256)])
@pytest.mark.parametrize("gga", [False, True])
def test_superblock_path_matches_dense_kernels(cuda, eps, sbp, gga):
    """Block-sparse K1/K2/K4 (GridBlocks) against the dense kernels on a spread-out molecule where the
    screening really drops shells: identical at eps = 0, within O(eps) otherwise."""
    from dqc_b200 import _lib
    from dqc_b200.grid.factory import get_predefined_grid
    zs = [6, 1, 8, 7, 1, 6]
    pos = [[0.0, 0.0, 0.0], [1.9, 0.3, -0.4], [9.0, 1.0, 0.5], [-8.0, -6.0, 2.0], [14.0, -3.0, -1.0], [0.5, 11.0, 7.0]]
    w, _ = util.make_wrapper(zs, pos, "def2-svp")
    nao, nb = w.nao(), len(w)
    grid = get_predefined_grid("sg2", zs, torch.tensor(pos, dtype=torch.float64, device=cuda), device=cuda)
    xyz, wts = grid.get_rgrid()[::7].contiguous(), grid.get_dvolume()[::7].contiguous()   # ~6 k points, atom-ordered
    ng = xyz.shape[0]
    db = w.device_basis(cuda)
    gb = _lib.GridBlocks(db, 0, nb, xyz, wts, 1 if gga else 0, sbp=sbp, eps=eps)
    if eps == 0.0:
        assert gb.kept_fraction == 1.0
    else:
        assert gb.kept_fraction < 0.95           # the screening is exercised
    ao = _lib.eval_gto(db, 0, nb, xyz, 1 if gga else 0)
    tol = 1e-13 if eps == 0.0 else 200 * eps
    dense = gb.dense_ao()
    # eps = 0: the same per-shell arithmetic in two kernels that are compiled with different register budgets (the
    # compiler's FMA contraction may differ): agreement to the last bits, not bitwise
    assert float((dense - ao[:, :ng, :nao]).abs().max()) <= (1e-14 if eps == 0.0 else 2 * eps * 10)
    dm = util.seeded_dm(nao, nao // 3, seed=4).to(cuda)
    rho_d, grad_d = _lib.rho(ao, _pad_dm(dm, ao.shape[2], cuda), gga)
    rho_s, grad_s = gb.rho(dm, gga)
    assert float((rho_s[:ng] - rho_d[:ng]).abs().max()) < tol * 10
    assert float(rho_s[ng:].abs().max()) == 0.0 if gb.ngl > ng else True
    if gga:
        assert float((grad_s[:, :ng] - grad_d[:, :ng]).abs().max()) < tol * 100
    g = torch.Generator().manual_seed(1)
    vr = torch.randn(ng, dtype=torch.float64, generator=g).to(cuda)
    vg = torch.randn(3, ng, dtype=torch.float64, generator=g).to(cuda) if gga else None

    def pad(t, n):
        out = torch.zeros(*t.shape[:-1], n, dtype=torch.float64, device=cuda)
        out[..., :ng] = t
        return out
    m_d = _lib.vxc_mat(ao, pad(wts, ao.shape[1]), pad(vr, ao.shape[1]), pad(vg, ao.shape[1]) if gga else None)[:nao, :nao]
    m_s = gb.vxc_mat(pad(vr, gb.ngl), pad(vg, gb.ngl) if gga else None)
    scale = float(m_d.abs().max())
    assert float((m_s - m_d).abs().max()) < max(tol * 100, 1e-12 * scale)


@pytest.mark.parametrize("nslice,tol", [(6, 5e-11), (5, 5e-9)])
@pytest.mark.parametrize("gga", [False, True])
@pytest.mark.parametrize("fused", [False, True])
def test_vxc_tcgen05_int8_matches_fp64_path(cuda, nslice, tol, gga, fused, monkeypatch):
    """K4 on tcgen05 (error-free sliced int8 GEMM, int32 TMEM accumulators) against the fp64 DMMA form of
    the same contraction on a real molecular grid: the slicing error bound is ~1e-12 (6 slices) / ~1e-10 (5).
    fused = the one-pass operand preparation (vb cut into int8 planes as it is formed, exponents from a bound that
    is verified while cutting; blocks whose bound came out more than 2^4 too large are cut again): at most 4 bits
    (a factor 16) behind the two-pass form with exact column maxima -- on a potential that is uncorrelated from
    point to point, the worst case for the bound."""
    from dqc_b200 import _lib
    from dqc_b200.utils import systems
    from dqc_b200.utils.config import config as cfg
    from dqc_b200.grid.factory import get_predefined_grid
    monkeypatch.setattr(cfg, "VXC_FUSED_VB", fused)
    if fused:
        tol = tol * 16
    zs, pos = systems.benzene()
    w, _ = util.make_wrapper(zs, pos.tolist(), "def2-svp")
    nb = len(w)
    grid = get_predefined_grid("sg2", zs, torch.tensor(pos, device=cuda), device=cuda)
    xyz, wts = grid.get_rgrid(), grid.get_dvolume()
    db = w.device_basis(cuda)
    deriv = 1 if gga else 0
    ref = _lib.GridBlocks(db, 0, nb, xyz, wts, deriv, sbp=512, eps=1e-12, i8_slices=0)
    gb = _lib.GridBlocks(db, 0, nb, xyz, wts, deriv, sbp=512, eps=1e-12, i8_slices=nslice)
    g = torch.Generator().manual_seed(0)
    # potentials spanning many orders of magnitude, like a real vxc on a molecular grid
    vr = (torch.randn(ref.ngl, dtype=torch.float64, generator=g) *
          10 ** (torch.rand(ref.ngl, dtype=torch.float64, generator=g) * 6 - 5)).to(cuda)
    vg = (torch.randn(3, ref.ngl, dtype=torch.float64, generator=g) * 0.3).to(cuda) if gga else None
    m_ref = ref.vxc_mat(vr, vg)
    m = gb.vxc_mat(vr, vg)
    scale = float(m_ref.abs().max())
    assert float((m - m_ref).abs().max()) < tol * max(scale, 1.0)
    assert (gb.colmax is not None) == fused
    m2 = gb.vxc_mat(vr, vg)                       # repeatable (atomics order noise only)
    assert float((m - m2).abs().max()) < 1e-13 * max(scale, 1.0)
    if fused:
        # the repair pass did run on this potential, and only on part of the blocks
        nfix = int(gb.fixflag.sum())
        assert 0 < nfix < gb.fixflag.numel()


@pytest.mark.parametrize("nslice,tol", [(6, 2e-10), (5, 2e-8)])
@pytest.mark.parametrize("gga", [False, True])
@pytest.mark.parametrize("mode", [1 << 17, 2 << 17, 3 << 17, (5 << 12) | (2 << 17)])
def test_rho_tcgen05_int8_matches_fp64_path(cuda, nslice, tol, gga, mode):
    """K2 on tcgen05 (point-stationary sliced int8 GEMM, fragment-layout epilogue) against the fp64 DMMA form of the
    same contraction, on a grid with several M tiles per superblock (benzene / cc-pVDZ: up to 128 kept AOs; the
    carbon cluster below: several hundred), without clusters and with the D_sb stream multicast over clusters of 2
    and 4 CTAs, and with a B cache too small for the tile (streaming slots)."""
    from dqc_b200 import _lib
    from dqc_b200.utils import systems
    from dqc_b200.grid.factory import get_predefined_grid
    zs, pos = systems.carbon_cluster(27)
    w, _ = util.make_wrapper(zs, pos.tolist(), "def2-svp")
    nb = len(w)
    grid = get_predefined_grid(3, zs, torch.tensor(pos, device=cuda), device=cuda)
    xyz, wts = grid.get_rgrid(), grid.get_dvolume()
    db = w.device_basis(cuda)
    deriv = 1 if gga else 0
    ref = _lib.GridBlocks(db, 0, nb, xyz, wts, deriv, sbp=512, eps=1e-12)
    gb = _lib.GridBlocks(db, 0, nb, xyz, wts, deriv, sbp=512, eps=1e-12, rho_i8_slices=nslice)
    assert gb.rho_bn == 128 and int(gb.nsp.max()) > 256
    dm = util.seeded_dm(w.nao(), int(sum(zs)) // 2, seed=1).to(cuda)
    r_ref, g_ref = ref.rho(dm, gga)
    lib = _lib.load()
    lib.b200qc_i8_mode(mode)
    try:
        r, g = gb.rho(dm, gga)
        r2, g2 = gb.rho(dm, gga)
    finally:
        lib.b200qc_i8_mode(0)
    scale = float(r_ref.abs().max())
    assert float((r - r_ref).abs().max()) < tol * scale
    assert torch.equal(r, r2)                        # fixed-order reductions, no atomics: bitwise repeatable
    if gga:
        assert float((g - g_ref).abs().max()) < tol * float(g_ref.abs().max())
        assert torch.equal(g, g2)


# ---- meta-GGA (reference: hcgto.py:183-186,420-438,473-489, libxc mgga_x_scan) ----
def _mgga_inputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    rho = 10 ** (torch.rand(n, dtype=torch.float64, generator=g) * 7 - 5)
    grad = torch.randn(3, n, dtype=torch.float64, generator=g) * rho ** (4.0 / 3) * 1.5
    tau_w = (grad * grad).sum(0) / (8 * rho)
    tau_unif = 0.3 * (3 * np.pi ** 2 * rho) ** (2.0 / 3) * rho
    alpha = 10 ** (torch.rand(n, dtype=torch.float64, generator=g) * 3 - 2)          # 0.01 .. 10, both branches of f(alpha)
    alpha = torch.where((alpha - 1).abs() < 0.02, alpha + 0.05, alpha)
    return rho, grad, torch.randn(n, dtype=torch.float64, generator=g), tau_w + alpha * tau_unif


@pytest.mark.parametrize("xcstr", ["mgga_x_scan", "mgga_x_scan + gga_c_pbe", "0.5*mgga_x_scan + 0.5*lda_x"])
def test_xc_mgga_unpol_matches_autograd_oracle(cuda, xcstr):
    from dqc_b200 import _lib
    from oracle import xc_ref
    rho, grad, lapl, tau = _mgga_inputs(4097, 7)
    e_ref, vr_ref, vg_ref, vl_ref, vk_ref = xc_ref.eval_unpol_mgga(xcstr, rho, grad, lapl, tau)
    e, vr, vg, vl, vk = _lib.xc_mgga_unpol(xc_ref.parse(xcstr), rho.to(cuda), grad.to(cuda).contiguous(), lapl.to(cuda),
                                           tau.to(cuda))
    rel = lambda a, b: float(((a.cpu() - b).abs() / (b.abs() + 1e-10 * b.abs().max())).max())
    assert rel(e, e_ref) < 1e-10
    assert rel(vr, vr_ref) < 1e-8
    assert rel(vk, vk_ref) < 1e-8
    assert float((vg.cpu() - vg_ref).abs().max() / vg_ref.abs().max()) < 1e-10
    assert float(vl.abs().max()) == 0.0 and float(vl_ref.abs().max()) == 0.0


def test_xc_mgga_pol_by_spin_scaling_matches_oracle(cuda):
    from dqc_b200 import get_xc, ValGrad, SpinParam
    from oracle import xc_ref
    ru, gu, lu, ku = _mgga_inputs(2001, 8)
    rd, gd, ld_, kd = _mgga_inputs(2001, 9)
    e_ref, vu_ref, vd_ref = xc_ref.eval_pol_mgga("mgga_x_scan", ru, rd, gu, gd, lu, ld_, ku, kd)
    xc = get_xc("mgga_x_scan")
    mk = lambda r, g, l, k: ValGrad(value=r.to(cuda), grad=g.to(cuda), lapl=l.to(cuda), kin=k.to(cuda))
    dens = SpinParam(u=mk(ru, gu, lu, ku), d=mk(rd, gd, ld_, kd))
    e = xc.get_edensityxc(dens)
    v = xc.get_vxc(dens)
    rel = lambda a, b: float(((a.cpu() - b).abs() / (b.abs() + 1e-10 * b.abs().max())).max())
    assert rel(e, e_ref) < 1e-10
    for got, ref in ((v.u, vu_ref), (v.d, vd_ref)):
        assert rel(got.value, ref[0]) < 1e-8 and rel(got.kin, ref[3]) < 1e-8
        assert float((got.grad.cpu() - ref[1]).abs().max() / ref[1].abs().max()) < 1e-10
    # unpolarised limit: both spins = rho / 2
    half = lambda t: 0.5 * t
    du = mk(half(ru), half(gu), half(lu), half(ku))
    e_u = xc.get_edensityxc(mk(ru, gu, lu, ku))
    e_p = xc.get_edensityxc(SpinParam(u=du, d=du))
    assert torch.allclose(e_u, e_p, rtol=1e-12, atol=0)


@pytest.mark.parametrize("which", ["h2o-def2svp", "highl"])
def test_mgga_densities_and_vxc_match_oracle(cuda, which):
    """Meta-GGA K2 (rho, grad rho, lapl rho, tau) and K4 (with the lapl and tau terms) on the superblock storage with five
    AO components, against oracle/fock_ref.py (hcgto.py:420-438, 473-489 restated) on a few thousand points."""
    from dqc_b200 import _lib
    from oracle import fock_ref, xc_ref
    xcstr = "mgga_x_scan"
    w, _ = util.make_wrapper(*util.H2O, "def2-svp") if which == "h2o-def2svp" else util.highl_wrapper()
    ngrid = 2500
    pts = util.random_points(ngrid, seed=13, span=2.5)
    wts = np.random.RandomState(6).uniform(0.0, 0.3, ngrid)
    h = fock_ref.RefHamilton(w, orthozer=False)
    h.setup_grid(pts, wts, xcstr)
    nao = w.nao()
    dm = util.seeded_dm(nao, max(1, nao // 4), seed=4)
    rho_ref, grad_ref, lapl_ref, kin_ref = h.dm2densinfo_mgga(dm)
    db = w.device_basis(cuda)
    gb = _lib.GridBlocks(db, 0, len(w), torch.tensor(pts, device=cuda), torch.tensor(wts, device=cuda), 2, sbp=512, eps=0.0)
    rho, grad, lapl, kin = gb.rho_mgga(dm.to(cuda))
    chk = lambda a, b, tol: float((a[..., :ngrid].cpu() - b).abs().max()) < tol * max(1.0, float(b.abs().max()))
    assert chk(rho, rho_ref, 1e-11) and chk(grad, grad_ref, 1e-10) and chk(lapl, lapl_ref, 1e-10) and chk(kin, kin_ref, 1e-10)
    # K4 from the oracle's potentials, then the whole chain on the device
    _, vr_ref, vg_ref, vl_ref, vk_ref = xc_ref.eval_unpol_mgga(xcstr, rho_ref, grad_ref, lapl_ref, kin_ref)
    # (a lapl-dependent potential too, so that the 2 vlapl lapl phi and 2 vlapl dphi dphi terms are exercised)
    vl_test = 0.05 * torch.sin(torch.arange(ngrid, dtype=torch.float64))
    mat_ref = h.vxc_from_potinfo_mgga(vr_ref, vg_ref, vl_test, vk_ref)
    pad = lambda t: torch.nn.functional.pad(t.to(cuda), (0, gb.ngl - ngrid)).contiguous()
    mat = gb.vxc_mat_mgga(pad(vr_ref), pad(vg_ref), pad(vl_test), pad(vk_ref))
    mat = 0.5 * (mat + mat.T)
    assert float((mat.cpu() - mat_ref).abs().max()) < 1e-10
    e, vr, vg, vl, vk = _lib.xc_mgga_unpol(xc_ref.parse(xcstr), rho, grad, lapl, kin)
    mat2 = gb.vxc_mat_mgga(vr, vg, vl, vk)
    mat2 = 0.5 * (mat2 + mat2.T)
    assert float((mat2.cpu() - h.vxc_from_potinfo_mgga(vr_ref, vg_ref, vl_ref, vk_ref)).abs().max()) < 1e-9
    exc_ref = float((xc_ref.eval_unpol_mgga(xcstr, rho_ref, grad_ref, lapl_ref, kin_ref)[0] * torch.tensor(wts)).sum())
    assert abs(float((e * gb.w).sum()) - exc_ref) < 1e-10
