"""Oracle numbers next to the PRODUCTION kernels at the BASELINE.json sizes.

The small-molecule parity tests (test_gpu_xcpath.py, test_gpu_hamilton.py) touch one N tile and one or two K steps
of the tcgen05 kernels.  Here the CPU oracle (oracle/cint_oracle.c AO values + the reference's per-iteration ops of
oracle/fock_ref.py, hcgto.py:371-495) is evaluated on slices of the real C60/def2-SVP and taxol-like/def2-SVP sg3
grids -- 16 384+ points each, cut at superblock boundaries so that the superblocks are exactly those of the full
run (nsp 400-840: 4-7 M tiles, up to 9 N tiles, 13-27 K steps) -- and compared with the default GPU path (superblock
storage, tcgen05 int8 density and Vxc kernels, orthogonalised basis) on exactly those points: rho, grad rho
pointwise, the partial Vxc matrix elementwise (bar 1e-6; asserted 1e-8) and the partial E_xc (bar 1e-8 Ha).
Benzene/cc-pVDZ (configs[1]) is checked at its FULL size: stored-regime J and K, density-fitted J, LDA Vxc on the
whole sg3 grid and the energies, all against the oracle."""
import numpy as np
import pytest
import torch
from tests import util

pytestmark = pytest.mark.gpu
dtype = torch.float64
B3LYP_SL = "0.08*lda_x + 0.72*gga_x_b88 + 0.19*lda_c_vwn_rpa + 0.81*gga_c_lyp"


class _SliceGrid(object):
    """A BaseGrid over a subset of the points of a molecular grid (positions and Becke-weighted volumes kept)."""
    coord_type = "cart"

    def __init__(self, rgrid, dvol):
        self._r, self._w = rgrid.contiguous(), dvol.contiguous()
        self.dtype, self.device = rgrid.dtype, rgrid.device

    def get_rgrid(self):
        return self._r

    def get_dvolume(self):
        return self._w

    def getparamnames(self, methodname, prefix=""):
        return []


def _grid_slices(ngrid, sbp, nsb_each, where):
    """Index of `nsb_each` whole superblocks starting at the superblock nearest to each fraction in `where`."""
    nsb = ngrid // sbp
    idx = []
    for f in where:
        s0 = min(max(int(f * nsb), 0), nsb - nsb_each)
        idx.append(torch.arange(s0 * sbp, (s0 + nsb_each) * sbp))
    return torch.cat(idx)


def _ao_basis_pair(h, ref, dm_ao, cuda):
    """The same AO-basis density expressed in the GPU's and in the oracle's orthogonal basis (X = U s^-1/2 is only
    defined up to rotations inside degenerate eigenspaces of S -- C60 has many), plus the map back to the AO basis."""
    Xg, Xr = h._orthozer._orthozer.cpu(), ref.X
    pg, pr = torch.linalg.pinv(Xg), torch.linalg.pinv(Xr)
    dm_g = (pg @ dm_ao @ pg.T).to(cuda)
    dm_r = pr @ dm_ao @ pr.T
    back = lambda m, p: p.T @ m @ p
    return dm_g, dm_r, (lambda m: back(m, pg)), (lambda m: back(m, pr))


@pytest.mark.parametrize("system,xcstr,nocc", [("c60", "gga_x_pbe + gga_c_pbe", 180),
                                               ("taxol_like", B3LYP_SL, 226)])
def test_production_xc_path_matches_oracle_on_fullsize_grid_slices(cuda, system, xcstr, nocc):
    from dqc_b200 import Mol, get_xc, config
    from dqc_b200.utils import systems
    from oracle import fock_ref
    assert config.RHO_I8_SLICES in (5, 6) and config.VXC_I8_SLICES in (5, 6)      # the tcgen05 kernels are the default
    zs, pos = getattr(systems, system)()
    mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=dtype)), basis="def2-svp", grid="sg3", device=cuda)
    mol.setup_grid()
    full = mol.get_grid()
    ngrid = full.get_rgrid().shape[0]
    # 4 x 8 superblocks: the inner shells of the first atom, two interior stretches and the tail of the grid
    sel = _grid_slices(ngrid, config.SB_POINTS, 8, (0.0, 0.31, 0.62, 1.0)).to(cuda)
    grid = _SliceGrid(full.get_rgrid()[sel], full.get_dvolume()[sel])
    h = mol.get_hamiltonian()
    h.setup_grid(grid, get_xc(xcstr))
    gb = h._gb
    assert gb.rho_i8_slices and gb.i8_slices
    assert int(gb.nsp.max()) >= 384 and gb.nsb == 32           # multi-tile, multi-K-step superblocks

    w, _ = util.make_wrapper(zs, np.asarray(pos).tolist(), "def2-svp")
    ref = fock_ref.RefHamilton(w, orthozer=True)
    assert ref.nao == h.nao
    ref.setup_grid(grid.get_rgrid().cpu().numpy(), grid.get_dvolume().cpu().numpy(), xcstr)
    dm_ao = util.seeded_dm(w.nao(), nocc, seed=11)
    dm_g, dm_r, back_g, back_r = _ao_basis_pair(h, ref, dm_ao, cuda)

    # rho, grad rho on every point of the slices
    dens = h._dm2densinfo(dm_g)
    rho_ref, grad_ref = ref.dm2densinfo(dm_r)
    rscale = float(rho_ref.abs().max())
    assert float((dens.value.cpu() - rho_ref).abs().max()) < 1e-8 * rscale
    assert float((dens.grad.cpu() - grad_ref).abs().max()) < 1e-8 * float(grad_ref.abs().max())
    # relative check where the density is not negligible
    big = rho_ref > 1e-6
    assert float(((dens.value.cpu() - rho_ref).abs() / rho_ref.abs())[big].max()) < 1e-6

    # partial Vxc (AO basis, elementwise) and partial E_xc of these points
    v_g = back_g(h.get_vxc(dm_g).fullmatrix().cpu())
    v_r = back_r(ref.get_vxc(dm_r))
    assert float((v_g - v_r).abs().max()) < 1e-8
    assert float(v_r.abs().max()) > 1e-3
    e_g, e_r = float(h.get_e_xc(dm_g)), float(ref.get_e_xc(dm_r))
    assert abs(e_g - e_r) < 1e-8
    assert abs(e_r) > 1e-2


@pytest.fixture(scope="module")
def benzene_oracle():
    """Benzene / cc-pVDZ on the CPU oracle: dense (ij|kl) (114^4), DF tensors, AO values on the full sg3 grid."""
    from dqc_b200.utils import systems
    zs, pos = systems.benzene()
    w, _ = util.make_wrapper(zs, pos.tolist(), "cc-pvdz")
    return zs, pos, w


def test_benzene_ccpvdz_fullsize_jk_and_vxc_match_oracle(cuda, benzene_oracle):
    """configs[1] at full size: 4-centre J and K (stored regime, the default for nao 114), LDA Vxc and E_xc on the
    whole 206 304-point grid, core Hamiltonian and energies -- GPU against oracle, elementwise."""
    from dqc_b200 import Mol, get_xc, _lib
    from oracle import fock_ref
    zs, pos, w = benzene_oracle
    xcstr = "lda_x + lda_c_pw"
    mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=dtype)), basis="cc-pvdz", grid="sg3", device=cuda)
    h = mol.get_hamiltonian()
    mol.setup_grid()
    grid = mol.get_grid()
    assert grid.get_rgrid().shape[0] == 206304
    h.setup_grid(grid, get_xc(xcstr))
    h.build()
    assert isinstance(h._jkplan, _lib.StoredERI) and h.nao == 114
    ref = fock_ref.RefHamilton(w, orthozer=True).build_eri()
    ref.setup_grid(grid.get_rgrid().cpu().numpy(), grid.get_dvolume().cpu().numpy(), xcstr)
    dm_ao = util.seeded_dm(114, 21, seed=5)
    dm_g, dm_r, back_g, back_r = _ao_basis_pair(h, ref, dm_ao, cuda)
    for name in ("get_elrep", "get_exchange", "get_vxc"):
        got = back_g(getattr(h, name)(dm_g).fullmatrix().cpu())
        want = back_r(getattr(ref, name)(dm_r))
        assert float((got - want).abs().max()) < 1e-8, name
    assert float((back_g(h.get_kinnucl().fullmatrix().cpu()) - back_r(ref.kinnucl_mat)).abs().max()) < 1e-8
    for name in ("get_e_hcore", "get_e_elrep", "get_e_exchange", "get_e_xc"):
        assert abs(float(getattr(h, name)(dm_g)) - float(getattr(ref, name)(dm_r))) < 1e-8, name
    # the fused build used by the SCF engines (one call, LDA + 4c J) equals the sum of the oracle's pieces
    f = back_g(h.get_fock_2e(dm_g).fullmatrix().cpu())
    assert float((f - back_r(ref.get_elrep(dm_r) + ref.get_vxc(dm_r))).abs().max()) < 1e-8


def test_benzene_ccpvdz_fullsize_direct_jk_matches_oracle(cuda, benzene_oracle):
    """The same J and K from the direct (never stored) Rys engine that larger molecules use."""
    from dqc_b200 import Mol, config, _lib
    from oracle import fock_ref
    zs, pos, w = benzene_oracle
    old = config.ERI_STORE_MAX_BYTES
    try:
        config.ERI_STORE_MAX_BYTES = 0
        mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=dtype)), basis="cc-pvdz", grid="sg3", device=cuda,
                  orthogonalize_basis=False)
        h = mol.get_hamiltonian().build()
    finally:
        config.ERI_STORE_MAX_BYTES = old
    assert isinstance(h._jkplan, _lib.JKPlan)
    ref = fock_ref.RefHamilton(w, orthozer=False).build_eri()
    dm = util.seeded_dm(114, 21, seed=6)
    for name in ("get_elrep", "get_exchange"):
        got = getattr(h, name)(dm.to(cuda)).fullmatrix().cpu()
        assert float((got - getattr(ref, name)(dm)).abs().max()) < 1e-8, name


def test_benzene_ccpvdz_fullsize_dfj_matches_oracle(cuda, benzene_oracle):
    """Density-fitted J of benzene/cc-pVDZ with the shipped aux set (naux 684): packed (ij|P), both GEMV passes."""
    from dqc_b200 import Mol
    from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
    from oracle import fock_ref
    zs, pos, w = benzene_oracle
    mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=dtype)), basis="cc-pvdz", grid="sg3", device=cuda)
    mol.densityfit(auxbasis="etb-jfit")
    h = mol.get_hamiltonian().build()
    aux, _ = util.make_wrapper(zs, pos.tolist(), "etb-jfit")
    bw, aw = LibcintWrapper.concatenate(w, aux)
    ref = fock_ref.RefHamilton(bw, auxwrapper=aw, orthozer=True).build_df()
    dm_ao = util.seeded_dm(114, 21, seed=7)
    dm_g, dm_r, back_g, back_r = _ao_basis_pair(h, ref, dm_ao, cuda)
    got = back_g(h.get_elrep(dm_g).fullmatrix().cpu())
    want = back_r(ref.get_elrep(dm_r))
    # the explicit inverse of (P|Q) (dfmol.py:48) amplifies rounding by its condition number on both sides
    assert float((got - want).abs().max()) < 1e-7
    assert abs(float(h.get_e_elrep(dm_g)) - float(ref.get_e_elrep(dm_r))) < 1e-7
