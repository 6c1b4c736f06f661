"""GPU tests of the operator surface (HamiltonCGTO / DFMol / B200XC / Mol / HF / KS):
 - Fock pieces at a fixed seeded density against the CPU oracle (oracle/fock_ref.py):
   1e-6 elementwise on matrices, 1e-8 Ha on energies (the north-star parity bar);
 - the reference's own golden numbers end to end (dqc/test/test_hf.py, test_ks.py,
   test_hamilton.py) with the tolerances those tests use."""
import numpy as np
import pytest
import torch
from tests import util

pytestmark = pytest.mark.gpu
dtype = torch.float64


def _mol(atomzs, pos, basis, cuda, **kw):
    from dqc_b200 import Mol
    return Mol((torch.tensor(atomzs), torch.tensor(pos, dtype=dtype)), basis=basis, device=cuda, **kw)


def _diatomic(atomzs, dist):
    return atomzs, (np.array([[-0.5, 0.0, 0.0], [0.5, 0.0, 0.0]]) * dist).tolist()


# ---------------------------------------------------------------------------------------------
# fixed-density parity against the oracle
@pytest.mark.parametrize("orthozer,stored", [(True, True), (False, True), (True, False)])
def test_jk_hcore_match_oracle_fixed_dm(cuda, orthozer, stored):
    from oracle import fock_ref
    from dqc_b200 import config, _lib
    zs, pos = util.H2O
    mol = _mol(zs, pos, "def2-svp", cuda, orthogonalize_basis=orthozer)
    old = config.ERI_STORE_MAX_BYTES
    try:
        config.ERI_STORE_MAX_BYTES = old if stored else 0      # stored-ERI GEMVs vs direct J/K plan
        h = mol.get_hamiltonian().build()
    finally:
        config.ERI_STORE_MAX_BYTES = old
    assert isinstance(h._jkplan, _lib.StoredERI if stored else _lib.JKPlan)
    w, _ = util.make_wrapper(zs, pos, "def2-svp")
    ref = fock_ref.RefHamilton(w, orthozer=orthozer).build_eri()
    assert h.nao == ref.nao
    g = torch.Generator().manual_seed(0)
    dm = util.seeded_dm(h.nao, 5, seed=1) + 0.05 * torch.randn(h.nao, h.nao, dtype=dtype, generator=g)  # not symmetric
    # the orthogonaliser is defined up to the sign/rotation of eigenvectors: compare in the AO basis
    Xg, Xr = h._orthozer._orthozer.cpu() if orthozer else torch.eye(h.nao, dtype=dtype), ref.X
    to_ao = lambda m, X: torch.linalg.pinv(X.T) @ m @ torch.linalg.pinv(X) if orthozer else m
    dm_ao = Xr @ dm @ Xr.T
    dm_g = (torch.linalg.pinv(Xg) @ dm_ao @ torch.linalg.pinv(Xg).T).to(cuda)
    for name in ("get_elrep", "get_exchange"):
        got = getattr(h, name)(dm_g).fullmatrix().cpu()
        want = getattr(ref, name)(dm)
        assert float((to_ao(got, Xg) - to_ao(want, Xr)).abs().max()) < 1e-8, name
    assert float((to_ao(h.get_kinnucl().fullmatrix().cpu(), Xg) - to_ao(ref.kinnucl_mat, Xr)).abs().max()) < 1e-9
    dms = 0.5 * (dm + dm.T)
    dms_g = 0.5 * (dm_g + dm_g.T)
    assert abs(float(h.get_e_hcore(dms_g)) - float(ref.get_e_hcore(dms))) < 1e-8
    assert abs(float(h.get_e_elrep(dms_g)) - float(ref.get_e_elrep(dms))) < 1e-8
    assert abs(float(h.get_e_exchange(dms_g)) - float(ref.get_e_exchange(dms))) < 1e-8


@pytest.mark.parametrize("xcstr", ["lda_x + lda_c_pw", "gga_x_pbe + gga_c_pbe", "mgga_x_scan"])
def test_vxc_exc_match_oracle_fixed_dm(cuda, xcstr):
    from dqc_b200 import get_xc
    from dqc_b200.grid.factory import get_predefined_grid
    from oracle import fock_ref
    zs, pos = util.H2O
    mol = _mol(zs, pos, "def2-svp", cuda, grid="sg2", orthogonalize_basis=False)
    h = mol.get_hamiltonian()
    mol.setup_grid()
    grid = mol.get_grid()
    h.setup_grid(grid, get_xc(xcstr))
    h.build()
    w, _ = util.make_wrapper(zs, pos, "def2-svp")
    ref = fock_ref.RefHamilton(w, orthozer=False)
    ref.setup_grid(grid.get_rgrid().cpu().numpy(), grid.get_dvolume().cpu().numpy(), xcstr)
    dm = util.seeded_dm(h.nao, 5, seed=2)
    got = h.get_vxc(dm.to(cuda)).fullmatrix().cpu()
    want = ref.get_vxc(dm)
    assert float((got - want).abs().max()) < 1e-8
    assert abs(float(h.get_e_xc(dm.to(cuda))) - float(ref.get_e_xc(dm))) < 1e-8
    # polarised: SpinParam in, SpinParam out
    from dqc_b200 import SpinParam
    du, dd = util.seeded_dm(h.nao, 5, seed=3) * 0.5, util.seeded_dm(h.nao, 4, seed=4) * 0.5
    gp = h.get_vxc(SpinParam(u=du.to(cuda), d=dd.to(cuda)))
    wu, wd = ref.get_vxc((du, dd))
    assert float((gp.u.fullmatrix().cpu() - wu).abs().max()) < 1e-8
    assert float((gp.d.fullmatrix().cpu() - wd).abs().max()) < 1e-8
    assert abs(float(h.get_e_xc(SpinParam(u=du.to(cuda), d=dd.to(cuda)))) - float(ref.get_e_xc((du, dd)))) < 1e-8


def test_becke_grid_matches_oracle_weights(cuda):
    from dqc_b200.grid.factory import get_predefined_grid
    from dqc_b200.utils.periodictable import atom_expected_radii
    from oracle import becke_ref
    zs, pos = util.CH4ISH
    grid = get_predefined_grid("sg2", zs, torch.tensor(pos, dtype=dtype, device=cuda), device=cuda)
    one = [get_predefined_grid("sg2", [z], torch.zeros(1, 3, dtype=dtype), device=torch.device("cpu")) for z in zs]
    owner = np.repeat(np.arange(len(zs)), [g.get_rgrid().shape[0] for g in one])
    dv = np.concatenate([g.get_dvolume().numpy() for g in one])
    wref = becke_ref.becke_weights(grid.get_rgrid().cpu().numpy(), owner, np.array(pos),
                                   np.array([atom_expected_radii[z] for z in zs]), "becke")
    assert np.allclose(grid.get_dvolume().cpu().numpy(), dv * wref, rtol=1e-11, atol=1e-300)


def test_dfj_operator_matches_oracle(cuda):
    from oracle import fock_ref
    from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
    zs, pos = util.H2O
    mol = _mol(zs, pos, "def2-svp", cuda).densityfit(auxbasis="etb-jfit")
    h = mol.get_hamiltonian().build()
    w, _ = util.make_wrapper(zs, pos, "def2-svp")
    aux, _ = util.make_wrapper(zs, pos, "etb-jfit")
    bw, aw = LibcintWrapper.concatenate(w, aux)
    ref = fock_ref.RefHamilton(bw, auxwrapper=aw, orthozer=True).build_df()
    dm_ao = util.seeded_dm(w.nao(), 5, seed=5)
    Xg, Xr = h._orthozer._orthozer.cpu(), ref.X
    dm_g = torch.linalg.pinv(Xg) @ dm_ao @ torch.linalg.pinv(Xg).T
    dm_r = torch.linalg.pinv(Xr) @ dm_ao @ torch.linalg.pinv(Xr).T
    got = h.get_elrep(dm_g.to(cuda)).fullmatrix().cpu()
    want = ref.get_elrep(dm_r)
    back = lambda m, X: torch.linalg.pinv(X.T) @ m @ torch.linalg.pinv(X)
    assert float((back(got, Xg) - back(want, Xr)).abs().max()) < 1e-8
    assert float((h.df.j2c.cpu() - ref.j2c).abs().max()) < 1e-10
    assert float((h.df.j3c.cpu() - ref.j3c).abs().max()) < 1e-10
    # the reference raises for exact exchange with density fitting (hcgto.py:229-230); here that is the
    # behaviour with the DF-K extension switched off
    from dqc_b200 import config
    old = config.DF_EXCHANGE
    try:
        config.DF_EXCHANGE = False
        with pytest.raises(RuntimeError):
            h.get_exchange(dm_g.to(cuda))
        with pytest.raises(RuntimeError):
            h.get_fock_2e(dm_g.to(cuda), exx=0.25, with_xc=False)
    finally:
        config.DF_EXCHANGE = old


# ---------------------------------------------------------------------------------------------
# operator-level behaviour the reference tests (dqc/test/test_hamilton.py)
def test_ao_orb2dm_batching(cuda):
    # test_hamilton.py:67-93
    zs, pos = _diatomic([1, 1], 1.0)
    h = _mol(zs, pos, "3-21g", cuda).get_hamiltonian().build()
    nao = h.nao
    norb = 2
    g = torch.Generator().manual_seed(1)
    ao_orb = torch.randn(3, 1, nao, norb, dtype=dtype, generator=g).to(cuda)
    w = torch.randn(2, norb, dtype=dtype, generator=g).to(cuda)
    dm = h.ao_orb2dm(ao_orb, w)
    assert list(dm.shape) == [3, 2, nao, nao]
    assert torch.allclose(dm[1, 0], h.ao_orb2dm(ao_orb[1, 0], w[0]))


def test_aodm2dens_golden_points(cuda):
    # test_hamilton.py:95-142: H2 ("H 0 0 0.8; H 0 0 -0.8") 3-21G HF density on 5 points, from PySCF
    from dqc_b200 import HF
    mol = _mol([1, 1], [[0.0, 0.0, 0.8], [0.0, 0.0, -0.8]], "3-21g", cuda)
    qc = HF(mol).run()
    dm = qc.aodm()
    pts = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 0.4], [0.0, 0.0, 0.8], [0.0, 0.0, -0.4], [0.0, 0.0, -0.8]],
                       dtype=dtype)
    dens = mol.get_hamiltonian().aodm2dens(dm, pts).cpu()
    true = torch.tensor([0.18742819, 0.23469519, 0.30250292, 0.23469519, 0.30250292], dtype=dtype)
    assert torch.allclose(dens, true)


def test_vext_constant_gives_overlap(cuda):
    # test_hamilton.py:144-155: a constant potential w gives w * S = w * I in the orthogonal basis
    mol = _mol(*_diatomic([1, 1], 1.0), "3-21g", cuda)
    mol.setup_grid()
    h = mol.get_hamiltonian()
    h.setup_grid(mol.get_grid())
    h.build()
    ngrid = mol.get_grid().get_rgrid().shape[0]
    w = 0.37
    mat = h.get_vext(torch.full((ngrid,), w, dtype=dtype, device=cuda)).fullmatrix()
    assert torch.allclose(mat, w * torch.eye(h.nao, dtype=dtype, device=cuda), rtol=4e-5, atol=4e-5)


# ---------------------------------------------------------------------------------------------
# end-to-end golden energies of the reference's tests
RHF = [([1, 1], 1.0, -1.07195346e+00), ([3, 3], 5.0, -1.47683688e+01), ([7, 7], 2.0, -1.08298897e+02),
       ([9, 9], 2.5, -1.97636373e+02), ([6, 8], 2.0, -1.12078732e+02)]


@pytest.mark.parametrize("atomzs,dist,etrue", RHF)
def test_rhf_energy_golden(cuda, atomzs, dist, etrue):
    # dqc/test/test_hf.py:18-51, rtol 1e-7
    from dqc_b200 import HF
    mol = _mol(*_diatomic(atomzs, dist), "3-21g", cuda)
    ene = HF(mol, restricted=True).run().energy()
    assert torch.allclose(ene, ene * 0 + etrue, rtol=1e-7)


@pytest.mark.parametrize("atomzs,dist,etrue", RHF[:2])
def test_uhf_same_as_rhf(cuda, atomzs, dist, etrue):
    # test_hf.py:163-174, rtol 1e-8
    from dqc_b200 import HF
    mol = _mol(*_diatomic(atomzs, dist), "3-21g", cuda)
    ene = HF(mol, restricted=False).run().energy()
    assert torch.allclose(ene, ene * 0 + etrue, rtol=1e-8)


@pytest.mark.parametrize("z,spin,etrue", [(1, 1, -4.96198609e-01), (3, 1, -7.38151326e+00),
                                          (5, 1, -2.43897617e+01), (8, 2, -7.43936572e+01)])
def test_uhf_atoms_golden(cuda, z, spin, etrue):
    # test_hf.py:141-189, rtol 1e-7
    from dqc_b200 import HF
    mol = _mol([z], [[0.0, 0.0, 0.0]], "3-21g", cuda, spin=spin)
    ene = HF(mol, restricted=False).run().energy()
    assert torch.allclose(ene, ene * 0 + etrue, atol=0.0, rtol=1e-7)


def test_uhf_no_golden(cuda):
    # test_hf.py:154-161,191-206: NO, -128.477807 Ha
    from dqc_b200 import HF
    mol = _mol(*_diatomic([7, 8], 2.0), "3-21g", cuda, spin=1)
    ene = HF(mol, restricted=False).run(fwd_options={"maxiter": 150}).energy()
    assert torch.allclose(ene, ene * 0 - 1.28477807e+02, rtol=1e-8, atol=0.0)


@pytest.mark.parametrize("xc,etrue", [("lda_x", -0.979143262), ("gga_x_pbe", -1.068217310366847)])
@pytest.mark.parametrize("grid", [3, "sg2"])
def test_rks_h2_golden(cuda, xc, etrue, grid):
    # dqc/test/test_ks.py:40-111: H2 6-311++G**, PySCF energies, atol 1.3e-3 (grids differ)
    from dqc_b200 import KS
    mol = _mol(*_diatomic([1, 1], 1.0), "6-311++G**", cuda, grid=grid)
    ene = KS(mol, xc=xc, restricted=True).run().energy()
    assert torch.allclose(ene, ene * 0 + etrue, atol=1.3e-3, rtol=0)


def test_rks_uks_scan_h2(cuda):
    """Converged meta-GGA KS energies.  The reference's only SCAN numbers (dqc/test/test_ks.py:56-62, 6-311++G**) are
    marked xfail for H2 there ("Psi4 and PySCF don't converge" -- the diffuse functions; the DIIS iteration here does not
    converge on that case either) and need basis tables for Li..F that are not embedded for the rest.  H2 / 3-21G instead:
    the CUDA path against the CPU oracle's own SCF on the same grid, and unrestricted against restricted."""
    from dqc_b200 import KS
    from oracle import fock_ref, scf_ref
    zs, pos = _diatomic([1, 1], 1.4)
    mol = _mol(zs, pos, "3-21g", cuda, grid=3)
    qc = KS(mol, xc="mgga_x_scan", restricted=True).run()
    e_r = qc.energy()
    assert qc.converged
    grid = mol.get_grid()
    w, _ = util.make_wrapper(zs, [list(map(float, p)) for p in pos], "3-21g")
    ref = fock_ref.RefHamilton(w).build_eri()
    ref.setup_grid(grid.get_rgrid().cpu().numpy(), grid.get_dvolume().cpu().numpy(), "mgga_x_scan")
    e_ref, _ = scf_ref.run_scf(ref, zs, np.array(pos, dtype=np.float64), 2, method="ks")
    assert abs(float(e_r) - e_ref) < 1e-7
    e_u = KS(_mol(zs, pos, "3-21g", cuda, grid=3), xc="mgga_x_scan", restricted=False).run().energy()
    assert torch.allclose(e_r, e_u, rtol=1e-8)


def test_uks_equals_rks_and_noxc(cuda):
    # test_ks.py:362-368 (UKS == RKS for a closed shell) and :246-259 (xc=None == 0*lda_x)
    from dqc_b200 import KS
    mk = lambda: _mol(*_diatomic([1, 1], 1.0), "3-21g", cuda, grid="sg2")
    e_r = KS(mk(), xc="lda_x", restricted=True).run().energy()
    e_u = KS(mk(), xc="lda_x", restricted=False).run().energy()
    assert torch.allclose(e_r, e_u, rtol=1e-8)
    e_none = KS(mk(), xc=None).run().energy()
    e_zero = KS(mk(), xc="0*lda_x").run().energy()
    assert torch.allclose(e_none, e_zero, rtol=1e-9)


def test_ks_df_close_to_nodf(cuda):
    # DF error of the even-tempered fitting set is far below the 1.1e-3 the reference allows (test_ks.py:442-464)
    from dqc_b200 import KS
    zs, pos = util.H2O
    e0 = KS(_mol(zs, pos, "def2-svp", cuda, grid="sg2"), xc="gga_x_pbe + gga_c_pbe").run().energy()
    e1 = KS(_mol(zs, pos, "def2-svp", cuda, grid="sg2").densityfit(auxbasis="etb-jfit"),
            xc="gga_x_pbe + gga_c_pbe").run().energy()
    assert abs(float(e0 - e1)) < 1.1e-3
    assert -76.4 < float(e0) < -76.1     # PBE/def2-SVP water is about -76.27 Ha (sanity, not a pin)


def test_hybrid_composition(cuda):
    # F = h + J + a K' + Vxc composed from the reference's own pieces (SURVEY 8a notes): a = 1 with
    # xc = None is exactly HF, a = 0 is plain KS
    from dqc_b200 import KS, HF
    mk = lambda: _mol(*_diatomic([1, 1], 1.0), "3-21g", cuda, grid="sg2")
    e_hf = HF(mk()).run().energy()
    e_x1 = KS(mk(), xc=None, exx_fraction=1.0).run().energy()
    assert torch.allclose(e_hf, e_x1, rtol=1e-9)
    e_b = KS(mk(), xc="0.8*lda_x + lda_c_pw", exx_fraction=0.2).run().energy()
    e_l = KS(mk(), xc="lda_x + lda_c_pw").run().energy()
    assert abs(float(e_b - e_l)) < 0.05 and float(e_b) != float(e_l)


def test_scf_lagged_convergence_check_gives_the_same_energy(cuda, monkeypatch):
    """The DIIS loop reads the convergence scalar of iteration k while iteration k + 1 is being enqueued (no stall of the
    device queue, config.SCF_CHECK_LAG = 1); the converged energy equals the synchronous loop's."""
    from dqc_b200 import HF
    from dqc_b200.utils.config import config as cfg
    es = []
    for lag in (0, 1, 2):
        monkeypatch.setattr(cfg, "SCF_CHECK_LAG", lag)
        qc = HF(_mol(*_diatomic([7, 7], 2.0), "3-21g", cuda), restricted=True).run()
        assert qc.converged
        es.append(float(qc.energy()))
    assert abs(es[0] - es[1]) < 1e-9 and abs(es[0] - es[2]) < 1e-9
    assert abs(es[0] - (-1.08298897e+02)) < 2e-5        # the reference's golden RHF energy of N2 / 3-21G (test_hf.py:18-51)


@pytest.mark.parametrize("gridname", ["sg2", "sg3", 3, 4])
def test_device_grid_is_bit_identical_to_the_host_construction(cuda, gridname):
    """Grid construction on the GPU: radial x Lebedev products, pruning and translation in
    b200qc_grid_assemble against the torch classes on the host (lebedev_grid.py / multiatoms_grid.py restated): the same
    points and radial x angular weights bit for bit, for both predefined families (Dasgupta and NWChem pruning)."""
    from dqc_b200.grid.factory import get_predefined_grid
    from dqc_b200.grid import factory
    zs, pos = util.H2O
    p = torch.tensor(pos, dtype=dtype)
    dev_grid = get_predefined_grid(gridname, zs, p.to(cuda), device=cuda)
    one = {z: get_predefined_grid(gridname, [z], torch.zeros(1, 3, dtype=dtype), device=torch.device("cpu")) for z in set(zs)}
    xyz_host = torch.cat([one[z].get_rgrid() + p[i] for i, z in enumerate(zs)])
    assert torch.equal(dev_grid.get_rgrid().cpu(), xyz_host)
    # single-atom grids have unit partition weights: their dvolume is the radial x angular weight itself
    dv_host = torch.cat([one[z].get_dvolume() for z in zs])
    w = dev_grid.get_dvolume().cpu() / dv_host
    assert float(w.max()) <= 1.0 + 1e-12 and float(w.min()) >= 0.0
    f = torch.exp(-0.5 * ((dev_grid.get_rgrid().cpu() - p[0]) ** 2).sum(-1))
    assert abs(float((f * dev_grid.get_dvolume().cpu()).sum()) - (2 * np.pi) ** 1.5) < 2e-3
