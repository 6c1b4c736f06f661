"""GPU parity of the Rys integral engine (int1e / int2c2e / int3c2e / int2e), the direct J/K
digestion and the density-fitted J against the CPU oracle (McMurchie-Davidson, a different
algorithm), all through the C-ABI.  Tolerance 1e-11 abs on integrals."""
import numpy as np
import pytest
import torch
from tests import util

pytestmark = pytest.mark.gpu


def _wrapper(which):
    if which == "h2o-def2svp":
        return util.make_wrapper(*util.H2O, "def2-svp")[0]
    if which == "highl":
        return util.highl_wrapper()[0]
    if which == "ch4ish-321g":
        return util.make_wrapper(*util.CH4ISH, "3-21g")[0]
    raise KeyError(which)


@pytest.mark.parametrize("which", ["h2o-def2svp", "highl", "ch4ish-321g"])
@pytest.mark.parametrize("kind", ["ovlp", "kin", "nuc"])
def test_int1e_matches_oracle(cuda, which, kind):
    from dqc_b200 import _lib
    from oracle import cint
    w = _wrapper(which)
    atm, bas, env = w.atm_bas_env
    nb = len(w)
    ref = cint.int1e(kind, atm, bas, env)
    got = _lib.int1e(w.device_basis(cuda), kind, (0, nb, 0, nb)).cpu().numpy()
    assert np.abs(got - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())
    # sub-blocks come out in the right place (dqc/test/test_libcint.py:145-199 subset consistency)
    if nb > 3:
        sub = _lib.int1e(w.device_basis(cuda), kind, (1, nb - 1, 2, nb)).cpu().numpy()
        loc = w.full_shell_to_aoloc
        assert np.abs(sub - ref[loc[1]:loc[nb - 1], loc[2]:loc[nb]]).max() < 1e-11 * max(1.0, np.abs(ref).max())


def test_int1e_rinv_matches_oracle(cuda):
    from dqc_b200 import _lib
    from oracle import cint
    w = _wrapper("h2o-def2svp")
    atm, bas, env = w.atm_bas_env
    nb = len(w)
    orig = np.array([0.3, -0.4, 0.5])
    env2 = env.copy()
    env2[4:7] = orig
    ref = cint.int1e("rinv", atm, bas, env2)
    got = _lib.int1e(w.device_basis(cuda), "rinv", (0, nb, 0, nb), orig).cpu().numpy()
    assert np.abs(got - ref).max() < 1e-11


def test_overlap_diagonal_is_one(cuda):
    from dqc_b200 import _lib
    w = _wrapper("highl")
    nb = len(w)
    s = _lib.int1e(w.device_basis(cuda), "ovlp", (0, nb, 0, nb))
    assert torch.allclose(torch.diagonal(s), torch.ones(w.nao(), dtype=torch.float64, device=cuda), atol=1e-12)


@pytest.mark.parametrize("which", ["c-etbjfit", "highl"])
def test_int2c2e_matches_oracle(cuda, which):
    from dqc_b200 import _lib
    from oracle import cint
    w = util.make_wrapper([6, 1], [[0.1, 0.2, -0.3], [1.2, -0.8, 0.9]], "etb-jfit")[0] if which == "c-etbjfit" \
        else _wrapper("highl")
    atm, bas, env = w.atm_bas_env
    nb = len(w)
    ref = cint.int2c2e(atm, bas, env)
    got = _lib.int2c2e(w.device_basis(cuda), (0, nb, 0, nb)).cpu().numpy()
    assert np.abs(got - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


def _df_wrappers(basis="def2-svp"):
    from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
    w = util.make_wrapper(*util.H2O, basis)[0]
    aux = util.make_wrapper(*util.H2O, "etb-jfit")[0]
    return LibcintWrapper.concatenate(w, aux)


def test_int3c2e_matches_oracle_dense_and_packed(cuda):
    from dqc_b200 import _lib
    from oracle import cint
    bw, aw = _df_wrappers()
    atm, bas, env = aw.atm_bas_env
    b0, b1 = bw.shell_idxs
    a0, a1 = aw.shell_idxs
    sl = (b0, b1, b0, b1, a0, a1)
    ref = cint.int3c2e(atm, bas, env, sl)
    db = aw.device_basis(cuda)
    got = _lib.int3c2e(db, sl).cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())
    packed = _lib.int3c2e_packed(db, sl)
    nao, naux = ref.shape[0], ref.shape[2]
    ii, jj = np.tril_indices(nao)
    assert np.abs(packed[:, :naux].cpu().numpy() - ref[ii, jj]).max() < 1e-11 * max(1.0, np.abs(ref).max())
    assert float(packed[:, naux:].abs().max()) == 0.0 if packed.shape[1] > naux else True
    # pack_tril of the dense tensor gives the same thing
    p2 = _lib.pack_tril(torch.tensor(ref, device=cuda))
    assert torch.allclose(p2, packed, atol=1e-11)


@pytest.mark.parametrize("which", ["h2o-def2svp", "ch4ish-321g"])
def test_int2e_matches_oracle(cuda, which):
    from dqc_b200 import _lib
    from oracle import cint
    w = _wrapper(which)
    atm, bas, env = w.atm_bas_env
    nb = len(w)
    ref = cint.int2e(atm, bas, env)
    got = _lib.int2e(w.device_basis(cuda), (0, nb) * 4).cpu().numpy()
    assert np.abs(got - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


def test_int2e_high_l_subset(cuda):
    """(ij|kl) with f and g functions on a shell subset (4 roots and more)."""
    from dqc_b200 import _lib
    from oracle import cint
    w = _wrapper("highl")
    atm, bas, env = w.atm_bas_env
    sl = (2, 4, 0, 2, 7, 8, 5, 6)   # (d f | s p | d' | s')
    ref = cint.int2e(atm, bas, env, sl)
    got = _lib.int2e(w.device_basis(cuda), sl).cpu().numpy()
    assert np.abs(got - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("which,nset", [("h2o-def2svp", 1), ("ch4ish-321g", 2)])
def test_jk_direct_matches_dense_contraction(cuda, which, nset):
    """K8 against the reference's dense einsums (hcgto.py:209,234) on oracle integrals."""
    from dqc_b200 import _lib
    from oracle import cint
    w = _wrapper(which)
    atm, bas, env = w.atm_bas_env
    nb, nao = len(w), w.nao()
    eri = torch.as_tensor(cint.int2e(atm, bas, env))
    dms = torch.stack([util.seeded_dm(nao, max(1, nao // 3), seed=s) for s in range(nset)])
    jref = torch.einsum("sij,ijkl->skl", dms, eri)
    kref = torch.einsum("sil,ijkl->sjk", dms, eri)
    vj, vk = _lib.jk_direct(w.device_basis(cuda), 0, nb, dms.to(cuda))
    assert float((vj.cpu() - jref).abs().max()) < 1e-10
    assert float((vk.cpu() - kref).abs().max()) < 1e-10
    # plan API with two "ranks": partial results add up
    plan = _lib.JKPlan(w.device_basis(cuda), 0, nb, 1e-14)
    assert plan.nquartets > 0
    j0, k0 = plan.run(dms.to(cuda), rank=0, world=2)
    j1, k1 = plan.run(dms.to(cuda), rank=1, world=2)
    assert float((j0 + j1 - jref.to(cuda)).abs().max()) < 1e-10
    assert float((k0 + k1 - kref.to(cuda)).abs().max()) < 1e-10


def test_jk_register_engine_matches_stored_eri_and_shared_engine(cuda, monkeypatch):
    """The register-resident quartet engine (csrc/jk_reg.cuh: every class pair of an s/p basis) on a 14-atom carbon
    cluster / 3-21G -- several primitive-count buckets, bras with more than one 128-ket work item, partial last rounds --
    against the stored-ERI GEMVs (Rys integrals pinned by the oracle above) and the shared-memory engine it replaces."""
    from dqc_b200 import _lib
    from dqc_b200.utils import systems
    zs, pos = systems.carbon_cluster(14)
    w, _ = util.make_wrapper(zs, pos.tolist(), "3-21g")
    nb, nao = len(w), w.nao()
    db = w.device_basis(cuda)
    dms = torch.stack([util.seeded_dm(nao, nao // 4, seed=1), util.seeded_dm(nao, nao // 5, seed=2)]).to(cuda)
    monkeypatch.setenv("B200QC_JK_ONE_BUCKET", "0")       # classes split by primitive-pair count, as on large systems
    plan = _lib.JKPlan(db, 0, nb, 1e-14)
    assert plan.nquartets_reg == plan.nquartets > 0       # s and p shells only: nothing left on the other engine
    vj, vk = plan.run(dms)
    monkeypatch.delenv("B200QC_JK_ONE_BUCKET")             # default for a system of this size: one bucket per (l_i, l_j)
    plan1 = _lib.JKPlan(db, 0, nb, 1e-14)
    assert plan1.nquartets == plan.nquartets
    vj1, vk1 = plan1.run(dms)
    assert float((vj1 - vj).abs().max()) < 1e-11 and float((vk1 - vk).abs().max()) < 1e-11
    js, ks = _lib.StoredERI(db, 0, nb).run(dms)
    assert float((vj - js).abs().max()) < 1e-10 and float((vk - ks).abs().max()) < 1e-10
    # J only / K only launches, and three "ranks" adding up
    j_only, _ = plan.run(dms, True, False)
    _, k_only = plan.run(dms, False, True)
    assert float((j_only - js).abs().max()) < 1e-10 and float((k_only - ks).abs().max()) < 1e-10
    parts = [plan.run(dms[:1], rank=r, world=3) for r in range(3)]
    assert float((sum(p[0] for p in parts) - js[:1]).abs().max()) < 1e-10
    assert float((sum(p[1] for p in parts) - ks[:1]).abs().max()) < 1e-10
    monkeypatch.setenv("B200QC_JK_NOREG", "1")
    old = _lib.JKPlan(db, 0, nb, 1e-14)
    assert old.nquartets_reg == 0 and old.nquartets == plan.nquartets
    oj, ok = old.run(dms)
    assert float((vj - oj).abs().max()) < 1e-10 and float((vk - ok).abs().max()) < 1e-10


def test_dfj_matches_oracle(cuda):
    """K9 against the reference's DF-J ops (dfmol.py:66-76) on oracle integrals."""
    from dqc_b200 import _lib
    from oracle import cint
    bw, aw = _df_wrappers()
    atm, bas, env = aw.atm_bas_env
    b0, b1 = bw.shell_idxs
    a0, a1 = aw.shell_idxs
    j3c = torch.as_tensor(cint.int3c2e(atm, bas, env, (b0, b1, b0, b1, a0, a1)))
    j2c = torch.as_tensor(cint.int2c2e(atm, bas, env, (a0, a1, a0, a1)))
    inv = torch.inverse(j2c)
    nao, naux = j3c.shape[0], j3c.shape[2]
    g = torch.Generator().manual_seed(3)
    dm = torch.randn(nao, nao, dtype=torch.float64, generator=g)   # deliberately not symmetric
    temp = torch.einsum("ij,ijl->l", dm, j3c)
    coef = torch.einsum("l,lk->k", temp, inv)
    jref = torch.einsum("k,ijk->ij", coef, j3c)
    db = aw.device_basis(cuda)
    packed = _lib.int3c2e_packed(db, (b0, b1, b0, b1, a0, a1))
    vj = _lib.dfj(packed, nao, naux, inv.to(cuda), dm.to(cuda))
    scale = float(jref.abs().max())
    assert float((vj.cpu() - jref).abs().max()) < 1e-10 * scale
    t1 = _lib.dfj_pass1(packed, nao, naux, dm.to(cuda))
    assert float((t1.cpu() - temp).abs().max()) < 1e-10 * float(temp.abs().max())
    vj2 = _lib.dfj_pass2(packed, nao, naux, coef.to(cuda))
    assert float((vj2.cpu() - jref).abs().max()) < 1e-10 * scale


def test_dfj_large_random(cuda):
    """K9 at a size with many row slabs / column blocks and an odd naux, against torch on the device."""
    from dqc_b200 import _lib
    g = torch.Generator().manual_seed(0)
    nao, naux = 150, 4099
    full = torch.randn(nao, nao, naux, dtype=torch.float64, generator=g)
    full = (full + full.transpose(0, 1)).to(cuda)
    inv = torch.randn(naux, naux, dtype=torch.float64, generator=g).to(cuda) / naux
    dm = torch.randn(nao, nao, dtype=torch.float64, generator=g).to(cuda)
    packed = _lib.pack_tril(full)
    assert packed.shape[1] == 4100
    vj = _lib.dfj(packed, nao, naux, inv, dm)
    temp = torch.einsum("ij,ijl->l", dm, full)
    ref = torch.einsum("k,ijk->ij", temp @ inv, full)
    assert float((vj - ref).abs().max() / ref.abs().max()) < 1e-12


@pytest.mark.parametrize("which", ["h2o-def2svp", "highl-sp"])
def test_multipole_integrals_match_grid_quadrature_and_symmetry(cuda, which):
    """int1e_r / int1e_rr / int1e_rrr (the electric-field terms of the core Hamiltonian, hcgto.py:118-127) from the
    overlap kernel's raw cartesian mode, against numerical quadrature of phi_i r_a r_b phi_j with the ORACLE's AO values
    on a Becke grid, and against the exact identities (symmetry in i, j and in the component indices; the dipole of an
    s-function pair at the same centre is the centre times the overlap)."""
    from dqc_b200.hamilton.intor import molintor as intor
    from dqc_b200.grid.factory import get_predefined_grid
    from oracle import cint
    zs, pos = util.H2O
    basis = "def2-svp" if which == "h2o-def2svp" else "3-21g"
    w, _ = util.make_wrapper(zs, [list(map(float, p)) for p in pos], basis)
    s = intor.overlap(w)
    orders = (1, 2) if which == "h2o-def2svp" else (1, 2, 3)
    grid = get_predefined_grid(4, zs, torch.tensor(pos, dtype=torch.float64, device=cuda), device=cuda)
    xyz, dv = grid.get_rgrid().cpu().numpy(), grid.get_dvolume().cpu().numpy()
    ao = cint.eval_gto(*w.atm_bas_env, xyz, 0)
    for n in orders:
        m = intor.int1e("r0" * n, w)
        assert m.shape == (3 ** n, w.nao(), w.nao())
        assert float((m - m.transpose(-2, -1)).abs().max()) < 1e-12
        mm = m.reshape(*([3] * n), w.nao(), w.nao())
        if n == 2:
            assert float((mm - mm.transpose(0, 1)).abs().max()) < 1e-12
        for ic in range(3 ** n):
            comps = np.unravel_index(ic, [3] * n)
            f = np.prod([xyz[:, d] for d in comps], axis=0)
            quad = (ao * (f * dv)[:, None]).T @ ao
            assert np.abs(m[ic].cpu().numpy() - quad).max() < 2e-5, (n, ic)
    # first s function of the oxygen with itself: <r> = R_O S
    d1 = intor.int1e("r0", w)
    for d in range(3):
        assert abs(float(d1[d, 0, 0]) - float(pos[0][d]) * float(s[0, 0])) < 1e-12


def test_efield_enters_the_core_hamiltonian(cuda):
    """Mol(..., efield=E): h = T + V + sum_d E_d <r_d> (hcgto.py:118-127) -- through HamiltonCGTO.build, and the HF energy
    responds to the field with the dipole moment (finite difference of E(field) = -mu . E to first order ... here just
    the sign convention: the core Hamiltonian difference equals the dipole matrix contracted with the field)."""
    from dqc_b200 import Mol
    from dqc_b200.hamilton.intor import molintor as intor
    zs, pos = util.H2O
    e = torch.tensor([0.01, -0.02, 0.03], dtype=torch.float64)
    mk = lambda ef: Mol((torch.tensor(zs), torch.tensor(pos, dtype=torch.float64)), basis="3-21g", device=cuda,
                        orthogonalize_basis=False, efield=ef).get_hamiltonian().build()
    h0, h1 = mk(None), mk(e)
    w, _ = util.make_wrapper(zs, [list(map(float, p)) for p in pos], "3-21g")
    dip = intor.int1e("r0", w)
    diff = h1.get_kinnucl().fullmatrix() - h0.get_kinnucl().fullmatrix()
    assert float((diff - torch.einsum("dab,d->ab", dip, e.to(dip.device))).abs().max()) < 1e-12


def test_dfj_row_skipping(cuda):
    """The row-skipping DF-J passes (b200qc_dfj_rowmask / _pass1_rows / _pass2_masked) give exactly what the plain passes
    give on a copy of (ij|P) whose masked pair rows are zeroed, and the mask is the row-maximum test it says it is."""
    from dqc_b200 import _lib
    bw, aw = _df_wrappers()
    b0, b1 = bw.shell_idxs
    a0, a1 = aw.shell_idxs
    db = aw.device_basis(cuda)
    packed = _lib.int3c2e_packed(db, (b0, b1, b0, b1, a0, a1))
    loc = aw.full_shell_to_aoloc
    nao, naux = int(loc[b1] - loc[b0]), int(loc[a1] - loc[a0])
    assert packed.shape[0] == nao * (nao + 1) // 2 and packed.shape[1] >= naux
    rowmax = packed[:, :naux].abs().amax(1)
    thresh = float(rowmax.median())
    mask = _lib.dfj_rowmask(packed, nao, naux, thresh)
    assert torch.equal(mask.bool(), rowmax < thresh) and 0 < int(mask.sum()) < mask.numel()
    zeroed = packed.clone()
    zeroed[mask.bool()] = 0.0
    g = torch.Generator().manual_seed(5)
    dm = torch.randn(nao, nao, dtype=torch.float64, generator=g).to(cuda)
    coef = torch.randn(naux, dtype=torch.float64, generator=g).to(cuda)
    rows = torch.nonzero(mask == 0).flatten().to(torch.int32)
    t_ref, t_m = _lib.dfj_pass1(zeroed, nao, naux, dm), _lib.dfj_pass1(packed, nao, naux, dm, rows)
    assert float((t_ref - t_m).abs().max()) <= 1e-13 * float(t_ref.abs().max())
    j_ref, j_m = _lib.dfj_pass2(zeroed, nao, naux, coef), _lib.dfj_pass2(packed, nao, naux, coef, mask)
    assert float((j_ref - j_m).abs().max()) <= 1e-13 * float(j_ref.abs().max())
