"""GPU tests of the tcgen05 int8 GEMM engine (csrc/gemm_i8.cuh) and of the density-fitted exact exchange built
on it.  The reference has no DF-K (hcgto.py:229-230 raises): parity is against the oracle's restatement of the
SURVEY 8a formula (oracle/fock_ref.py:get_exchange_df) and, loosely, against the 4-centre K it approximates."""
import numpy as np
import pytest
import torch
from tests import util

pytestmark = pytest.mark.gpu
dtype = torch.float64


def _rand(shape, seed, cuda, spread=3.0):
    """Entries spanning several orders of magnitude (exercises the per-row scales)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, dtype=dtype, generator=g)
    return (x * torch.exp(spread * torch.randn(shape[0], 1, dtype=dtype, generator=g))).to(cuda)


@pytest.mark.parametrize("M,N,K,S", [(128, 64, 32, 6), (300, 130, 1000, 6), (129, 65, 33, 6), (1, 1, 1, 6),
                                     (257, 200, 515, 5), (64, 840, 4096, 6)])
def test_gemm_f64emu_matches_fp64(cuda, M, N, K, S):
    from dqc_b200 import _lib
    x, y = _rand((M, K), 1, cuda), _rand((N, K), 2, cuda)
    got = _lib.gemm_f64emu(x, y, nslice=S)
    want = x @ y.T
    # worst-case truncation error of the scheme: 4 (S + 1) 2^(-7 S) K rowmax(x) rowmax(y), doubled
    bound = x.abs().max(1).values[:, None] * y.abs().max(1).values[None, :] * K * (S + 1) * 2.0 ** (-7 * S + 3)
    assert bool(((got - want).abs() <= bound + 1e-300).all())
    tol = 1e-10 if S == 6 else 1e-8
    assert float(((got - want).abs() / (x.abs() @ y.abs().T + 1e-300)).max()) < tol


def test_gemm_split_k_chunks_and_atomics(cuda):
    """Long K: several K chunks, every chunk with its own row scales, atomically accumulated."""
    from dqc_b200 import _lib
    M, N, K = 200, 100, 9000
    x, y = _rand((M, K), 3, cuda), _rand((N, K), 4, cuda)
    got = _lib.gemm_f64emu(x, y, nslice=6, kchunk=2048)      # 5 chunks, the last one partial
    want = x @ y.T
    assert float(((got - want).abs() / (x.abs() @ y.abs().T)).max()) < 1e-11


def test_slicer_row_contiguous_and_lower_triangle_mode(cuda):
    """The strided (row-contiguous) slicer and the symmetric lower-triangle tile mode: Z Z^T from Z^T storage."""
    from dqc_b200 import _lib
    n, K = 300, 777
    zt = _rand((K, n), 5, cuda, spread=0.5)                   # element (r, k) of Z at zt[k][r]: sr = 1, sk = n
    a = _lib.I8Operand("A", 1, n, K, 6, device=cuda).fill(zt, 0, 1, n)
    b = _lib.I8Operand("B", 1, n, K, 6, device=cuda).fill(zt, 0, 1, n)
    out = torch.zeros(n, n, dtype=dtype, device=cuda)
    _lib.gemm_i8(a, b, out, 0, n, n, n, mode=2, alpha=0.5)
    want = 0.5 * zt.T @ zt
    got = torch.tril(out) + torch.tril(out, -1).T
    # the scheme's error is relative to (row max) x (row max) x K, not to the result
    rmax = zt.abs().max(0).values
    assert float(((got - want).abs() / (rmax[:, None] * rmax[None, :] * K)).max()) < 2e-12
    # tiles strictly above the diagonal band were skipped
    assert float(out[:128, 256:].abs().max()) == 0.0


def _df_setup(cuda, zs, pos, basis, aux, orthozer):
    from dqc_b200 import Mol
    from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
    from oracle import fock_ref
    mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=dtype)), basis=basis, device=cuda,
              orthogonalize_basis=orthozer).densityfit(auxbasis=aux)
    h = mol.get_hamiltonian().build()
    w, _ = util.make_wrapper(zs, pos, basis)
    auxw, _ = util.make_wrapper(zs, pos, aux)
    bw, aw = LibcintWrapper.concatenate(w, auxw)
    ref = fock_ref.RefHamilton(bw, auxwrapper=aw, orthozer=orthozer).build_df()
    return mol, h, ref


@pytest.mark.parametrize("zs_pos", [util.H2O, util.CH4ISH])
def test_dfk_matches_oracle_fixed_dm(cuda, zs_pos):
    zs, pos = zs_pos
    mol, h, ref = _df_setup(cuda, zs, pos, "def2-svp", "etb-jfit", False)
    nocc = sum(zs) // 2
    dm = util.seeded_dm(h.nao, nocc, seed=3)
    got = h.get_exchange(dm.to(cuda)).fullmatrix().cpu()
    want = ref.get_exchange_df(dm)
    assert float((got - want).abs().max()) < 1e-8
    # indefinite, non-symmetric input: eigen-decomposition route with both signs
    g = torch.Generator().manual_seed(5)
    dm2 = dm + 0.1 * torch.randn(h.nao, h.nao, dtype=dtype, generator=g)
    got2 = h.get_exchange(dm2.to(cuda)).fullmatrix().cpu()
    assert float((got2 - ref.get_exchange_df(dm2)).abs().max()) < 1e-8
    # energy
    assert abs(float(h.get_e_exchange(dm.to(cuda))) - float(0.5 * torch.einsum("ij,ji->", want, dm))) < 1e-8


def test_dfk_orbital_tag_and_spin(cuda):
    """Densities made by ao_orb2dm skip the eigen-decomposition; both routes give the same K; UHF = per-spin K[2 D_s]."""
    from dqc_b200.utils.datastruct import SpinParam
    zs, pos = util.H2O
    mol, h, ref = _df_setup(cuda, zs, pos, "def2-svp", "etb-jfit", True)
    g = torch.Generator().manual_seed(11)
    q, _ = torch.linalg.qr(torch.randn(h.nao, h.nao, dtype=dtype, generator=g))
    orb = q[:, :5].to(cuda)
    wu, wd = torch.tensor([1.0, 1.0, 1.0, 1.0, 1.0], dtype=dtype, device=cuda), \
        torch.tensor([1.0, 1.0, 1.0, 0.5, 0.0], dtype=dtype, device=cuda)
    du, dd = h.ao_orb2dm(orb, wu), h.ao_orb2dm(orb, wd)
    assert hasattr(du, "_b200_orb")
    k_tag = h.get_exchange(du).fullmatrix()
    k_eig = h.get_exchange(du.clone()).fullmatrix()
    assert float((k_tag - k_eig).abs().max()) < 1e-10
    # an in-place edit invalidates the tag (tensor versions are part of it): K is linear in D
    du_half = h.ao_orb2dm(orb, wu)
    du_half.mul_(0.5)
    assert float((h.get_exchange(du_half).fullmatrix() - 0.5 * k_tag).abs().max()) < 1e-10
    ks = h.get_exchange(SpinParam(u=du, d=dd))
    # compare in the AO basis (the orthogonaliser's eigenvectors are defined up to sign / rotation)
    Xg, Xr = h._orthozer._orthozer.cpu(), ref.X
    to_ao = lambda m, X: torch.linalg.pinv(X.T) @ m @ torch.linalg.pinv(X)
    for got, d in ((ks.u, du), (ks.d, dd)):
        d_ao = Xg @ d.cpu() @ Xg.T
        d_ref = torch.linalg.pinv(Xr) @ d_ao @ torch.linalg.pinv(Xr).T
        want = ref.get_exchange_df(2 * d_ref)
        assert float((to_ao(got.fullmatrix().cpu(), Xg) - to_ao(want, Xr)).abs().max()) < 1e-8


def test_dfk_approximates_four_centre_k(cuda):
    """The fitted K against the exact 4-centre K of the same molecule (what the extension approximates)."""
    from dqc_b200 import Mol
    zs, pos = util.H2O
    mol, h, _ = _df_setup(cuda, zs, pos, "def2-svp", "etb-jfit", False)
    h4 = Mol((torch.tensor(zs), torch.tensor(pos, dtype=dtype)), basis="def2-svp", device=cuda,
             orthogonalize_basis=False).get_hamiltonian().build()
    dm = util.seeded_dm(h.nao, 5, seed=0).to(cuda)
    k_df, k_4c = h.get_exchange(dm).fullmatrix(), h4.get_exchange(dm).fullmatrix()
    e_df, e_4c = float(h.get_e_exchange(dm)), float(h4.get_e_exchange(dm))
    assert float((k_df - k_4c).abs().max()) < 5e-2
    assert abs(e_df - e_4c) < 1e-2 * abs(e_4c)


def test_hybrid_fock_with_df(cuda):
    """get_fock_2e(exx) with density fitting = J + exx K' + Vxc from the individual members."""
    from dqc_b200 import get_xc
    zs, pos = util.H2O
    mol, h, _ = _df_setup(cuda, zs, pos, "def2-svp", "etb-jfit", True)
    mol.setup_grid()
    h.setup_grid(mol.get_grid(), get_xc("0.75 * gga_x_pbe + gga_c_pbe"))
    dm = util.seeded_dm(h.nao, 5, seed=2).to(cuda)
    f = h.get_fock_2e(dm, exx=0.25).fullmatrix()
    want = h.get_elrep(dm).fullmatrix() + 0.25 * h.get_exchange(dm).fullmatrix() + h.get_vxc(dm).fullmatrix()
    assert float((f - want).abs().max()) < 1e-10


def test_hybrid_scf_with_df_exchange_against_four_centre_exchange(cuda):
    """End to end through the KS driver: B3LYP composition (0.08 Slater + 0.72 B88 + 0.19 VWN-RPA + 0.81 LYP + 0.2 K).
    The density-fitted build (DF-J + DF-K on tcgen05) and the 4-centre build (stored-ERI J, K) converge to energies
    that differ only by the fitting error of the J-fit aux basis; the open-shell driver agrees with the closed-shell
    one through the per-spin K[2 D_s] route."""
    from dqc_b200 import Mol, KS
    zs, pos = util.H2O
    xc = "0.08*lda_x + 0.72*gga_x_b88 + 0.19*lda_c_vwn_rpa + 0.81*gga_c_lyp"
    mk = lambda: Mol((torch.tensor(zs), torch.tensor(pos, dtype=dtype)), basis="def2-svp", grid="sg2", device=cuda)
    e_4c = float(KS(mk(), xc=xc, exx_fraction=0.2).run().energy())
    e_df = float(KS(mk().densityfit(auxbasis="etb-jfit"), xc=xc, exx_fraction=0.2).run().energy())
    e_df_u = float(KS(mk().densityfit(auxbasis="etb-jfit"), xc=xc, exx_fraction=0.2, restricted=False).run().energy())
    assert -76.5 < e_4c < -76.0                      # B3LYP/def2-SVP water is about -76.3 Ha
    assert abs(e_df - e_4c) < 2e-3
    assert abs(e_df_u - e_df) < 1e-7
