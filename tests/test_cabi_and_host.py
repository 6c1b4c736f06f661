"""CPU-side tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/b200qc.h
declares (no compute call is made without a GPU), the product path refuses to run without CUDA,
and the host-side logic (basis reader, libcint packing, occupations, grids, XC algebra, work
splitting) behaves like the reference's."""
import ctypes
import os
import re
import numpy as np
import pytest
import torch
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dtype = torch.float64


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200qc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200qc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from dqc_b200 import _lib
    so = os.path.join(ROOT, "dqc_b200", "libb200qc.so")
    assert os.path.exists(so), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(so)
    names = _header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libb200qc.so does not export %s" % n
    # the ctypes table of the binding covers the header exactly
    assert sorted(_lib.exported_symbols()) == names
    _lib.load(require_cuda=False)
    assert _lib.load(False).b200qc_version() >= 100
    assert _lib.launch_count() == 0


def test_binding_argument_counts_match_the_header():
    """Every prototype of include/b200qc.h against the ctypes table of dqc_b200/_lib.py: same number of arguments,
    pointer arguments bound as pointers, 64-bit sizes as c_int64 (an ABI drift here corrupts the stack silently)."""
    from dqc_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "b200qc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = re.findall(r"\b(b200qc_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S)
    assert len(protos) >= 30
    seen = set()
    for name, args in protos:
        seen.add(name)
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        _, argtypes = _lib._SIGS[name]
        assert len(params) == len(argtypes), "%s: header has %d arguments, the binding %d" % (name, len(params), len(argtypes))
        for prm, ct in zip(params, argtypes):
            if "*" in prm:
                assert ct in (ctypes.c_void_p, ctypes.c_char_p), "%s: '%s' must be bound as a pointer" % (name, prm)
            elif prm.startswith("int64_t"):
                assert ct is ctypes.c_int64, "%s: '%s' must be bound as c_int64" % (name, prm)
            elif prm.startswith("double"):
                assert ct is ctypes.c_double, "%s: '%s' must be bound as c_double" % (name, prm)
            elif prm.startswith("int "):
                assert ct is ctypes.c_int, "%s: '%s' must be bound as c_int" % (name, prm)
    assert seen == set(_lib._SIGS)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from dqc_b200 import _lib, Mol
    with pytest.raises(_lib.B200QCError):
        _lib.load(require_cuda=True)
    with pytest.raises(_lib.B200QCError):
        Mol("H 0 0 0; H 0 0 1.4", basis="3-21G")     # the Hamiltonian needs the CUDA path


def test_product_path_never_imports_the_oracle():
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "dqc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "liboracle" in src:
                    bad.append(os.path.join(dp, f))
    assert not bad, "product files touching oracle/: %s" % bad


# ---- basis ingestion / packing (dqc/api/loadbasis.py:11-152, lcintwrap.py:24-123) ----
def test_loadbasis_grammar_and_normalisation():
    from dqc_b200 import loadbasis
    b = loadbasis("6:3-21G")
    assert [s.angmom for s in b] == [0, 0, 1, 0, 1]          # S, SP -> s + p, SP -> s + p
    assert [len(s.alphas) for s in b] == [3, 2, 2, 1, 1]
    from dqc_b200.utils.datastruct import gaussian_int
    for s in b:                                               # radial self-overlap == 1 (datastruct.py:34-61)
        l2 = 2 * s.angmom + 2
        pair = gaussian_int(l2, s.alphas.unsqueeze(-1) + s.alphas.unsqueeze(-2))
        assert abs(float(torch.einsum("a,ab,b", s.coeffs, pair, s.coeffs)) - 1.0) < 1e-12
    with pytest.raises(FileNotFoundError):
        loadbasis("92:3-21G")


def test_libcint_layout():
    w, pos = util.make_wrapper(*util.H2O, "3-21g")
    atm, bas, env = w.atm_bas_env
    assert atm.shape == (3, 6) and bas.shape[1] == 8 and atm.dtype == np.int32
    assert list(atm[:, 0]) == [8, 1, 1] and all(atm[:, 2] == 1)
    assert atm[0, 1] == 20                                                   # 20-slot env header (lcintwrap.py:37)
    assert np.allclose(env[atm[1, 1]:atm[1, 1] + 3], pos[1].numpy())
    assert all(bas[:, 3] == 1) and all(bas[:, 4] == 0)
    assert all(bas[:, 6] == bas[:, 5] + bas[:, 2])                           # coefficients follow exponents
    assert list(w.full_shell_to_aoloc) == list(np.concatenate([[0], np.cumsum(2 * bas[:, 1] + 1)]))
    sub = w[1:3]
    assert sub.shell_idxs == (1, 3) and sub.parent is w and len(sub) == 2
    assert sub.nao() == int(w.full_shell_to_aoloc[3] - w.full_shell_to_aoloc[1])


def test_concatenate_wrappers():
    from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
    w1, _ = util.make_wrapper(*util.H2O, "3-21g")
    w2, _ = util.make_wrapper(*util.H2O, "etb-jfit")
    a, b = LibcintWrapper.concatenate(w1, w2)
    assert a.parent is b.parent and a.shell_idxs == (0, len(w1)) and b.shell_idxs == (len(w1), len(w1) + len(w2))
    assert a.parent.natoms == 6
    c, d, e = LibcintWrapper.concatenate(w1, w2, w1)                         # repeated parent is not duplicated
    assert c.shell_idxs == e.shell_idxs and c.parent.natoms == 6
    assert LibcintWrapper.concatenate(w1, w1[0:2])[1].shell_idxs == (0, 2)


# ---- occupations, nuclear repulsion (mol.py:252-260,395-443; test_system.py) ----
def test_occupations_and_parsing():
    from dqc_b200.system.mol import _get_nelecs_spin, _get_orb_weights
    from dqc_b200.api.parser import parse_moldesc
    zs, pos = parse_moldesc("H 1 0 0; Be -1 0 0.5")
    assert list(zs) == [1, 4] and pos.shape == (2, 3)
    n, spin, frac = _get_nelecs_spin(torch.tensor(8), None, 0)
    assert (int(n), int(spin), frac) == (8, 0, False)
    ow, owu, owd = _get_orb_weights(torch.tensor(8), 2, False, dtype, torch.device("cpu"))
    assert ow.tolist() == [2, 2, 2, 1, 1] and owu.tolist() == [1] * 5 and owd.tolist() == [1] * 3
    ow, owu, owd = _get_orb_weights(torch.tensor(1), 1, False, dtype, torch.device("cpu"))
    assert ow.tolist() == [1] and owd.tolist() == [0]
    with pytest.raises(AssertionError):
        _get_nelecs_spin(torch.tensor(8), 1, 0)


def test_nuclei_energy_formula():
    from oracle import fock_ref
    # dqc/test/test_system.py:61-73: two protons 1.5 apart with Z = 2, 3 -> 2*3/1.5 = 4
    assert abs(fock_ref.nuclei_energy([2, 3], [[0, 0, 0], [1.5, 0, 0]]) - 4.0) < 1e-14
    z = torch.tensor([2.0, 3.0], dtype=dtype)
    pos = torch.tensor([[0, 0, 0], [1.5, 0, 0]], dtype=dtype)
    r12 = torch.cdist(pos, pos) + torch.diag(torch.full((2,), float("inf"), dtype=dtype))
    assert abs(float((z[:, None] * z[None, :] / r12).sum() * 0.5) - 4.0) < 1e-14


# ---- grids (dqc/test/test_grid.py:16-104) ----
@pytest.mark.parametrize("spec", [3, 4, "sg2", "sg3"])
def test_predefined_atomic_grid_integrates_gaussian(spec):
    from dqc_b200.grid.factory import get_predefined_grid
    g = get_predefined_grid(spec, [6], torch.zeros(1, 3, dtype=dtype), device=torch.device("cpu"))
    r = g.get_rgrid().norm(dim=-1)
    val = float((torch.exp(-r * r * 0.5) * g.get_dvolume()).sum())
    assert abs(val - 2 * np.sqrt(2 * np.pi) * np.pi) < 1e-8 * 16
    assert g.coord_type == "cart"


def test_radial_and_lebedev_grids():
    from dqc_b200.grid.radial_grid import RadialGrid
    from dqc_b200.grid.lebedev_grid import LebedevGrid
    for integ in ("chebyshev", "chebyshev2", "uniform"):
        for tf in ("logm3", "de2", "treutlerm4"):
            rg = RadialGrid(200, integ, tf)
            r = rg.get_rgrid().reshape(-1)
            val = float((torch.exp(-r * r * 0.5) * rg.get_dvolume()).sum())
            assert abs(val - 2 * np.sqrt(2 * np.pi) * np.pi) < 2e-4 * 16, (integ, tf)
    g = LebedevGrid(RadialGrid(100, "chebyshev", "logm3"), prec=13)
    xyz = g.get_rgrid()
    f = torch.exp(-(xyz ** 2).sum(-1) * 0.5) * (1 + xyz[:, 0] * xyz[:, 1])     # odd part integrates to 0
    assert abs(float((f * g.get_dvolume()).sum()) - 2 * np.sqrt(2 * np.pi) * np.pi) < 1e-5 * 16


def test_two_atom_becke_oracle_integrates_two_gaussians():
    # test_grid.py:80-104 with the oracle's Becke weights (the product path does them on the GPU)
    from dqc_b200.grid.factory import get_predefined_grid
    from oracle import becke_ref
    pos = np.array([[-0.7, 0.0, 0.0], [0.9, 0.0, 0.0]])
    one = get_predefined_grid("sg2", [1], torch.zeros(1, 3, dtype=dtype), device=torch.device("cpu"))
    pts = np.concatenate([one.get_rgrid().numpy() + p for p in pos])
    dv = np.concatenate([one.get_dvolume().numpy()] * 2)
    owner = np.repeat([0, 1], one.get_rgrid().shape[0])
    w = becke_ref.becke_weights(pts, owner, pos)
    f = sum(np.exp(-((pts - p) ** 2).sum(-1) * 0.5) for p in pos)
    assert abs((f * dv * w).sum() - 2 * 2 * np.sqrt(2 * np.pi) * np.pi) < 3e-3 * 32


# ---- XC algebra and parsing (base_xc.py:183-268, getxc.py:38-59) ----
def test_get_xc_algebra_without_gpu():
    from dqc_b200 import get_xc, ValGrad, SpinParam
    from dqc_b200.xc import B200XC, BaseXC, AddBaseXC
    xc = get_xc("lda_x + 0.5*gga_c_pbe")
    assert isinstance(xc, B200XC) and xc.family == 2
    assert xc.terms == [(1.0, "lda_x"), (0.5, "gga_c_pbe")]
    assert get_xc("2*lda_x").terms == [(2.0, "lda_x")] and get_xc("lda_x*2").terms == [(2.0, "lda_x")]
    with pytest.raises(NotImplementedError):
        get_xc("hyb_gga_xc_b3lyp")
    assert get_xc("mgga_x_scan").family == 4 and get_xc("mgga_x_scan + gga_c_pbe").family == 4
    with pytest.raises(NotImplementedError):
        get_xc("mgga_c_scan")

    class MyLDA(BaseXC):          # user functional: default get_vxc goes through autograd (base_xc.py:39-125)
        @property
        def family(self):
            return 1

        def get_edensityxc(self, densinfo):
            if isinstance(densinfo, SpinParam):
                return 0.5 * (self.get_edensityxc(densinfo.u * 2) + self.get_edensityxc(densinfo.d * 2))
            return -0.7 * densinfo.value ** (4.0 / 3)
    rho = torch.rand(7, dtype=dtype) + 0.1
    v = MyLDA().get_vxc(ValGrad(value=rho))
    assert torch.allclose(v.value, -0.7 * 4.0 / 3 * rho ** (1.0 / 3))
    vp = (MyLDA() + MyLDA() * 2).get_vxc(SpinParam(u=ValGrad(value=rho), d=ValGrad(value=rho * 0.5)))
    assert torch.allclose(vp.u.value, 3 * -0.7 * 4.0 / 3 * (2 * rho) ** (1.0 / 3))      # spin scaling, (a + 2a)
    assert torch.allclose(vp.d.value, 3 * -0.7 * 4.0 / 3 * rho ** (1.0 / 3))
    assert isinstance(MyLDA() + xc, AddBaseXC)


def test_linear_operator_shim():
    from dqc_b200.utils.linop import LinearOperator
    a = torch.randn(4, 4, dtype=dtype)
    a = a + a.T
    op = LinearOperator.m(a, is_hermitian=True) + LinearOperator.m(2 * a, is_hermitian=True)
    assert op.is_hermitian and torch.allclose(op.fullmatrix(), 3 * a) and tuple(op.shape) == (4, 4)
    x = torch.randn(4, 3, dtype=dtype)
    assert torch.allclose(op.mm(x), 3 * a @ x) and torch.allclose(op.mv(x[:, 0]), 3 * a @ x[:, 0])


def test_split_rows_covers_everything():
    from dqc_b200.utils.dist import split_rows
    for n in (0, 1, 127, 128, 1000, 1060440):
        for world in (1, 2, 4, 8):
            segs = [split_rows(n, world, r, 128) for r in range(world)]
            assert segs[0][0] == 0 and segs[-1][1] == n
            for (a, b), (c, d) in zip(segs[:-1], segs[1:]):
                assert b == c and a <= b and (b % 128 == 0 or b == n)


def test_benchmark_geometries():
    from dqc_b200.utils import systems
    zs, pos = systems.c60()
    d = np.linalg.norm(pos[:, None] - pos[None], axis=-1) * 0.52917721092
    nb = np.sort(d, axis=1)[:, 1:4]
    assert len(zs) == 60 and np.allclose(nb[:, 0], 1.40, atol=1e-9) and np.allclose(nb[:, 1:], 1.45, atol=1e-9)
    zs, pos = systems.taxol_like()
    assert sorted(zs).count(6) == 47 and zs.count(1) == 51 and zs.count(7) == 1 and zs.count(8) == 14
    zs2, pos2 = systems.taxol_like()
    assert np.array_equal(pos, pos2)
    assert len(systems.carbon_cluster(36)[0]) == 36


def test_dfk_split_k_plan():
    """Host logic of the density-fitted exchange: the split-K chunks of stage 2 cover the contraction exactly, stay
    inside the exact-integer-accumulation limit and are multiples of the MMA K step."""
    from dqc_b200.df.dfmol import dfk_chunking
    for nao, nl, nocc, lim in [(840, 3420, 180, 32768), (1123, 4299, 226, 32768), (24, 84, 5, 32768), (7, 1, 1, 32768),
                               (840, 428, 180, 32768), (3010, 12255, 645, 32768), (840, 3420, 180, 65536)]:
        npad = (nocc + 63) // 64 * 64
        pc, nchunk, kc, ktot, k_last = dfk_chunking(nao, nl, npad, lim)
        assert pc >= 1 and kc == pc * npad and kc <= lim and kc % 32 == 0
        assert ktot == nl * npad and (nchunk - 1) * kc + k_last == ktot and 0 < k_last <= kc and k_last % 32 == 0
        assert nchunk == -(-nl // pc)


def test_equilibrium_solvers_on_a_contraction_map():
    """The fixed-point solvers of the SCF driver (scf_qccalc.equilibrium: DIIS default, Broyden-1, simple) on
    y = A y + b with ||A|| < 1, against the linear solve; and the convergence flag when maxiter is too small."""
    import warnings
    from dqc_b200.qccalc.scf_qccalc import equilibrium, ConvergenceWarning
    g = torch.Generator().manual_seed(0)
    a = torch.randn(12, 12, dtype=dtype, generator=g)
    a = 0.5 * a / torch.linalg.matrix_norm(a, 2)
    b = torch.randn(12, dtype=dtype, generator=g)
    exact = torch.linalg.solve(torch.eye(12, dtype=dtype) - a, b)
    for method in ("diis", "broyden1", "simple"):
        info = {}
        y = equilibrium(lambda v: a @ v + b, torch.zeros(12, dtype=dtype), method=method, maxiter=200, f_tol=1e-12, info=info)
        assert info["converged"] and float((y - exact).abs().max()) < 1e-10, method
    info = {}
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        equilibrium(lambda v: a @ v + b, torch.zeros(12, dtype=dtype), method="simple", maxiter=2, f_tol=1e-12, info=info)
    assert not info["converged"] and any(issubclass(r.category, ConvergenceWarning) for r in rec)


def test_jk_unit_table_is_in_sync_with_its_generator(tmp_path):
    """csrc/jk_reg_units.inc (class table, per-unit entry definitions, dispatcher of the register-resident J/K engine)
    is generated by tools/gen_jk_units.py: the committed file must be what the generator writes, every class must
    satisfy li >= lj, lk >= ll, (li, lj) >= (lk, ll), and the passes must divide the bra components."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_jk_units", os.path.join(root, "tools", "gen_jk_units.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    ncart = lambda l: (l + 1) * (l + 2) // 2
    seen = set()
    for li, lj, lk, ll, npass, parts, _ in gen.CLASSES:
        assert li >= lj and lk >= ll and (li, lj) >= (lk, ll) and (li, lj, lk, ll) not in seen
        seen.add((li, lj, lk, ll))
        assert (ncart(li) * ncart(lj)) % npass == 0 and npass % parts == 0
        assert (li + lj + lk + ll) // 2 + 1 <= 5
    out = str(tmp_path / "jk_reg_units.inc")
    gen.main(out)
    committed = open(os.path.join(root, "dqc_b200", "csrc", "jk_reg_units.inc")).read()
    assert open(out).read() == committed, "run python tools/gen_jk_units.py and commit csrc/jk_reg_units.inc"


def test_refined_rys_table_reproduces_the_base_table():
    """b200qc_rys_refine (host arithmetic; the form the register-resident J/K engine reads: intervals of width 1/2,
    Chebyshev degree 9) against the base table (width 1, degree 13) at random x: roots and weights agree to 2e-15."""
    import ctypes
    import os
    import numpy as np
    from numpy.polynomial import chebyshev as C
    from dqc_b200 import _lib
    lib = _lib.load(require_cuda=False)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rng = np.random.RandomState(3)
    with np.load(os.path.join(root, "dqc_b200", "data", "rys_table.npz")) as z:
        nmax, h, deg, xmax = z["meta"]
        for n in (1, 2, 3, 5):
            base = np.ascontiguousarray(z["coef_%d" % n], dtype=np.float64)      # (nint, 2n, deg + 1)
            nint, nf = base.shape[0], base.shape[1]
            fine = np.empty((nint * 2, nf, 10), dtype=np.float64)
            rc = lib.b200qc_rys_refine(base.ctypes.data_as(ctypes.c_void_p), nint, nf, int(deg), 2, 10,
                                       fine.ctypes.data_as(ctypes.c_void_p))
            assert rc == 0
            x = np.concatenate([rng.uniform(0.0, float(xmax), 4000), [0.0, 0.5, 1.0, float(xmax) - 1e-9]])
            it = np.minimum((x / h).astype(int), nint - 1)
            t = 2.0 * (x - it * h) / h - 1.0
            itf = np.minimum((x / (h / 2)).astype(int), 2 * nint - 1)
            tf = 2.0 * (x - itf * (h / 2)) / (h / 2) - 1.0
            for f in range(nf):
                vb = np.array([C.chebval(t[i], base[it[i], f]) for i in range(x.size)])
                vf = np.array([C.chebval(tf[i], fine[itf[i], f]) for i in range(x.size)])
                assert np.abs(vf - vb).max() <= 2e-15 * max(1.0, 0.0) + 2e-15 * np.abs(vb).max(), (n, f)
                assert (np.abs(vf - vb) / np.abs(vb)).max() < 5e-14, (n, f)      # also relatively, for the small roots
