"""Self-checks of the integral oracle (oracle/cint_oracle.c) that need no reference library (the
reference's own integral tests compare against a live PySCF, dqc/test/test_libcint.py:83-199, which is absent here):
closed forms for s-type primitives (Boys F0), normalisation S_mu_mu = 1 for every l <= 4, the 8-fold symmetry of
(ij|kl), symmetry of the one-electron and 2-/3-centre integrals, and shell-subset consistency in the style of
test_libcint.py:145-199.  CPU tests."""
import math
import numpy as np
import pytest
import torch
from scipy.special import erf
from tests import util
from oracle import cint
from dqc_b200.utils.datastruct import AtomCGTOBasis, CGTOBasis
from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper

dtype = torch.float64


def _s_primitives(alphas, centres, zs):
    abs_ = [AtomCGTOBasis(atomz=z, bases=[CGTOBasis(angmom=0, alphas=torch.tensor([a], dtype=dtype),
                                                    coeffs=torch.tensor([1.0], dtype=dtype))],
                          pos=torch.tensor(c, dtype=dtype)) for a, c, z in zip(alphas, centres, zs)]
    return LibcintWrapper(abs_)


def _f0(t):
    return 1.0 if t < 1e-14 else 0.5 * math.sqrt(math.pi / t) * erf(math.sqrt(t))


def test_s_type_closed_forms():
    al = [0.8, 1.7, 0.35, 2.4]
    ctr = np.array([[0.0, 0.0, 0.0], [0.9, -0.4, 0.3], [-0.7, 1.1, 0.2], [0.2, 0.5, -1.3]])
    zs = [1, 2, 3, 4]
    w = _s_primitives(al, ctr.tolist(), zs)
    atm, bas, env = w.atm_bas_env
    n = [(2 * a / math.pi) ** 0.75 for a in al]            # normalised s Gaussians
    S, T, V = (cint.int1e(k, atm, bas, env) for k in ("ovlp", "kin", "nuc"))
    eri = cint.int2e(atm, bas, env)
    for i in range(4):
        for j in range(4):
            a, b = al[i], al[j]
            p, mu = a + b, a * b / (a + b)
            r2 = float(((ctr[i] - ctr[j]) ** 2).sum())
            kab = math.exp(-mu * r2)
            s_ref = n[i] * n[j] * (math.pi / p) ** 1.5 * kab
            assert abs(S[i, j] - s_ref) < 1e-13
            assert abs(T[i, j] - s_ref * mu * (3 - 2 * mu * r2)) < 1e-13
            pc = (a * ctr[i] + b * ctr[j]) / p
            v_ref = -sum(z * n[i] * n[j] * 2 * math.pi / p * kab * _f0(p * float(((pc - c) ** 2).sum()))
                         for z, c in zip(zs, ctr))
            assert abs(V[i, j] - v_ref) < 1e-12
            for k in range(4):
                for l in range(4):
                    c, d = al[k], al[l]
                    q = c + d
                    kcd = math.exp(-c * d / q * float(((ctr[k] - ctr[l]) ** 2).sum()))
                    qc = (c * ctr[k] + d * ctr[l]) / q
                    t = p * q / (p + q) * float(((pc - qc) ** 2).sum())
                    ref = n[i] * n[j] * n[k] * n[l] * 2 * math.pi ** 2.5 / (p * q * math.sqrt(p + q)) * kab * kcd * _f0(t)
                    assert abs(eri[i, j, k, l] - ref) < 1e-12


@pytest.mark.parametrize("which", ["h2o-321g", "highl"])
def test_normalisation_and_symmetries(which):
    w = util.make_wrapper(*util.H2O, "3-21g")[0] if which == "h2o-321g" else util.highl_wrapper()[0]
    atm, bas, env = w.atm_bas_env
    S = cint.int1e("ovlp", atm, bas, env)
    assert np.abs(np.diag(S) - 1.0).max() < 1e-12           # wfnormalize_ + the s/p solid-harmonic constants
    for kind in ("ovlp", "kin", "nuc"):
        m = cint.int1e(kind, atm, bas, env)
        assert np.abs(m - m.T).max() < 1e-12
    j2c = cint.int2c2e(atm, bas, env)
    assert np.abs(j2c - j2c.T).max() < 1e-11 and np.linalg.eigvalsh(j2c).min() > 0      # a Coulomb metric
    nb = len(bas)
    sub = (0, min(nb, 4))
    eri = cint.int2e(atm, bas, env, sub * 4)
    for perm in ((1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1), (1, 0, 3, 2), (3, 2, 1, 0)):
        assert np.abs(eri - eri.transpose(perm)).max() < 1e-11
    j3c = cint.int3c2e(atm, bas, env, sub * 2 + (0, nb))
    assert np.abs(j3c - j3c.transpose(1, 0, 2)).max() < 1e-11


def test_shell_subsets_are_blocks_of_the_full_tensors():
    w, _ = util.make_wrapper(*util.CH4ISH, "3-21g")
    atm, bas, env = w.atm_bas_env
    loc = cint.ao_loc_sph(bas)
    nb = len(bas)
    a, b, c, d = (1, 4), (0, 3), (2, nb), (3, 5)
    sl = lambda s: slice(int(loc[s[0]]), int(loc[s[1]]))
    for kind in ("ovlp", "kin", "nuc"):
        full = cint.int1e(kind, atm, bas, env)
        assert np.abs(cint.int1e(kind, atm, bas, env, a + c) - full[sl(a), sl(c)]).max() < 1e-13
    full2 = cint.int2c2e(atm, bas, env)
    assert np.abs(cint.int2c2e(atm, bas, env, b + d) - full2[sl(b), sl(d)]).max() < 1e-13
    full4 = cint.int2e(atm, bas, env)
    assert np.abs(cint.int2e(atm, bas, env, a + b + c + d) - full4[sl(a), sl(b), sl(c), sl(d)]).max() < 1e-13
    full3 = cint.int3c2e(atm, bas, env, (0, nb) * 3)
    assert np.abs(cint.int3c2e(atm, bas, env, a + b + d) - full3[sl(a), sl(b), sl(d)]).max() < 1e-13
    # (ij|P) of the 3-centre routine equals the 4-centre integral with a unit s function of zero exponent? no such
    # shell exists in the basis: instead contract consistency, (ij|P) with P an ordinary shell equals (ij|P 1) only
    # through int2c2e/int3c2e conventions -- checked against the GPU kernels in tests/test_gpu_integrals.py


def test_ao_gradient_is_the_derivative_of_the_value():
    w, _ = util.highl_wrapper()
    atm, bas, env = w.atm_bas_env
    pts = util.random_points(40, seed=3, span=1.5)
    g = cint.eval_gto(atm, bas, env, pts, 1)
    h = 1e-5
    for d in range(3):
        e = np.zeros(3)
        e[d] = h
        fd = (cint.eval_gto(atm, bas, env, pts + e, 0) - cint.eval_gto(atm, bas, env, pts - e, 0)) / (2 * h)
        assert np.abs(fd - g[d]).max() < 1e-8


def test_ao_laplacian_is_the_divergence_of_the_gradient():
    """eval_gto(deriv=2) (the oracle of eval_laplgto) against central differences of the analytic gradient, l = 0..4."""
    w, _ = util.highl_wrapper()
    atm, bas, env = w.atm_bas_env
    rng = np.random.RandomState(11)
    pts = rng.uniform(-1.5, 1.5, size=(40, 3))
    lap = cint.eval_gto(atm, bas, env, pts, 2)
    h = 1e-4
    fd = np.zeros_like(lap)
    for d in range(3):
        e = np.zeros(3)
        e[d] = h
        fd += (cint.eval_gto(atm, bas, env, pts + e, 1)[d] - cint.eval_gto(atm, bas, env, pts - e, 1)[d]) / (2 * h)
    scale = np.abs(lap).max()
    assert np.abs(lap - fd).max() < 1e-6 * scale
