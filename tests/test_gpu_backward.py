"""First slice of the backward pass: derivative integrals from the Rys kernels' raw cartesian mode,
autograd Functions of S / T / V / (ij|kl) / AO values with respect to the atomic positions -- checked against finite
differences of the CPU ORACLE at displaced geometries (the reference checks the same things with gradcheck against
libcint, dqc/test/test_libcint.py:201-466) and, end to end, the analytic RHF force against finite differences of the
converged energy."""
import numpy as np
import pytest
import torch
from tests import util

pytestmark = pytest.mark.gpu
dtype = torch.float64


def _wrapper_with_grad(zs, pos, basis):
    from dqc_b200.api.loadbasis import loadbasis
    from dqc_b200.utils.datastruct import AtomCGTOBasis
    from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
    p = torch.tensor(pos, dtype=dtype, requires_grad=True)
    ab = [AtomCGTOBasis(atomz=z, bases=loadbasis("%d:%s" % (z, basis)), pos=p[i]) for i, z in enumerate(zs)]
    return LibcintWrapper(ab), p


def _oracle_ints(zs, pos, basis, kind):
    from oracle import cint
    w, _ = util.make_wrapper(zs, pos, basis)
    atm, bas, env = w.atm_bas_env
    if kind == "eri":
        return cint.int2e(atm, bas, env)
    return cint.int1e(kind, atm, bas, env)


# ((ij|kl) with d shells needs the (f d|d d) class in cartesians, 2160 components: above the 1312 a lane group of the Rys
# kernel holds; the two-electron derivative is checked on s and p shells)
@pytest.mark.parametrize("basis,kind", [("3-21g", k) for k in ("ovlp", "kin", "nuc", "eri")] +
                         [("def2-svp", k) for k in ("ovlp", "kin", "nuc")])
def test_integral_position_gradients_match_oracle_finite_differences(cuda, basis, kind):
    """d/dR of sum_ij G_ij I_ij(R), G random, through the autograd Functions (derivative integrals on the GPU) against
    central differences of the oracle's integrals at displaced geometries: H2O, s, p and d shells."""
    from dqc_b200.hamilton.intor import molintor as intor
    zs, pos = util.H2O
    pos = [list(map(float, p)) for p in pos]
    w, p = _wrapper_with_grad(zs, pos, basis)
    fn = {"ovlp": intor.overlap, "kin": intor.kinetic, "nuc": intor.nuclattr, "eri": intor.elrep}[kind]
    val = fn(w)
    assert val.requires_grad
    g = torch.Generator().manual_seed(4)
    G = torch.randn(val.shape, dtype=dtype, generator=g).to(val.device)
    (grad,) = torch.autograd.grad((val * G).sum(), p)
    ref0 = _oracle_ints(zs, pos, basis, kind)
    assert float((val.detach().cpu() - torch.as_tensor(ref0)).abs().max()) < 1e-10
    h = 1e-4
    Gn = G.cpu().numpy()
    for ia in range(len(zs)):
        for d in range(3):
            pp = [list(r) for r in pos]
            pm = [list(r) for r in pos]
            pp[ia][d] += h
            pm[ia][d] -= h
            fd = ((_oracle_ints(zs, pp, basis, kind) - _oracle_ints(zs, pm, basis, kind)) * Gn).sum() / (2 * h)
            assert abs(float(grad[ia, d]) - fd) < 2e-6 * max(1.0, abs(fd)), (ia, d, float(grad[ia, d]), fd)


def test_ip_integrals_antisymmetry_and_translation(cuda):
    """<d phi_i|phi_j> + <phi_i|d phi_j> = 0 (integration by parts), and the sum over all centres of dV/dR vanishes
    (translational invariance) -- identities libcint's ip integrals satisfy."""
    from dqc_b200.hamilton.intor import molintor as intor
    zs, pos = util.H2O
    w, p = _wrapper_with_grad(zs, [list(map(float, q)) for q in pos], "def2-svp")
    ip = intor.int1e("ipovlp", w)
    assert ip.shape == (3, w.nao(), w.nao())
    assert float((ip + ip.transpose(-2, -1)).abs().max()) < 1e-12
    v = intor.nuclattr(w)
    G = torch.ones_like(v)
    (grad,) = torch.autograd.grad((v * G).sum(), p)
    assert float(grad.sum(0).abs().max()) < 1e-9


def test_eval_gto_backward(cuda):
    """AO values differentiated with respect to the grid points and the atomic positions, against central differences
    of the oracle's eval_gto."""
    from dqc_b200.hamilton.intor import gtoeval
    from oracle import cint
    zs, pos = util.H2O
    pos = [list(map(float, q)) for q in pos]
    w, p = _wrapper_with_grad(zs, pos, "def2-svp")
    pts = util.random_points(50, seed=2, span=2.0)
    r = torch.tensor(pts, dtype=dtype, device=cuda, requires_grad=True)
    ao = gtoeval.eval_gto(w, r, to_transpose=True)
    g = torch.Generator().manual_seed(1)
    G = torch.randn(ao.shape, dtype=dtype, generator=g).to(cuda)
    gr, gp = torch.autograd.grad((ao * G).sum(), (r, p))
    Gn = G.cpu().numpy()

    def f(pos_, pts_):
        ww, _ = util.make_wrapper(zs, pos_, "def2-svp")
        return (cint.eval_gto(*ww.atm_bas_env, pts_, 0) * Gn).sum()
    h = 1e-5
    for ia in range(3):
        for d in range(3):
            pp = [list(q) for q in pos]
            pm = [list(q) for q in pos]
            pp[ia][d] += h
            pm[ia][d] -= h
            fd = (f(pp, pts) - f(pm, pts)) / (2 * h)
            assert abs(float(gp[ia, d]) - fd) < 1e-6 * max(1.0, abs(fd))
    for k in (0, 17, 49):
        for d in range(3):
            qp, qm = pts.copy(), pts.copy()
            qp[k, d] += h
            qm[k, d] -= h
            fd = (f(pos, qp) - f(pos, qm)) / (2 * h)
            assert abs(float(gr[k, d]) - fd) < 1e-6 * max(1.0, abs(fd))


def _rhf(zs, pos, basis, cuda):
    from dqc_b200 import Mol, HF
    mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=dtype)), basis=basis, device=cuda, orthogonalize_basis=False)
    qc = HF(mol, restricted=True).run(fwd_options={"maxiter": 200, "f_tol": 1e-11})
    assert qc.converged
    return mol, qc


@pytest.mark.parametrize("zs,pos,basis", [([1, 1], [[-0.7, 0.0, 0.0], [0.7, 0.0, 0.0]], "3-21g"),
                                          ([8, 1, 1], None, "3-21g")])
def test_rhf_force_matches_finite_difference_of_the_energy(cuda, zs, pos, basis):
    """Analytic RHF gradient: autograd of  Tr[D h(R)] + 1/2 sum (D_ij D_kl - 1/2 D_ik D_jl)(ij|kl)(R) - Tr[W S(R)] + E_nn(R)
    at the converged D, W (energy-weighted density) through the position backward of S, T, V and (ij|kl) -- against
    central differences of the converged SCF energy of the CUDA path."""
    from dqc_b200.hamilton.intor import molintor as intor
    if pos is None:
        pos = [list(map(float, q)) for q in util.H2O[1]]
    mol, qc = _rhf(zs, pos, basis, cuda)
    w, p = _wrapper_with_grad(zs, pos, basis)
    S, T, V, eri = intor.overlap(w), intor.kinetic(w), intor.nuclattr(w), intor.elrep(w)
    # converged orbitals in the AO basis from the Fock matrix of the converged density
    with torch.no_grad():
        h = (T + V).detach()
        dm_ao = qc.aodm().to(S.device)
        if dm_ao.shape[-1] != S.shape[-1]:
            pytest.skip("density not in the AO basis")
        J = torch.einsum("ijkl,kl->ij", eri.detach(), dm_ao)
        K = torch.einsum("ikjl,kl->ij", eri.detach(), dm_ao)
        F = h + J - 0.5 * K
        sv, su = torch.linalg.eigh(S.detach())
        X = su / sv.sqrt()
        e, c = torch.linalg.eigh(X.t() @ F @ X)
        C = X @ c
        nocc = int(sum(zs)) // 2
        D = 2 * C[:, :nocc] @ C[:, :nocc].t()
        W = 2 * (C[:, :nocc] * e[:nocc]) @ C[:, :nocc].t()
        assert float((D - dm_ao).abs().max()) < 1e-6
    z = torch.tensor(zs, dtype=dtype)
    rij = (p.unsqueeze(0) - p.unsqueeze(1)).norm(dim=-1) + torch.eye(len(zs), dtype=dtype)
    enn = 0.5 * ((z.unsqueeze(0) * z.unsqueeze(1)) * (1 - torch.eye(len(zs), dtype=dtype)) / rij).sum()
    D_, W_ = D.to(S.device), W.to(S.device)
    energy = ((T + V) * D_).sum() + 0.5 * torch.einsum("ijkl,ij,kl->", eri, D_, D_) \
        - 0.25 * torch.einsum("ijkl,ik,jl->", eri, D_, D_) - (S * W_).sum() + enn.to(S.device)
    (force,) = torch.autograd.grad(energy, p)
    hstep = 1e-3
    checks = [(0, 0), (1, 0)] if len(zs) == 2 else [(0, 1), (1, 0), (2, 2)]
    for ia, d in checks:
        pp = [list(q) for q in pos]
        pm = [list(q) for q in pos]
        pp[ia][d] += hstep
        pm[ia][d] -= hstep
        ep = float(_rhf(zs, pp, basis, cuda)[1].energy())
        em = float(_rhf(zs, pm, basis, cuda)[1].energy())
        fd = (ep - em) / (2 * hstep)
        assert abs(float(force[ia, d]) - fd) < 2e-5, (ia, d, float(force[ia, d]), fd)
    assert float(force.sum(0).abs().max()) < 1e-7        # no net force
