"""DF-K on the tcgen05 int8 engine against a plain fp64 (torch / cuBLAS) evaluation of the same formula at full
size (GPU): max |dK'| and dE_K for 6 and 5 slices.  Usage: python tools/check_dfk_fp64.py [c60|benzene|taxol_like]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import Mol, config  # noqa: E402
from dqc_b200.utils import systems  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c60"
dev = torch.device("cuda:0")
zs, pos = getattr(systems, name)()
mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=torch.float64)), basis="def2-svp", device=dev,
          orthogonalize_basis=False).densityfit(auxbasis="etb-jfit")
h = mol.get_hamiltonian().build()
nao, naux = h.nao, h.df._naux
nocc = int(sum(zs)) // 2
g = torch.Generator().manual_seed(0)
q, _ = torch.linalg.qr(torch.randn(nao, nao, dtype=torch.float64, generator=g))
orb = q[:, :nocc].to(dev)
dm = h.ao_orb2dm(orb, torch.full((nocc,), 2.0, dtype=torch.float64, device=dev))
cw = orb * (2.0 ** 0.5)
# fp64 reference: B = (ij|Q) L^-T, Y_i = B_i^T-contracted with cw, K = sum Y Y^T
chol = torch.linalg.cholesky(h.df._j2c)
linv = torch.linalg.solve_triangular(chol, torch.eye(naux, dtype=torch.float64, device=dev), upper=False)
j3c = h.df._j3c_packed
ii, jj = torch.tril_indices(nao, nao, device=dev)
y = torch.empty(nao, naux * nocc, dtype=torch.float64, device=dev)
tri = torch.zeros(nao, nao, dtype=torch.long, device=dev)
tri[ii, jj] = torch.arange(ii.shape[0], device=dev)
tri = torch.maximum(tri, tri.t())
for i0 in range(0, nao, 24):
    i1 = min(nao, i0 + 24)
    rows = tri[i0:i1].reshape(-1)
    b = torch.matmul(j3c[rows, :naux], linv.t()).reshape(i1 - i0, nao, naux)          # (i, j, P)
    y[i0:i1] = torch.einsum("ijp,jo->ipo", b, cw).reshape(i1 - i0, -1)
kref = -0.5 * torch.matmul(y, y.t())
del y
eref = float(0.5 * torch.einsum("ij,ji->", kref, dm))
print("%s nao %d naux %d nocc %d: fp64 E_K %.10f Ha, |K'|max %.3e, cond(j2c) %.2e" % (
    name, nao, naux, nocc, eref, float(kref.abs().max()), float(torch.linalg.cond(h.df._j2c))))
for S in (6, 5):
    config.DFK_I8_SLICES = S
    h.df._k_planes = None
    k = h.get_exchange(dm).fullmatrix()
    print("S = %d: max|dK'| %.3e  dE_K %.3e Ha" % (S, float((k - kref).abs().max()),
                                                 float(0.5 * torch.einsum("ij,ji->", k, dm)) - eref))
