"""Accuracy of the sliced-integer XC kernels at full size: E_xc, N_el and the Vxc matrix of C60 (or another system)
with S = 5 / 6 int8 slices against the fp64 DMMA kernels at the same seeded density (GPU).
Usage: python tools/check_exc_slices.py [c60|taxol_like|benzene]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import Mol, get_xc, config  # noqa: E402
from dqc_b200.utils import systems  # noqa: E402
from tests import util  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c60"
dev = torch.device("cuda:0")
zs, pos = getattr(systems, name)()
out = {}
for rs, vs in ((0, 0), (6, 6), (5, 5)):
    config.RHO_I8_SLICES, config.VXC_I8_SLICES = rs, vs
    mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=torch.float64)), basis="def2-svp", grid="sg3", device=dev,
              orthogonalize_basis=False)
    h = mol.get_hamiltonian()
    mol.setup_grid()
    h.setup_grid(mol.get_grid(), get_xc("gga_x_pbe + gga_c_pbe"))
    dm = util.seeded_dm(h.nao, int(sum(zs)) // 2, seed=0).to(dev)
    exc = float(h.get_e_xc(dm))
    rho = h._dm2densinfo(dm)
    nel = float((rho.value * h.dvolume).sum())
    vxc = h.get_vxc(dm).fullmatrix()
    out[(rs, vs)] = (exc, nel, vxc, rho.value, rho.grad)
    del h, mol
    torch.cuda.empty_cache()
e0, n0, v0, r0, g0 = out[(0, 0)]
print("%s: fp64 E_xc %.10f Ha, N_el %.10f, |Vxc|max %.3e" % (name, e0, n0, float(v0.abs().max())))
for k in ((6, 6), (5, 5)):
    e, n, v, r, g = out[k]
    print("S = %d: dE_xc %.3e Ha  dN_el %.3e  max|dVxc| %.3e  max|drho|/rho %.3e  max|dgrad| %.3e" % (
        k[0], e - e0, n - n0, float((v - v0).abs().max()), float(((r - r0).abs() / (r0.abs() + 1e-10)).max()),
        float((g - g0).abs().max())))
