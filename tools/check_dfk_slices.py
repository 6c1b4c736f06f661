"""DF-K with 5 against 6 int8 slices at full size (GPU): max |dK| and the exchange-energy difference at a seeded
closed-shell density.  Usage: python tools/check_dfk_slices.py [c60|taxol_like|benzene]"""
import os
import sys
import time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import Mol, config  # noqa: E402
from dqc_b200.utils import systems  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c60"
dev = torch.device("cuda:0")
zs, pos = getattr(systems, name)()
mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=torch.float64)), basis="def2-svp", device=dev).densityfit(
    auxbasis="etb-jfit")
h = mol.get_hamiltonian().build()
nocc = int(sum(zs)) // 2
g = torch.Generator().manual_seed(0)
q, _ = torch.linalg.qr(torch.randn(h.nao, h.nao, dtype=torch.float64, generator=g))
dm = h.ao_orb2dm(q[:, :nocc].to(dev), torch.full((nocc,), 2.0, dtype=torch.float64, device=dev))
res = {}
for S in (6, 5):
    config.DFK_I8_SLICES = S
    h.df._k_planes = None
    k = h.get_exchange(dm).fullmatrix()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        k = h.get_exchange(dm).fullmatrix()
    torch.cuda.synchronize()
    res[S] = (k, float(0.5 * torch.einsum("ij,ji->", k, dm)), (time.time() - t0) / 3 * 1e3)
k6, e6, t6 = res[6]
k5, e5, t5 = res[5]
print("%s nao %d: |K'|max %.3e  E_K %.10f Ha;  S=5 vs S=6: max|dK'| %.3e  dE_K %.3e Ha;  %.1f ms vs %.1f ms" % (
    name, h.nao, float(k6.abs().max()), e6, float((k5 - k6).abs().max()), e5 - e6, t5, t6))
