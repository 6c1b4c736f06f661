"""Bring-up check of the tcgen05 int8 density kernel against the fp64 DMMA path (GPU)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import _lib
from dqc_b200.utils import systems
from dqc_b200.grid.factory import get_predefined_grid
from tests import util

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "benzene"
zs, pos = getattr(systems, name)()
w, _ = util.make_wrapper(zs, pos.tolist(), "def2-svp")
nb, nao = len(w), w.nao()
grid = get_predefined_grid("sg3", zs, torch.tensor(pos, device=dev), device=dev)
xyz, wts = grid.get_rgrid(), grid.get_dvolume()
db = w.device_basis(dev)
dm = util.seeded_dm(nao, max(1, int(sum(zs)) // 2), seed=0).to(dev)
ref = _lib.GridBlocks(db, 0, nb, xyz, wts, 1, sbp=512, eps=1e-12)
r0, g0 = ref.rho(dm, True)
torch.cuda.synchronize()
print("ref rho max %.3e nel %.8f" % (float(r0.abs().max()), float((r0 * ref.w).sum())))
for S in (6, 5):
    gb = _lib.GridBlocks(db, 0, nb, xyz, wts, 1, sbp=512, eps=1e-12, rho_i8_slices=S)
    r1, g1 = gb.rho(dm, True)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        gb.rho(dm, True)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / 3
    t0 = time.time()
    for _ in range(3):
        ref.rho(dm, True)
    torch.cuda.synchronize()
    dr = (time.time() - t0) / 3
    rel = ((r1 - r0).abs() / (r0.abs() + 1e-10)).max()
    print("S %d: rho max abs diff %.3e  max rel %.3e  grad diff %.3e  nel diff %.3e  (%.2f ms vs ref %.2f ms)" % (
        S, float((r1 - r0).abs().max()), float(rel), float((g1 - g0).abs().max()),
        float(((r1 - r0) * ref.w).sum().abs()), dt * 1e3, dr * 1e3))
    del gb
