"""Wide superblocks (> 1024 kept AOs): tcgen05 int8 density / Vxc kernels against the fp64 DMMA kernels on a dense
carbon cluster (GPU).  Usage: python tools/check_wide_sb.py [natom]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import _lib  # noqa: E402
from dqc_b200.grid.factory import get_grid  # noqa: E402
from tests import util  # noqa: E402
from tools.sweep import cluster  # noqa: E402

natom = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda:0")
zs, pos = cluster(natom)
w, _ = util.make_wrapper(zs, pos.tolist(), "def2-svp")
nb, nao = len(w), w.nao()
grid = get_grid(zs, torch.tensor(pos, dtype=torch.float64), nr=20, nang=26, truncate=None, device=dev)
xyz, wts = grid.get_rgrid(), grid.get_dvolume()
db = w.device_basis(dev)
dm = util.seeded_dm(nao, 3 * natom, seed=0).to(dev)
ref = _lib.GridBlocks(db, 0, nb, xyz, wts, 1, sbp=512, eps=1e-12)
r0, g0 = ref.rho(dm, True)
v = torch.randn(ref.ngl, dtype=torch.float64, device=dev)
vg = torch.randn(3, ref.ngl, dtype=torch.float64, device=dev) * 0.1
m0 = ref.vxc_mat(v, vg)
print("nao %d max nsp %d ngrid %d" % (nao, ref.max_nsp, xyz.shape[0]))
for S in (6, 5):
    gb = _lib.GridBlocks(db, 0, nb, xyz, wts, 1, sbp=512, eps=1e-12, rho_i8_slices=S, i8_slices=S)
    r1, g1 = gb.rho(dm, True)
    m1 = gb.vxc_mat(v, vg)
    rel = float(((r1 - r0).abs() / (r0.abs() + 1e-10)).max())
    dv = float((m1 - m0).abs().max() / m0.abs().max())
    print("S %d: rho max rel diff %.3e  grad max abs diff %.3e  vxc rel diff %.3e" % (S, rel, float((g1 - g0).abs().max()), dv))
    assert rel < (1e-9 if S == 6 else 1e-7) and dv < (1e-9 if S == 6 else 1e-7)
    del gb
print("ok")
