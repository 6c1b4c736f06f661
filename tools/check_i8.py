"""Bring-up check of the tcgen05 int8 Vxc GEMM against the fp64 DMMA path (GPU)."""
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
ref = _lib.GridBlocks(db, 0, nb, xyz, wts, 1, sbp=512, eps=1e-12)
g = torch.Generator().manual_seed(0)
vr = torch.randn(ref.ngl, dtype=torch.float64, generator=g).to(dev)
vg = (torch.randn(3, ref.ngl, dtype=torch.float64, generator=g) * 0.3).to(dev)
m_ref = ref.vxc_mat(vr, vg)
torch.cuda.synchronize()
print("ref |M|max %.3e" % float(m_ref.abs().max()))
for S in (6,):
    for variant in (0,):
        _lib.load().b200qc_i8_debug_variant(variant)
        try:
            gb = _lib.GridBlocks(db, 0, nb, xyz, wts, 1, sbp=512, eps=1e-12, i8_slices=S, i8_variant=variant)
            m = gb.vxc_mat(vr, vg)
            torch.cuda.synchronize()
            t0 = time.time()
            for _ in range(3):
                m = gb.vxc_mat(vr, vg)
            torch.cuda.synchronize()
            dt = (time.time() - t0) / 3
            print("S %d variant %d: max abs diff %.3e  (%.2f ms per call; ref path:" % (S, variant, float((m - m_ref).abs().max()), dt * 1e3), end=" ")
            t0 = time.time()
            for _ in range(3):
                ref.vxc_mat(vr, vg)
            torch.cuda.synchronize()
            print("%.2f ms)" % ((time.time() - t0) / 3 * 1e3))
            del gb
        except Exception as e:
            print("S %d variant %d failed: %s" % (S, variant, e))
            raise
