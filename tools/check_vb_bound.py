"""How loose is the column-maximum bound of the fused vb slicer on a real potential?  Prints the histogram of
log2(bound exponent / exact exponent) over all (superblock, column) pairs and the flagged fraction per threshold.
Usage: python tools/check_vb_bound.py [c60|taxol|benzene]"""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import _lib, config  # noqa: E402
from dqc_b200.utils import systems  # noqa: E402
from dqc_b200.grid.factory import get_predefined_grid  # noqa: E402
from tests import util  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c60"
    dev = torch.device("cuda:0")
    zs, pos = {"c60": systems.c60, "benzene": systems.benzene, "taxol": systems.taxol_like}[name]()
    basis = "cc-pvdz" if name == "benzene" else "def2-svp"
    w, _ = util.make_wrapper(zs, pos.tolist(), basis)
    grid = get_predefined_grid("sg3", zs, torch.tensor(pos, device=dev), device=dev)
    xyz, wts = grid.get_rgrid(), grid.get_dvolume()
    db = w.device_basis(dev)
    gb = _lib.GridBlocks(db, 0, len(w), xyz, wts, 1, sbp=config.SB_POINTS, eps=config.AO_SCREEN,
                         i8_slices=config.VXC_I8_SLICES, rho_i8_slices=config.RHO_I8_SLICES)
    dm = util.seeded_dm(w.nao(), max(1, int(sum(zs)) // 2), seed=0).to(dev)
    rho, grad = gb.rho(dm, True)
    _, vr, vg = _lib.xc_unpol([(1.0, "gga_x_pbe"), (1.0, "gga_c_pbe")], rho, grad)
    lib = _lib.load()
    lib.b200qc_i8_mode(1 << 23)            # bound only, no repair
    m_f = gb.vxc_mat(vr, vg)
    torch.cuda.synchronize()
    s_f = gb.bscale.clone()
    flags7 = int(gb.fixflag.sum())
    lib.b200qc_i8_mode(0)
    m_d = gb.vxc_mat(vr, vg)
    flags4 = int(gb.fixflag.sum())
    colmax, gb.colmax = gb.colmax, None
    m_u = gb.vxc_mat(vr, vg)
    torch.cuda.synchronize()
    s_u = gb.bscale.clone()
    gb.colmax = colmax
    nblk = int(((gb.nsp + 63) // 64).sum())
    print("blocks %d  flagged without repair: %d (%.1f%%)  at 2^4: %d (%.1f%%)" % (nblk, flags7, 100.0 * flags7 / nblk, flags4, 100.0 * flags4 / nblk))
    ok = (s_u > 0) & (s_f > 0)
    bits = torch.log2(s_f[ok] / s_u[ok]).round().to(torch.int64).cpu().numpy()
    h = np.bincount(np.clip(bits, 0, 20))
    print("log2(bound / exact) histogram over columns (0..20+):", h.tolist())
    print("cumulative fraction <= k bits:", np.round(np.cumsum(h) / h.sum(), 3).tolist())
    # flagged 64-column blocks under other (looseness, absolute) thresholds, from the two exponent sets
    S = config.VXC_I8_SLICES
    lb = torch.log2(s_f / s_u.clamp_min(1e-300)).cpu().numpy()
    eb = torch.log2(s_f.clamp_min(1e-300)).cpu().numpy()
    valid = (s_u > 0).cpu().numpy()
    blk = np.repeat(np.arange(nblk), 64)[:lb.shape[0]]
    for L in (4, 5, 6, 8):
        row = []
        for A in (-49, -46, -44, -42, -40):
            bad = valid & (lb > L) & (eb > A + 7 * S)
            row.append("%.1f%%" % (100.0 * np.unique(blk[bad]).size / nblk))
        print("loose > 2^%d, abs step > 2^A for A = -49, -46, -44, -42, -40: flagged blocks" % L, row)
    sc = float(m_u.abs().max())
    print("max |fused(no repair) - unfused| = %.2e, |fused(2^4) - unfused| = %.2e (scale %.2e)" % (
        float((m_f - m_u).abs().max()), float((m_d - m_u).abs().max()), sc))


if __name__ == "__main__":
    main()
