"""Times the superblock XC kernels (K2 density, K4 Vxc) of the production path at a BASELINE workload size under the
library's timing-experiment switches; also the target of ncu captures (one process, no DF set-up).
Usage: python tools/prof_sb.py [c60|taxol|benzene] [iters] [what: rho|vxc|both] [modes: comma list of variant:mode]
  variant (b200qc_i8_debug_variant): 0 normal, 1 no AO loads in the K2 epilogue, 2 no MMAs, 3 no epilogue work
  mode (b200qc_i8_mode): bits 12..16 = number of B-cache slots of the point-stationary K2"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import _lib, config  # noqa: E402
from dqc_b200.utils import systems  # noqa: E402
from dqc_b200.grid.factory import get_predefined_grid  # noqa: E402
from tests import util  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c60"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    what = sys.argv[3] if len(sys.argv) > 3 else "both"
    modes = [tuple(int(x) for x in m.split(":")) for m in (sys.argv[4] if len(sys.argv) > 4 else "0:0").split(",")]
    dev = torch.device("cuda:0")
    zs, pos = {"c60": systems.c60, "benzene": systems.benzene, "taxol": systems.taxol_like}[name]()
    basis = "cc-pvdz" if name == "benzene" else "def2-svp"
    w, _ = util.make_wrapper(zs, pos.tolist(), basis)
    nao = w.nao()
    grid = get_predefined_grid("sg3", zs, torch.tensor(pos, device=dev), device=dev)
    xyz, wts = grid.get_rgrid(), grid.get_dvolume()
    db = w.device_basis(dev)
    gb = _lib.GridBlocks(db, 0, len(w), xyz, wts, 1, sbp=config.SB_POINTS, eps=config.AO_SCREEN,
                         i8_slices=config.VXC_I8_SLICES, rho_i8_slices=config.RHO_I8_SLICES)
    nsp = gb.nsp
    print("system %s nao %d ngrid %d nsb %d  nsp mean %.0f max %d  >512: %.2f  >640: %.2f  rho_bn %s fused_vb %s" % (
        name, nao, xyz.shape[0], gb.nsb, nsp.mean(), nsp.max(), (nsp > 512).mean(), (nsp > 640).mean(),
        getattr(gb, "rho_bn", None), gb.colmax is not None))
    dm = util.seeded_dm(nao, max(1, int(sum(zs)) // 2), seed=0).to(dev)
    lib = _lib.load()
    terms = [(1.0, "gga_x_pbe"), (1.0, "gga_c_pbe")]
    for variant, mode in modes:
        lib.b200qc_i8_debug_variant(variant)
        lib.b200qc_i8_mode(mode)
        for it in range(iters):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            if what in ("rho", "both"):
                rho, grad = gb.rho(dm, True)
            e[1].record()
            if what in ("vxc", "both"):
                if what == "vxc":
                    rho, grad = torch.rand(gb.ngl, dtype=torch.float64, device=dev), torch.randn(3, gb.ngl, dtype=torch.float64, device=dev)
                _, vr, vg = _lib.xc_unpol(terms, rho.abs() + 1e-10, grad)
                mat = gb.vxc_mat(vr, vg)
            e[2].record()
            torch.cuda.synchronize()
            print("variant %d mode %d iter %d: rho (gather+K2) %.3f ms | xc+vxc %.3f ms | nelec %.8f" % (
                variant, mode, it, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]),
                float((rho * gb.w).sum()) if what != "vxc" else 0.0))
    lib.b200qc_i8_debug_variant(0)
    lib.b200qc_i8_mode(0)


if __name__ == "__main__":
    main()
