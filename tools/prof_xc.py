"""Times the XC grid path (K1..K4) at a BASELINE workload size; also the target of ncu captures.
Usage: python tools/prof_xc.py [c60|benzene|h2o|taxol] [lda|pbe] [iters]"""
import os
import sys
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import _lib  # noqa: E402
from dqc_b200.utils import systems  # noqa: E402
from dqc_b200.grid.factory import get_predefined_grid  # noqa: E402
from tests import util  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c60"
    xc = sys.argv[2] if len(sys.argv) > 2 else "pbe"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    dev = torch.device("cuda:0")
    zs, pos = {"c60": systems.c60, "benzene": systems.benzene, "h2o": systems.h2o,
               "taxol": systems.taxol_like}[name]()
    basis = {"c60": "def2-svp", "benzene": "cc-pvdz", "h2o": "def2-svp", "taxol": "def2-svp"}[name]
    terms = [(1.0, "gga_x_pbe"), (1.0, "gga_c_pbe")] if xc == "pbe" else [(1.0, "lda_x"), (1.0, "lda_c_pw")]
    gga = xc == "pbe"
    w, _ = util.make_wrapper(zs, pos.tolist(), basis)
    nao = w.nao()
    t0 = time.time()
    grid = get_predefined_grid("sg3", zs, torch.tensor(pos, device=dev), device=dev)
    torch.cuda.synchronize()
    xyz, wts = grid.get_rgrid(), grid.get_dvolume()
    ngrid = xyz.shape[0]
    print("system %s nao %d ngrid %d grid build %.2fs" % (name, nao, ngrid, time.time() - t0))
    db = w.device_basis(dev)

    def ev():
        a = torch.cuda.Event(enable_timing=True)
        a.record()
        return a

    e0 = ev()
    ao = _lib.eval_gto(db, 0, len(w), xyz, 1 if gga else 0)
    e1 = ev()
    torch.cuda.synchronize()
    ncomp, ngl, ld = ao.shape
    t_ao = e0.elapsed_time(e1)  # includes the zero fill
    print("K1 ao eval (+memset): %.2f ms, %.1f GB written -> %.0f GB/s" %
          (t_ao, ao.numel() * 8 / 1e9, 2 * ao.numel() * 8 / 1e6 / t_ao))
    dm = torch.zeros(ld, ld, dtype=torch.float64, device=dev)
    dm[:nao, :nao] = util.seeded_dm(nao, max(1, nao // 5), seed=0).to(dev)
    wpad = torch.zeros(ngl, dtype=torch.float64, device=dev)
    wpad[:ngrid] = wts
    for it in range(iters):
        e0 = ev()
        rho, grad = _lib.rho(ao, dm, gga)
        e1 = ev()
        e, vr, vg = _lib.xc_unpol(terms, rho, grad)
        e2 = ev()
        mat = _lib.vxc_mat(ao, wpad, vr, vg)
        e3 = ev()
        torch.cuda.synchronize()
        t2, t3, t4 = e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)
        fl = 2.0 * ngl * ld * ld
        print("iter %d: K2 rho %.2f ms (%.1f TF/s, %.0f GB/s) | K3 xc %.3f ms | K4 vxc %.2f ms (%.1f TF/s) | "
              "nelec %.6f exc %.8f" % (it, t2, fl / t2 / 1e9, ncomp * ngl * ld * 8 / t2 / 1e6, t3, t4,
                                       fl / t4 / 1e9, float((rho * wpad).sum()), float((e * wpad).sum())))


if __name__ == "__main__":
    main()
