"""One-off kernels of the XC path at a BASELINE size (GPU): K1 AO evaluation on the superblock layout, Becke weights.
usage: python tools/time_k1.py [c60|taxol|benzene]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import _lib, config
from dqc_b200.utils import systems
from dqc_b200.grid.factory import get_predefined_grid
from tests import util

name = sys.argv[1] if len(sys.argv) > 1 else "c60"
dev = torch.device("cuda:0")
zs, pos = {"c60": systems.c60, "benzene": systems.benzene, "taxol": systems.taxol_like}[name]()
basis = "cc-pvdz" if name == "benzene" else "def2-svp"
w, _ = util.make_wrapper(zs, pos.tolist(), basis)
_lib.load()
grid = get_predefined_grid("sg3", zs, torch.tensor(pos, device=dev), device=dev)
xyz, wts = grid.get_rgrid(), grid.get_dvolume()
db = w.device_basis(dev)
res = {"system": name, "nao": w.nao(), "ngrid": int(xyz.shape[0])}
for it in range(2):
    _lib.profile_enable(True)
    gb = _lib.GridBlocks(db, 0, len(w), xyz, wts, 1, sbp=config.SB_POINTS, eps=config.AO_SCREEN,
                         i8_slices=config.VXC_I8_SLICES, rho_i8_slices=config.RHO_I8_SLICES)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    res["ao_bytes"] = float(gb.ao_bytes)
    res["ao_eval_ms_run%d" % it] = prof["ao_eval_kernel"][1]
    res["ao_eval_GBps_run%d" % it] = gb.ao_bytes / (prof["ao_eval_kernel"][1] * 1e-3) / 1e9
    del gb
print(json.dumps(res))
