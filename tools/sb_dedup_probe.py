"""How many superblocks share their kept-AO index list (candidates for one gathered / sliced D_sb per distinct list)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import _lib, config
from dqc_b200.utils import systems
from dqc_b200.grid.factory import get_predefined_grid
from tests import util
dev = torch.device("cuda:0")
for name in ("c60", "taxol_like"):
    zs, pos = getattr(systems, name)()
    w, _ = util.make_wrapper(zs, pos.tolist(), "def2-svp")
    _lib.load()
    grid = get_predefined_grid("sg3", zs, torch.tensor(pos, device=dev), device=dev)
    xyz = grid.get_rgrid().contiguous()
    db = w.device_basis(dev)
    flags = _lib.ao_screen(db, 0, len(w), xyz, config.SB_POINTS, config.AO_SCREEN, 1).cpu().numpy()
    uniq, inv, cnt = np.unique(flags, axis=0, return_inverse=True, return_counts=True)
    loc = db.ao_loc.astype(np.int64)
    sizes = loc[1:] - loc[:-1]
    nsp = np.maximum(64, ((flags.astype(bool) * sizes[None, :]).sum(1) + 63) // 64 * 64)
    nsp_u = np.maximum(64, ((uniq.astype(bool) * sizes[None, :]).sum(1) + 63) // 64 * 64)
    # consecutive duplicates (same list as the previous superblock)
    consec = int((np.abs(np.diff(flags.astype(np.int8), axis=0)).sum(1) == 0).sum())
    print(json.dumps({"system": name, "nsb": int(flags.shape[0]), "distinct_lists": int(uniq.shape[0]),
                      "same_as_previous": consec, "sum_nsp2_all": float((nsp.astype(float) ** 2).sum()),
                      "sum_nsp2_distinct": float((nsp_u.astype(float) ** 2).sum())}))
