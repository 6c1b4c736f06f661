"""Consistency checks of the direct J/K build on mid-size systems (GPU): run-to-run determinism,
rank-partition additivity, and agreement with the stored-ERI GEMVs."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import _lib
from dqc_b200.utils import systems
from tests import util

dev = torch.device("cuda:0")
for name, basis in (("benzene", "cc-pvdz"), ("carbon_cluster", "def2-svp"), ("taxol_like", "3-21g")):
    zs, pos = systems.carbon_cluster(14) if name == "carbon_cluster" else getattr(systems, name)()
    w, _ = util.make_wrapper(zs, pos.tolist(), basis)
    nb, nao = len(w), w.nao()
    db = w.device_basis(dev)
    dm = torch.stack([util.seeded_dm(nao, nao // 4, seed=1), util.seeded_dm(nao, nao // 5, seed=2)]).to(dev)
    t0 = time.time()
    plan = _lib.JKPlan(db, 0, nb, 1e-13)
    torch.cuda.synchronize(); t1 = time.time()
    j1, k1 = plan.run(dm)
    torch.cuda.synchronize(); t2 = time.time()
    j2, k2 = plan.run(dm)
    ja, ka = plan.run(dm, rank=0, world=3); jb, kb = plan.run(dm, rank=1, world=3); jc, kc = plan.run(dm, rank=2, world=3)
    torch.cuda.synchronize()
    print("%s/%s nao %d quartets %d plan %.2fs run %.2fs" % (name, basis, nao, plan.nquartets, t1 - t0, t2 - t1))
    print("  nan:", bool(torch.isnan(j1).any() or torch.isnan(k1).any()),
          " rerun diff J %.2e K %.2e" % (float((j1 - j2).abs().max()), float((k1 - k2).abs().max())),
          " 3-rank sum diff J %.2e K %.2e" % (float((ja + jb + jc - j1).abs().max()), float((ka + kb + kc - k1).abs().max())),
          " |J|max %.2e |K|max %.2e" % (float(j1.abs().max()), float(k1.abs().max())))
    if nao % 2 == 0 and 16 * nao ** 4 < 60e9:
        st = _lib.StoredERI(db, 0, nb)
        js, ks = st.run(dm)
        print("  vs stored ERI: J %.2e K %.2e" % (float((js - j1).abs().max()), float((ks - k1).abs().max())))
        del st
