"""Throughput of the direct 4-centre J/K build (GPU): contracted shell quartets per second of the register-resident
engine (csrc/jk_reg.cuh) and of the shared-memory engine (csrc/jk.cuh, B200QC_JK_NOREG=1) on the same plan, with the
agreement between the two.  usage: python tools/bench_jk.py [system:basis ...]   (default taxol_like:3-21g)"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_b200 import _lib
from dqc_b200.utils import systems
from tests import util

dev = torch.device("cuda:0")


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


for spec in (sys.argv[1:] or ["taxol_like:3-21g"]):
    name, basis = spec.split(":")[:2]
    with_shared = not spec.endswith(":noshared")
    zs, pos = systems.carbon_cluster(int(name[14:])) if name.startswith("carbon_cluster") else getattr(systems, name)()
    w, _ = util.make_wrapper(zs, pos.tolist(), basis)
    nb, nao = len(w), w.nao()
    db = w.device_basis(dev)
    dm = util.seeded_dm(nao, nao // 4, seed=1).unsqueeze(0).to(dev)
    res = {"system": name, "basis": basis, "nao": nao, "nshell": nb}
    out = {}
    for label, env in (("register", None), ("shared", "1")):
        if label == "shared" and not with_shared:
            continue
        if env is None:
            os.environ.pop("B200QC_JK_NOREG", None)
        else:
            os.environ["B200QC_JK_NOREG"] = env
        t0 = time.time()
        plan = _lib.JKPlan(db, 0, nb, 1e-13)
        torch.cuda.synchronize()
        res["plan_s_" + label] = round(time.time() - t0, 3)
        res["quartets"] = plan.nquartets
        res["quartets_reg_" + label] = plan.nquartets_reg
        reps = 3 if label == "register" else 1
        for mode, (wj, wk) in (("jk", (True, True)), ("j", (True, False)), ("k", (False, True))):
            if label == "shared" and mode != "jk":
                continue
            ms, o = timed(lambda: plan.run(dm, wj, wk), reps)
            res["%s_%s_ms" % (label, mode)] = round(ms, 2)
            res["%s_%s_quartets_per_s" % (label, mode)] = float("%.3g" % (plan.nquartets / (ms * 1e-3)))
            if mode == "jk":
                out[label] = o
        if label == "register":
            for world in (2, 4, 8):     # one rank's share of a sharded build, timed on this GPU
                ms, _ = timed(lambda: plan.run(dm, True, True, rank=0, world=world), 3)
                res["register_jk_ms_rank0_of_%d" % world] = round(ms, 2)
        del plan
    if with_shared:
      res["max_abs_diff_J"] = float((out["register"][0] - out["shared"][0]).abs().max())
      res["max_abs_diff_K"] = float((out["register"][1] - out["shared"][1]).abs().max())
    res["max_abs_J"] = float(out["register"][0].abs().max())
    print(json.dumps(res))
