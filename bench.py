#!/usr/bin/env python
"""Fock-build benchmark (BASELINE.json metric): wall-ms per SCF-iteration Fock build.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

A "step" is ONE Fock build at a fixed density: F_2e = J[D] (+ a K'[D]) + Vxc[D] through
``HamiltonCGTO.get_fock_2e`` -- every kernel, the basis changes and (N > 1) the single packed
all-reduce.  Default workload = the configuration the metric is quoted on: C60 / def2-SVP / PBE with
density fitting on the sg3 grid (1 060 440 points, nao 840).  Inputs (AO values 28.5 GB, packed
(ij|P) 9.7 GB) are far larger than the 126 MB L2, so no flush is needed between iterations.

Prints ONE JSON line (rank 0).  ``value`` = device-timed ms per step, max over ranks; ``e2e`` = the
same build called with HOST buffers (pinned density in, Fock matrix out, copies inside the timed
region); ``roofline`` = the dominant kernel timed live with CUDA events inside the timed region;
``cpu_baseline`` = the reference-equivalent PyTorch-CPU ops of oracle/fock_ref.py on a bounded sample.
``--impl reference`` times only that CPU path (rank 0), on the same workload and metric.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

B3LYP_SL = "0.08*lda_x + 0.72*gga_x_b88 + 0.19*lda_c_vwn_rpa + 0.81*gga_c_lyp"
WORKLOADS = {
    # name: (system, basis, xc, df aux basis or None, exx fraction, grid)
    "c60-pbe-df": ("c60", "def2-svp", "gga_x_pbe + gga_c_pbe", "etb-jfit", 0.0, "sg3"),
    "benzene-lda-4c": ("benzene", "cc-pvdz", "lda_x + lda_c_pw", None, 0.0, "sg3"),
    "h2o-pbe-df": ("h2o", "def2-svp", "gga_x_pbe + gga_c_pbe", "etb-jfit", 0.0, "sg3"),
    "h2o-hf": ("h2o", "sto-3g", None, None, 1.0, "sg3"),
    "taxol-like-pbe0-4c": ("taxol_like", "def2-svp", "0.75*gga_x_pbe + gga_c_pbe", None, 0.25, "sg3"),
    "taxol-like-pbe-df": ("taxol_like", "def2-svp", "gga_x_pbe + gga_c_pbe", "etb-jfit", 0.0, "sg3"),
    # hybrids with density-fitted J AND K (the DF-K extension: two batched tcgen05 int8 GEMMs per build)
    "c60-pbe0-df": ("c60", "def2-svp", "0.75*gga_x_pbe + gga_c_pbe", "etb-jfit", 0.25, "sg3"),
    "benzene-pbe0-df": ("benzene", "cc-pvdz", "0.75*gga_x_pbe + gga_c_pbe", "etb-jfit", 0.25, "sg3"),
    "taxol-like-pbe0-df": ("taxol_like", "def2-svp", "0.75*gga_x_pbe + gga_c_pbe", "etb-jfit", 0.25, "sg3"),
    # B3LYP as libxc composes it (0.08 Slater + 0.72 B88 + 0.19 VWN-RPA + 0.81 LYP + 0.20 exact exchange)
    "c60-b3lyp-df": ("c60", "def2-svp", B3LYP_SL, "etb-jfit", 0.20, "sg3"),
    "taxol-like-b3lyp-df": ("taxol_like", "def2-svp", B3LYP_SL, "etb-jfit", 0.20, "sg3"),
    # nbasis / grid sweep (BASELINE configs[4]): simple-cubic carbon clusters, 14 AOs and ~17.7e3 sg3 points per atom --
    # nao 504 / 1008 / 2002 with 0.64e6 / 1.27e6 / 2.5e6 points, through the same sharded Fock build (1, 2, 4, 8 GPUs)
    "cluster36-pbe-df": ("carbon_cluster_36", "def2-svp", "gga_x_pbe + gga_c_pbe", "etb-jfit", 0.0, "sg3"),
    "cluster72-pbe-df": ("carbon_cluster_72", "def2-svp", "gga_x_pbe + gga_c_pbe", "etb-jfit", 0.0, "sg3"),
    "cluster143-pbe-df": ("carbon_cluster_143", "def2-svp", "gga_x_pbe + gga_c_pbe", "etb-jfit", 0.0, "sg3"),
    # meta-GGA (SCAN exchange) through the fp64 superblock engine
    "benzene-scan-4c": ("benzene", "cc-pvdz", "mgga_x_scan", None, 0.0, "sg3"),
    # the same with the direct 4-centre J/K engine (no density fitting): seconds per build -- the weak kernel of
    # DESIGN.md section 7; use --steps 1 --warmup 3
    "taxol-like-b3lyp-4c": ("taxol_like", "def2-svp", B3LYP_SL, None, 0.20, "sg3"),
}
METRIC = "fock_build_wall_ms_per_scf_iter"


def geometry(name):
    from dqc_b200.utils import systems
    if name.startswith("carbon_cluster_"):
        return systems.carbon_cluster(int(name.rsplit("_", 1)[1]))
    return getattr(systems, name)()


def seeded_orb(nao, nocc):
    """C = first nocc columns of qr(randn(nao, nao)) under manual_seed(0), on the host."""
    g = torch.Generator().manual_seed(0)
    q, _ = torch.linalg.qr(torch.randn(nao, nao, dtype=torch.float64, generator=g))
    return q[:, :nocc].contiguous()


def seeded_dm(nao, nocc, device):
    """D = 2 C C^T of the seeded orbitals."""
    c = seeded_orb(nao, nocc)
    return (2 * c @ c.T).to(device)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ---------------------------------------------------------------------------------------------
# Workload description shared by both arms (the driver compares the two `config` objects)
BASIS_PROVENANCE = {
    "sto-3g": "embedded table pinned by the H2O RHF energy of the Crawford project (tests/test_oracle_golden.py)",
    "cc-pvdz": "embedded table; H and O pinned by a literature H2O RHF energy to 2e-6 Ha, C typed and unverified by energy",
    "def2-svp": "embedded table typed from the published one; NOT verified against an independent energy (no def2-SVP "
                "energy is known offline to better than 1e-3 Ha)",
    "3-21g": "embedded table pinned by the reference's RHF/UHF golden energies (rtol 1e-7)",
}


def workload_config(workload, world):
    """Everything that defines the measured problem -- identical for `--impl b200` and `--impl reference`."""
    from dqc_b200.grid.factory import get_predefined_grid
    from dqc_b200.api.loadbasis import loadbasis
    sysname, basis, xc, aux, exx, gridname = WORKLOADS[workload]
    zs, pos = geometry(sysname)
    shells = {z: loadbasis("%d:%s" % (z, basis)) for z in set(zs)}
    nao = sum(sum(2 * b.angmom + 1 for b in shells[z]) for z in zs)
    naux = None
    if aux is not None:
        ashells = {z: loadbasis("%d:%s" % (z, aux)) for z in set(zs)}
        naux = sum(sum(2 * b.angmom + 1 for b in ashells[z]) for z in zs)
    ngrid = 0
    if xc is not None:
        for z in set(zs):
            g1 = get_predefined_grid(gridname, [z], torch.zeros(1, 3, dtype=torch.float64), device=torch.device("cpu"))
            ngrid += g1.get_rgrid().shape[0] * zs.count(z)
    return {"workload": "%s: %s / %s / xc=%s / %s / grid %s" % (
        workload, sysname, basis, xc, ("DF-J aux=" + aux) if aux else "4-centre J/K", gridname),
        "exx_fraction": exx, "nao": nao, "ngrid": ngrid, "naux": naux, "natoms": len(zs),
        "density": "D = 2 C C^T, C = first nocc columns of qr(randn(nao, nao)) under manual_seed(0)",
        "basis_tables": BASIS_PROVENANCE.get(basis, "embedded table"),
        "geometry": "synthetic 113-atom C47H51NO14 skeleton (dqc_b200/utils/systems.py: no Taxol structure file offline)"
                    if sysname == "taxol_like" else "closed-formula geometry (dqc_b200/utils/systems.py)",
        "l2_policy": "inputs (AO values, (ij|P)) larger than L2; no flush",
        "parallelism": "grid rows + aux shells (or J/K work items) sharded over %d GPU(s), one packed all-reduce" % world}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's per-iteration torch-CPU ops (oracle/fock_ref.py) at the FULL problem size
def _host_mem_available():
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:
        return 0


def cpu_reference_step_factory(workload, grid_sample=None, aux_sample=None, j3c_oracle_seconds=25.0):
    """Builds the CPU problem with the ORACLE only (no CUDA kernel on this path) and returns (step_fn, info).

    Full size by default: AO values / gradients of every grid point (oracle/cint_oracle.c) and the dense
    (nao, nao, naux) 3-centre tensor, exactly the tensors the reference keeps (hcgto.py:168-186, dfmol.py:38-55);
    step_fn() is then one real reference-equivalent Fock build and nothing is extrapolated.  The one-off 3-centre
    integrals would take minutes on the oracle, so only the first aux shells (about `j3c_oracle_seconds` of work) are
    integrated and the remaining aux columns repeat them cyclically -- the per-iteration einsums are dense and do not
    depend on the values.  When the host lacks the memory (about 60 GB at C60) the sample falls back to
    `grid_sample` points / `aux_sample` aux functions with linear extrapolation, and says so."""
    from oracle import fock_ref, cint
    from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
    from dqc_b200.grid.factory import get_predefined_grid
    from tests import util
    sysname, basis, xc, aux, exx, gridname = WORKLOADS[workload]
    zs, pos = geometry(sysname)
    torch.set_num_threads(os.cpu_count() or 1)
    w, _ = util.make_wrapper(zs, pos.tolist(), basis)
    nao = w.nao()
    info = {"nao": nao, "cores": torch.get_num_threads()}
    per_atom = {}
    pts = []
    for z, p in zip(zs, pos):
        if z not in per_atom:
            g1 = get_predefined_grid(gridname, [z], torch.zeros(1, 3, dtype=torch.float64), device=torch.device("cpu"))
            per_atom[z] = g1.get_rgrid().numpy()
        pts.append(per_atom[z] + p)
    pts = np.concatenate(pts) if xc is not None else np.zeros((0, 3))
    ngrid_full = int(pts.shape[0])
    naux_full = 0
    auxw = None
    if aux is not None:
        auxw, _ = util.make_wrapper(zs, pos.tolist(), aux)
        naux_full = auxw.nao()
    ncomp = 1 if xc is None else (4 if any(t in xc for t in ("gga", "b88", "lyp", "pbe")) else 1)
    need = 8.0 * ngrid_full * nao * (ncomp + 1) + 8.0 * nao * nao * naux_full + 4e9
    avail = _host_mem_available()
    full = (grid_sample is None and aux_sample is None) and (avail == 0 or avail > need * 1.1)
    if not full:
        grid_sample = grid_sample or 65536
        aux_sample = aux_sample or 96
        info["fallback"] = "host memory %.0f GB < %.0f GB needed for the full-size tensors" % (avail / 1e9, need / 1e9)
    stride = 1 if full else max(1, ngrid_full // grid_sample)
    sample = np.ascontiguousarray(pts[::stride])
    info.update(ngrid_full=ngrid_full, ngrid_sample=int(sample.shape[0]), full=full)
    h = fock_ref.RefHamilton(w, orthozer=True)
    if xc is not None:
        # weights do not influence the timing: a constant stands in for the Becke-weighted volumes
        h.setup_grid(sample, np.full(sample.shape[0], 1e-3), xc)
    if aux is not None:
        bw, aw = LibcintWrapper.concatenate(w, auxw)
        atm, bas, env = aw.atm_bas_env
        a0, a1 = aw.shell_idxs
        b0, b1 = bw.shell_idxs
        loc = aw.full_shell_to_aoloc
        # integrate aux shells until the time budget (full mode) or the sample size (fallback) is reached
        target = naux_full if full else aux_sample
        cols, a_lo, t0 = [], a0, time.perf_counter()
        while a_lo < a1 and sum(c.shape[-1] for c in cols) < target:
            a_hi = a_lo + 1
            while a_hi < a1 and loc[a_hi] - loc[a_lo] < 64:
                a_hi += 1
            cols.append(torch.as_tensor(cint.int3c2e(atm, bas, env, (b0, b1, b0, b1, a_lo, a_hi))))
            a_lo = a_hi
            if time.perf_counter() - t0 > j3c_oracle_seconds:
                break
        blk = torch.cat(cols, dim=-1)
        n_int = int(blk.shape[-1])
        if full and n_int < naux_full:
            j3c = torch.empty(nao, nao, naux_full, dtype=torch.float64)
            for c0 in range(0, naux_full, n_int):
                c1 = min(naux_full, c0 + n_int)
                j3c[:, :, c0:c1] = blk[:, :, :c1 - c0]
        else:
            j3c = blk[:, :, :target] if not full else blk
        del blk, cols
        n_used = int(j3c.shape[-1])
        # the (P|Q) metric of the same number of functions (values irrelevant for the timing of temp @ inv)
        a_hi = a0 + 1
        while a_hi < a1 and loc[a_hi] - loc[a0] < min(n_used, 256):
            a_hi += 1
        j2c_small = torch.as_tensor(cint.int2c2e(atm, bas, env, (a0, a_hi, a0, a_hi)))
        inv = torch.zeros(n_used, n_used, dtype=torch.float64)
        ns = j2c_small.shape[0]
        inv_small = torch.inverse(j2c_small)
        for c0 in range(0, n_used, ns):
            c1 = min(n_used, c0 + ns)
            inv[c0:c1, c0:c1] = inv_small[:c1 - c0, :c1 - c0]
        h.j3c, h.j2c, h.inv_j2c = j3c, None, inv
        info.update(naux_full=naux_full, naux_sample=n_used, naux_integrated=n_int)
        if exx != 0.0:
            info["note"] = "the reference has no density-fitted exchange (hcgto.py:229-230 raises): K is not in the CPU step"
    elif exx != 0.0 or xc is None:
        info["note"] = "dense-ERI J/K of this workload is not part of the CPU step (nao^4 tensor)"
    nocc = max(1, int(sum(zs)) // 2)
    dm = seeded_dm(h.nao, min(nocc, h.nao), torch.device("cpu"))

    def step():
        t = {}
        t0 = time.perf_counter()
        if h.j3c is not None:
            h.get_elrep(dm)
        t["dfj_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        if xc is not None:
            h.get_vxc(dm)
        t["xc_s"] = time.perf_counter() - t0
        return t
    return step, info


def run_cpu_reference(workload, steps, warmup, **kw):
    step, info = cpu_reference_step_factory(workload, **kw)
    for _ in range(warmup):
        step()
    acc = {"dfj_s": 0.0, "xc_s": 0.0}
    for _ in range(steps):
        t = step()
        for k in acc:
            acc[k] += t[k]
    dfj = acc["dfj_s"] / steps
    xcs = acc["xc_s"] / steps
    if info["full"]:
        total = xcs + dfj
        sample = "FULL size, measured: XC ops on all %d grid points" % info["ngrid_full"]
        if info.get("naux_full"):
            sample += ", DF-J ops on the dense (nao, nao, %d) tensor (%d aux functions integrated by the oracle, the " \
                      "other columns repeat them: dense einsums, timing independent of the values)" % (
                          info["naux_full"], info["naux_integrated"])
        sample += "; nothing extrapolated"
    else:
        total = xcs * info["ngrid_full"] / max(1, info["ngrid_sample"])
        if info.get("naux_full"):
            total += dfj * info["naux_full"] / info["naux_sample"]
        sample = "FALLBACK (%s): XC ops on %d of %d grid points" % (info.get("fallback"), info["ngrid_sample"], info["ngrid_full"])
        if info.get("naux_full"):
            sample += ", DF-J ops on %d of %d aux functions" % (info["naux_sample"], info["naux_full"])
        sample += "; per-step time extrapolated linearly to the full sizes"
    info.update(xc_ms=xcs * 1e3, dfj_ms=dfj * 1e3)
    return total * 1e3, info, sample


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c60-pbe-df", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    sysname, basis, xc, aux, exx, gridname = WORKLOADS[args.workload]
    config = workload_config(args.workload, world)      # the same object in both arms
    impl_config = {}                                    # how THIS arm runs it (kernels, storage, set-up time)

    if args.impl == "reference":
        if rank != 0:
            return
        ms, info, sample = run_cpu_reference(args.workload, max(1, args.steps), args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "impl_config": {"xc_ms": info["xc_ms"], "dfj_ms": info["dfj_ms"]},
                "cpu_baseline": {"value": ms, "unit": "ms", "cores": info["cores"], "kind": "port", "sample": sample},
                "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from dqc_b200 import Mol, get_xc, _lib
    from dqc_b200.utils.dist import get_context
    from dqc_b200.utils.config import config as qc_config
    ctx = get_context()
    assert ctx.world == world

    zs, pos = geometry(sysname)
    t_setup = time.perf_counter()
    mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=torch.float64)), basis=basis, grid=gridname, device=dev)
    if aux is not None:
        mol.densityfit(auxbasis=aux)
    h = mol.get_hamiltonian()
    if xc is not None:
        mol.setup_grid()
        h.setup_grid(mol.get_grid(), get_xc(xc))
    h.build()
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    nao = h.nao
    nocc = max(1, int(sum(zs)) // 2)
    # the density the SCF engine would hand over: D = C w C^T through ao_orb2dm (hcgto.py:272-281), w = 2
    orb_host = seeded_orb(nao, min(nocc, nao)).pin_memory()
    occ = torch.full((orb_host.shape[1],), 2.0, dtype=torch.float64, device=dev)
    dm = h.ao_orb2dm(orb_host.to(dev), occ)
    ngrid = int(mol.get_grid().get_rgrid().shape[0]) if xc is not None else 0
    if h.df is not None:
        impl_config["dfj_pair_rows_read"] = "%.3f of the (ij|P) pair rows (the others are below %g in every column: skipped)" % (
            getattr(h.df, "row_kept_fraction", 1.0), qc_config.DFJ_ROW_SKIP)
    if aux is None:
        impl_config["jk_engine"] = type(h._jkplan).__name__ + (
            " (both dense (ij|kl) layouts resident in HBM, J/K = GEMVs)" if type(h._jkplan).__name__ == "StoredERI"
            else " (Schwarz-screened direct build, %d unique shell quartets)" % h._jkplan.nquartets)
    assert (config["nao"], config["ngrid"], config["naux"]) == (
        h._nao_ao, ngrid, h.df._naux if h.df is not None else None), "workload_config disagrees with the built system"
    impl_config.update(nao_orthogonal=nao, setup_s=round(t_setup, 2))

    def step():
        return h.get_fock_2e(dm, exx=exx, with_xc=xc is not None).fullmatrix()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    # ---- timed region: device time, CUDA events on the current stream, max over ranks ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    _lib.profile_enable(True)
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        fock = step()
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t[0])
    ms_per_step = total_ms / args.steps

    # ---- e2e: host buffers in and out, copies inside the timed region ----
    # host side of an SCF iteration: occupied orbitals in (pinned), Fock matrix out (pinned); D = C w C^T is
    # formed on the device by ao_orb2dm like the engine does (scp2dm), then the same get_fock_2e call
    fock_host = torch.empty(nao, nao, dtype=torch.float64).pin_memory()
    orb_dev = torch.empty_like(orb_host, device=dev)
    barrier()
    e0.record()
    for _ in range(args.steps):
        orb_dev.copy_(orb_host, non_blocking=True)
        f = h.get_fock_2e(h.ao_orb2dm(orb_dev, occ), exx=exx, with_xc=xc is not None).fullmatrix()
        fock_host.copy_(f, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    e2e_ms /= args.steps
    # same build twice: equal up to the summation-order noise of the fp64 atomics (J/K digestion, Vxc
    # scatter) amplified by the orthogonaliser X = U s^-1/2
    dev_diff = float((fock_host.to(dev) - fock).abs().max())
    assert dev_diff <= 1e-8 * max(1.0, float(fock.abs().max())), "e2e and device-resident Fock differ: %g" % dev_diff

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (timed live above) ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fpk:
            peaks = json.load(fpk)
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    dmma_peak = _lib.peak_fp64_dmma(20000)
    # int8 tensor peak: measured here with the library's own tcgen05.mma.kind::i8 issue-rate microbenchmark
    # (MEASURED_PEAKS.json has no int8 entry; round 1 assumed 2 x bf16_tflops)
    i8_peak = max(_lib.peak_i8_mma(40000) for _ in range(3))
    i8_src = ("tcgen05.mma.kind::i8 issue-rate microbenchmark run in this process (b200qc_peak_i8_mma: M = 128, N = 256, "
              "K = 32, operands resident in shared memory, best of 3); for comparison 2 x bf16_tflops of "
              "MEASURED_PEAKS.json = %.0f" % (2.0 * peaks.get("bf16_tflops", 1590.0)))
    gb = h._gb if xc is not None else None
    ncomp = {1: 1, 2: 4, 4: 5}[h.xcfamily] if xc is not None else 1
    xc_flops = gb.flops_per_pass if gb is not None else 0.0      # 2 * sum_sb SBP * nsp^2 (K2 and K4 GEMM each)
    if gb is not None:
        impl_config.update(sb_points=gb.sbp, ao_screen=gb.eps, kept_ao_fraction=round(gb.kept_fraction, 4),
                      vxc_gemm=("tcgen05 int8 x%d slices" % gb.i8_slices) if gb.i8_slices else "fp64 DMMA",
                      vxc_operand_prep="fused vb slicer" if getattr(gb, "colmax", None) is not None else "fp64 vb + slicer",
                      rho_form=("point-stationary (64-point phi tiles resident in shared memory)"
                                if getattr(gb, "rho_bn", 0) == 128 else "row-tile streaming") if gb.rho_i8_slices else "fp64",
                      rho_gemm=("tcgen05 int8 x%d slices" % gb.rho_i8_slices) if gb.rho_i8_slices else "fp64 DMMA",
                      ao_resident_gb=round(gb.ao_bytes / 1e9, 2),
                      dense_equiv_flops_per_pass=2.0 * ngrid * h._nao_ao ** 2 / world)
    kern = {k: {"launches": c, "ms_per_launch": ms / c} for k, (c, ms) in prof.items()}
    dominant = max(prof.items(), key=lambda kv: kv[1][1])[0] if prof else None
    roofline = None
    def gemm_roofline(kname):
        """K2 / K4 GEMM: fp64 DMMA form -> fraction of the measured DMMA peak; tcgen05 int8 (Ozaki) form ->
        int8 ops (S (S + 1) / 2 slice products) against 2 x the measured bf16 tensor peak."""
        t = kern[kname]["ms_per_launch"] * 1e-3
        nsl = (gb.rho_i8_slices if kname == "rho_kernel" else gb.i8_slices) if gb is not None else 0
        if nsl:
            nprod = nsl * (nsl + 1) // 2
            ops = xc_flops * nprod
            peak = i8_peak
            # the same launch seen from HBM: every operand once (fp64 AO rows of the epilogue, int8 planes, outputs)
            nsp_ = gb.nsp.astype(np.float64)
            if kname == "rho_kernel":
                hb = gb.ao_bytes + nsl * gb.sbp * nsp_.sum() + nsl * (nsp_ ** 2).sum() + 8.0 * ncomp * gb.ngl
            else:
                hb = nsl * gb.sbp * (np.ceil(nsp_ / 128) * 128).sum() + nsl * gb.sbp * nsp_.sum() + 8.0 * (nsp_ ** 2).sum()
            hbm_view = {"algorithmic_bytes_per_launch": hb, "achieved_GBps": hb / t / 1e9, "peak_GBps": hbm_peak,
                        "frac": hb / t / 1e9 / hbm_peak,
                        "measured_traffic_frac_of_peak": (traffic.get(kname) / t / 1e9 / hbm_peak) if traffic.get(kname) else None}
            tensor_view = {"achieved": ops / t / 1e12, "peak": peak, "unit": "TOP/s (int8)", "frac": ops / t / 1e12 / peak,
                           "peak_source": i8_src, "algorithmic_ops_per_launch": ops, "int8_slice_products": nprod,
                           "fp64_equivalent_tflops": xc_flops / t / 1e12,
                           "fp64_equivalent_vs_dmma_peak": xc_flops / t / 1e12 / dmma_peak}
            # the binding roofline is the one with the larger lower bound on the launch time: algorithmic bytes at the
            # measured HBM rate against algorithmic int8 ops at the measured tcgen05 rate
            if hb / (hbm_peak * 1e9) > ops / (peak * 1e12):
                return {"kernel": kname, "bound": "hbm", "achieved": hb / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hb / t / 1e9 / hbm_peak, "traffic": traffic.get(kname), "peak_source": hbm_src,
                        "algorithmic_bytes_per_launch": hb,
                        "lower_bounds_ms": {"hbm": hb / (hbm_peak * 1e9) * 1e3, "tensor": ops / (peak * 1e12) * 1e3},
                        "tensor_view": tensor_view}
            out = {"kernel": kname, "bound": "tensor", "traffic": traffic.get(kname), "hbm_view": hbm_view,
                   "lower_bounds_ms": {"hbm": hb / (hbm_peak * 1e9) * 1e3, "tensor": ops / (peak * 1e12) * 1e3}}
            out.update(tensor_view)
            return out
        # meta-GGA: four GEMMs per pass (phi and the three gradient components on the left, hcgto.py:427-429, 485-489)
        ngemm = 4 if (xc is not None and h.xcfamily == 4) else 1
        ach = ngemm * xc_flops / t / 1e12
        return {"kernel": kname, "bound": "tensor", "achieved": ach, "peak": dmma_peak, "unit": "TFLOP/s",
                "frac": ach / dmma_peak, "traffic": traffic.get(kname),
                "peak_source": "fp64 DMMA (mma.sync m8n8k4) issue-rate microbenchmark run in this process "
                               "(b200qc_peak_fp64_dmma); MEASURED_PEAKS.json has no fp64 entry",
                "algorithmic_flops_per_launch": ngemm * xc_flops}
    def dfk_roofline():
        """DF-K: both tcgen05 GEMM launches of a build together.  Algorithmic fp64 flops: stage 1
        2 nao^2 naux nocc, stage 2 (symmetric) nao^2 naux nocc; each is S (S + 1) / 2 int8 products."""
        nsl = h.df._k_S
        nprod = nsl * (nsl + 1) // 2
        c, ms = prof["gemm_i8_kernel"]
        nocc_ = orb_host.shape[1]
        fl = 3.0 * h._nao_ao ** 2 * h.df._naux_local * nocc_ * (c / (2.0 * args.steps))   # per build
        t = ms / args.steps * 1e-3
        peak = i8_peak
        return {"kernel": "gemm_i8_kernel", "bound": "tensor", "achieved": fl * nprod / t / 1e12, "peak": peak,
                "unit": "TOP/s (int8)", "frac": fl * nprod / t / 1e12 / peak, "traffic": traffic.get("gemm_i8_kernel"),
                "peak_source": i8_src,
                "algorithmic_ops_per_launch": fl * nprod / 2.0, "launches_per_build": c / args.steps,
                "int8_slice_products": nprod, "fp64_equivalent_tflops": fl / t / 1e12,
                "fp64_equivalent_vs_dmma_peak": fl / t / 1e12 / dmma_peak}
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as ft:
            traffic = json.load(ft).get(args.workload, {})
    except Exception:
        pass
    if dominant in ("vxc_gemm_kernel", "rho_kernel"):
        roofline = gemm_roofline(dominant)
    elif dominant in ("dfj_pass1_kernel", "dfj_pass2_kernel"):
        nb = h.df._j3c_packed.numel() * 8.0
        ach = nb / (kern[dominant]["ms_per_launch"] * 1e-3) / 1e9
        roofline = {"kernel": dominant, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": traffic.get(dominant), "peak_source": hbm_src,
                    "algorithmic_bytes_per_launch": nb}
    elif dominant == "gemv_rows_kernel":
        nb = 8.0 * h._nao_ao ** 4
        ach = nb / (kern[dominant]["ms_per_launch"] * 1e-3) / 1e9
        roofline = {"kernel": dominant, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": None, "peak_source": hbm_src,
                    "algorithmic_bytes_per_launch": nb}
    elif dominant == "gemm_i8_kernel":
        roofline = dfk_roofline()
    elif dominant == "jk_kernel":
        # direct 4-centre J/K (csrc/jk_reg.cuh): fp64-pipe bound.  achieved = the plan's operation count (Rys roots,
        # 2-D tables and root sums once per primitive quartet + the tile contractions, b200qc_jkplan_flops) over the
        # summed launch time of the class kernels; peak = the DFMA issue rate measured in this process.
        jk_ms = sum(ms for k, (c, ms) in prof.items() if k == "jk_kernel") / args.steps
        hybrid = exx != 0.0
        fl = h._jkplan.flops(True, hybrid)
        fma_peak = _lib.peak_fp64_fma(20000)
        ach = fl / (jk_ms * 1e-3) / 1e12
        roofline = {"kernel": dominant, "bound": "fp64-alu", "achieved": ach, "peak": fma_peak, "unit": "TFLOP/s",
                    "frac": ach / fma_peak, "traffic": traffic.get(dominant),
                    "peak_source": "DFMA issue-rate peak measured in this process (b200qc_peak_fp64_fma); "
                                   "MEASURED_PEAKS.json has no fp64 entry",
                    "algorithmic_flops_per_build": fl,
                    "contracted_shell_quartets_per_s": h._jkplan.nquartets / (jk_ms * 1e-3),
                    "quartets": h._jkplan.nquartets, "quartets_register_engine": h._jkplan.nquartets_reg,
                    "note": "jk_kernel time = the whole region of class-pair kernels (4 side streams), bracketed once"}
    # secondary rooflines (HBM-bound kernels) for context
    extra = {}
    if "dfj_pass1_kernel" in kern and h.df is not None and h.df._j3c_packed is not None:
        nb = h.df._j3c_packed.numel() * 8.0 * getattr(h.df, "row_kept_fraction", 1.0)   # rows the passes really read
        for k in ("dfj_pass1_kernel", "dfj_pass2_kernel"):
            extra[k] = {"GB/s": nb / (kern[k]["ms_per_launch"] * 1e-3) / 1e9, "frac_of_hbm": nb / (
                kern[k]["ms_per_launch"] * 1e-3) / 1e9 / hbm_peak}
    if "vxc_vb_kernel" in kern:
        if getattr(gb, "colmax", None) is not None:     # fused: fp64 AO values in, int8 planes out
            nb = gb.ao_bytes + gb.i8_slices * gb.sbp * float(gb.nsp.sum())
        else:                                           # fp64 vb out (sliced by a second kernel)
            nb = (ncomp + 1) * gb.ao_bytes / ncomp
        extra["vxc_vb_kernel"] = {"GB/s": nb / (kern["vxc_vb_kernel"]["ms_per_launch"] * 1e-3) / 1e9,
                                  "frac_of_hbm": nb / (kern["vxc_vb_kernel"]["ms_per_launch"] * 1e-3) / 1e9 / hbm_peak}
    if "gemm_i8_kernel" in kern and h.df is not None and getattr(h.df, "_k_planes", None) is not None:
        r = dfk_roofline()
        extra["gemm_i8_kernel"] = {kk: r[kk] for kk in ("achieved", "peak", "unit", "frac", "fp64_equivalent_tflops")}
    for k in ("rho_kernel", "vxc_gemm_kernel"):
        if k in kern and gb is not None:
            r = gemm_roofline(k)
            extra[k] = {kk: r[kk] for kk in ("bound", "achieved", "peak", "unit", "frac", "lower_bounds_ms") if kk in r}
            tv = r.get("tensor_view", r)
            if "fp64_equivalent_tflops" in tv:
                extra[k]["fp64_equivalent_tflops"] = tv["fp64_equivalent_tflops"]
                extra[k]["int8_tops"] = tv["achieved"]
                extra[k]["frac_of_int8_peak"] = tv["frac"]

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        ms, info, sample = run_cpu_reference(args.workload, 2, 1)
        cpu_baseline = {"value": ms, "unit": "ms", "cores": info["cores"], "kind": "port", "sample": sample}

    xc_ms = sum(ms for k, (c, ms) in prof.items() if k in ("rho_kernel", "xc_kernel", "vxc_vb_kernel", "vxc_gemm_kernel",
                                                           "slab_reduce_kernel", "sb_gather_dm_kernel",
                                                           "sb_slice_kernel")) / args.steps
    line = {"metric": METRIC, "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "impl_config": impl_config, "clocks": clocks,
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(orb_host.numel() * 8),
                    "d2h_bytes_per_step": int(fock_host.numel() * 8)},
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "xc_grid_points_per_s": (ngrid / (xc_ms * 1e-3)) if xc_ms > 0 else None,
            "fp64_dmma_peak_tflops": dmma_peak, "int8_mma_peak_tops": i8_peak, "kernels": kern,
            "kernel_rooflines": extra}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
