#!/usr/bin/env python
"""nbasis / grid sweep (BASELINE.json configs[4], SURVEY 8d "C5"): carbon-only simple-cubic clusters, spacing
2.7 bohr, def2-SVP (14 AOs per atom), unpruned (nr, nang) grids sized to hit ~5e4 / 5e5 / 5e6 points; seeded
density D = 2 C C^T.  For every (nao, ngrid) the XC kernels (K2 density, K3 functional, K4 Vxc) are timed with the
library's own CUDA events (median-free mean over `--iters` builds after 3 warm-ups) and reported against their
rooflines; for every nao that fits, the density-fitted J (HBM-bound GEMV passes) and K (tcgen05 GEMMs).

    python tools/sweep.py --out profiles/r01_sweep.json [--natoms 8,36,72] [--grids 5e4,5e5,5e6]

One GPU; combinations whose resident AO values would not fit are skipped with the reason recorded (the BASELINE
config runs them on 8 GPUs, where every rank holds 1/8 of the grid)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

NANGS = [26, 50, 110, 194, 302, 434, 590]
MEM_LIMIT = 130e9


def cluster(natom, spacing=2.7):
    n = int(np.ceil(natom ** (1.0 / 3)))
    pts = np.array([[i, j, k] for i in range(n) for j in range(n) for k in range(n)], dtype=np.float64)
    pts -= pts.mean(0)
    order = np.argsort((pts ** 2).sum(1), kind="stable")
    return [6] * natom, pts[order[:natom]] * spacing


def pick_grid(natom, target):
    """(nr, nang) with natom * nr * nang closest to target, nr in 20..150."""
    best = None
    for nang in NANGS:
        nr = int(round(target / (natom * nang)))
        if nr < 20 or nr > 150:
            continue
        err = abs(natom * nr * nang - target) / target
        if best is None or err < best[0]:
            best = (err, nr, nang)
    if best is None:
        nang = NANGS[0] if target / (natom * NANGS[0]) < 20 else NANGS[-1]
        return max(20, min(150, int(round(target / (natom * nang))))), nang
    return best[1], best[2]


def seeded_dm(nao, nocc, dev):
    g = torch.Generator().manual_seed(0)
    q, _ = torch.linalg.qr(torch.randn(nao, nao, dtype=torch.float64, generator=g))
    c = q[:, :nocc]
    return c.to(dev), (2 * c @ c.T).to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--natoms", default="8,36,72,143,215")
    ap.add_argument("--grids", default="5e4,5e5,5e6")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--df-max-atoms", type=int, default=72)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_sweep.json"))
    args = ap.parse_args()
    assert torch.cuda.is_available(), "the sweep needs a CUDA device"
    dev = torch.device("cuda:0")
    from dqc_b200 import Mol, get_xc, _lib
    from dqc_b200.grid.factory import get_grid
    from dqc_b200.utils.config import config
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm, i8peak = peaks.get("hbm_gbs", 6650.0), 2.0 * peaks.get("bf16_tflops", 1590.0)
    dmma = _lib.peak_fp64_dmma(20000)
    xc = get_xc("gga_x_pbe + gga_c_pbe")
    rows, dfrows = [], []

    def dump():
        out = {"what": "nbasis/grid sweep on 1 x B200 (tools/sweep.py)",
               "peaks": {"hbm_gbs": hbm, "int8_tops": i8peak, "fp64_dmma_tflops": dmma}, "xc": rows, "df": dfrows}
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)

    for natom in [int(x) for x in args.natoms.split(",")]:
        zs, pos = cluster(natom)
        t0 = time.perf_counter()
        mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=torch.float64)), basis="def2-svp", device=dev)
        h = mol.get_hamiltonian()
        nao, nao2 = h._nao_ao, h.nao
        nocc = 3 * natom
        orb, dm = seeded_dm(nao2, min(nocc, nao2), dev)
        s0, s1 = h.libcint_wrapper.shell_idxs
        sizes = np.diff(h._devbasis.ao_loc[s0:s1 + 1]).astype(np.float64)
        for target in [float(x) for x in args.grids.split(",")]:
            nr, nang = pick_grid(natom, target)
            grid = get_grid(zs, torch.tensor(pos, dtype=torch.float64), nr=nr, nang=nang, truncate=None,
                            radgrid_transform="sg3-dasgupta", device=dev)
            ngrid = int(grid.get_rgrid().shape[0])
            row = {"natom": natom, "nao": nao, "nr": nr, "nang": nang, "ngrid": ngrid}
            flags = _lib.ao_screen(h._devbasis, s0, s1, grid.get_rgrid().to(dev).contiguous(), config.SB_POINTS,
                                   config.AO_SCREEN, 1).cpu().numpy()
            nsp = np.maximum(64.0, np.ceil((flags * sizes[None, :]).sum(1) / 64.0) * 64.0)
            est = 72.0 * config.SB_POINTS * nsp.sum() + 6.0 * 2 * (nsp ** 2).sum()
            row["resident_bytes_estimate"] = est
            if est > MEM_LIMIT:
                row["skipped"] = "resident AO values + int8 planes ~%.0f GB exceed one GPU (8-GPU config)" % (est / 1e9)
                rows.append(row)
                print(json.dumps(row), flush=True)
                continue
            h.setup_grid(grid, xc)
            gb = h._gb
            for _ in range(3):
                h._vxc_ao_partial(dm)
            torch.cuda.synchronize()
            _lib.profile_enable(True)
            for _ in range(args.iters):
                h._vxc_ao_partial(dm)
            prof = _lib.profile_read()
            _lib.profile_enable(False)
            ms = {k: v[1] / args.iters for k, v in prof.items()}
            fl = gb.flops_per_pass
            t_rho, t_vx = ms.get("rho_kernel", 0.0) * 1e-3, ms.get("vxc_gemm_kernel", 0.0) * 1e-3
            xc_ms = sum(ms.get(k, 0.0) for k in ("rho_kernel", "xc_kernel", "vxc_vb_kernel", "vxc_gemm_kernel",
                                                 "sb_gather_dm_kernel", "sb_slice_kernel"))
            row.update(kept_ao_fraction=round(gb.kept_fraction, 4), mean_nsp=float(nsp.mean()),
                       ao_resident_gb=round(gb.ao_bytes / 1e9, 3), kernel_ms={k: round(v, 4) for k, v in ms.items()},
                       xc_ms_per_iter=xc_ms, xc_grid_points_per_s=ngrid / (xc_ms * 1e-3),
                       dense_equiv_flops=2.0 * 2.0 * ngrid * nao ** 2, sparse_flops=2.0 * fl,
                       rho_fp64_equiv_tflops=fl / t_rho / 1e12, vxc_fp64_equiv_tflops=fl / t_vx / 1e12,
                       rho_int8_frac_of_peak=21 * fl / t_rho / 1e12 / i8peak,
                       vxc_int8_frac_of_peak=21 * fl / t_vx / 1e12 / i8peak,
                       vxc_vb_hbm_gbs=1.25 * gb.ao_bytes / (ms.get("vxc_vb_kernel", 1e9) * 1e-3) / 1e9,
                       vxc_vb_frac_of_hbm=1.25 * gb.ao_bytes / (ms.get("vxc_vb_kernel", 1e9) * 1e-3) / 1e9 / hbm)
            rows.append(row)
            print(json.dumps(row), flush=True)
            del gb
            h._gb = None
            torch.cuda.empty_cache()
        if natom <= args.df_max_atoms:
            mol.densityfit(auxbasis="etb-jfit")
            hd = mol.get_hamiltonian().build()
            naux = hd.df._naux
            dmd = hd.ao_orb2dm(orb, torch.full((orb.shape[1],), 2.0, dtype=torch.float64, device=dev))
            for _ in range(3):
                hd.get_fock_2e(dmd, exx=0.25, with_xc=False)
            torch.cuda.synchronize()
            _lib.profile_enable(True)
            for _ in range(args.iters):
                hd.get_fock_2e(dmd, exx=0.25, with_xc=False)
            prof = _lib.profile_read()
            _lib.profile_enable(False)
            ms = {k: v[1] / args.iters for k, v in prof.items()}
            nb = hd.df._j3c_packed.numel() * 8.0
            flk = 3.0 * nao ** 2 * naux * orb.shape[1]
            tk = ms.get("gemm_i8_kernel", 1e9) * 1e-3
            r = {"natom": natom, "nao": nao, "naux": naux, "nocc": int(orb.shape[1]),
                 "kernel_ms": {k: round(v, 4) for k, v in ms.items()},
                 "dfj_pass1_gbs": nb / (ms["dfj_pass1_kernel"] * 1e-3) / 1e9,
                 "dfj_pass2_gbs": nb / (ms["dfj_pass2_kernel"] * 1e-3) / 1e9,
                 "dfj_pass1_frac_of_hbm": nb / (ms["dfj_pass1_kernel"] * 1e-3) / 1e9 / hbm,
                 "dfj_pass2_frac_of_hbm": nb / (ms["dfj_pass2_kernel"] * 1e-3) / 1e9 / hbm,
                 "dfk_fp64_equiv_tflops": flk / tk / 1e12, "dfk_int8_frac_of_peak": 21 * flk / tk / 1e12 / i8peak,
                 "dfk_vs_dmma_peak": flk / tk / 1e12 / dmma}
            dfrows.append(r)
            print(json.dumps(r), flush=True)
            del hd
            torch.cuda.empty_cache()
        dump()
        print("natom %d done in %.1f s" % (natom, time.perf_counter() - t0), file=sys.stderr, flush=True)
    dump()


if __name__ == "__main__":
    main()
