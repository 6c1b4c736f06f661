"""Build dqc_b200/data/lebedev.npz from the Lebedev-Laikov quadrature tables.

The Lebedev-Laikov rules are published mathematical constants (Lebedev & Laikov, Doklady
Mathematics 59 (1999) 477).  The reference ships them as 32 text files
(dqc/datasets/lebedevquad/lebedev_NNN.txt, columns: phi[deg] theta[deg] weight, sum(w)=1; read by
dqc/grid/lebedev_grid.py:33-60).  This script reads those public tables where they lie (run in
the build container only; /root/reference is absent on the GPU box) and stores them, keeping the
point order so that the product grid has the reference's point ordering, as one compressed
binary: key "p%03d" -> float64 (npts, 3) = (phi_rad, theta_rad, w).
"""
import glob
import os
import sys
import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/dqc/datasets/lebedevquad"
out = os.path.join(os.path.dirname(__file__), "..", "dqc_b200", "data", "lebedev.npz")
tab = {}
for f in sorted(glob.glob(os.path.join(src, "lebedev_*.txt"))):
    prec = int(os.path.basename(f)[8:11])
    a = np.loadtxt(f)
    a[:, :2] *= np.pi / 180.0   # degrees -> radians, as lebedev_grid.py:27
    assert abs(a[:, 2].sum() - 1.0) < 1e-10
    tab["p%03d" % prec] = a
np.savez_compressed(out, **tab)
print("wrote", out, {k: v.shape[0] for k, v in tab.items()})
