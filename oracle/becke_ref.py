"""ORACLE (test infrastructure only): Becke partition weights on CPU, restating
dqc/grid/multiatoms_grid.py:173-273 (dense form of the same arithmetic: the reference's sparse
bookkeeping only skips cells it then fills with zero)."""
import numpy as np


def becke_weights(xyz, owner, atompos, radii=None, ratom_adjust="becke"):
    """xyz (ngrid,3); owner (ngrid,) atom index of each point; atompos (natoms,3);
    radii (natoms,) or None.  Returns (ngrid,) weights P_owner / sum_k P_k."""
    xyz = np.asarray(xyz, dtype=np.float64)
    atompos = np.asarray(atompos, dtype=np.float64)
    nat = atompos.shape[0]
    rd = atompos[None, :, :] - atompos[:, None, :] + np.eye(nat)[:, :, None]
    ratoms = np.linalg.norm(rd, axis=-1)                      # [i, j] = |R_j - R_i| (+eye)
    aij = None
    if radii is not None:
        rad = np.asarray(radii, dtype=np.float64)
        if ratom_adjust == "treutler":
            rad = rad ** 0.5
        uij = (rad[None, :] - rad[:, None]) / (rad[None, :] + rad[:, None])
        aij = np.clip(uij / (uij * uij - 1), -0.45, 0.45)
    out = np.empty(xyz.shape[0])
    for s in range(0, xyz.shape[0], 8192):
        p = xyz[s:s + 8192]
        rg = np.linalg.norm(p[None, :, :] - atompos[:, None, :], axis=-1)   # (nat, n)
        mu = (rg[None, :, :] - rg[:, None, :]) / ratoms[:, :, None]           # [i, j, g] = (r_j - r_i)/R_ij
        if aij is not None:
            mu = mu + aij[:, :, None] * (1 - mu * mu)
        keep = np.all(mu < 0.74, axis=0)                                     # (nat_j, n)
        f = mu
        for _ in range(3):
            f = 0.5 * f * (3 - f * f)
        sfun = 0.5 * (1.0 + 1e-12 - f)
        sfun[np.arange(nat), np.arange(nat), :] += 0.5
        P = sfun.prod(axis=0) * keep                                          # (nat_j, n)
        own = owner[s:s + 8192]
        out[s:s + 8192] = P[own, np.arange(p.shape[0])] / P.sum(axis=0)
    return out
