"""ORACLE (test infrastructure only): SCF on top of oracle/fock_ref.py, used to pin the oracle
against the reference's golden energies (dqc/test/test_hf.py:18-51,141-206, test_ks.py:40-111).
The fixed point is the reference's (dqc/qccalc/hf.py:93-119, ks.py:110-187: F -> lowest
eigenvectors -> D = C w C^T -> F'); it is reached here with plain DIIS, which changes the path,
not the converged answer."""
import numpy as np
import torch
from oracle import fock_ref


def occ_weights(nelec, spin):
    """(restricted weights, up weights, down weights) as dqc/system/mol.py:421-443."""
    ndn = (nelec - spin) // 2
    nup = ndn + spin
    wu = torch.ones(nup, dtype=torch.float64)
    wd_full = torch.zeros(nup, dtype=torch.float64)
    wd_full[:ndn] = 1
    wd = torch.ones(ndn, dtype=torch.float64) if ndn > 0 else torch.zeros(1, dtype=torch.float64)
    return wu + wd_full, wu, wd


def _eig_dm(h, F, w):
    n = len(w)
    if h.orthozer:
        ev, C = torch.linalg.eigh(F)
    else:
        L = torch.linalg.cholesky(h.S)
        Li = torch.linalg.inv(L)
        ev, Cp = torch.linalg.eigh(Li @ F @ Li.T)
        C = Li.T @ Cp
    return h.ao_orb2dm(C[:, :n], w)


def run_scf(h, atomzs, atompos, nelec, spin=0, method="hf", restricted=None, maxiter=200, tol=1e-10):
    """Returns (energy, dm) with dm a tensor (restricted) or (dm_u, dm_d)."""
    polarized = (spin != 0) if restricted is None else (not restricted)
    w, wu, wd = occ_weights(nelec, spin)
    enn = fock_ref.nuclei_energy(atomzs, atompos)
    hcore = h.kinnucl_mat

    def fock(dm):
        dmtot = dm[0] + dm[1] if polarized else dm
        J = h.get_elrep(dmtot)
        if method == "hf":
            if polarized:
                return [hcore + J + h.get_exchange(2 * dm[0]), hcore + J + h.get_exchange(2 * dm[1])]
            return [hcore + J + h.get_exchange(dm)]
        if h.xcstr is None:
            return [hcore + J] * (2 if polarized else 1)
        if polarized:
            vu, vd = h.get_vxc((dm[0], dm[1]))
            return [hcore + J + vu, hcore + J + vd]
        return [hcore + J + h.get_vxc(dm)]

    def energy(dm):
        dmtot = dm[0] + dm[1] if polarized else dm
        e = h.get_e_hcore(dmtot) + h.get_e_elrep(dmtot)
        if method == "hf":
            e = e + h.get_e_exchange((dm[0], dm[1]) if polarized else dm)
        elif h.xcstr is not None:
            e = e + h.get_e_xc((dm[0], dm[1]) if polarized else dm)
        return float(e) + enn

    ws = [wu, wd] if polarized else [w]
    # dm0 = "1e": zero density -> core Fock -> density (scf_qccalc.py:88-91)
    dm = [_eig_dm(h, hcore, wi) for wi in ws]
    S = h.olp_mat
    fs, errs = [], []
    elast = None
    for it in range(maxiter):
        F = fock(dm if polarized else dm[0])
        err = torch.cat([(Fi @ di @ S - S @ di @ Fi).reshape(-1) for Fi, di in zip(F, dm)])
        fs.append(torch.stack(F))
        errs.append(err)
        fs, errs = fs[-8:], errs[-8:]
        if len(fs) > 1:
            n = len(fs)
            B = torch.zeros(n + 1, n + 1, dtype=torch.float64)
            for a in range(n):
                for b in range(n):
                    B[a, b] = errs[a] @ errs[b]
            B[n, :n] = B[:n, n] = -1
            rhs = torch.zeros(n + 1, dtype=torch.float64)
            rhs[n] = -1
            try:
                c = torch.linalg.solve(B, rhs)[:n]
                Fm = sum(ci * fi for ci, fi in zip(c, fs))
            except Exception:
                Fm = fs[-1]
        else:
            Fm = fs[-1]
        dm = [_eig_dm(h, Fm[i], ws[i]) for i in range(len(ws))]
        e = energy(dm if polarized else dm[0])
        if elast is not None and abs(e - elast) < tol and float(err.abs().max()) < 1e-7:
            break
        elast = e
    return e, (tuple(dm) if polarized else dm[0])
