"""Multipole one-electron integrals <phi_i| r_a r_b ... |phi_j> (origin at zero): libcint's int1e_r, int1e_rr, int1e_rrr
as the reference asks for them through ``intor.int1e("r0" * n, wrapper)`` for the electric-field terms of the core
Hamiltonian (dqc/hamilton/hcgto.py:118-127, namemgr.py:22-24: 3, 9, 27 components).

Built from the overlap kernel of libb200qc.so in its raw cartesian mode:  r_d = (r - A)_d + A_d  on the centre A of the
bra function, and (r - A)_d x^a y^b z^c = the cartesian component with exponent a + e_d of the SAME radial function one
angular momentum higher.  So <r_a r_b phi_i|phi_j> is a fixed combination of overlap blocks over helper shells with
l + 1, l + 2 (same exponents and coefficients) and the shell itself, followed by the kernels' own cartesian ->
real-spherical matrices.  l + n must stay within the kernels' range (l <= 4): dipoles and quadrupoles for s..d shells,
octupoles for s and p."""
from itertools import product
from typing import List
import numpy as np
import torch
from dqc_b200 import _lib
from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
from dqc_b200.hamilton.intor.deriv import _cart_components, _ncart, _cuda_device

__all__ = ["multipole1e"]


class _RaisedBasis(object):
    """Cartesian helper basis: group k (k = 0..order) holds every shell of the wrapper with angular momentum l + k."""

    def __init__(self, wrapper: LibcintWrapper, order: int, device: torch.device):
        atm, bas, env = wrapper.parent.atm_bas_env
        s0, s1 = wrapper.shell_idxs
        self.ls = [int(bas[s, 1]) for s in range(s0, s1)]
        if max(self.ls) + order > _lib.LMAX:
            raise NotImplementedError("multipole integrals of order %d need l + %d <= %d" % (order, order, _lib.LMAX))
        rows = []
        for k in range(order + 1):
            for s in range(s0, s1):
                r = list(bas[s])
                r[1] = int(r[1]) + k
                rows.append(r)
        nbas = np.array(rows, dtype=np.int32)
        n = s1 - s0
        self.n, self.order = n, order
        self.ao_loc = np.concatenate([[0], np.cumsum([_ncart(int(r[1])) for r in nbas])]).astype(np.int32)
        self.db = _lib.DeviceBasis(atm, nbas, np.asarray(env, dtype=np.float64), self.ao_loc, spherical=False, device=device)
        self.centres = [np.array(env[int(atm[int(bas[s, 0]), 1]):int(atm[int(bas[s, 0]), 1]) + 3]) for s in range(s0, s1)]
        nsph = int(sum(2 * l + 1 for l in self.ls))
        nco = int(self.ao_loc[n])
        T = np.zeros((nsph, nco))
        isph = 0
        for k, l in enumerate(self.ls):
            T[isph:isph + 2 * l + 1, int(self.ao_loc[k]):int(self.ao_loc[k + 1])] = _lib.c2s_matrix(l)
            isph += 2 * l + 1
        self.T = torch.as_tensor(T, dtype=torch.float64, device=device)

    def group(self, k: int):
        return (k * self.n, (k + 1) * self.n)

    def ncart(self, k: int) -> int:
        return int(self.ao_loc[(k + 1) * self.n] - self.ao_loc[k * self.n])


def multipole1e(order: int, wrapper: LibcintWrapper) -> torch.Tensor:
    """(3^order, nao, nao): <phi_i| r_{d1} ... r_{d_order} |phi_j>, components in C order of (d1, ..., d_order)."""
    assert order in (1, 2, 3)
    dev = _cuda_device(wrapper)
    cache = wrapper.__dict__.setdefault("_b200_raised_basis", {})
    key = (str(dev), order)
    if key not in cache:
        cache[key] = _RaisedBasis(wrapper, order, dev)
    b = cache[key]
    nco = b.ncart(0)
    with torch.cuda.device(dev):
        # overlap blocks <cart_(l+k) i | cart_l j>, k = 0..order
        blocks = [_lib.int1e(b.db, "ovlp", (*b.group(k), *b.group(0))) for k in range(order + 1)]
        out = torch.zeros((3 ** order, nco, nco), dtype=torch.float64, device=dev)
        # gather matrices per component tuple: prod_d ((r - A)_d + A_d) expanded over the subsets that are "raised"
        for ic, comps in enumerate(product(range(3), repeat=order)):
            for mask in product((0, 1), repeat=order):          # 1 = raise by this coordinate, 0 = multiply by A_d
                k = sum(mask)
                g = np.zeros((nco, b.ncart(k)))
                for sh, l in enumerate(b.ls):
                    fac = 1.0
                    for d, m in zip(comps, mask):
                        if not m:
                            fac *= b.centres[sh][d]
                    if fac == 0.0:
                        continue
                    co = int(b.ao_loc[sh])
                    ck = int(b.ao_loc[k * b.n + sh]) - int(b.ao_loc[k * b.n])
                    idx = {c: i for i, c in enumerate(_cart_components(l + k))}
                    for i, comp in enumerate(_cart_components(l)):
                        up = list(comp)
                        for d, m in zip(comps, mask):
                            if m:
                                up[d] += 1
                        g[co + i, ck + idx[tuple(up)]] += fac
                out[ic] += torch.as_tensor(g, device=dev) @ blocks[k]
        return torch.matmul(torch.matmul(b.T, out), b.T.t())
