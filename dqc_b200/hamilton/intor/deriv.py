"""Derivative integrals and the backward pass of the integral front-end with respect to the atomic positions.

The reference differentiates its integrals through libcint's derivative kernels: ``_Int2cFunction`` /
``_Int4cFunction.backward`` ask the name manager for ``int1e_ipovlp``, ``int1e_ipkin``, ``int1e_ipnuc``,
``int1e_iprinv``, ``int2e_ip1`` ... (dqc/hamilton/intor/molintor.py:178-578, namemgr.py:71-105) and ``_EvalGTO.backward``
for ``GTOval_ip`` (gtoeval.py:124-193).  Here the same quantities come from the Rys kernels of libb200qc.so run in
their RAW CARTESIAN mode (``b200qc_basis_set_cartesian``) on a helper basis that holds, for every shell
(l, {a_p, c_p}) of the wrapper,

    O: the shell itself,   U: (l + 1, {a_p, 2 a_p c_p}),   D: (l - 1, {a_p, c_p})   (l > 0)

because   d/dx [x^a y^b z^c sum_p c_p exp(-a_p r^2)] = a x^(a-1) y^b z^c sum_p c_p exp(..) - x^(a+1) y^b z^c sum_p 2 a_p c_p exp(..).
An "ip" block is therefore a fixed integer combination of U and D blocks, followed by the kernels' own
cartesian -> real-spherical matrices (``b200qc_c2s_matrix``).  ``ip`` differentiates with respect to the ELECTRON
coordinate, as in libcint; the derivative with respect to the centre of the function is its negative.

This is the first slice of SURVEY 8f rank 1: gradients with respect to atomic positions of S, T, V (basis-function and
operator parts), (ij|kl) and the AO values.  Gradients with respect to exponents / contraction coefficients, second
derivatives (``ipip``), and the backward of the grid-contraction kernels raise NotImplementedError.
"""
from typing import Dict, List, Optional, Tuple
import numpy as np
import torch
from dqc_b200 import _lib
from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper

__all__ = ["DerivBasis", "ip1e", "ip2e", "Int1eFunction", "Int2eFunction", "EvalGTOFunction"]


def _cart_components(l: int) -> List[Tuple[int, int, int]]:
    # libcint order: lx descending, then ly descending
    return [(x, y, l - x - y) for x in range(l, -1, -1) for y in range(l - x, -1, -1)]


def _ncart(l: int) -> int:
    return (l + 1) * (l + 2) // 2


class DerivBasis(object):
    """Cartesian helper basis (groups O, U, D) of the shells of ``wrapper`` and the constant matrices that turn its raw
    blocks into spherical ``ip`` integrals."""

    def __init__(self, wrapper: LibcintWrapper, device: torch.device):
        atm, bas, env = wrapper.parent.atm_bas_env
        s0, s1 = wrapper.shell_idxs
        if int(bas[s0:s1, 1].max()) + 1 > _lib.LMAX:
            raise NotImplementedError("derivative integrals need l + 1 <= %d" % _lib.LMAX)
        env = list(env)
        rows_o, rows_u, rows_d = [], [], []
        self.ls = [int(bas[s, 1]) for s in range(s0, s1)]
        for s in range(s0, s1):
            _, l, ng, _, _, pe, pc, _ = (int(v) for v in bas[s])
            rows_o.append(list(bas[s]))
            up = list(bas[s])
            up[1] = l + 1
            up[6] = len(env)
            env.extend(2.0 * env[pe + k] * env[pc + k] for k in range(ng))
            rows_u.append(up)
            if l > 0:
                dn = list(bas[s])
                dn[1] = l - 1
                rows_d.append(dn)
        nbas = np.array(rows_o + rows_u + rows_d, dtype=np.int32)
        n = s1 - s0
        self.sl_o, self.sl_u, self.sl_d = (0, n), (n, 2 * n), (2 * n, len(nbas))
        ao_loc = np.concatenate([[0], np.cumsum([_ncart(int(r[1])) for r in nbas])]).astype(np.int32)
        self.db = _lib.DeviceBasis(atm, nbas, np.array(env, dtype=np.float64), ao_loc, spherical=False, device=device)
        self.device = device
        off = lambda sl: int(ao_loc[sl[0]])
        nco = int(ao_loc[n] - ao_loc[0])
        ncu = int(ao_loc[2 * n] - ao_loc[n])
        ncd = int(ao_loc[-1] - ao_loc[2 * n])
        nsph = int(sum(2 * l + 1 for l in self.ls))
        T = np.zeros((nsph, nco))
        gup = np.zeros((3, nco, ncu))
        gdn = np.zeros((3, nco, max(ncd, 1)))
        isph, d_shell = 0, 0
        c2s: Dict[int, np.ndarray] = {}
        for k, l in enumerate(self.ls):
            if l not in c2s:
                c2s[l] = _lib.c2s_matrix(l)
            co = int(ao_loc[k]) - off(self.sl_o)
            cu = int(ao_loc[n + k]) - off(self.sl_u)
            T[isph:isph + 2 * l + 1, co:co + _ncart(l)] = c2s[l]
            isph += 2 * l + 1
            up_idx = {c: i for i, c in enumerate(_cart_components(l + 1))}
            dn_idx = {c: i for i, c in enumerate(_cart_components(l - 1))} if l > 0 else {}
            cd = int(ao_loc[2 * n + d_shell]) - off(self.sl_d) if l > 0 else 0
            for i, comp in enumerate(_cart_components(l)):
                for d in range(3):
                    up = list(comp)
                    up[d] += 1
                    gup[d, co + i, cu + up_idx[tuple(up)]] = -1.0
                    if comp[d] > 0:
                        dn = list(comp)
                        dn[d] -= 1
                        gdn[d, co + i, cd + dn_idx[tuple(dn)]] = float(comp[d])
            if l > 0:
                d_shell += 1
        tt = lambda a: torch.as_tensor(a, dtype=torch.float64, device=device)
        self.T, self.gup, self.gdn = tt(T), tt(gup), tt(gdn)
        self.has_d = ncd > 0

    def _ip_cart(self, iu: torch.Tensor, idn: Optional[torch.Tensor]) -> torch.Tensor:
        """(ncartU, ...), (ncartD, ...) raw blocks -> (3, ncartO, ...) derivative on the first index."""
        flat_u = iu.reshape(iu.shape[0], -1)
        out = torch.matmul(self.gup, flat_u)
        if idn is not None:
            out = out + torch.matmul(self.gdn, idn.reshape(idn.shape[0], -1))
        return out.reshape(3, -1, *iu.shape[1:])


def _deriv_basis(wrapper: LibcintWrapper, device: torch.device) -> DerivBasis:
    cache = wrapper.__dict__.setdefault("_b200_deriv_basis", {})
    key = str(device)
    if key not in cache:
        cache[key] = DerivBasis(wrapper, device)
    return cache[key]


def _cuda_device(wrapper: LibcintWrapper) -> torch.device:
    dev = wrapper.device
    if dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def ip1e(shortname: str, wrapper: LibcintWrapper, rinv_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(3, nao, nao): <d_d phi_i| op |phi_j> for op = "ovlp", "kin", "nuc", "rinv" -- libcint's int1e_ipovlp, int1e_ipkin,
    int1e_ipnuc, int1e_iprinv (derivative with respect to the electron coordinate, acting on the bra)."""
    if shortname not in ("ovlp", "kin", "nuc", "rinv"):
        raise NotImplementedError("int1e_ip%s is not built" % shortname)
    dev = _cuda_device(wrapper)
    b = _deriv_basis(wrapper, dev)
    orig = None if rinv_pos is None else rinv_pos.detach().cpu().numpy()
    with torch.cuda.device(dev):
        iu = _lib.int1e(b.db, shortname, (*b.sl_u, *b.sl_o), orig)
        idn = _lib.int1e(b.db, shortname, (*b.sl_d, *b.sl_o), orig) if b.has_d else None
        cart = b._ip_cart(iu, idn)                                     # (3, ncartO, ncartO)
        return torch.matmul(torch.matmul(b.T, cart), b.T.t())


def ip2e(wrapper: LibcintWrapper) -> torch.Tensor:
    """(3, nao, nao, nao, nao): (d_d phi_i phi_j | phi_k phi_l) -- libcint's int2e_ip1.  Dense: small molecules only."""
    dev = _cuda_device(wrapper)
    b = _deriv_basis(wrapper, dev)
    with torch.cuda.device(dev):
        iu = _lib.int2e(b.db, (*b.sl_u, *b.sl_o, *b.sl_o, *b.sl_o))
        idn = _lib.int2e(b.db, (*b.sl_d, *b.sl_o, *b.sl_o, *b.sl_o)) if b.has_d else None
        cart = b._ip_cart(iu, idn)                                     # (3, cO, cO, cO, cO)
        return torch.einsum("ia,jb,kc,ld,xabcd->xijkl", b.T, b.T, b.T, b.T, cart)


def _scatter_atoms(per_ao: torch.Tensor, ao_to_atom: torch.Tensor, natoms: int) -> torch.Tensor:
    """(3, nao) contributions of every AO -> (natoms, 3) summed over the AOs of each atom."""
    out = torch.zeros(natoms, 3, dtype=per_ao.dtype, device=per_ao.device)
    out.index_add_(0, ao_to_atom.to(per_ao.device), per_ao.t().contiguous())
    return out


class Int1eFunction(torch.autograd.Function):
    """S, T or V of a wrapper as a function of the atomic positions (natoms, 3).

    dI_ij/dR_A = -<d phi_i|op|phi_j> [i on A] - <d phi_j|op|phi_i> [j on A]   (+ for op = nuc, from the operator:
    -Z_A (<d phi_i| 1/r_A |phi_j> + <d phi_j| 1/r_A |phi_i>), by translational invariance)."""

    @staticmethod
    def forward(ctx, allpos: torch.Tensor, wrapper: LibcintWrapper, shortname: str):
        from dqc_b200.hamilton.intor import molintor
        ctx.wrapper, ctx.shortname, ctx.pos_device = wrapper, shortname, allpos.device
        return molintor._int1e_nograd(shortname, wrapper)

    @staticmethod
    def backward(ctx, gout: torch.Tensor):
        w, name = ctx.wrapper, ctx.shortname
        if w.fracz:
            raise NotImplementedError("position gradients with fractional nuclear charges are not built")
        gs = gout + gout.transpose(-2, -1)
        ip = ip1e(name, w)
        natoms = w.parent.natoms
        grad = _scatter_atoms(-(ip * gs.unsqueeze(0)).sum(-1), w.ao_to_atom(), natoms)
        if name == "nuc":
            for ia, ab in enumerate(w.parent.atombases):
                ipr = ip1e("rinv", w, rinv_pos=ab.pos)
                grad[ia] += -float(ab.atomz) * (ipr * gs.unsqueeze(0)).sum((-2, -1))
        return grad.to(ctx.pos_device), None, None


class Int2eFunction(torch.autograd.Function):
    """(ij|kl) as a function of the atomic positions; backward through int2e_ip1 on all four indices."""

    @staticmethod
    def forward(ctx, allpos: torch.Tensor, wrapper: LibcintWrapper):
        from dqc_b200.hamilton.intor import molintor
        ctx.wrapper, ctx.pos_device = wrapper, allpos.device
        return molintor._int2e_nograd(wrapper)

    @staticmethod
    def backward(ctx, gout: torch.Tensor):
        w = ctx.wrapper
        gs = gout + gout.permute(1, 0, 2, 3) + gout.permute(2, 3, 0, 1) + gout.permute(3, 2, 0, 1)
        ip = ip2e(w)
        per_ao = -(ip * gs.unsqueeze(0)).sum((-3, -2, -1))            # (3, nao)
        return _scatter_atoms(per_ao, w.ao_to_atom(), w.parent.natoms).to(ctx.pos_device), None


class EvalGTOFunction(torch.autograd.Function):
    """AO values (ngrid, nao) as a function of the grid points (ngrid, 3) and the atomic positions (natoms, 3):
    d phi_mu(r - R_A)/dr = grad phi_mu, d/dR_A = -grad phi_mu (the first-order part of gtoeval.py:124-193)."""

    @staticmethod
    def forward(ctx, rgrid: torch.Tensor, allpos: torch.Tensor, wrapper: LibcintWrapper):
        from dqc_b200.hamilton.intor import gtoeval
        ao = gtoeval.eval_gto_padded(wrapper, rgrid.detach(), 1)[:, :rgrid.shape[0], :wrapper.nao()]
        ctx.wrapper, ctx.pos_device, ctx.grid_device = wrapper, allpos.device, rgrid.device
        ctx.save_for_backward(ao[1:].clone())
        ctx.needs = (ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return ao[0].clone()

    @staticmethod
    def backward(ctx, gout: torch.Tensor):
        (dao,) = ctx.saved_tensors                                     # (3, ngrid, nao)
        w = ctx.wrapper
        g = gout.to(dao.device)
        grad_r = grad_p = None
        if ctx.needs[0]:
            grad_r = torch.einsum("gm,dgm->gd", g, dao).to(ctx.grid_device)
        if ctx.needs[1]:
            per_ao = -torch.einsum("gm,dgm->dm", g, dao)
            grad_p = _scatter_atoms(per_ao, w.ao_to_atom(), w.parent.natoms).to(ctx.pos_device)
        return grad_r, grad_p, None
