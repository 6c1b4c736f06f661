"""AO values on a grid with the reference's names (dqc/hamilton/intor/gtoeval.py:60-73):
``eval_gto`` / ``eval_gradgto`` return (nao, ngrid) / (3, nao, ngrid), or (ngrid, nao) /
(3, ngrid, nao) with ``to_transpose=True``; ``eval_gto_padded`` hands out the padded buffer the
Fock-build kernels read in place."""
import torch
from dqc_b200 import _lib
from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
from dqc_b200.hamilton.intor.molintor import _device

__all__ = ["eval_gto", "eval_gradgto", "eval_laplgto", "eval_gto_padded"]


def eval_gto_padded(wrapper: LibcintWrapper, rgrid: torch.Tensor, deriv: int) -> torch.Tensor:
    """(ncomp, ngrid_ld, ao_ld) zero-padded; ncomp = 1 (values), 4 (values, d/dx, d/dy, d/dz) or 5 (+ Laplacian)."""
    dev = rgrid.device if rgrid.is_cuda else _device(wrapper)
    db = wrapper.device_basis(dev)
    s0, s1 = wrapper.shell_idxs
    return _lib.eval_gto(db, s0, s1, rgrid.to(dev).to(torch.float64), deriv)


def eval_gto(wrapper: LibcintWrapper, rgrid: torch.Tensor, *, to_transpose: bool = False) -> torch.Tensor:
    """AO values; differentiable (first order) with respect to the grid points and the atomic positions when either is
    part of an autograd graph (the reference's _EvalGTO.backward, gtoeval.py:124-193)."""
    pos = wrapper.parent.params[2]
    if torch.is_grad_enabled() and (rgrid.requires_grad or pos.requires_grad):
        from dqc_b200.hamilton.intor import deriv
        dev = rgrid.device if rgrid.is_cuda else _device(wrapper)
        ao = deriv.EvalGTOFunction.apply(rgrid.to(dev).to(torch.float64), pos, wrapper)
        return ao if to_transpose else ao.transpose(-2, -1)
    ao = eval_gto_padded(wrapper, rgrid, 0)[0, :rgrid.shape[0], :wrapper.nao()]
    return ao.contiguous() if to_transpose else ao.transpose(-2, -1).contiguous()


def eval_gradgto(wrapper: LibcintWrapper, rgrid: torch.Tensor, *, to_transpose: bool = False) -> torch.Tensor:
    ao = eval_gto_padded(wrapper, rgrid, 1)[1:, :rgrid.shape[0], :wrapper.nao()]
    return ao.contiguous() if to_transpose else ao.transpose(-2, -1).contiguous()


def eval_laplgto(wrapper: LibcintWrapper, rgrid: torch.Tensor, *, to_transpose: bool = False) -> torch.Tensor:
    """Laplacian of the AOs (the reference sums the xx, yy, zz components of GTOval_sph_deriv2, gtoeval.py:66-73)."""
    ao = eval_gto_padded(wrapper, rgrid, 2)[4, :rgrid.shape[0], :wrapper.nao()]
    return ao.contiguous() if to_transpose else ao.transpose(-2, -1).contiguous()
