"""Containers that travel along the Fock-build path.

Same names, fields and behaviour as the reference's dqc/utils/datastruct.py:28-184
(CGTOBasis.wfnormalize_ :34-61, AtomCGTOBasis :63-67, DensityFitInfo :73-76, SpinParam :78-132,
ValGrad :134-184) so the operator surface built on them is a drop-in."""
from __future__ import annotations
from dataclasses import dataclass
from typing import Callable, Generic, List, Optional, TypeVar, Union
import math
import torch

__all__ = ["CGTOBasis", "AtomCGTOBasis", "ValGrad", "SpinParam", "DensityFitInfo", "ZType",
           "is_z_float", "gaussian_int"]

T = TypeVar("T")
ZType = Union[int, float, torch.Tensor]


def is_z_float(a: ZType) -> bool:
    return a.is_floating_point() if isinstance(a, torch.Tensor) else isinstance(a, float)


def gaussian_int(n: int, alpha):
    """int_0^inf x^n exp(-alpha x^2) dx = Gamma((n+1)/2) / (2 alpha^((n+1)/2))
    (reference: dqc/utils/misc.py:53-56)."""
    h = 0.5 * (n + 1)
    return math.gamma(h) / (2 * alpha ** h)


@dataclass
class CGTOBasis:
    angmom: int
    alphas: torch.Tensor  # (nprim,)
    coeffs: torch.Tensor  # (nprim,)
    normalized: bool = False

    def wfnormalize_(self) -> "CGTOBasis":
        """Radial normalisation exactly as dqc/utils/datastruct.py:34-61 (libcint's CINTgto_norm
        per primitive, then the contracted shell is rescaled to unit radial self-overlap)."""
        if self.normalized:
            return self
        l2 = 2 * self.angmom + 2
        c = self.coeffs / torch.sqrt(gaussian_int(l2, 2 * self.alphas))
        pair = gaussian_int(l2, self.alphas.unsqueeze(-1) + self.alphas.unsqueeze(-2))
        c = c / torch.sqrt(torch.einsum("a,ab,b", c, pair, c))
        self.coeffs = c
        self.normalized = True
        return self


@dataclass
class AtomCGTOBasis:
    atomz: ZType
    bases: List[CGTOBasis]
    pos: torch.Tensor  # (3,)


@dataclass
class DensityFitInfo:
    method: str
    auxbases: List[AtomCGTOBasis]


@dataclass
class SpinParam(Generic[T]):
    """Pair of spin-up / spin-down values (reference :78-132)."""
    u: T
    d: T

    def sum(a):
        return a.u + a.d if isinstance(a, SpinParam) else a

    def reduce(a, fcn: Callable):
        return fcn(a.u, a.d) if isinstance(a, SpinParam) else a

    @staticmethod
    def apply_fcn(fcn: Callable, *a):
        assert len(a) > 0
        if isinstance(a[0], SpinParam):
            return SpinParam(u=fcn(*[x.u for x in a]), d=fcn(*[x.d for x in a]))
        return fcn(*a)


@dataclass
class ValGrad:
    """Local density / potential bundle: value (*BD, nr), grad (*BD, 3, nr), lapl, kin
    (reference :134-184)."""
    value: torch.Tensor
    grad: Optional[torch.Tensor] = None
    lapl: Optional[torch.Tensor] = None
    kin: Optional[torch.Tensor] = None

    def __add__(self, b: "ValGrad") -> "ValGrad":
        opt = lambda x, y: x + y if x is not None else None
        return ValGrad(self.value + b.value, opt(self.grad, b.grad), opt(self.lapl, b.lapl),
                       opt(self.kin, b.kin))

    def __mul__(self, f) -> "ValGrad":
        if isinstance(f, torch.Tensor):
            assert f.numel() == 1, "ValGrad multiplication with tensor can only be done with 1-element tensor"
        opt = lambda x: x * f if x is not None else None
        return ValGrad(self.value * f, opt(self.grad), opt(self.lapl), opt(self.kin))
