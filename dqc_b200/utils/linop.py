"""Minimal stand-in for the slice of ``xitorch`` the Fock-build callers use (xitorch is an
un-vendored dependency of the reference and absent here): a matrix-backed ``LinearOperator`` with
``LinearOperator.m(mat, is_hermitian)``, ``+``, ``.fullmatrix()``, ``.mm/.mv``, shape/dtype/device
(call sites: dqc/hamilton/hcgto.py:192-250, dqc/qccalc/hf.py:97-103,184,200, ks.py:178-182), and
``EditableModule`` (only ``getparamnames`` plumbing is needed on the forward path)."""
from typing import List, Sequence
import torch

__all__ = ["LinearOperator", "EditableModule"]


class EditableModule(object):
    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        return []


class LinearOperator(EditableModule):
    def __init__(self, mat: torch.Tensor, is_hermitian: bool = False):
        self._mat = mat
        self.is_hermitian = is_hermitian

    @classmethod
    def m(cls, mat: torch.Tensor, is_hermitian=None) -> "LinearOperator":
        if is_hermitian is None:
            is_hermitian = False
        return cls(mat, bool(is_hermitian))

    @property
    def shape(self) -> Sequence[int]:
        return self._mat.shape

    @property
    def dtype(self) -> torch.dtype:
        return self._mat.dtype

    @property
    def device(self) -> torch.device:
        return self._mat.device

    def fullmatrix(self) -> torch.Tensor:
        return self._mat

    def mm(self, x: torch.Tensor) -> torch.Tensor:
        return torch.matmul(self._mat, x)

    def mv(self, x: torch.Tensor) -> torch.Tensor:
        return torch.matmul(self._mat, x.unsqueeze(-1)).squeeze(-1)

    def __add__(self, other: "LinearOperator") -> "LinearOperator":
        return LinearOperator(self._mat + other.fullmatrix(), self.is_hermitian and other.is_hermitian)

    def __sub__(self, other: "LinearOperator") -> "LinearOperator":
        return LinearOperator(self._mat - other.fullmatrix(), self.is_hermitian and other.is_hermitian)

    def __mul__(self, f) -> "LinearOperator":
        return LinearOperator(self._mat * f, self.is_hermitian)

    __rmul__ = __mul__

    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        return [prefix + "_mat"]
