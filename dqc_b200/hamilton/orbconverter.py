"""Basis orthogonalisation, restating dqc/hamilton/orbconverter.py:67-163:
X = U_{lambda > 1e-6} lambda^{-1/2} from the overlap eigen-decomposition (cuSOLVER syevd through
torch.linalg.eigh on the device); operators go in as X^T M X, densities come back as X D X^T."""
from typing import List
import torch

__all__ = ["OrbitalOrthogonalizer", "IdentityOrbConverter"]


class OrbitalOrthogonalizer(object):
    def __init__(self, ovlp: torch.Tensor, threshold: float = 1e-6):
        ovlp_eival, ovlp_eivec = torch.linalg.eigh(ovlp)
        acc_idx = ovlp_eival > threshold
        orthozer = ovlp_eivec[..., acc_idx] * (ovlp_eival[acc_idx]) ** (-0.5)  # (nao, nao2)
        self._orthozer = orthozer.contiguous()

    def nao(self) -> int:
        return self._orthozer.shape[-1]

    def convert2(self, mat: torch.Tensor) -> torch.Tensor:
        return torch.matmul(self._orthozer.transpose(-2, -1), torch.matmul(mat, self._orthozer))

    def convert4(self, mat: torch.Tensor) -> torch.Tensor:
        # index by index: same numbers as the reference's 5-operand einsum, O(nao^5) instead of O(nao^8)
        x = self._orthozer
        mat = torch.einsum("ijkl,im->mjkl", mat, x)
        mat = torch.einsum("mjkl,jn->mnkl", mat, x)
        mat = torch.einsum("mnkl,kp->mnpl", mat, x)
        return torch.einsum("mnpl,lq->mnpq", mat, x)

    def convert_ortho_orb(self, orb: torch.Tensor) -> torch.Tensor:
        return torch.matmul(self._orthozer, orb)

    def unconvert_dm(self, dm: torch.Tensor) -> torch.Tensor:
        return torch.matmul(torch.matmul(self._orthozer, dm), self._orthozer.transpose(-2, -1))

    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        return [prefix + "_orthozer"]


class IdentityOrbConverter(object):
    def __init__(self, ovlp: torch.Tensor):
        self._nao = ovlp.shape[-1]

    def nao(self) -> int:
        return self._nao

    def convert2(self, mat: torch.Tensor) -> torch.Tensor:
        return mat

    def convert4(self, mat: torch.Tensor) -> torch.Tensor:
        return mat

    def convert_ortho_orb(self, orb: torch.Tensor) -> torch.Tensor:
        return orb

    def unconvert_dm(self, dm: torch.Tensor) -> torch.Tensor:
        return dm

    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        return []
