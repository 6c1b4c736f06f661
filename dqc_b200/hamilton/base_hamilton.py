"""Operator surface of the Hamiltonian -- the drop-in boundary.  Same members, argument meaning and
return conventions as dqc/hamilton/base_hamilton.py:11-279 (operators are Hermitian LinearOperators
in the orthogonalised basis; batch dimensions broadcast)."""
from __future__ import annotations
from abc import abstractmethod, abstractproperty
from typing import List, Optional, Tuple, Union
import torch
from dqc_b200.utils.linop import EditableModule, LinearOperator
from dqc_b200.utils.datastruct import SpinParam
from dqc_b200.grid.base_grid import BaseGrid

__all__ = ["BaseHamilton"]


class BaseHamilton(EditableModule):
    # ---- properties ----
    @abstractproperty
    def nao(self) -> int:
        """Number of (orthogonalised) atomic-orbital basis functions."""

    @abstractproperty
    def kpts(self) -> torch.Tensor:
        """k-points (nkpts, ndim); TypeError for an isolated molecule."""

    @abstractproperty
    def df(self):
        """The density-fitting object attached to this Hamiltonian, or None."""

    # ---- setups ----
    @abstractmethod
    def build(self) -> "BaseHamilton":
        """Construct the one-off ingredients (integrals, fitting tensors)."""

    @abstractmethod
    def setup_grid(self, grid: BaseGrid, xc=None) -> None:
        """Evaluate the AOs (and the derivatives the xc family needs) on the grid."""

    # ---- Fock components ----
    @abstractmethod
    def get_nuclattr(self) -> LinearOperator:
        pass

    @abstractmethod
    def get_kinnucl(self) -> LinearOperator:
        pass

    @abstractmethod
    def get_overlap(self) -> LinearOperator:
        pass

    @abstractmethod
    def get_elrep(self, dm: torch.Tensor) -> LinearOperator:
        """Coulomb operator of dm (*BD, nao, nao)."""

    @abstractmethod
    def get_exchange(self, dm: Union[torch.Tensor, SpinParam[torch.Tensor]]):
        """Exact-exchange operator (-1/2 K for a restricted density; per spin K[2 D_s] for SpinParam)."""

    @abstractmethod
    def get_vext(self, vext: torch.Tensor) -> LinearOperator:
        """Operator of an external potential sampled on the grid (*BR, ngrid)."""

    @abstractmethod
    def get_vxc(self, dm: Union[torch.Tensor, SpinParam[torch.Tensor]]):
        """Exchange-correlation potential operator (SpinParam in -> SpinParam out)."""

    # ---- density-matrix interface ----
    @abstractmethod
    def ao_orb2dm(self, orb: torch.Tensor, orb_weight: torch.Tensor) -> torch.Tensor:
        """orb (*BO, nao, norb), orb_weight (*BW, norb) -> dm (*BOW, nao, nao)."""

    @abstractmethod
    def aodm2dens(self, dm: torch.Tensor, xyz: torch.Tensor) -> torch.Tensor:
        """Density of dm (*BD, nao, nao) at xyz (*BR, ndim) -> (*BRD)."""

    # ---- energies ----
    @abstractmethod
    def get_e_hcore(self, dm: torch.Tensor) -> torch.Tensor:
        pass

    @abstractmethod
    def get_e_elrep(self, dm: torch.Tensor) -> torch.Tensor:
        pass

    @abstractmethod
    def get_e_exchange(self, dm: Union[torch.Tensor, SpinParam[torch.Tensor]]) -> torch.Tensor:
        pass

    @abstractmethod
    def get_e_xc(self, dm: Union[torch.Tensor, SpinParam[torch.Tensor]]) -> torch.Tensor:
        pass

    # ---- variational-SCF parametrisation (outside the Fock-build path) ----
    def ao_orb_params2dm(self, ao_orb_params, ao_orb_coeffs, orb_weight, with_penalty=None):
        raise NotImplementedError("orbital parametrisation belongs to the variational SCF mode, "
                                  "which is outside the B200 Fock-build path (DESIGN.md, out of scope)")

    def dm2ao_orb_params(self, dm: torch.Tensor, norb: int) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError("orbital parametrisation belongs to the variational SCF mode, "
                                  "which is outside the B200 Fock-build path (DESIGN.md, out of scope)")

    @abstractmethod
    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        pass
