"""System interface (subset of dqc/system/base_system.py:10-139 the SCF engines call)."""
from abc import abstractmethod
from typing import List, Union
import torch
from dqc_b200.utils.linop import EditableModule
from dqc_b200.utils.datastruct import SpinParam, ZType

__all__ = ["BaseSystem"]


class BaseSystem(EditableModule):
    @abstractmethod
    def densityfit(self, method=None, auxbasis=None) -> "BaseSystem":
        pass

    @abstractmethod
    def get_hamiltonian(self):
        pass

    @abstractmethod
    def get_orbweight(self, polarized: bool = False) -> Union[torch.Tensor, SpinParam[torch.Tensor]]:
        pass

    @abstractmethod
    def get_nuclei_energy(self) -> torch.Tensor:
        pass

    @abstractmethod
    def setup_grid(self) -> None:
        pass

    @abstractmethod
    def get_grid(self):
        pass

    @abstractmethod
    def requires_grid(self) -> bool:
        pass
