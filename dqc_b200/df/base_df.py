"""Density-fitting interface, same members as dqc/df/base_df.py:7-46."""
from abc import abstractmethod, abstractproperty
from typing import List
import torch
from dqc_b200.utils.linop import EditableModule, LinearOperator

__all__ = ["BaseDF"]


class BaseDF(EditableModule):
    @abstractmethod
    def build(self) -> "BaseDF":
        pass

    @abstractmethod
    def get_elrep(self, dm: torch.Tensor) -> LinearOperator:
        pass

    @abstractproperty
    def j2c(self) -> torch.Tensor:
        pass

    @abstractproperty
    def j3c(self) -> torch.Tensor:
        pass

    @abstractmethod
    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        pass
