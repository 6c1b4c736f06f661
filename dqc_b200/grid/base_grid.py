"""Grid interface, same members as dqc/grid/base_grid.py:6-53."""
from abc import abstractmethod, abstractproperty
from typing import List
import torch


class BaseGrid(object):
    @abstractproperty
    def dtype(self) -> torch.dtype:
        pass

    @abstractproperty
    def device(self) -> torch.device:
        pass

    @abstractproperty
    def coord_type(self) -> str:
        """"cart" for (x, y, z) rows or "radial" for a single r column."""
        pass

    @abstractmethod
    def get_dvolume(self) -> torch.Tensor:
        """Integration weights, shape (ngrid,)."""
        pass

    @abstractmethod
    def get_rgrid(self) -> torch.Tensor:
        """Grid positions, shape (ngrid, ndim)."""
        pass

    @abstractmethod
    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        pass
