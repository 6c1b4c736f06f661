"""Radial x Lebedev product grids (reference: dqc/grid/lebedev_grid.py:11-102).  Angular tables come
from dqc_b200/data/lebedev.npz (see tools/make_lebedev_table.py): rows (phi, theta, w), sum w = 1,
the 4 pi lives in the radial weight; the product is radial-major."""
import os
from typing import Dict, List, Sequence
import numpy as np
import torch
from dqc_b200.grid.base_grid import BaseGrid
from dqc_b200.grid.radial_grid import RadialGrid

__all__ = ["LebedevGrid", "TruncatedLebedevGrid"]

_TABLE: Dict[str, np.ndarray] = {}


def load_lebedev(prec: int) -> np.ndarray:
    if not _TABLE:
        path = os.path.join(os.path.dirname(os.path.realpath(__file__)), "..", "data", "lebedev.npz")
        with np.load(path) as z:
            _TABLE.update({k: z[k] for k in z.files})
    key = "p%03d" % prec
    assert key in _TABLE, "The Lebedev table of precision %d does not exist" % prec
    return _TABLE[key]


class LebedevGrid(BaseGrid):
    def __init__(self, radgrid: RadialGrid, prec: int) -> None:
        self._dtype, self._device = radgrid.dtype, radgrid.device
        assert (prec % 2 == 1) and (3 <= prec <= 131), "Precision must be an odd number between 3 and 131"
        tab = torch.tensor(load_lebedev(prec), dtype=self._dtype, device=self._device)
        phi, theta, wang = tab[:, 0], tab[:, 1], tab[:, 2]
        assert radgrid.coord_type == "radial"
        r = radgrid.get_rgrid().unsqueeze(-1)  # (nr, 1, 1) -> broadcast (nr, 1)
        r = r.reshape(-1, 1)
        rs = r * torch.sin(theta)
        x = (rs * torch.cos(phi)).reshape(-1, 1)
        y = (rs * torch.sin(phi)).reshape(-1, 1)
        z = (r * torch.cos(theta)).reshape(-1, 1)
        self._xyz = torch.cat((x, y, z), dim=-1)
        self._dvolume = (radgrid.get_dvolume().unsqueeze(-1) * wang).reshape(-1)

    def get_rgrid(self) -> torch.Tensor:
        return self._xyz

    def get_dvolume(self) -> torch.Tensor:
        return self._dvolume

    @property
    def coord_type(self) -> str:
        return "cart"

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        if methodname == "get_rgrid":
            return [prefix + "_xyz"]
        if methodname == "get_dvolume":
            return [prefix + "_dvolume"]
        raise KeyError("Invalid methodname: %s" % methodname)


class TruncatedLebedevGrid(LebedevGrid):
    """Pruned atomic grid: consecutive radial slices, each with its own angular order."""

    def __init__(self, radgrids: Sequence[RadialGrid], precs: Sequence[int]):
        assert len(radgrids) == len(precs) and len(precs) > 0
        self.lebedevs = [LebedevGrid(rg, p) for rg, p in zip(radgrids, precs)]
        self._dtype, self._device = self.lebedevs[0].dtype, self.lebedevs[0].device
        self._xyz = torch.cat([g.get_rgrid() for g in self.lebedevs], dim=0)
        self._dvolume = torch.cat([g.get_dvolume() for g in self.lebedevs], dim=0)
