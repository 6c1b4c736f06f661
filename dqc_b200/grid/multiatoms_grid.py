"""Multi-centre (Becke) molecular grid.

Semantics restated from dqc/grid/multiatoms_grid.py:8-57,158-273: points are the atomic grids
shifted to their atoms and concatenated in atom order; the weight of a point of atom ``a`` is
``P_a / sum_k P_k`` with ``P_j = prod_{i != j} s(mu_ij)``, ``mu_ij = (r_j - r_i)/R_ij`` plus the
hetero-nuclear shift ``a_ij (1 - mu^2)`` (``a_ij = clamp(u/(u^2-1), +-0.45)``,
``u = (rad_j - rad_i)/(rad_j + rad_i)``, rad = radii or sqrt(radii) for Treutler), three
iterations of ``f <- f (3 - f^2)/2``, ``s = (1 + 1e-12 - f)/2``, and cells with any
``mu_ij >= 0.74`` dropped (:231-234).

The O(natoms^2 ngrid) weight evaluation is the ``becke_weights`` CUDA kernel (one thread per grid
point, atoms staged in shared memory); the reference does it as a per-atom torch loop on CPU.
"""
from typing import List, Optional
import torch
from dqc_b200.grid.base_grid import BaseGrid

__all__ = ["BeckeGrid"]


class BeckeGrid(BaseGrid):
    def __init__(self, atomgrid: Optional[List[BaseGrid]], atompos: torch.Tensor,
                 atomradii: Optional[torch.Tensor] = None, ratom_adjust: str = "becke", prebuilt=None) -> None:
        """prebuilt = (xyz, dvol_atoms, owner, counts): the points already assembled on the device
        (b200qc_grid_assemble, dqc_b200/grid/factory.py); atomgrid is then not used."""
        if ratom_adjust not in ("becke", "treutler"):
            raise ValueError("Unknown atom adjustment: %s. Available: ['becke', 'treutler']" % ratom_adjust)
        self._device = atompos.device
        dev = self._device
        owner_pre = None
        if prebuilt is not None:
            self._rgrid, dvol_atoms, owner_pre, counts = prebuilt
            self._dtype = self._rgrid.dtype
        else:
            assert atompos.shape[0] == len(atomgrid), "The lengths of atomgrid and atompos must be the same"
            assert len(atomgrid) > 0
            self._dtype = atomgrid[0].dtype
            pts = [gr.get_rgrid().to(dev) + pos for gr, pos in zip(atomgrid, atompos)]
            self._rgrid = torch.cat(pts, dim=0).contiguous()
            dvol_atoms = torch.cat([gr.get_dvolume().to(dev) for gr in atomgrid], dim=0)
            counts = [p.shape[0] for p in pts]
        self._atom_ngrids = counts

        natoms = atompos.shape[0]
        if natoms == 1:
            # P_a / P_a == 1 exactly; nothing to evaluate
            w = torch.ones_like(dvol_atoms)
        else:
            from dqc_b200 import _lib
            if atomradii is not None:
                rad = atomradii.to(dev) if ratom_adjust == "becke" else atomradii.to(dev) ** 0.5
                uij = (rad - rad.unsqueeze(1)) / (rad + rad.unsqueeze(1))  # [i, j] = (r_j - r_i)/(r_j + r_i)
                aij = torch.clamp(uij / (uij * uij - 1), min=-0.45, max=0.45).contiguous()
            else:
                aij = None
            owner = owner_pre if owner_pre is not None else torch.repeat_interleave(
                torch.arange(natoms, device=dev, dtype=torch.int32), torch.tensor(counts, device=dev))
            w = _lib.becke_weights(self._rgrid, owner, atompos.contiguous(), aij)
        self._dvolume = dvol_atoms * w

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    @property
    def coord_type(self):
        return "cart"

    def get_dvolume(self) -> torch.Tensor:
        return self._dvolume

    def get_rgrid(self) -> torch.Tensor:
        return self._rgrid

    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        if methodname == "get_rgrid":
            return [prefix + "_rgrid"]
        if methodname == "get_dvolume":
            return [prefix + "_dvolume"]
        raise KeyError("Invalid methodname: %s" % methodname)
