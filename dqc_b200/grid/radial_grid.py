"""Radial quadratures: an integrator on [-1, 1] composed with an x -> r map.
Restates dqc/grid/radial_grid.py:10-196 (dV = 4 pi r^2 (dr/dx) w, :45-48; integrators :82-120;
DE2 :143-162, LogM3 :164-175, TreutlerM4 :177-196)."""
from __future__ import annotations
from typing import Tuple, Union
import numpy as np
import torch
from dqc_b200.grid.base_grid import BaseGrid

__all__ = ["RadialGrid", "SlicedRadialGrid", "DE2Transformation", "LogM3Transformation",
           "TreutlerM4Transformation"]


def get_xw_integration(n: int, kind: str) -> Tuple[np.ndarray, np.ndarray]:
    kind = kind.lower()
    if kind == "uniform":  # trapezoid on linspace(-1, 1, n)
        x = np.linspace(-1, 1, n)
        w = np.full(n, x[1] - x[0])
        w[0] *= 0.5
        w[-1] *= 0.5
        return x, w
    i = np.arange(n, 0, -1)
    th = i * np.pi / (n + 1.0)
    s = np.sin(th)
    if kind == "chebyshev2":  # Gauss-Chebyshev of the 2nd kind (A&S p. 889)
        return np.cos(th), np.pi / (n + 1.0) * s
    if kind == "chebyshev":  # Becke-style mapped Chebyshev, JCP 108, 3226 eq (9)-(10)
        x = (n + 1.0 - 2 * i) / (n + 1.0) + 2 / np.pi * (1 + 2.0 / 3 * s * s) * np.cos(th) * s
        return x, 16.0 / (3 * (n + 1.0)) * s ** 4
    raise RuntimeError("Unknown grid_integrator: %s. Available: %s" % (kind, ["chebyshev", "chebyshev2", "uniform"]))


class BaseGridTransform(object):
    def x2r(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def get_drdx(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class DE2Transformation(BaseGridTransform):
    """r = exp(alpha x' - exp(-x')), x' affine in x (Mitani & Yoshioka, TCA 131, 1169 eq. 31)."""

    def __init__(self, alpha: float = 1.0, rmin: float = 1e-7, rmax: float = 20):
        assert rmin < 1.0
        self.alpha = alpha
        self.xmin = -np.log(-np.log(rmin))
        self.xmax = np.log(rmax) / alpha

    def _xnew(self, x):
        return 0.5 * (x * (self.xmax - self.xmin) + (self.xmax + self.xmin))

    def x2r(self, x):
        xn = self._xnew(x)
        return torch.exp(self.alpha * xn - torch.exp(-xn))

    def get_drdx(self, x):
        xn = self._xnew(x)
        return self.x2r(x) * (self.alpha + torch.exp(-xn)) * (0.5 * (self.xmax - self.xmin))


class LogM3Transformation(BaseGridTransform):
    def __init__(self, ra: float = 1.0, eps: float = 1e-15):
        self.ra, self.eps, self.ln2 = ra, eps, np.log(2.0 + eps)

    def x2r(self, x):
        return self.ra * (1 - torch.log1p(-x + self.eps) / self.ln2)

    def get_drdx(self, x):
        return self.ra / self.ln2 / (1 - x + self.eps)


class TreutlerM4Transformation(BaseGridTransform):
    def __init__(self, xi: float = 1.0, alpha: float = 0.6, eps: float = 1e-15):
        self._xi, self._alpha, self._eps, self._ln2 = xi, alpha, eps, np.log(2.0 + eps)

    def x2r(self, x):
        a = 1.0 + self._eps
        return self._xi / self._ln2 * (a + x) ** self._alpha * (self._ln2 - torch.log1p(-x + self._eps))

    def get_drdx(self, x):
        a = 1.0 + self._eps
        fac = self._xi / self._ln2 * (a + x) ** self._alpha
        return fac * self._alpha / (a + x) * (self._ln2 - torch.log1p(-x + self._eps)) + fac / (1 - x + self._eps)


def get_grid_transform(s: Union[str, BaseGridTransform]) -> BaseGridTransform:
    if isinstance(s, BaseGridTransform):
        return s
    table = {"logm3": LogM3Transformation, "de2": DE2Transformation, "treutlerm4": TreutlerM4Transformation}
    if s.lower() not in table:
        raise RuntimeError("Unknown grid transformation: %s" % s)
    return table[s.lower()]()


class RadialGrid(BaseGrid):
    def __init__(self, ngrid: int, grid_integrator: str = "chebyshev",
                 grid_transform: Union[str, BaseGridTransform] = "logm3",
                 dtype: torch.dtype = torch.float64, device: torch.device = torch.device("cpu")):
        self._dtype, self._device = dtype, device
        tf = get_grid_transform(grid_transform)
        x_np, w_np = get_xw_integration(ngrid, grid_integrator)
        x = torch.as_tensor(x_np, dtype=dtype, device=device)
        w = torch.as_tensor(w_np, dtype=dtype, device=device)
        r = tf.x2r(x)
        self.rgrid = r.unsqueeze(-1)
        self.dvolume = (4 * np.pi * r * r) * (tf.get_drdx(x) * w)

    @property
    def coord_type(self):
        return "radial"

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    def get_dvolume(self) -> torch.Tensor:
        return self.dvolume

    def get_rgrid(self) -> torch.Tensor:
        return self.rgrid

    def __getitem__(self, key) -> "RadialGrid":
        if isinstance(key, slice):
            return SlicedRadialGrid(self, key)
        raise KeyError("Indexing for RadialGrid is not defined")

    def getparamnames(self, methodname: str, prefix: str = ""):
        if methodname == "get_dvolume":
            return [prefix + "dvolume"]
        if methodname == "get_rgrid":
            return [prefix + "rgrid"]
        raise KeyError("getparamnames for %s is not set" % methodname)


class SlicedRadialGrid(RadialGrid):
    def __init__(self, obj: RadialGrid, key: slice):
        self._dtype, self._device = obj.dtype, obj.device
        self.dvolume = obj.dvolume[key]
        self.rgrid = obj.rgrid[key]
