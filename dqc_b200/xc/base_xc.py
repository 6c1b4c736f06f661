"""XC functional interface -- same surface as dqc/xc/base_xc.py:8-268: ``family`` (1 LDA, 2 GGA,
4 meta-GGA), ``get_edensityxc(densinfo)`` -> energy per unit volume (*BD, nr), ``get_vxc(densinfo)``
-> ValGrad (or SpinParam[ValGrad]) with ``value = de/drho`` and ``grad = de/d(grad rho)``, and the
``+`` / ``*`` algebra.  The default ``get_vxc`` differentiates ``get_edensityxc`` with autograd (the
route user-defined functionals take); the built-in functionals (dqc_b200/xc/b200xc.py) override it
with the fused CUDA kernel."""
from abc import abstractmethod, abstractproperty
from typing import List, Union
import torch
from dqc_b200.utils.linop import EditableModule
from dqc_b200.utils.datastruct import ValGrad, SpinParam

__all__ = ["BaseXC", "AddBaseXC", "MulBaseXC"]


def _leaves(densinfo, family):
    names = ["value"] + (["grad"] if family >= 2 else [])
    infos = [densinfo] if isinstance(densinfo, ValGrad) else [densinfo.u, densinfo.d]
    return infos, names


class BaseXC(EditableModule):
    @abstractproperty
    def family(self) -> int:
        pass

    @abstractmethod
    def get_edensityxc(self, densinfo: Union[ValGrad, SpinParam[ValGrad]]) -> torch.Tensor:
        pass

    def get_vxc(self, densinfo):
        if self.family not in (1, 2):
            raise NotImplementedError("Default vxc for family %d is not implemented" % self.family)
        infos, names = _leaves(densinfo, self.family)
        # differentiate w.r.t. fresh leaves holding the same numbers
        new_infos = []
        params = []
        for info in infos:
            kw = {}
            for nm in names:
                t = getattr(info, nm).detach().clone().requires_grad_(True)
                kw[nm] = t
                params.append(t)
            new_infos.append(ValGrad(**kw))
        dinfo = new_infos[0] if isinstance(densinfo, ValGrad) else SpinParam(u=new_infos[0], d=new_infos[1])
        with torch.enable_grad():
            edens = self.get_edensityxc(dinfo)
        grads = torch.autograd.grad(edens, params, grad_outputs=torch.ones_like(edens), allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, params)]
        outs = []
        k = len(names)
        for i in range(len(infos)):
            outs.append(ValGrad(**{nm: grads[i * k + j] for j, nm in enumerate(names)}))
        return outs[0] if isinstance(densinfo, ValGrad) else SpinParam(u=outs[0], d=outs[1])

    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        return []

    # ---- algebra ----
    def __add__(self, other):
        return AddBaseXC(self, other)

    def __mul__(self, other):
        return MulBaseXC(self, other)

    def __rmul__(self, other):
        return MulBaseXC(self, other)


class AddBaseXC(BaseXC):
    def __init__(self, a: BaseXC, b: BaseXC) -> None:
        self.a, self.b = a, b
        self._family = max(a.family, b.family)

    @property
    def family(self) -> int:
        return self._family

    def get_vxc(self, densinfo):
        av, bv = self.a.get_vxc(densinfo), self.b.get_vxc(densinfo)
        if isinstance(densinfo, ValGrad):
            return _add_potinfo(av, bv)
        return SpinParam(u=_add_potinfo(av.u, bv.u), d=_add_potinfo(av.d, bv.d))

    def get_edensityxc(self, densinfo):
        return self.a.get_edensityxc(densinfo) + self.b.get_edensityxc(densinfo)


def _add_potinfo(a: ValGrad, b: ValGrad) -> ValGrad:
    # an LDA term has no grad field: treat it as zero (ValGrad.__add__ would drop b's gradient)
    grad = a.grad if b.grad is None else (b.grad if a.grad is None else a.grad + b.grad)
    return ValGrad(value=a.value + b.value, grad=grad)


class MulBaseXC(BaseXC):
    def __init__(self, a: BaseXC, b) -> None:
        self.a, self.b = a, b

    @property
    def family(self) -> int:
        return self.a.family

    def get_vxc(self, densinfo):
        av = self.a.get_vxc(densinfo)
        if isinstance(densinfo, ValGrad):
            return av * self.b
        return SpinParam(u=av.u * self.b, d=av.d * self.b)

    def get_edensityxc(self, densinfo):
        return self.a.get_edensityxc(densinfo) * self.b
