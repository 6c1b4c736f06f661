"""Built-in LDA / GGA functionals evaluated by the K3 CUDA kernel (b200qc_xc_unpol / _pol).

Plays the role of the reference's libxc bridge (dqc/xc/libxc.py:24-242 + libxc_wrapper.py:380-413):
same inputs (rho, grad rho; SpinParam for polarised), same outputs (energy per unit VOLUME;
potential ``value = de/drho``, ``grad = 2 (de/dsigma) grad rho``, polarised
``grad_u = 2 vs_uu grad_u + vs_ud grad_d``).  A sum of scaled functionals is ONE object holding the
term list, so ``"gga_x_pbe + gga_c_pbe"`` is a single kernel launch."""
from typing import List, Tuple, Union
import torch
from dqc_b200 import _lib
from dqc_b200.xc.base_xc import BaseXC, AddBaseXC, MulBaseXC
from dqc_b200.utils.datastruct import ValGrad, SpinParam

__all__ = ["B200XC", "get_b200xc"]


class B200XC(BaseXC):
    def __init__(self, terms: List[Tuple[float, str]]):
        for _, name in terms:
            if name not in _lib.FUNC_IDS:
                raise NotImplementedError(
                    "functional '%s' is not built into the B200 path; available: %s (user functionals: "
                    "subclass BaseXC, they run through autograd)" % (name, sorted(_lib.FUNC_IDS)))
        self.terms = [(float(c), n) for c, n in terms]
        self._family = max(_lib.FUNC_FAMILY[n] for _, n in terms)

    @property
    def family(self) -> int:
        return self._family

    # ---- kernel plumbing: flatten the batch dimensions, one launch per batch entry ----
    def _run(self, densinfo, want_e: bool, want_v: bool):
        gga = self._family == 2
        if isinstance(densinfo, ValGrad):
            rho = densinfo.value
            bshape, n = rho.shape[:-1], rho.shape[-1]
            rho2 = rho.reshape(-1, n).contiguous()
            grad2 = densinfo.grad.reshape(-1, 3, n).contiguous() if gga else None
            es, vrs, vgs = [], [], []
            for b in range(rho2.shape[0]):
                e, vr, vg = _lib.xc_unpol(self.terms, rho2[b], grad2[b] if gga else None, want_e, want_v)
                es.append(e); vrs.append(vr); vgs.append(vg)
            e = torch.stack(es).reshape(*bshape, n) if want_e else None
            if not want_v:
                return e, None
            vr = torch.stack(vrs).reshape(*bshape, n)
            vg = torch.stack(vgs).reshape(*bshape, 3, n) if gga else None
            return e, ValGrad(value=vr, grad=vg)
        ru, rd = densinfo.u.value, densinfo.d.value
        bshape, n = ru.shape[:-1], ru.shape[-1]
        rho2 = torch.stack([ru.reshape(-1, n), rd.reshape(-1, n)], dim=1).contiguous()       # (B, 2, n)
        grad2 = torch.stack([densinfo.u.grad.reshape(-1, 3, n), densinfo.d.grad.reshape(-1, 3, n)],
                            dim=1).contiguous() if gga else None                              # (B, 2, 3, n)
        es, vrs, vgs = [], [], []
        for b in range(rho2.shape[0]):
            e, vr, vg = _lib.xc_pol(self.terms, rho2[b], grad2[b] if gga else None, want_e, want_v)
            es.append(e); vrs.append(vr); vgs.append(vg)
        e = torch.stack(es).reshape(*bshape, n) if want_e else None
        if not want_v:
            return e, None
        vr = torch.stack(vrs)                       # (B, 2, n)
        vg = torch.stack(vgs) if gga else None      # (B, 2, 3, n)
        mk = lambda s: ValGrad(value=vr[:, s].reshape(*bshape, n),
                               grad=vg[:, s].reshape(*bshape, 3, n) if gga else None)
        return e, SpinParam(u=mk(0), d=mk(1))

    def get_edensityxc(self, densinfo: Union[ValGrad, SpinParam[ValGrad]]) -> torch.Tensor:
        return self._run(densinfo, True, False)[0]

    def get_vxc(self, densinfo):
        return self._run(densinfo, False, True)[1]

    def get_edens_vxc(self, densinfo):
        """Both in one launch (the reference calls libxc twice)."""
        return self._run(densinfo, True, True)

    # ---- algebra keeps everything in one term list ----
    def __add__(self, other):
        if isinstance(other, B200XC):
            return B200XC(self.terms + other.terms)
        return AddBaseXC(self, other)

    def __mul__(self, other):
        if isinstance(other, (int, float)):
            return B200XC([(c * float(other), n) for c, n in self.terms])
        return MulBaseXC(self, other)

    __rmul__ = __mul__

    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        return []


def get_b200xc(name: str) -> B200XC:
    return B200XC([(1.0, name.lower())])
