"""Kohn-Sham DFT on the B200 Fock-build path.  Engine semantics restated from dqc/qccalc/ks.py:45-238:
F = (T + V + vext) + J[D_total] + Vxc[D]; E = E_core + E_J + E_xc + E_nn.  ``exx_fraction`` adds
a * K' for a hybrid composition (F = h + J + a K' + Vxc[semi-local]), which the reference itself
cannot express (ks.py:176-187, getxc.py:29-36; SURVEY section 8a notes) -- default 0 = the
reference's behaviour."""
from typing import Optional, Union
import torch
from dqc_b200.api.getxc import get_xc
from dqc_b200.qccalc.scf_qccalc import SCF_QCCalc, BaseSCFEngine
from dqc_b200.qccalc.hf import _HFEngine
from dqc_b200.utils.datastruct import SpinParam
from dqc_b200.xc.base_xc import BaseXC

__all__ = ["KS"]


class KS(SCF_QCCalc):
    def __init__(self, system, xc: Union[str, BaseXC, None], restricted: Optional[bool] = None,
                 variational: bool = False, exx_fraction: float = 0.0):
        super().__init__(_KSEngine(system, xc, restricted, exx_fraction), variational)


class _KSEngine(BaseSCFEngine):
    def __init__(self, system, xc, restricted: Optional[bool] = None, exx_fraction: float = 0.0):
        self.xc: Optional[BaseXC] = get_xc(xc) if isinstance(xc, str) else xc
        self._system = system
        self._exx = float(exx_fraction)
        self.hamilton = system.get_hamiltonian()
        if self.xc is not None or system.requires_grid():
            system.setup_grid()
            self.hamilton.setup_grid(system.get_grid(), self.xc)
        self.hf_engine = _HFEngine(system, restricted=restricted, build_grid_if_necessary=False)
        self._polarized = self.hf_engine.polarized
        self.orb_weight = system.get_orbweight(polarized=self._polarized)
        self.norb = SpinParam.apply_fcn(lambda w: int(w.shape[-1]), self.orb_weight)
        self.knvext_linop = self.hamilton.get_kinnucl()

    def get_system(self):
        return self._system

    @property
    def shape(self):
        return self.knvext_linop.shape

    @property
    def dtype(self):
        return self.knvext_linop.dtype

    @property
    def device(self):
        return self.knvext_linop.device

    @property
    def polarized(self):
        return self._polarized

    def set_eigen_options(self, eigen_options) -> None:
        self.hf_engine.set_eigen_options(eigen_options)

    def dm2scp(self, dm) -> torch.Tensor:
        fock = self._dm2fock(dm)
        if isinstance(dm, torch.Tensor):
            return fock.fullmatrix()
        return torch.cat((fock.u.fullmatrix().unsqueeze(0), fock.d.fullmatrix().unsqueeze(0)), dim=0)

    def scp2dm(self, scp: torch.Tensor):
        return self.hf_engine.scp2dm(scp)

    def scp2scp(self, scp: torch.Tensor) -> torch.Tensor:
        return self.dm2scp(self.scp2dm(scp))

    def dm2energy(self, dm) -> torch.Tensor:
        dmtot = SpinParam.sum(dm)
        e = self.hamilton.get_e_hcore(dmtot) + self.hamilton.get_e_elrep(dmtot)
        if self.xc is not None:
            e = e + self.hamilton.get_e_xc(dm)
        if self._exx != 0.0:
            e = e + self._exx * self.hamilton.get_e_exchange(dm)
        return e + self._system.get_nuclei_energy()

    def _dm2fock(self, dm):
        v = self.hamilton.get_fock_2e(dm, exx=self._exx, with_xc=self.xc is not None)
        return SpinParam.apply_fcn(lambda vi: self.knvext_linop + vi, v)

    def getparamnames(self, methodname: str, prefix: str = ""):
        return []
