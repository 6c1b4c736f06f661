"""Hartree-Fock on the B200 Fock-build path.  Engine semantics restated from dqc/qccalc/hf.py:42-316:
F = (T + V) + J[D_total] + K'[D] with K' = -1/2 K (restricted) or -1/2 K[2 D_s] per spin;
F -> lowest ``norb`` eigenvectors (exact eigendecomposition, cuSOLVER through torch) -> D = C w C^T."""
from typing import Any, Dict, Optional, Union
import torch
from dqc_b200.qccalc.scf_qccalc import SCF_QCCalc, BaseSCFEngine
from dqc_b200.utils.datastruct import SpinParam

__all__ = ["HF"]


class HF(SCF_QCCalc):
    def __init__(self, system, restricted: Optional[bool] = None, variational: bool = False):
        super().__init__(_HFEngine(system, restricted), variational)


class _HFEngine(BaseSCFEngine):
    def __init__(self, system, restricted: Optional[bool] = None, build_grid_if_necessary: bool = True):
        self._polarized = (bool(system.spin != 0) if restricted is None else not restricted)
        self._system = system
        self.hamilton = system.get_hamiltonian()
        if build_grid_if_necessary and system.requires_grid():
            system.setup_grid()
            self.hamilton.setup_grid(system.get_grid())
        self.hamilton.build()
        self.orb_weight = system.get_orbweight(polarized=self._polarized)
        self.norb = SpinParam.apply_fcn(lambda w: int(w.shape[-1]), self.orb_weight)
        self.knvext_linop = self.hamilton.get_kinnucl()
        self.eigen_options: Dict[str, Any] = {}

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

    def set_eigen_options(self, eigen_options: Dict[str, Any]) -> None:
        self.eigen_options = eigen_options

    # ---- scp <-> dm ----
    def dm2scp(self, dm) -> torch.Tensor:
        fock = self._dm2fock(dm)
        if isinstance(dm, torch.Tensor):
            return fock.fullmatrix()
        return torch.cat((fock.u.fullmatrix().unsqueeze(0), fock.d.fullmatrix().unsqueeze(0)), dim=0)

    def scp2dm(self, scp: torch.Tensor):
        if not self._polarized:
            return self._fock2dm(scp, self.orb_weight, self.norb)
        return SpinParam(u=self._fock2dm(scp[0], self.orb_weight.u, self.norb.u),
                         d=self._fock2dm(scp[1], self.orb_weight.d, self.norb.d))

    def scp2scp(self, scp: torch.Tensor) -> torch.Tensor:
        return self.dm2scp(self.scp2dm(scp))

    def dm2energy(self, dm) -> torch.Tensor:
        dmtot = SpinParam.sum(dm)
        e_core = self.hamilton.get_e_hcore(dmtot)
        e_elrep = self.hamilton.get_e_elrep(dmtot)
        e_exch = self.hamilton.get_e_exchange(dm)
        return e_core + e_elrep + e_exch + self._system.get_nuclei_energy()

    # ---- pieces ----
    def _dm2fock(self, dm):
        v2e = self.hamilton.get_fock_2e(dm, exx=1.0, with_xc=False)
        return SpinParam.apply_fcn(lambda v: self.knvext_linop + v, v2e)

    def _fock2dm(self, fock: torch.Tensor, orb_weight: torch.Tensor, norb: int) -> torch.Tensor:
        return self.hamilton.ao_orb2dm(self.diagonalize(fock, norb)[1], orb_weight)

    def diagonalize(self, fock: torch.Tensor, norb: int):
        """Lowest `norb` eigenpairs of the (symmetrised) Fock matrix; the overlap is the identity in
        the orthogonalised basis, otherwise F C = S C e is reduced with a Cholesky factor."""
        fock = (fock + fock.transpose(-2, -1)) * 0.5
        ovlp = self.hamilton.get_overlap().fullmatrix()
        eye = torch.eye(ovlp.shape[-1], dtype=ovlp.dtype, device=ovlp.device)
        if torch.allclose(ovlp, eye, atol=1e-10):
            ev, c = torch.linalg.eigh(fock)
        else:
            L = torch.linalg.cholesky(ovlp)
            Linv = torch.linalg.inv(L)
            ev, cp = torch.linalg.eigh(Linv @ fock @ Linv.transpose(-2, -1))
            c = Linv.transpose(-2, -1) @ cp
        return ev[..., :norb], c[..., :norb]

    def getparamnames(self, methodname: str, prefix: str = ""):
        return []
