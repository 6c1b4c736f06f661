"""Basis packer: turns a list of AtomCGTOBasis into the libcint ``(atm, bas, env)`` triplet.

Same layout and public surface as the reference's LibcintWrapper
(dqc/hamilton/intor/lcintwrap.py:24-123): ``env`` has a 20-slot header (slot 4..6 = origin of the
``rinv`` operator), then per atom ``(x, y, z, 0.0)``, then per shell ``nprim`` exponents followed by
``nprim`` normalised coefficients; ``atm`` row = ``(int(Z), ptr_xyz, 1, ptr_xyz+3, 0, 0)``; ``bas`` row
= ``(iatom, l, nprim, 1, 0, ptr_exp, ptr_coef, 0)``; ``shell_to_aoloc`` are prefix sums of the
spherical (2l+1) or cartesian shell sizes.  The CUDA kernels consume exactly these arrays
(uploaded once per wrapper through ``b200qc_basis_upload``), which is what makes the C-ABI a
drop-in for the ``dqclibs`` entry points.
"""
from __future__ import annotations
import copy
from typing import Dict, List, Optional, Tuple
import numpy as np
import torch
from dqc_b200.utils.datastruct import AtomCGTOBasis

__all__ = ["LibcintWrapper", "SubsetLibcintWrapper", "PTR_RINV_ORIG", "NDIM"]

PTR_RINV_ORIG = 4
PTR_ENV_START = 20
NDIM = 3


def _shell_size(l: int, spherical: bool) -> int:
    # what CINTcgto_spheric / CINTcgto_cart return for nctr = 1 (reference :376-383)
    return 2 * l + 1 if spherical else (l + 1) * (l + 2) // 2


class LibcintWrapper(object):
    def __init__(self, atombases: List[AtomCGTOBasis], spherical: bool = True, lattice=None) -> None:
        if lattice is not None:
            raise NotImplementedError("periodic systems are outside the B200 Fock-build path")
        self._atombases = atombases
        self._spherical = spherical
        self._natoms = len(atombases)
        self._fracz = False
        self._lattice = None
        self.dtype = atombases[0].bases[0].alphas.dtype
        self.device = atombases[0].bases[0].alphas.device

        atm, bas = [], []
        env: List[float] = [0.0] * PTR_ENV_START
        ptr = PTR_ENV_START
        allpos, allalphas, allcoeffs = [], [], []
        angmoms: List[int] = []
        shell_to_atom: List[int] = []
        ngauss_at_shell: List[int] = []
        gauss_to_shell: List[int] = []
        for iatom, ab in enumerate(atombases):
            assert ab.pos.numel() == NDIM
            z = ab.atomz
            if isinstance(z, float) or (isinstance(z, torch.Tensor) and z.is_floating_point()):
                self._fracz = True
            atm.append([int(z), ptr, 1, ptr + NDIM, 0, 0])
            env.extend(float(x) for x in ab.pos.detach().cpu())
            env.append(0.0)
            ptr += NDIM + 1
            allpos.append(ab.pos.unsqueeze(0))
            for shell in ab.bases:
                assert shell.alphas.shape == shell.coeffs.shape and shell.alphas.ndim == 1
                shell.wfnormalize_()
                ng = len(shell.alphas)
                bas.append([iatom, shell.angmom, ng, 1, 0, ptr, ptr + ng, 0])
                env.extend(float(x) for x in shell.alphas.detach().cpu())
                env.extend(float(x) for x in shell.coeffs.detach().cpu())
                ptr += 2 * ng
                allalphas.append(shell.alphas)
                allcoeffs.append(shell.coeffs)
                angmoms.extend([shell.angmom] * ng)
                gauss_to_shell.extend([len(ngauss_at_shell)] * ng)
                ngauss_at_shell.append(ng)
                shell_to_atom.append(iatom)

        self._allpos_params = torch.cat(allpos, dim=0)
        self._allalphas_params = torch.cat(allalphas, dim=0)
        self._allcoeffs_params = torch.cat(allcoeffs, dim=0)
        self._allangmoms = torch.tensor(angmoms, dtype=torch.int32)
        self._gauss_to_shell = torch.tensor(gauss_to_shell, dtype=torch.int32)
        self._atm = np.array(atm, dtype=np.int32, order="C")
        self._bas = np.array(bas, dtype=np.int32, order="C")
        self._env = np.array(env, dtype=np.float64, order="C")

        nshells = len(bas)
        sizes = [_shell_size(b[1], spherical) for b in bas]
        self._shell_to_aoloc = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        self._shell_idxs = (0, nshells)
        self._ngauss_at_shell_list = ngauss_at_shell
        self._ao_to_shell = torch.tensor(np.repeat(np.arange(nshells), sizes), dtype=torch.long)
        self._ao_to_atom = torch.tensor(np.repeat(np.array(shell_to_atom, dtype=np.int64), sizes),
                                        dtype=torch.long)
        self._dev_handle = None  # lazily created device copy (see device_basis())

    # ---- identity / bookkeeping (same names as the reference) ----
    @property
    def parent(self) -> "LibcintWrapper":
        return self

    @property
    def natoms(self) -> int:
        return self._natoms

    @property
    def fracz(self) -> bool:
        return self._fracz

    @property
    def lattice(self):
        return self._lattice

    @property
    def spherical(self) -> bool:
        return self._spherical

    @property
    def atombases(self) -> List[AtomCGTOBasis]:
        return self._atombases

    @property
    def atm_bas_env(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        return self._atm, self._bas, self._env

    @property
    def full_angmoms(self) -> torch.Tensor:
        return self._allangmoms

    @property
    def params(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        return self._allcoeffs_params, self._allalphas_params, self._allpos_params

    @property
    def shell_idxs(self) -> Tuple[int, int]:
        return self._shell_idxs

    @property
    def full_shell_to_aoloc(self) -> np.ndarray:
        return self._shell_to_aoloc

    @property
    def full_gauss_to_shell(self) -> torch.Tensor:
        return self._gauss_to_shell

    @property
    def full_ao_to_atom(self) -> torch.Tensor:
        return self._ao_to_atom

    @property
    def full_ao_to_shell(self) -> torch.Tensor:
        return self._ao_to_shell

    @property
    def ngauss_at_shell(self) -> List[int]:
        return self._ngauss_at_shell_list

    def __len__(self) -> int:
        return self.shell_idxs[1] - self.shell_idxs[0]

    def nao(self) -> int:
        loc = self.full_shell_to_aoloc
        return int(loc[self.shell_idxs[1]] - loc[self.shell_idxs[0]])

    def ao_idxs(self) -> Tuple[int, int]:
        loc = self.full_shell_to_aoloc
        return int(loc[self.shell_idxs[0]]), int(loc[self.shell_idxs[1]])

    def ao_to_atom(self) -> torch.Tensor:
        return self.full_ao_to_atom[slice(*self.ao_idxs())]

    def ao_to_shell(self) -> torch.Tensor:
        return self.full_ao_to_shell[slice(*self.ao_idxs())]

    def __getitem__(self, inp) -> "LibcintWrapper":
        assert isinstance(inp, slice)
        assert inp.step is None or inp.step == 1
        assert inp.start is not None or inp.stop is not None
        nshells = self.shell_idxs[1]
        start = 0 if inp.start is None else inp.start
        stop = nshells if inp.stop is None else inp.stop
        if start < 0:
            start += nshells
        if stop < 0:
            stop += nshells
        return SubsetLibcintWrapper(self, slice(start, stop))

    @staticmethod
    def concatenate(*wrappers: "LibcintWrapper") -> Tuple["LibcintWrapper", ...]:
        """One environment holding the atoms of every distinct parent, returned as subset views in
        the callers' order (reference :298-361); used by density fitting (dqc/df/dfmol.py:30-33)."""
        parents: List[LibcintWrapper] = []
        index: Dict[int, int] = {}
        which: List[int] = []
        offsets = [0]
        for w in wrappers:
            p = w.parent
            if id(p) not in index:
                index[id(p)] = len(parents)
                parents.append(p)
                offsets.append(offsets[-1] + len(p))
            which.append(index[id(p)])
        assert len(parents) > 0
        if len(parents) == 1:
            return tuple(wrappers)
        p0 = parents[0]
        atombases = copy.copy(p0.atombases)
        for p in parents[1:]:
            assert p.spherical == p0.spherical
            atombases.extend(p.atombases)
        grand = LibcintWrapper(atombases, spherical=p0.spherical)
        out = []
        for w, ip in zip(wrappers, which):
            s0, s1 = w.shell_idxs
            out.append(grand[s0 + offsets[ip]: s1 + offsets[ip]])
        return tuple(out)

    # ---- device side ----
    def device_basis(self, device: Optional[torch.device] = None):
        """Upload (once) and return the device-resident copy of (atm, bas, env, ao_loc) used by
        every CUDA entry point.  Raises if the CUDA library is missing: there is no CPU fallback."""
        from dqc_b200 import _lib
        par = self.parent
        if par._dev_handle is None:
            par._dev_handle = _lib.DeviceBasis(par._atm, par._bas, par._env, par._shell_to_aoloc,
                                               par._spherical, device)
        return par._dev_handle


class SubsetLibcintWrapper(LibcintWrapper):
    """Contiguous shell range of a parent wrapper sharing its environment (reference :386-433)."""

    def __init__(self, parent: LibcintWrapper, subset: slice):
        self._parent = parent
        self._shell_idxs = (subset.start, subset.stop)

    @property
    def parent(self) -> LibcintWrapper:
        return self._parent

    @property
    def shell_idxs(self) -> Tuple[int, int]:
        return self._shell_idxs

    def __getitem__(self, inp):
        raise NotImplementedError("Indexing of SubsetLibcintWrapper is not implemented")

    def __getattr__(self, name):
        return getattr(self._parent, name)
