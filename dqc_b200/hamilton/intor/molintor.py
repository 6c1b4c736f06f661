"""Integral front-end with the reference's names and output orientation
(dqc/hamilton/intor/molintor.py:12-130): every function takes LibcintWrapper objects (possibly
subsets of one concatenated environment) and returns a device tensor computed by the Rys kernels of
libb200qc.so -- there is no CPU path."""
from typing import Optional
import torch
from dqc_b200 import _lib
from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper

__all__ = ["int1e", "int2c2e", "int3c2e", "int2e", "overlap", "kinetic", "nuclattr", "elrep",
           "coul2c", "coul3c", "coul3c_packed"]


def _same_env(*ws: LibcintWrapper):
    p = ws[0].parent
    for w in ws[1:]:
        if w.parent is not p:
            raise RuntimeError("wrappers must share one environment: use LibcintWrapper.concatenate first "
                               "(the reference concatenates on the fly, lcintwrap.py:298-361)")
    return p


def _device(wrapper: LibcintWrapper) -> torch.device:
    dev = wrapper.device
    if dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else dev
    return dev


def _wants_pos_grad(wrapper: LibcintWrapper) -> bool:
    """True when the atomic positions behind the wrapper are part of an autograd graph; exponents / coefficients in a
    graph are refused (their backward is not built: the result would silently carry no gradient)."""
    coeffs, alphas, pos = wrapper.parent.params
    if torch.is_grad_enabled() and (coeffs.requires_grad or alphas.requires_grad):
        raise NotImplementedError("gradients with respect to basis exponents / coefficients are not built "
                                  "(SURVEY 8f rank 1: only the atomic positions so far); detach them")
    return torch.is_grad_enabled() and pos.requires_grad


def _int1e_nograd(shortname: str, wrapper: LibcintWrapper, other: Optional[LibcintWrapper] = None,
                  rinv_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    other = wrapper if other is None else other
    db = wrapper.device_basis(_device(wrapper))
    orig = None if rinv_pos is None else rinv_pos.detach().cpu().numpy()
    return _lib.int1e(db, shortname, (*wrapper.shell_idxs, *other.shell_idxs), orig)


def int1e(shortname: str, wrapper: LibcintWrapper, other: Optional[LibcintWrapper] = None, *,
          rinv_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """2-centre 1-electron integrals <wrapper| op |other>: shortname in "ovlp", "kin", "nuc", "rinv", or their first
    derivatives with respect to the electron coordinate of the bra, "ipovlp", "ipkin", "ipnuc", "iprinv" -> (3, nao, nao)
    (libcint's int1e_ip*, molintor.py:178-300).  With atomic positions that require grad the plain integrals are
    differentiable (first order, positions only)."""
    same = other is None or other is wrapper
    other = wrapper if other is None else other
    _same_env(wrapper, other)
    if shortname in ("ipovlp", "ipkin", "ipnuc", "iprinv"):
        from dqc_b200.hamilton.intor import deriv
        if not same:
            raise NotImplementedError("derivative integrals between two different shell ranges are not built")
        return deriv.ip1e(shortname[2:], wrapper, rinv_pos=rinv_pos)
    if shortname in ("r0", "r0r0", "r0r0r0"):
        # multipole integrals of the electric-field terms (hcgto.py:118-127): 3, 9, 27 components
        from dqc_b200.hamilton.intor import multipole
        if not same:
            raise NotImplementedError("multipole integrals between two different shell ranges are not built")
        return multipole.multipole1e(len(shortname) // 2, wrapper)
    if shortname not in ("ovlp", "kin", "nuc", "rinv"):
        raise NotImplementedError("int1e_%s is outside the Fock-build path" % shortname)
    if shortname == "rinv":
        assert rinv_pos is not None and rinv_pos.numel() == 3, "rinv_pos must be given for rinv"
    if shortname != "rinv" and _wants_pos_grad(wrapper):
        from dqc_b200.hamilton.intor import deriv
        if not same:
            raise NotImplementedError("position gradients of integrals between two different shell ranges are not built")
        return deriv.Int1eFunction.apply(wrapper.parent.params[2], wrapper, shortname)
    return _int1e_nograd(shortname, wrapper, other, rinv_pos)


def int2c2e(shortname: str, wrapper: LibcintWrapper, other: Optional[LibcintWrapper] = None) -> torch.Tensor:
    if shortname not in ("r12", "ar12"):
        raise NotImplementedError("int2c2e_%s is outside the Fock-build path" % shortname)
    other = wrapper if other is None else other
    _same_env(wrapper, other)
    db = wrapper.device_basis(_device(wrapper))
    return _lib.int2c2e(db, (*wrapper.shell_idxs, *other.shell_idxs))


def int3c2e(shortname: str, wrapper: LibcintWrapper, other1: Optional[LibcintWrapper] = None,
            other2: Optional[LibcintWrapper] = None) -> torch.Tensor:
    if shortname not in ("ar12",):
        raise NotImplementedError("int3c2e_%s is outside the Fock-build path" % shortname)
    other1 = wrapper if other1 is None else other1
    other2 = wrapper if other2 is None else other2
    _same_env(wrapper, other1, other2)
    db = wrapper.device_basis(_device(wrapper))
    return _lib.int3c2e(db, (*wrapper.shell_idxs, *other1.shell_idxs, *other2.shell_idxs))


def _int2e_nograd(wrapper: LibcintWrapper, others=None) -> torch.Tensor:
    others = [wrapper] * 3 if others is None else others
    db = wrapper.device_basis(_device(wrapper))
    sl = (*wrapper.shell_idxs, *others[0].shell_idxs, *others[1].shell_idxs, *others[2].shell_idxs)
    return _lib.int2e(db, sl)


def int2e(shortname: str, wrapper: LibcintWrapper, other1: Optional[LibcintWrapper] = None,
          other2: Optional[LibcintWrapper] = None, other3: Optional[LibcintWrapper] = None) -> torch.Tensor:
    """(ij|kl), or with "ipar12b" its derivative with respect to the electron coordinate of i -> (3, nao, nao, nao, nao)
    (libcint's int2e_ip1).  With atomic positions that require grad (ij|kl) is differentiable (first order)."""
    same = all(o is None or o is wrapper for o in (other1, other2, other3))
    others = [wrapper if o is None else o for o in (other1, other2, other3)]
    _same_env(wrapper, *others)
    if shortname == "ipar12b":
        from dqc_b200.hamilton.intor import deriv
        if not same:
            raise NotImplementedError("derivative integrals between different shell ranges are not built")
        return deriv.ip2e(wrapper)
    if shortname not in ("ar12b",):
        raise NotImplementedError("int2e_%s is outside the Fock-build path" % shortname)
    if _wants_pos_grad(wrapper):
        from dqc_b200.hamilton.intor import deriv
        if not same:
            raise NotImplementedError("position gradients of integrals between different shell ranges are not built")
        return deriv.Int2eFunction.apply(wrapper.parent.params[2], wrapper)
    return _int2e_nograd(wrapper, others)


def overlap(wrapper, other=None):
    return int1e("ovlp", wrapper, other=other)


def kinetic(wrapper, other=None):
    return int1e("kin", wrapper, other=other)


def nuclattr(wrapper: LibcintWrapper, other: Optional[LibcintWrapper] = None) -> torch.Tensor:
    """sum_A -Z_A <mu| 1/|r - R_A| |nu>.  Integer charges go through int1e_nuc; fractional charges
    are summed explicitly from rinv integrals like the reference (molintor.py:102-112)."""
    if not wrapper.fracz:
        return int1e("nuc", wrapper, other=other)
    res = None
    par = wrapper.parent
    for ab in par.atombases:
        y = int1e("rinv", wrapper, other=other, rinv_pos=ab.pos) * (-float(ab.atomz))
        res = y if res is None else res + y
    return res


def elrep(wrapper, other1=None, other2=None, other3=None):
    return int2e("ar12b", wrapper, other1, other2, other3)


def coul2c(wrapper, other=None):
    return int2c2e("r12", wrapper, other)


def coul3c(wrapper, other1=None, other2=None):
    return int3c2e("ar12", wrapper, other1, other2)


def coul3c_packed(wrapper: LibcintWrapper, auxwrapper: LibcintWrapper, aux_slice=None) -> torch.Tensor:
    """(ij|P) stored once per AO pair i >= j: (nao (nao + 1) / 2, ld) with ld = naux rounded up to
    even -- the resident tensor of the density-fitted J (half of the reference's j3c).  aux_slice =
    (s0, s1) restricts P to a shell sub-range (aux-sharded multi-GPU layout)."""
    _same_env(wrapper, auxwrapper)
    db = wrapper.device_basis(_device(wrapper))
    a0, a1 = auxwrapper.shell_idxs if aux_slice is None else aux_slice
    return _lib.int3c2e_packed(db, (*wrapper.shell_idxs, *wrapper.shell_idxs, a0, a1))
