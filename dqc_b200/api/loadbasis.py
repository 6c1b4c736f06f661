"""gaussian94 basis reader with the reference's grammar (dqc/api/loadbasis.py:11-152):
'!'/blank header lines skipped, first other line is the element line, each block
``TYPE nprim scale`` followed by nprim rows ``alpha c1 [c2 ...]`` (Fortran 'D' exponents accepted),
every coefficient column becomes its own shell sharing the exponents, ``SP`` -> one s + one p,
stop at ``**``; every shell is normalised on load.  There is no network here, so instead of
basis_set_exchange the tables live in dqc_b200/data/basis/<name>/<ZZ>.gaussian94."""
import os
from typing import List
import torch
from dqc_b200.utils.datastruct import CGTOBasis

__all__ = ["loadbasis", "has_basis"]

_SPDF = {c: l for l, c in enumerate("spdfghi")}
_DATADIR = os.path.join(os.path.dirname(os.path.realpath(__file__)), "..", "data", "basis")


def loadbasis(cmd: str, dtype: torch.dtype = torch.double,
              device: torch.device = torch.device("cpu"), requires_grad: bool = False) -> List[CGTOBasis]:
    path = cmd if os.path.exists(cmd) else _basis_path(cmd)
    with open(path, "r") as f:
        lines = f.read().split("\n")
    # header: drop comments/blank lines and the element line
    while lines:
        line = lines.pop(0)
        if line == "" or line.startswith("!"):
            continue
        break
    shells: List[CGTOBasis] = []
    while lines:
        head = lines.pop(0)
        if head.startswith("**"):
            break
        if head.strip() == "":
            continue
        typ, nprim = head.split()[0], int(head.split()[1])
        if nprim == 0:
            raise RuntimeError("Zero line on basis %s" % path)
        rows = [[float(x.replace("D", "E").replace("d", "e")) for x in lines.pop(0).split()]
                for _ in range(nprim)]
        alphas = torch.tensor([r[0] for r in rows], dtype=dtype, device=device, requires_grad=requires_grad)
        ncol = len(rows[0]) - 1
        for icol, l in enumerate(_expand_angmoms(typ, ncol)):
            c = torch.tensor([r[1 + icol] for r in rows], dtype=dtype, device=device,
                             requires_grad=requires_grad)
            shells.append(CGTOBasis(angmom=l, alphas=alphas, coeffs=c).wfnormalize_())
    return shells


def _expand_angmoms(s: str, n: int) -> List[int]:
    if len(s) != n:
        if n % len(s) != 0:
            raise RuntimeError("Do not know how to read orbital %s with %d coefficient columns" % (s, n))
        s = s * (n // len(s))
    return [_SPDF[c] for c in s.lower()]


def _normalize_basisname(name: str) -> str:
    b = name.lower().replace("+", "p").replace("*", "s")
    for ch in "(),":
        b = b.replace(ch, "_")
    return b


def has_basis(z: int, name: str) -> bool:
    """True when the table of basis `name` for element `z` is embedded under data/basis."""
    return os.path.exists(os.path.join(_DATADIR, _normalize_basisname(name.strip()), "%02d.gaussian94" % int(z)))


def _basis_path(cmd: str) -> str:
    z, raw = cmd.split(":")
    name = _normalize_basisname(raw.strip())
    path = os.path.join(_DATADIR, name, "%02d.gaussian94" % int(z))
    if not os.path.exists(path):
        have = sorted(os.listdir(_DATADIR))
        raise FileNotFoundError(
            "basis '%s' for Z=%d is not embedded (no network to fetch it). Embedded sets: %s; "
            "or pass a gaussian94 file path." % (raw, int(z), have))
    return path
