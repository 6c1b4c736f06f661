"""Molecule descriptor parser; same accepted inputs as dqc/api/parser.py:8-62
("H 0 0 0; H 0 0 1.4" strings in Bohr, or (atomzs, atompos) tuples)."""
from typing import Tuple
import torch
from dqc_b200.utils.periodictable import get_atomz

__all__ = ["parse_moldesc"]


def parse_moldesc(moldesc, dtype: torch.dtype = torch.float64,
                  device: torch.device = torch.device("cpu")) -> Tuple[torch.Tensor, torch.Tensor]:
    if isinstance(moldesc, str):
        rows = [ln.split() for ln in moldesc.split(";") if ln.strip()]
        zs = torch.tensor([get_atomz(r[0].strip()) for r in rows], device=device)
        pos = torch.tensor([[float(x) for x in r[1:]] for r in rows], dtype=dtype, device=device)
    else:
        zs_raw, pos_raw = moldesc
        assert len(zs_raw) == len(pos_raw), "Mismatch length of atomz and atompos"
        assert len(zs_raw) > 0, "Empty atom list"
        zs = zs_raw.to(device) if isinstance(zs_raw, torch.Tensor) else \
            torch.tensor([get_atomz(a) for a in zs_raw], device=device)
        pos = pos_raw.to(dtype).to(device) if isinstance(pos_raw, torch.Tensor) else \
            torch.as_tensor(pos_raw, dtype=dtype, device=device)
    if zs.is_floating_point():
        zs = zs.to(dtype)
    return zs, pos
