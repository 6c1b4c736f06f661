"""Small helpers with the reference's names (dqc/utils/misc.py:58-66 logger, :53-56 gaussian_int;
dqc/utils/safeops.py occnumber)."""
import torch
from dqc_b200.utils.config import config
from dqc_b200.utils.datastruct import gaussian_int  # noqa: F401

__all__ = ["logger", "occnumber"]


class _Logger(object):
    def log(self, s: str, vlevel: int = 0):
        if config.VERBOSE > vlevel:
            print(s, flush=True)


logger = _Logger()


def occnumber(a, n=None, dtype=torch.double, device=torch.device("cpu")) -> torch.Tensor:
    """Occupation vector (1, ..., 1, frac) summing to ``a`` (dqc/utils/safeops.py:21-110 forward
    semantics): floor(a) ones, then the fractional remainder; padded with zeros to length n."""
    a = float(a)
    nfloor = int(a // 1)
    frac = a - nfloor
    vals = [1.0] * nfloor + ([frac] if frac > 1e-12 else [])
    if n is not None:
        assert n >= len(vals), "n is too small for the occupation number"
        vals = vals + [0.0] * (n - len(vals))
    return torch.tensor(vals, dtype=dtype, device=device)
