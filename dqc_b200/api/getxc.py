"""``get_xc("lda_x + 0.5*gga_c_pbe")`` -> BaseXC, same expression grammar as the reference
(dqc/api/getxc.py:38-59: every identifier becomes a functional object and the expression is
evaluated with the BaseXC ``+`` / ``*`` algebra); hybrids are rejected like there (:29-36)."""
import re
from dqc_b200.xc.base_xc import BaseXC
from dqc_b200.xc.b200xc import get_b200xc

__all__ = ["get_xc", "get_libxc"]


def get_libxc(name: str) -> BaseXC:
    if name.lower().startswith("hyb_"):
        raise NotImplementedError("Hybrid functionals are not supported through get_xc (as in the reference); "
                                  "compose exact exchange explicitly (qccalc.KS(..., exx_fraction=...))")
    return get_b200xc(name)


def get_xc(xcstr: str) -> BaseXC:
    pattern = r"([a-zA-Z_$][a-zA-Z_$0-9]*)"
    new_xcstr = re.sub(pattern, r'get_libxc("\1")', xcstr)
    return eval(new_xcstr, {"get_libxc": get_libxc, "__builtins__": {}})
