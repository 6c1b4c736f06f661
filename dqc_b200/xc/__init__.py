from dqc_b200.xc.base_xc import BaseXC, AddBaseXC, MulBaseXC  # noqa: F401
from dqc_b200.xc.b200xc import B200XC, get_b200xc  # noqa: F401
