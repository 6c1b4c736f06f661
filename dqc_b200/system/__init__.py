from dqc_b200.system.base_system import BaseSystem  # noqa: F401
from dqc_b200.system.mol import Mol  # noqa: F401
