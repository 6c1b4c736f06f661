from dqc_b200.hamilton.base_hamilton import BaseHamilton  # noqa: F401
from dqc_b200.hamilton.hcgto import HamiltonCGTO  # noqa: F401
