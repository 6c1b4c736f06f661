from dqc_b200.grid.base_grid import BaseGrid  # noqa: F401
from dqc_b200.grid.radial_grid import RadialGrid  # noqa: F401
from dqc_b200.grid.lebedev_grid import LebedevGrid, TruncatedLebedevGrid  # noqa: F401
from dqc_b200.grid.multiatoms_grid import BeckeGrid  # noqa: F401
from dqc_b200.grid.factory import get_grid, get_predefined_grid  # noqa: F401
