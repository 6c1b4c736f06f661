from dqc_b200.df.base_df import BaseDF  # noqa: F401
from dqc_b200.df.dfmol import DFMol  # noqa: F401
