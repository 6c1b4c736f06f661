from dqc_b200.qccalc.scf_qccalc import SCF_QCCalc, equilibrium  # noqa: F401
from dqc_b200.qccalc.hf import HF  # noqa: F401
from dqc_b200.qccalc.ks import KS  # noqa: F401
