"""b200qc: the SCF Fock-build path of diffqc/dqc on a B200, behind dqc's operator surface.
Importing the package does not need a GPU; constructing a Hamiltonian / running a kernel does."""
from dqc_b200.utils.datastruct import CGTOBasis, AtomCGTOBasis, SpinParam, ValGrad, DensityFitInfo  # noqa: F401
from dqc_b200.utils.config import config  # noqa: F401
from dqc_b200.api.loadbasis import loadbasis  # noqa: F401
from dqc_b200.api.getxc import get_xc  # noqa: F401
from dqc_b200.system.mol import Mol  # noqa: F401
from dqc_b200.qccalc.hf import HF  # noqa: F401
from dqc_b200.qccalc.ks import KS  # noqa: F401

__version__ = "0.1.0"
