from dqc_b200.utils.datastruct import *  # noqa: F401,F403
from dqc_b200.utils.config import config  # noqa: F401
