from dqc_b200.hamilton.intor.lcintwrap import *  # noqa: F401,F403
from dqc_b200.hamilton.intor.molintor import *  # noqa: F401,F403
from dqc_b200.hamilton.intor.gtoeval import *  # noqa: F401,F403
