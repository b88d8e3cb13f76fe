from .module import Module, ModuleList, Sequential, ModuleDict  # noqa: F401
from .layers import *  # noqa: F401,F403
from .parallel import *  # noqa: F401,F403
