"""Stand-in for `gymnasium` (>= 1.1).  TEST INFRASTRUCTURE -- see oracle/refshim/README.md."""
from . import envs, error, spaces, utils, vector, wrappers
from .core import ActionWrapper, Env, ObservationWrapper, RewardWrapper, Wrapper
from .envs.registration import make, make_vec, register, registry
from .spaces import Space

__version__ = "1.1.0+refshim"
