"""r2p2_b200 -- B200-native batched robot-step core behind r2p2's plugin surface."""
from ._lib import R2P2Error, build, lib  # noqa: F401
from .sim import ACT, CTRL, BatchSim, expand_sonars  # noqa: F401

__version__ = "0.1.0"
