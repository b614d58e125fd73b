"""B200-native RAISR hot path of saturdaycoder/oclComputerVision."""
from .raisr import ClRaisr, get_elapsed_ms  # noqa: F401
from .interpolation import clUtility  # noqa: F401

__all__ = ["ClRaisr", "clUtility", "get_elapsed_ms"]
