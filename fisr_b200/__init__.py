"""fisr_b200: B200-native FISRnet hot path (hand-written sm_100a CUDA behind a C ABI).

``FISRnet`` mirrors the reference class surface (FISRnet.py:14-17); ``Engine`` is the thin
device context underneath it.  The CPU checker package is never imported from here.
"""
from ._lib import FisrError, LIB_PATH  # noqa: F401
from .engine import Engine, param_inventory  # noqa: F401
from .FISRnet import FISRnet  # noqa: F401
