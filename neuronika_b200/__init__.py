"""neuronika_b200 -- B200-native dense forward/backward hot path of neuronika.

Device tensors live in HBM; every operator is a hand-written sm_100a CUDA kernel reached through
the C ABI in include/nk_b200.h (libnk_b200.so).  Importing this package without the built library
raises ImportError: there is no CPU fallback."""
from . import _lib
from ._lib import NkError
from .device import BF16, F32, CuArray, Device
from . import ops
from . import variable, nn, optim
from .variable import (Reduction, Var, VarDiff, from_ndarray, full, ones, rand, set_fusion, zeros)

__all__ = ["Device", "CuArray", "F32", "BF16", "NkError", "ops", "variable", "nn", "optim", "Var", "VarDiff",
           "Reduction", "zeros", "ones", "full", "rand", "from_ndarray", "set_fusion"]
