"""blockmaze_b200 -- Python (ctypes) face of libzkb200.so, the B200-native Groth16 prover behind BlockMaze's libzk* C-ABI.

Everything computes on the GPU through the C-ABI declared in include/zkb200.h; there is no CPU fallback and importing
this package on a machine without the built library raises immediately.
"""
from .api import *  # noqa: F401,F403
from .api import lib, init, ProvingKey, ZkError  # noqa: F401
