"""B200-native GraphSLAM back end (association + Gauss-Newton step) behind the reference's
Slam/Cone interface.  Python here is plumbing for tests and bench: the product is the CUDA library
(csrc/, libslam_b200.so) behind the C ABI of include/slam_b200.h and the C++ Slam mirror in
csrc/host/."""
from . import synth  # noqa: F401
from . import capi  # noqa: F401
from ._build import build  # noqa: F401
from .capi import Context, SymbolicAnalysis, SlamB200Error  # noqa: F401
