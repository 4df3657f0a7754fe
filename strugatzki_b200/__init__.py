"""strugatzki_b200 -- B200-native engine for Strugatzki's feature-space similarity hot path
(FeatureCorrelation / FeatureSegmentation / SelfSimilarity) behind the reference's Config/Processor API.

The arithmetic lives in hand-written sm_100a CUDA (strugatzki_b200/csrc) behind the C ABI declared in
include/strugatzki_b200.h; this package is the host-side mirror of the reference's interface.  There is no
CPU fallback: compute calls raise NativeError when the library or a B200 is missing.
"""
from ._native import Aborted, NativeError, build  # noqa: F401

__version__ = "0.1.0"
