"""hankb200 — host-side mirror of the reference's model API over libhankb200.so (CUDA, sm_100a).

The compute path is the C ABI in include/hankb200.h; this package only marshals arrays and mirrors
the reference's call signatures (BackwardIteration / ForwardIteration / JVP / NewtonRaphsonHANK) so
that parity tests read like the reference's own scripts.  No CPU fallback exists.
"""
from ._lib import HankError, load  # noqa: F401
from .household import HouseholdBlock  # noqa: F401
