"""dmvae - B200-native trajectory-VAE hot path (host side).

Python mirror of the reference's interface for this path; all arithmetic runs
in hand-written sm_100a kernels behind ``libdmvae.so`` (``include/dmvae.h``).
"""
from ._lib import DmvaeError, LIB_PATH  # noqa: F401
from .model import ConditionalTrajectoryVAE  # noqa: F401

__all__ = ["ConditionalTrajectoryVAE", "DmvaeError", "LIB_PATH"]
