"""rl_on_manifold_b200 — batched ATACOM tangent-space projection on NVIDIA B200 (sm_100a).

Drop-in for the one hot path of PuzeLiu/rl_on_manifold: AtacomEnvWrapper.step_action_function
(atacom/atacom.py:123-139) and the helpers it reaches, over batches of 10^4-10^5 independent
environments.  The compute lives in hand-written CUDA (csrc/) behind a C ABI
(include/atacom_b200.h); this package is the host-side mirror of the reference's interface.
"""
from . import _lib                                   # raises if libatacom_b200.so is missing
from ._lib import AtacomParams, AtacomError, default_params, version
from . import projection, synthetic

__all__ = ["AtacomParams", "AtacomError", "default_params", "version", "projection", "synthetic"]
