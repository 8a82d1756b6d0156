"""solorl_b200 — B200-native batched Solo8/Solo12 environment step.

Host-side mirror of the reference's env interface (``baseEnv.py``, ``solo.py``,
``agents/ppo/envs.py``) over hand-written sm_100a CUDA kernels reached through the
C-ABI of ``include/solo_b200.h``.  No CPU fallback.
"""
from .model import SoloModel  # noqa: F401
from .abi import default_params, params_from_config  # noqa: F401

__all__ = ["SoloModel", "default_params", "params_from_config"]
