"""Stand-in for the few gym names the reference env adapters import (gym is not installed here).

TEST INFRASTRUCTURE ONLY.  Semantics that matter for parity: spaces.Box casts low/high to its
dtype (float32 for the single-agent env, which fixes the dtype of the normalisation arithmetic,
reference gym_highway/envs/highway_env.py:51 and :131-140).
"""
from . import spaces, utils, error  # noqa: F401
from . import envs  # noqa: F401


from .core import Env, Wrapper  # noqa: F401,E402
