"""gym_highway_b200 — B200-native batched simulator for gym-highway's per-step hot path.

Public surface (mirrors the reference's names):
  HighwayVecEnv / HighwayMAVecEnv          baselines VecEnv API over N environments on one GPU
  HighwayEnv / HighwayEnvContinuous        gym per-env API (gym_highway.envs)
  MultiAgentEnv / MultiAgentEnvContinuous  gym per-env multi-agent API (gym_highway.multiagent_envs)
  make_vec_env                             baselines.common.cmd_util.make_vec_env for the ids below
Everything steps through libhighway_b200.so (hand-written sm_100a CUDA behind a C ABI); importing an
env class without that library raises — there is no CPU fallback.
"""
__version__ = "0.1.0"

# gym ids the reference registers (gym_highway/__init__.py:6-17, models/config.py:111-126)
ENV_IDS = {
    'Highway-v0': ('sa', False),
    'Highway-train-v0': ('sa', False),
    'Highway-cont-train-v0': ('sa', True),
    'HighwayMultiagent-v0': ('ma', False),
    'MA-Highway-train-v0': ('ma', False),
    'MA-Highway-cont-train-v0': ('ma', True),
}


def __getattr__(name):
    # lazy so that `import gym_highway_b200` (and the CPU-only tests of the host logic) need no GPU
    if name in ("HighwayVecEnv", "VecEnv", "AlreadySteppingError", "NotSteppingError"):
        from . import vec_env
        return getattr(vec_env, name)
    if name in ("HighwayMAVecEnv",):
        from . import vec_ma_env
        return getattr(vec_ma_env, name)
    if name in ("HighwayEnv", "HighwayEnvContinuous"):
        from . import envs
        return getattr(envs, name)
    if name in ("MultiAgentEnv", "MultiAgentEnvContinuous"):
        from . import multiagent_envs
        return getattr(multiagent_envs, name)
    if name == "make_vec_env":
        from .cmd_util import make_vec_env
        return make_vec_env
    raise AttributeError(name)
