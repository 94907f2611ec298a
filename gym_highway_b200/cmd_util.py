"""`make_vec_env` with the signature trainers already call (baselines/common/cmd_util.py:23-55).

The reference forks `num_env` processes (SubprocVecEnv / SubprocVecMAEnv), seeds env i with
`seed + 10000*mpi_rank + i` and wraps each in a Monitor; here one GPU-resident VecEnv holds all
`num_env` environments and env i of rank r draws from the Philox stream of its GLOBAL id
`start_index + r*num_env + i` (the reference's fixed 10000 stride would make the id ranges of two ranks
overlap as soon as a VecEnv holds more than 10000 environments - most of rank r+1's environments would
then replay rank r's obstacle spawns and, under DeviceRunner, its sampled actions); Monitor's episode
accounting is fused into the step kernel.
"""
import os

from . import ENV_IDS


def make_vec_env(env_id, env_type=None, num_env=1, seed=None, wrapper_kwargs=None, start_index=0, reward_scale=1.0,
                 flatten_dict_observations=True, gamestate=None, isMultiAgent=False, num_agents=5, device=None,
                 env_kwargs=None, return_numpy=False):
    if env_id not in ENV_IDS:
        raise KeyError("unknown highway env id %r (known: %s)" % (env_id, ", ".join(sorted(ENV_IDS))))
    if reward_scale != 1.0:
        raise NotImplementedError("reward_scale is a retro wrapper option and does not apply to the highway envs")
    kind, continuous = ENV_IDS[env_id]
    rank = int(os.environ.get("RANK", "0"))
    seed = 0 if seed is None else int(seed)
    kw = dict(env_kwargs or {})
    env_id0 = int(start_index) + rank * int(num_env)  # disjoint global env ids across ranks for any num_env
    if kind == 'ma' or isMultiAgent:
        from .vec_ma_env import HighwayMAVecEnv
        return HighwayMAVecEnv(num_env, num_agents=num_agents, continuous=continuous, seed=seed, device=device,
                               world_config=kw, env_id0=env_id0, return_numpy=return_numpy)
    from .vec_env import HighwayVecEnv
    return HighwayVecEnv(num_env, continuous=continuous, seed=seed, device=device, env_id0=env_id0,
                         return_numpy=return_numpy, **kw)
