"""gym per-env API of the multi-agent highway env, one environment on the GPU.

Mirrors gym_highway/multiagent_envs/highway_env.py:11-149 (MultiAgentEnv) and highway_ma_c_env.py:12-82
(MultiAgentEnvContinuous): list-valued spaces, `env.n`, `step(action_n) -> (obs_n, reward_n, done_n,
{'n': info_n})` with one entry per car.  No auto-reset at this level.
"""
import numpy as np

from .vec_ma_env import HighwayMAVecEnv


class MultiAgentEnv(object):
    metadata = {'render.modes': ['human', 'rgb_array']}
    _continuous = False

    def __init__(self, world_config=None, num_agents=1, reset_callback=None, reward_callback=None,
                 observation_callback=None, info_callback=None, done_callback=None, shared_reward=False,
                 seed=0, device=None):
        if any(cb is not None for cb in (reset_callback, reward_callback, observation_callback, info_callback, done_callback)):
            raise NotImplementedError("scenario callbacks are fixed to the highway Scenario (simple_base.py)")
        self._vec = HighwayMAVecEnv(1, num_agents=num_agents, continuous=self._continuous, seed=seed, device=device,
                                    world_config=world_config, shared_reward=shared_reward, return_numpy=True,
                                    auto_reset=False)
        self.n = self._vec.num_agents
        self.shared_reward = shared_reward
        self.action_space = self._vec.action_space
        self.observation_space = self._vec.observation_space

    def step(self, action_n):
        if self._continuous:
            act = np.asarray(action_n, np.float32).reshape(1, self.n, 2)
        else:
            act = np.asarray([np.argmax(a) for a in action_n], np.int32).reshape(1, self.n)  # highway_env.py:78
        obs, rew, done, infos = self._vec.step(act)
        info = infos[0]
        return [obs[0, a] for a in range(self.n)], [float(r) for r in rew[0]], [bool(d) for d in done[0]], {'n': info['n']}

    def reset(self):
        obs = self._vec.reset()
        return [obs[0, a] for a in range(self.n)]

    def seed(self, seed):
        self._vec.seed(seed)

    def close(self):
        self._vec.close()


class MultiAgentEnvContinuous(MultiAgentEnv):
    _continuous = True
