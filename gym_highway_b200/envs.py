"""gym per-env API of the single-agent highway env, one environment on the GPU.

Same names, constructor keywords, return types and episode semantics as
gym_highway/envs/highway_env.py:22-147 (HighwayEnv) and highway_env_cont.py:10-55 (HighwayEnvContinuous):
`reset() -> ob`, `step(a) -> (ob float32[18], reward float, done bool, info dict)`, `seed(s)`, `close()`.
There is no auto-reset at this level (that is the VecEnv's job): after `done` the caller resets.
A batch of one is the least efficient way to use a GPU; this class exists so that code written against
`gym.make('Highway-v0')` keeps working, e.g. for evaluation.
"""
import numpy as np

from .vec_env import HighwayVecEnv

ACTION_LOOKUP = {0: 'LEFT', 1: 'RIGHT', 2: 'ACCELERATE', 3: 'MAINTAIN'}  # highway_env.py:14-20


class HighwayEnv(object):
    metadata = {'render.modes': ['human']}
    reward_range = (-float('inf'), float('inf'))
    spec = None
    _continuous = False

    def __init__(self, manual=False, inf_obs=True, save=False, render=False, real_time=False, seed=0, device=None):
        self.__version__ = "0.0.1"
        self._vec = HighwayVecEnv(1, continuous=self._continuous, seed=seed, device=device, manual=manual, inf_obs=inf_obs,
                                  save=save, render=render, real_time=real_time, return_numpy=True, auto_reset=False)
        self.action_space = self._vec.action_space
        self.observation_space = self._vec.observation_space

    @property
    def unwrapped(self):
        return self

    def step(self, action):
        obs, rew, done, infos = self._vec.step([action])
        info = dict(infos[0])
        info.pop('episode', None)
        # the reference returns python's round(reward, 3): recover it from the float32 buffer
        return obs[0], round(float(rew[0]), 3), bool(done[0]), info

    def reset(self):
        return self._vec.reset()[0]

    def _get_state(self):
        return self._vec.observe()[0]

    def seed(self, seed):
        self._vec.seed(seed)

    def close(self):
        self._vec.close()

    def render(self, mode='human'):
        raise NotImplementedError("rendering belongs to the interactive pygame front-end")


class HighwayEnvContinuous(HighwayEnv):
    _continuous = True

    def step(self, action):
        obs, rew, done, infos = self._vec.step(np.asarray(action, np.float32).reshape(1, 2))
        info = dict(infos[0])
        info.pop('episode', None)
        return obs[0], round(float(rew[0]), 3), bool(done[0]), info
