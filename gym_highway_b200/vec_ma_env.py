"""GPU-resident vectorised MULTI-AGENT highway environments behind the fork's multi-agent VecEnv API.

Mirrors DummyVecMAEnv / SubprocVecMAEnv (baselines/common/vec_env/dummy_vec_ma_env.py:6-104,
subproc_vec_ma_env.py:5-118): observations [N, A, 4A+14] float32, rewards [N, A] float32, dones [N, A]
bool, infos N x {'n': [A dicts]}; an environment is reset as soon as ANY of its cars is done and the
returned observation is then the post-reset one.  Per-env semantics are those of MultiAgentEnv /
MultiAgentEnvContinuous (gym_highway/multiagent_envs/highway_env.py, highway_ma_c_env.py) with the
`Action` name fix the discrete env needs.
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _lib
from .spaces import Box, Discrete
from .vec_env import AlreadySteppingError, LazyInfos, NotSteppingError, VecEnv, _as_ptr, _run_time_table


class HighwayMAVecEnv(VecEnv):
    """N multi-agent highway environments (A <= 5 policy cars + scripted obstacles) per kernel launch.

    `num_agents` cars start in the reference's "X" formation (simple_base.py:32-38, at most five).
    Actions: int[N, A] action indices, or [N, A, 4] score vectors (arg-maxed like the reference does,
    highway_env.py:77-78), or float[N, A, 2] when `continuous`.
    """

    def __init__(self, num_envs, num_agents=5, continuous=False, seed=0, device=None, world_config=None,
                 shared_reward=False, num_slots=_lib.MA_MAX_SLOTS, env_id0=0, return_numpy=False, auto_reset=True,
                 track_stats=True, exact_steering=False, kernel="auto", extended_agents=False):
        cfg = dict(manual=False, inf_obs=True, save=False, render=False, real_time=False)
        cfg.update(world_config or {})
        if cfg['manual'] or cfg['save'] or cfg['render']:
            raise NotImplementedError("manual / save / render belong to the interactive pygame front-end")
        limit = _lib.MA_MAX_AGENTS_EXT if extended_agents else _lib.MA_MAX_AGENTS
        if not 1 <= num_agents <= limit:
            # the reference raises IndexError in Scenario.observations beyond five cars.  extended_agents=True
            # opens the repo's own 6..8-car extension: an invented continuation of the start grid, specified only by the CPU
            # restatement (oracle/ma_oracle.c) - there is no reference behaviour to be equal to.
            raise IndexError("the reference's start grid defines %d cars; num_agents=%d%s" % (
                _lib.MA_MAX_AGENTS, num_agents, "" if extended_agents else " (extended_agents=True allows up to %d, without reference parity)" % _lib.MA_MAX_AGENTS_EXT))
        if num_agents > _lib.MA_MAX_AGENTS and kernel == "group":
            raise ValueError("the 6..8-car extension is stepped by the thread-per-environment kernel only")
        if not 3 <= num_slots <= _lib.MA_MAX_SLOTS:
            raise ValueError("num_slots must be in 3..%d" % _lib.MA_MAX_SLOTS)
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.HighwayLibraryError("HighwayMAVecEnv needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.n = self.num_agents = int(num_agents)
        self.num_slots = int(num_slots)
        self.continuous = bool(continuous)
        self.shared_reward = bool(shared_reward)
        self.params = _lib.sim_params(cfg['real_time'])
        self.return_numpy = bool(return_numpy)
        self._flags = ((_lib.FLAG_CONTINUOUS if continuous else 0) | (0 if cfg['inf_obs'] else _lib.FLAG_NO_INF_OBS)
                       | (0 if auto_reset else _lib.FLAG_NO_AUTORESET) | (_lib.FLAG_SHARED_REWARD if shared_reward else 0)
                       | (_lib.FLAG_EXACT_STEERING if exact_steering else 0)
                       | (_lib.FLAG_MA_THREAD_PER_ENV if kernel == "thread" else 0)
                       | (_lib.FLAG_MA_LANE_PER_CAR if kernel == "group" else 0))
        if kernel not in ("auto", "group", "thread"):
            raise ValueError("kernel must be 'auto' (chosen by batch size, default), 'group' (a lane per car) or 'thread' (a thread per "
                             "environment, cars in registers); the results are identical")
        self._seed = int(seed) & (2 ** 64 - 1)
        self._env_id0 = int(env_id0)
        A, K = self.num_agents, self.num_slots
        self.obs_dim = 4 * A + 14
        obs_space = [Box(-np.inf, np.inf, shape=(self.obs_dim,), dtype=np.float32) for _ in range(A)]
        act_space = [(Box(np.array([-1.0, -1.0]), np.array([1.0, 1.0]), dtype=np.float32) if continuous else Discrete(4))
                     for _ in range(A)]
        VecEnv.__init__(self, int(num_envs), obs_space, act_space)
        n, dev = self.num_envs, self.device
        with torch.cuda.device(dev):
            self.cars = torch.zeros(A * 9, n, dtype=torch.float32, device=dev)
            self.obst = torch.zeros(K, n, 4, dtype=torch.float32, device=dev)
            self.head = torch.zeros(10, n, dtype=torch.int32, device=dev)
            self._out = [dict(obs=torch.zeros(n, A, self.obs_dim, dtype=torch.float32, device=dev),
                              rew=torch.zeros(n, A, dtype=torch.float32, device=dev),
                              done=torch.zeros(n, A, dtype=torch.uint8, device=dev),
                              term=torch.zeros(n, 5, dtype=torch.float64, device=dev)) for _ in range(2)]
            self.stats = torch.zeros(8, dtype=torch.int64, device=dev) if track_stats else None
        self._state_c = _lib.HwMaState(self.cars.data_ptr(), self.obst.data_ptr(), self.head.data_ptr(), A, K)
        self._out_c = [_lib.HwMaOut(o['obs'].data_ptr(), o['rew'].data_ptr(), o['done'].data_ptr(), o['term'].data_ptr(),
                                    self.stats.data_ptr() if track_stats else 0) for o in self._out]
        self._cur = 0
        self._serial = 0
        self._pending = None
        self._keep = None
        self._run_time = _run_time_table(self.params)
        self._last_rew = None
        self.tstart = time.time()
        self._host = None
        self.reset()

    def _stream(self):
        return C.c_void_p(_lib.raw_stream(self._dev_index))

    def seed(self, seed=None):
        if seed is not None:
            self._seed = int(seed) & (2 ** 64 - 1)
        return [self._seed]

    def _coerce_actions(self, actions):
        n, A = self.num_envs, self.num_agents
        a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions))
        if self.continuous:
            if a.numel() != n * A * 2:
                raise AssertionError("actions of shape {} cannot match {} environments x {} cars".format(tuple(a.shape), n, A))
            a = a.reshape(n, A, 2).to(device=self.device, dtype=torch.float32, non_blocking=True)
        else:
            if a.numel() == n * A * 4 and (a.dim() >= 2 and a.shape[-1] == 4):
                a = a.reshape(n, A, 4).to(self.device).argmax(dim=-1)  # np.argmax(action_n[id]), highway_env.py:78
            elif a.numel() != n * A:
                raise AssertionError("actions of shape {} cannot match {} environments x {} cars".format(tuple(a.shape), n, A))
            a = a.reshape(n, A).to(device=self.device, dtype=torch.int32, non_blocking=True)
        return a.contiguous()

    def _wrap(self, t, as_bool=False):
        if self.return_numpy:
            x = t.cpu().numpy()
            return x.astype(bool) if as_bool else x
        return t.view(torch.bool) if as_bool else t  # the kernels write 0 / 1 bytes: a view, not a conversion launch per step

    def reset(self):
        assert not self.closed
        self._pending = None
        self._serial += 1
        out = self._out[self._cur]
        with _lib.launch_ctx(self._dev_index, "hw_ma_reset"):
            rc = self._lib.hw_ma_reset(C.byref(self._state_c), None, _as_ptr(out['obs']), self.num_envs, self._stream())
        _lib.check(rc, "hw_ma_reset")
        return self._wrap(out['obs'])

    def observe(self):
        obs = torch.empty(self.num_envs, self.num_agents, self.obs_dim, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self._lib.hw_ma_observe(C.byref(self._state_c), _as_ptr(obs), self.num_envs, self._stream())
        _lib.check(rc, "hw_ma_observe")
        return self._wrap(obs)

    def step_async(self, actions):
        assert not self.closed
        if self._pending is not None:
            raise AlreadySteppingError()
        a = self._coerce_actions(actions)
        self._cur ^= 1
        self._serial += 1
        with _lib.launch_ctx(self._dev_index, "hw_ma_step"):
            rc = self._lib.hw_ma_step(C.byref(self._state_c), _as_ptr(a), C.byref(self._out_c[self._cur]),
                                      C.byref(self.params), self._seed, self._env_id0, self.num_envs, self._flags,
                                      self._stream())
        _lib.check(rc, "hw_ma_step")
        self._keep = a
        self._pending = self._cur

    def step_wait(self):
        if self._pending is None:
            raise NotSteppingError()
        out = self._out[self._pending]
        self._pending = None
        infos = LazyInfos(self.num_envs, self._make_info_fetcher(out))
        return self._wrap(out['obs']), self._wrap(out['rew']), self._wrap(out['done'], True), infos

    def step_into(self, actions, obs, rew, done):
        """step() with caller-provided output tensors (float32 [N, A, 4A+14], float32 [N, A], uint8 [N, A] on this device,
        contiguous): the kernel writes straight into them - e.g. rows of a rollout's [T, N, A] buffers.  `actions` must
        already be an int32 [N, A] (float32 [N, A, 2] when continuous) tensor on this device.  Asynchronous; no infos
        (use episode_statistics())."""
        assert not self.closed and self._pending is None
        self._serial += 1
        out = _lib.HwMaOut(obs.data_ptr(), rew.data_ptr(), done.data_ptr(), self._out[0]['term'].data_ptr(),
                           self.stats.data_ptr() if self.stats is not None else 0)
        with _lib.launch_ctx(self._dev_index, "hw_ma_step"):
            rc = self._lib.hw_ma_step(C.byref(self._state_c), _as_ptr(actions), C.byref(out), C.byref(self.params),
                                      self._seed, self._env_id0, self.num_envs, self._flags, self._stream())
        _lib.check(rc, "hw_ma_step")

    def step_host(self, actions):
        n, A = self.num_envs, self.num_agents
        if self._host is None:
            ashape, adt = ((n, A, 2), torch.float32) if self.continuous else ((n, A), torch.int32)
            self._host = dict(act=torch.empty(ashape, dtype=adt).pin_memory(),
                              obs=torch.empty(n, A, self.obs_dim, dtype=torch.float32).pin_memory(),
                              rew=torch.empty(n, A, dtype=torch.float32).pin_memory(),
                              done=torch.empty(n, A, dtype=torch.uint8).pin_memory(),
                              d_act=torch.empty(ashape, dtype=adt, device=self.device))
        h = self._host
        h['act'].copy_(torch.as_tensor(np.asarray(actions)).reshape(h['act'].shape))
        self._cur ^= 1
        self._serial += 1
        with _lib.launch_ctx(self._dev_index, "hw_ma_step_host"):
            rc = self._lib.hw_ma_step_host(C.byref(self._state_c), _as_ptr(h['act']), _as_ptr(h['d_act']),
                                           C.byref(self._out_c[self._cur]), _as_ptr(h['obs']), _as_ptr(h['rew']),
                                           _as_ptr(h['done']), C.byref(self.params), self._seed, self._env_id0,
                                           n, self._flags, self._stream())
        _lib.check(rc, "hw_ma_step_host")
        return h['obs'].numpy(), h['rew'].numpy(), h['done'].numpy().view(np.bool_)

    def _make_info_fetcher(self, out):
        serial = self._serial
        rtab = lambda tick: float(self._run_time[min(int(tick), len(self._run_time) - 1)])

        def fetch(episodes_only):
            if self._serial - serial > (1 if episodes_only else 0):
                raise RuntimeError("the infos of a step must be read before the %s step() (they are built lazily from "
                                   "device tensors that a later step overwrites)" % ("second following" if episodes_only else "next"))
            done = out['done'].cpu().numpy().astype(bool)
            any_done = done.any(axis=1)
            term = out['term'].cpu().numpy()
            now = round(time.time() - self.tstart, 6)
            if episodes_only:
                return [dict(env=int(i), r=round(float(term[i, 3]), 6), l=int(term[i, 4]), t=now) for i in np.nonzero(any_done)[0]]
            rew = out['rew'].cpu().numpy()
            head = self.head.cpu().numpy().view(np.uint32)
            infos = []
            for i in range(self.num_envs):
                if any_done[i]:
                    rt, nco, nca = rtab(term[i, 0]), int(term[i, 1]), int(term[i, 2])
                else:
                    rt, nco, nca = rtab(head[0, i] & 0xFFFF), int(head[4, i]), int(head[5, i])
                info = {'n': [{"rewards": float(rew[i, a]), "run_time": rt, "num_obs_collisions": nco,
                               "num_agent_collisions": nca} for a in range(self.num_agents)]}
                if any_done[i]:
                    info['episode'] = {"r": round(float(term[i, 3]), 6), "l": int(term[i, 4]), "t": now}
                infos.append(info)
            return infos
        return fetch

    def episode_statistics(self, reset=False, group=None):
        if self.stats is None:
            raise RuntimeError("constructed with track_stats=False")
        s = self.stats.clone()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(s, op=torch.distributed.ReduceOp.SUM, group=group)
        if reset:
            self.stats.zero_()
        v = s.cpu().tolist()
        eps = v[0]
        return dict(episodes=eps, mean_return=(v[2] / 1000.0 / eps) if eps else float('nan'),
                    mean_length=(v[1] / eps) if eps else float('nan'), obs_collisions=v[3], timeouts=v[4],
                    agent_collisions=v[5], dropped_spawns=v[6])

    def diagnostics(self):
        """sticky per-env flags: bit 0 = a spawn was dropped (all slots live), bit 1 = a lane had no obstacle to observe"""
        return (self.head[3].view(torch.int32) >> 16) & 0xFF  # bits 24..28 of the word hold the persistent per-car done flags

    def state_dict(self):
        return dict(cars=self.cars.clone(), obst=self.obst.clone(), head=self.head.clone())

    def load_state_dict(self, sd):
        for k in ("cars", "obst", "head"):
            getattr(self, k).copy_(sd[k].to(self.device))
