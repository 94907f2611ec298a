"""GPU-resident vectorised highway environments behind the baselines `VecEnv` API.

Mirrors (names, argument meaning, return shapes, auto-reset and error behaviour):
  VecEnv ABC                     baselines/common/vec_env/__init__.py:26-135
  DummyVecEnv / SubprocVecEnv    baselines/common/vec_env/dummy_vec_env.py:6-82, subproc_vec_env.py:5-112
  Monitor episode accounting     baselines/bench/monitor.py:51-79
  make_env seeding               baselines/common/cmd_util.py:58-77 (env i is seeded seed + i)

One `step()` is one CUDA kernel launch over structure-of-arrays torch tensors (state never leaves the
GPU).  torch is plumbing here: it owns the device memory and the stream; the arithmetic is in
csrc/hw_sa.cu behind the C ABI of include/highway_b200.h.
"""
import ctypes as C
import time
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import _lib
from .spaces import Box, Discrete, highway_obs_bounds


class AlreadySteppingError(Exception):
    """step_async() called while a step is pending (vec_env/__init__.py:4-12)"""

    def __init__(self):
        Exception.__init__(self, 'already running an async step')


class NotSteppingError(Exception):
    """step_wait() called with no step pending (vec_env/__init__.py:15-23)"""

    def __init__(self):
        Exception.__init__(self, 'not running an async step')


class VecEnv(ABC):
    """An abstract asynchronous, vectorized environment (same contract as the reference's ABC)."""
    closed = False
    viewer = None
    metadata = {'render.modes': ['human', 'rgb_array']}

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def step_async(self, actions):
        pass

    @abstractmethod
    def step_wait(self):
        pass

    def close_extras(self):
        pass

    def close(self):
        if self.closed:
            return
        self.close_extras()
        self.closed = True

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def render(self, mode='human'):
        raise NotImplementedError("rendering is out of scope of the batched simulator")

    @property
    def unwrapped(self):
        return self


def _run_time_table(params):
    """run_time after t ticks, accumulated like the reference does (`self.run_time += dt`)"""
    out = np.zeros(params.timeout_tick + params.ticks_per_step + 2, np.float64)
    t = 0.0
    for i in range(1, len(out)):
        t += params.dt
        out[i] = t
    return out


class LazyInfos(object):
    """The `infos` sequence of a step, materialised on first access.

    Building N python dicts per step is the one part of the VecEnv contract that cannot scale to
    10^6 environments, so the tensors are only brought to the host when a trainer actually indexes
    or iterates the sequence.  `episodes()` is the cheap path trainers want: only the environments
    whose episode ended, as Monitor-style dicts.
    """

    def __init__(self, n, fetch):
        self._n = n
        self._fetch = fetch
        self._list = None

    def _materialise(self):
        if self._list is None:
            self._list = self._fetch(False)
        return self._list

    def episodes(self):
        return self._fetch(True)

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        return self._materialise()[i]

    def __iter__(self):
        return iter(self._materialise())

    def copy(self):
        return list(self._materialise())


def _as_ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class HighwayVecEnv(VecEnv):
    """N single-agent 3-lane highway environments stepped by one kernel launch.

    Drop-in for `make_vec_env(Config.env_id | env_cont_id, ...)`: the observation is the normalised
    float32[18] vector of HighwayEnv._get_state, the action space is Discrete(4) (0 LEFT, 1 RIGHT,
    2 ACCELERATE, 3 MAINTAIN) or Box(2) when `continuous`.  Constructor keywords mirror
    HighwayEnv.__init__(manual, inf_obs, save, render, real_time); manual/save/render are GUI and
    logging features of the interactive simulator and are rejected.

    Outputs are torch CUDA tensors by default (`return_numpy=True` gives host numpy arrays like the
    reference's VecEnvs).  Output buffers are double-buffered: a returned tensor stays valid until
    the second following step().
    """

    def __init__(self, num_envs, continuous=False, seed=0, device=None, manual=False, inf_obs=True, save=False,
                 render=False, real_time=False, env_id0=0, return_numpy=False, auto_reset=True, track_stats=True,
                 exact_steering=False):
        if manual or save or render:
            raise NotImplementedError("manual / save / render belong to the interactive pygame front-end")
        if num_envs < 1:
            raise ValueError("num_envs must be positive")
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.HighwayLibraryError("HighwayVecEnv needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.continuous = bool(continuous)
        self.real_time = bool(real_time)
        self.params = _lib.sim_params(real_time)
        self.return_numpy = bool(return_numpy)
        self._flags = ((_lib.FLAG_CONTINUOUS if continuous else 0) | (0 if inf_obs else _lib.FLAG_NO_INF_OBS)
                       | (0 if auto_reset else _lib.FLAG_NO_AUTORESET)
                       | (_lib.FLAG_EXACT_STEERING if exact_steering else 0))
        self._seed = int(seed) & (2 ** 64 - 1)
        self._env_id0 = int(env_id0)
        low, high = highway_obs_bounds(4, np.float32)
        obs_space = Box(low, high, dtype=np.float32)
        act_space = (Box(np.array([-1.0, -1.0]), np.array([1.0, 1.0]), dtype=np.float32) if continuous
                     else Discrete(4))
        VecEnv.__init__(self, int(num_envs), obs_space, act_space)

        n, dev = self.num_envs, self.device
        with torch.cuda.device(dev):
            self.car_a = torch.zeros(n, 4, dtype=torch.float32, device=dev)
            self.car_b = torch.zeros(n, 4, dtype=torch.float32, device=dev)
            self.obst = torch.zeros(3, n, 4, dtype=torch.float32, device=dev)
            self.stat = torch.zeros(n, 4, dtype=torch.int32, device=dev)
            self._out = [dict(obs=torch.zeros(n, _lib.SA_OBS_DIM, dtype=torch.float32, device=dev),
                              rew=torch.zeros(n, dtype=torch.float32, device=dev),
                              done=torch.zeros(n, dtype=torch.uint8, device=dev),
                              term=torch.zeros(n, 4, dtype=torch.int32, device=dev)) for _ in range(2)]
            self.stats = torch.zeros(8, dtype=torch.int64, device=dev) if track_stats else None
        self._state_c = _lib.HwSaState(self.car_a.data_ptr(), self.car_b.data_ptr(), self.obst.data_ptr(),
                                       self.stat.data_ptr())
        self._out_c = [_lib.HwSaOut(o['obs'].data_ptr(), o['rew'].data_ptr(), o['done'].data_ptr(),
                                    o['term'].data_ptr(), self.stats.data_ptr() if track_stats else 0)
                       for o in self._out]
        self._cur = 0
        self._serial = 0  # launches that changed the state so far (lazy infos check it)
        self._pending = None
        self._actions_keepalive = None
        self._run_time = _run_time_table(self.params)
        self.tstart = time.time()
        self._host = None  # pinned host buffers for the host-buffer path, allocated on first use
        self.reset()

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(_lib.raw_stream(self._dev_index))

    def seed(self, seed=None):
        """env i draws its respawn uniforms from Philox keyed by (seed, env_id0 + i) — the batched
        equivalent of `env.seed(seed + subrank)` (cmd_util.py:74)"""
        if seed is not None:
            self._seed = int(seed) & (2 ** 64 - 1)
        return [self._seed]

    def _coerce_actions(self, actions):
        n = self.num_envs
        want = torch.float32 if self.continuous else torch.int32
        shape = (n, 2) if self.continuous else (n,)
        if isinstance(actions, torch.Tensor):
            a = actions
        else:
            arr = np.asarray(actions)
            # a single non-batched action is accepted when there is one env (dummy_vec_env.py:33-44)
            if n == 1 and arr.shape != shape and arr.size == (2 if self.continuous else 1):
                arr = arr.reshape(shape)
            a = torch.from_numpy(np.ascontiguousarray(arr, np.float32 if self.continuous else np.int64))
        if a.numel() != n * (2 if self.continuous else 1):
            raise AssertionError("actions {} is either not a list or has a wrong size - cannot match to {} "
                                 "environments".format(tuple(a.shape), n))
        a = a.reshape(shape)
        if a.device != self.device or a.dtype != want or not a.is_contiguous():
            a = a.to(device=self.device, dtype=want, non_blocking=True).contiguous()
        return a

    def _wrap(self, t, as_bool=False):
        if self.return_numpy:
            a = t.cpu().numpy()
            return a.astype(bool) if as_bool else a
        return t.view(torch.bool) if as_bool else t  # the kernels write 0 / 1 bytes: a view, not a conversion launch per step

    # ------------------------------------------------------------------ VecEnv API
    def reset(self):
        """reset every environment; returns obs [N, 18] float32"""
        assert not self.closed
        self._pending = None
        self._serial += 1
        out = self._out[self._cur]
        with _lib.launch_ctx(self._dev_index, "hw_sa_reset"):
            rc = self._lib.hw_sa_reset(C.byref(self._state_c), None, _as_ptr(out['obs']), self.num_envs, self._stream())
        _lib.check(rc, "hw_sa_reset")
        return self._wrap(out['obs'])

    def reset_done(self, mask):
        """reset the environments whose mask entry is non-zero (only needed with auto_reset=False)"""
        mask = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
        assert mask.numel() == self.num_envs
        with _lib.launch_ctx(self._dev_index, "hw_sa_reset"):
            rc = self._lib.hw_sa_reset(C.byref(self._state_c), _as_ptr(mask), None, self.num_envs, self._stream())
        _lib.check(rc, "hw_sa_reset")
        return self.observe()

    def observe(self):
        """observation of the current state (HighwayEnv._get_state), no side effects"""
        obs = torch.empty(self.num_envs, _lib.SA_OBS_DIM, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self._lib.hw_sa_observe(C.byref(self._state_c), _as_ptr(obs), self.num_envs, self._stream())
        _lib.check(rc, "hw_sa_observe")
        return self._wrap(obs)

    def step_async(self, actions):
        assert not self.closed
        if self._pending is not None:
            raise AlreadySteppingError()
        a = self._coerce_actions(actions)
        self._cur ^= 1
        self._serial += 1
        with _lib.launch_ctx(self._dev_index, "hw_sa_step"):
            rc = self._lib.hw_sa_step(C.byref(self._state_c), _as_ptr(a), C.byref(self._out_c[self._cur]),
                                      C.byref(self.params), self._seed, self._env_id0, self.num_envs, self._flags,
                                      self._stream())
        _lib.check(rc, "hw_sa_step")
        self._actions_keepalive = a
        self._pending = self._cur

    def step_wait(self):
        if self._pending is None:
            raise NotSteppingError()
        out = self._out[self._pending]
        self._pending = None
        infos = LazyInfos(self.num_envs, self._make_info_fetcher(out))
        return self._wrap(out['obs']), self._wrap(out['rew']), self._wrap(out['done'], True), infos

    def step_into(self, actions, obs, rew, done):
        """step() with caller-provided output tensors (float32 [N, 18], float32 [N], uint8 [N] on this device, contiguous):
        the kernel writes the post-step observation, reward and done flag straight into them - e.g. rows of a rollout's
        [T, N] buffers - instead of the env's own double-buffered outputs.  `actions` must already be an int32 [N]
        (float32 [N, 2] when continuous) tensor on this device.  Asynchronous; no infos (use episode_statistics())."""
        assert not self.closed and self._pending is None
        self._serial += 1
        out = _lib.HwSaOut(obs.data_ptr(), rew.data_ptr(), done.data_ptr(), self._out[0]['term'].data_ptr(),
                           self.stats.data_ptr() if self.stats is not None else 0)
        with _lib.launch_ctx(self._dev_index, "hw_sa_step"):
            rc = self._lib.hw_sa_step(C.byref(self._state_c), _as_ptr(actions), C.byref(out), C.byref(self.params),
                                      self._seed, self._env_id0, self.num_envs, self._flags, self._stream())
        _lib.check(rc, "hw_sa_step")

    def rollout_random(self, steps, action_ctr0=0):
        """`steps` env.steps with in-kernel uniform random actions in one launch (benchmark / warm-up);
        returns the last step's (obs, rew, done)"""
        self._cur ^= 1
        self._serial += 1
        with _lib.launch_ctx(self._dev_index, "hw_sa_rollout_random"):
            rc = self._lib.hw_sa_rollout_random(C.byref(self._state_c), C.byref(self._out_c[self._cur]),
                                                C.byref(self.params), self._seed, self._env_id0, int(action_ctr0),
                                                int(steps), self.num_envs, self._flags, self._stream())
        _lib.check(rc, "hw_sa_rollout_random")
        out = self._out[self._cur]
        return self._wrap(out['obs']), self._wrap(out['rew']), self._wrap(out['done'], True)

    def step_host(self, actions, compact=False):
        """Host-buffer step through hw_sa_step_host: numpy actions in, numpy (obs, rew, done) out, with the
        host<->device copies and the stream drain inside the call.

        `compact=True` (HW_FLAG_COMPACT_OBS): the observation comes back as [N, 14] - the four columns that are the constant 0.5
        (11, 13, 15, 17: the lateral velocities, always zero in this simulator) never cross PCIe, which is what bounds this call;
        `expand_obs` restores the reference's [N, 18] layout on the host when a consumer wants it."""
        n = self.num_envs
        if self._host is None:
            self._host = dict(
                act=torch.empty((n, 2) if self.continuous else (n,),
                                dtype=torch.float32 if self.continuous else torch.int32).pin_memory(),
                obs=torch.empty(n, _lib.SA_OBS_DIM, dtype=torch.float32).pin_memory(),
                rew=torch.empty(n, dtype=torch.float32).pin_memory(),
                done=torch.empty(n, dtype=torch.uint8).pin_memory(),
                d_act=torch.empty((n, 2) if self.continuous else (n,),
                                  dtype=torch.float32 if self.continuous else torch.int32, device=self.device),
                aux=torch.cuda.Stream(device=self.device))  # second stream of the two-half pipeline, lives with the env
        h = self._host
        h['act'].copy_(torch.as_tensor(np.asarray(actions)).reshape(h['act'].shape))
        self._cur ^= 1
        self._serial += 1
        # compact rows reuse the same device and host buffers: [N, 14] is a prefix of their storage
        out = self._out_c[self._cur]
        h_obs = h['obs'].view(-1)[:n * _lib.SA_OBS_DIM_COMPACT].view(n, _lib.SA_OBS_DIM_COMPACT) if compact else h['obs']
        with _lib.launch_ctx(self._dev_index, "hw_sa_step_host"):
            rc = self._lib.hw_sa_step_host(C.byref(self._state_c), _as_ptr(h['act']), _as_ptr(h['d_act']),
                                           C.byref(out), _as_ptr(h_obs), _as_ptr(h['rew']),
                                           _as_ptr(h['done']), C.byref(self.params), self._seed, self._env_id0,
                                           n, self._flags | (_lib.FLAG_COMPACT_OBS if compact else 0), self._stream(),
                                           C.c_void_p(h['aux'].cuda_stream))
        _lib.check(rc, "hw_sa_step_host")
        return h_obs.numpy(), h['rew'].numpy(), h['done'].numpy().view(np.bool_)  # the kernel writes 0 / 1

    @staticmethod
    def expand_obs(compact_obs, out=None):
        """[N, 14] compact observation rows -> the reference's [N, 18] layout (columns 11, 13, 15, 17 = 0.5)"""
        c = np.asarray(compact_obs)
        if out is None:
            out = np.empty((c.shape[0], _lib.SA_OBS_DIM), np.float32)
        out[:, 11:18:2] = 0.5
        out[:, list(_lib.SA_COMPACT_COLUMNS)] = c
        return out

    # ------------------------------------------------------------------ infos / statistics
    def _make_info_fetcher(self, out):
        serial = self._serial
        rt = lambda tick: float(self._run_time[min(int(tick), len(self._run_time) - 1)])

        def fetch(episodes_only):
            # terminal rows come from this step's own (double-buffered) outputs; the non-terminal rows are read from the
            # live state, so they are only this step's while no later step has been launched
            if self._serial - serial > (1 if episodes_only else 0):
                raise RuntimeError("the infos of a step must be read before the %s step() (they are built lazily from "
                                   "device tensors that a later step overwrites)" % ("second following" if episodes_only else "next"))
            done = out['done'].cpu().numpy().astype(bool)
            term = out['term'].cpu().numpy()
            now = round(time.time() - self.tstart, 6)
            if episodes_only:
                idx = np.nonzero(done)[0]
                return [dict(env=int(i), r=round(term[i, 2] / 1000.0, 6), l=int(term[i, 3]), t=now) for i in idx]
            stat = self.stat.cpu().numpy()
            tick = (self.car_b[:, 3].contiguous().view(torch.int32) & 0xFFFF).cpu().numpy()
            infos = []
            for i in range(self.num_envs):
                if done[i]:
                    info = {"run_time": rt(term[i, 0]), "num_obs_collisions": int(term[i, 1]),
                            "num_agent_collisions": 0,
                            "episode": {"r": round(term[i, 2] / 1000.0, 6), "l": int(term[i, 3]), "t": now}}
                else:
                    info = {"run_time": rt(tick[i]), "num_obs_collisions": int(stat[i, 0]),
                            "num_agent_collisions": 0}
                infos.append(info)
            return infos
        return fetch

    def episode_statistics(self, reset=False, group=None):
        """(episodes, mean return, mean length, collisions, time-outs) accumulated by the kernel since the
        last reset of the counters; summed across ranks with one all_reduce when torch.distributed is
        initialised (replaces Monitor + the trainers' deque means / MPI stats all-reduce)."""
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
                    mean_length=(v[1] / eps) if eps else float('nan'), obs_collisions=v[3], timeouts=v[4])

    # ------------------------------------------------------------------ state exchange (tests, checkpoints)
    def state_dict(self):
        return dict(car_a=self.car_a.clone(), car_b=self.car_b.clone(), obst=self.obst.clone(), stat=self.stat.clone())

    def load_state_dict(self, sd):
        for k in ("car_a", "car_b", "obst", "stat"):
            getattr(self, k).copy_(sd[k].to(self.device))
