"""ctypes front-end of the multi-agent CPU oracle (oracle/ma_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

from .oracle import (FLAG_CONTINUOUS, FLAG_NO_AUTORESET, FLAG_NO_INF_OBS, FLAG_QUANTISE_F32, _p, lib, sim_params)

FLAG_SHARED_REWARD = 16
DEFAULT_K = 16


class MaState(object):
    """canonical multi-agent state for n environments of A cars and up to K live obstacles"""

    FIELDS = ("car", "lane", "cflags", "tick", "leader", "nobs", "obst", "olane", "ofresh", "newest",
              "nc_obs", "nc_agent", "spawn_ctr", "ep_ret", "ep_len", "overflow", "done_mask")

    def __init__(self, n, A, K=DEFAULT_K):
        self.n, self.A, self.K = n, A, K
        self.car = np.zeros((n, A, 8), np.float64)   # px, py, rx, vx, vy, acc, steer, cruise
        self.lane = np.zeros((n, A), np.int32)
        self.cflags = np.zeros((n, A), np.int32)     # 1 left, 2 right, 4 do_accelerate, 8 do_maintain
        self.tick = np.zeros(n, np.int32)
        self.leader = np.zeros(n, np.int32)
        self.nobs = np.zeros(n, np.int32)
        self.obst = np.zeros((n, K, 4), np.float64)  # px, rx, vx, init_vx ; slots 0..nobs-1 in insertion order
        self.olane = np.zeros((n, K), np.int32)
        self.ofresh = np.zeros((n, K), np.int32)     # rect still the pygame default (0,0,76,38)
        self.newest = np.zeros((n, 3), np.int32)     # slot of lane_max_obs[lane], -1 once it has been retired
        self.nc_obs = np.zeros(n, np.int32)
        self.nc_agent = np.zeros(n, np.int32)
        self.spawn_ctr = np.zeros(n, np.uint32)
        self.ep_ret = np.zeros(n, np.float64)
        self.ep_len = np.zeros(n, np.int32)
        self.overflow = np.zeros(n, np.int32)
        self.done_mask = np.zeros(n, np.int32)        # bit a: car a finished in an earlier step (persists until reset, no auto-reset only)

    def copy(self):
        o = MaState(self.n, self.A, self.K)
        for f in self.FIELDS:
            setattr(o, f, getattr(self, f).copy())
        return o

    def select(self, idx):
        o = MaState(len(idx), self.A, self.K)
        for f in self.FIELDS:
            setattr(o, f, np.ascontiguousarray(getattr(self, f)[idx]))
        return o

    def quantise_f32(self):
        self.car = self.car.astype(np.float32).astype(np.float64)
        self.obst = self.obst.astype(np.float32).astype(np.float64)
        return self


def _state_args(st):
    return [C.c_int(st.n), C.c_int(st.A), C.c_int(st.K), _p(st.car, C.c_double), _p(st.lane, C.c_int32),
            _p(st.cflags, C.c_int32), _p(st.tick, C.c_int32), _p(st.leader, C.c_int32), _p(st.nobs, C.c_int32),
            _p(st.obst, C.c_double), _p(st.olane, C.c_int32), _p(st.ofresh, C.c_int32), _p(st.newest, C.c_int32),
            _p(st.nc_obs, C.c_int32), _p(st.nc_agent, C.c_int32)]


def ma_reset(st, mask=None, want_obs=True):
    obs = np.zeros((st.n, st.A, 4 * st.A + 14), np.float32) if want_obs else None
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
    L = lib()
    L.hw_oracle_ma_reset.restype = C.c_int
    rc = L.hw_oracle_ma_reset(*(_state_args(st) + [_p(st.ep_ret, C.c_double), _p(st.ep_len, C.c_int32),
                                                   _p(st.overflow, C.c_int32), _p(mask, C.c_uint8), _p(obs, C.c_float)]))
    assert rc == 0
    if mask is None:
        st.done_mask[:] = 0
    else:
        st.done_mask[np.asarray(mask).astype(bool)] = 0
    return obs


def ma_step(st, actions, continuous=False, real_time=False, seed=0, env_id0=0, inf_obs=True, autoreset=True,
            quantise=False, shared_reward=False, uniforms=None, want_next=False, nthreads=1):
    """advance `st` in place by one env.step of every environment"""
    n, A, K = st.n, st.A, st.K
    dt, tps, timeout = sim_params(real_time)
    flags = ((FLAG_CONTINUOUS if continuous else 0) | (0 if inf_obs else FLAG_NO_INF_OBS)
             | (0 if autoreset else FLAG_NO_AUTORESET) | (FLAG_QUANTISE_F32 if quantise else 0)
             | (FLAG_SHARED_REWARD if shared_reward else 0))
    act_d = act_c = None
    if continuous:
        act_c = np.ascontiguousarray(np.asarray(actions, np.float32).astype(np.float64).reshape(n, A, 2))
    else:
        act_d = np.ascontiguousarray(np.asarray(actions).reshape(n, A), np.int32)
    obs = np.zeros((n, A, 4 * A + 14), np.float32)
    rew = np.zeros((n, A), np.float64)
    done = np.zeros((n, A), np.uint8)
    term = np.zeros((n, 5), np.float64)
    u_used = np.zeros(n, np.int32)
    u_stride = 0
    if uniforms is not None:
        uniforms = np.ascontiguousarray(uniforms, np.float64).reshape(n, -1)
        u_stride = uniforms.shape[1]
    nxt = {}
    if want_next:
        nxt = dict(next_car=np.zeros((n, A, 8), np.float64), next_meta=np.zeros((2, n, A), np.int32),
                   next_obst=np.zeros((n, K, 4), np.float64), next_ometa=np.zeros((2, n, K), np.int32))
    L = lib()
    L.hw_oracle_ma_step.restype = C.c_int
    rc = L.hw_oracle_ma_step(*(_state_args(st) + [
        _p(st.spawn_ctr, C.c_uint32), _p(st.ep_ret, C.c_double), _p(st.ep_len, C.c_int32), _p(st.overflow, C.c_int32),
        _p(st.done_mask, C.c_int32), _p(act_d, C.c_int32), _p(act_c, C.c_double), _p(obs, C.c_float), _p(rew, C.c_double), _p(done, C.c_uint8),
        _p(term, C.c_double), C.c_double(dt), C.c_int(tps), C.c_int(timeout), C.c_uint64(seed), C.c_uint64(env_id0),
        C.c_int(flags), _p(uniforms, C.c_double), C.c_int(u_stride), _p(u_used, C.c_int32),
        _p(nxt.get("next_car"), C.c_double), _p(nxt.get("next_meta"), C.c_int32), _p(nxt.get("next_obst"), C.c_double),
        _p(nxt.get("next_ometa"), C.c_int32), C.c_int(nthreads)]))
    assert rc == 0
    out = dict(obs=obs, rew=rew, done=done.astype(bool), term=term, u_used=u_used)
    out.update(nxt)
    return out
