"""ctypes front-end of the CPU oracle (oracle/sa_oracle.c, oracle/ma_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never imported by gym_highway_b200 (the product).

State is exchanged in a *canonical* unpacked layout (float64 / int32 numpy arrays, one row per
environment); gym_highway_b200.layout converts between it and the packed fp32 SoA the kernels use.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhighway_oracle.so")

FLAG_CONTINUOUS = 1
FLAG_NO_INF_OBS = 2
FLAG_NO_AUTORESET = 4
FLAG_QUANTISE_F32 = 8

LANE_C = (1.25, 3.75, 6.25)


def build(force=False):
    """compile the C restatement with gcc (seconds)"""
    srcs = [os.path.join(_HERE, f) for f in ("sa_oracle.c", "ma_oracle.c", "philox.h", "tan_small.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s", "clean", "all"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.hw_oracle_sa_step.restype = C.c_int
        _lib.hw_oracle_sa_reset.restype = C.c_int
        _lib.hw_oracle_sa_observe.restype = C.c_int
    return _lib


def set_tan(poly):
    """which tangent the restatement evaluates: False = libm's (the reference's, default), True = the kernels' polynomial
    (diagnosis of soak mismatches only, see tan_small.h)"""
    lib().hw_oracle_set_tan(1 if poly else 0)


def sim_params(real_time=False):
    """(dt, ticks_per_step, timeout_tick) exactly as the reference derives them
    (multi_lane_sim.py:358,371,888,910-913; highway_env.py:101)."""
    ticks = 60.0 if real_time else 36.0
    dt = 1.0 / ticks
    run_time, t = 0.0, 0
    while not run_time >= 60:
        run_time += dt
        t += 1
    return dt, int(ticks / 4), t


def _p(a, ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


class SaState(object):
    """canonical single-agent state for n environments"""

    FIELDS = ("car", "lane", "flags", "tick", "obst", "ncoll", "spawn_ctr", "ep_ret_milli", "ep_len")

    def __init__(self, n):
        self.n = n
        self.car = np.zeros((n, 7), np.float64)      # px, py, rx, vx, acc, steer, cruise
        self.lane = np.zeros(n, np.int32)
        self.flags = np.zeros(n, np.int32)           # 1 left, 2 right, 4 do_accelerate, 8 do_maintain
        self.tick = np.zeros(n, np.int32)
        self.obst = np.zeros((n, 3, 4), np.float64)  # px, rx, vx, init_vx ; slot k drives in lane k+1
        self.ncoll = np.zeros(n, np.int32)
        self.spawn_ctr = np.zeros(n, np.uint32)
        self.ep_ret_milli = np.zeros(n, np.int64)
        self.ep_len = np.zeros(n, np.int32)

    def copy(self):
        o = SaState(self.n)
        for f in self.FIELDS:
            setattr(o, f, getattr(self, f).copy())
        return o

    def quantise_f32(self):
        self.car = self.car.astype(np.float32).astype(np.float64)
        self.obst = self.obst.astype(np.float32).astype(np.float64)
        return self

    def select(self, idx):
        o = SaState(len(idx))
        for f in self.FIELDS:
            setattr(o, f, np.ascontiguousarray(getattr(self, f)[idx]))
        return o


def sa_reset(state, mask=None, want_obs=True):
    obs = np.zeros((state.n, 18), np.float32) if want_obs else None
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
    lib().hw_oracle_sa_reset(C.c_int(state.n), _p(state.car, C.c_double), _p(state.lane, C.c_int32),
                             _p(state.flags, C.c_int32), _p(state.tick, C.c_int32), _p(state.obst, C.c_double),
                             _p(state.ncoll, C.c_int32), _p(state.ep_ret_milli, C.c_int64),
                             _p(state.ep_len, C.c_int32), _p(mask, C.c_uint8), _p(obs, C.c_float))
    return obs


def sa_observe(state):
    obs = np.zeros((state.n, 18), np.float32)
    lib().hw_oracle_sa_observe(C.c_int(state.n), _p(state.car, C.c_double), _p(state.obst, C.c_double),
                               _p(obs, C.c_float))
    return obs


def sa_step(state, actions, continuous=False, real_time=False, seed=0, env_id0=0, inf_obs=True,
            autoreset=True, quantise=False, uniforms=None, want_next=False, nthreads=1):
    """advance `state` in place by one env.step; returns dict(obs, rew, done, term[, next_*, u_used])"""
    n = state.n
    dt, tps, timeout = sim_params(real_time)
    flags = ((FLAG_CONTINUOUS if continuous else 0) | (0 if inf_obs else FLAG_NO_INF_OBS)
             | (0 if autoreset else FLAG_NO_AUTORESET) | (FLAG_QUANTISE_F32 if quantise else 0))
    act_d = act_c = None
    if continuous:
        # fp32-representable actions evaluated in fp64 (reference-era NumPy promoted to float64)
        act_c = np.ascontiguousarray(np.asarray(actions, np.float32).astype(np.float64).reshape(n, 2))
    else:
        act_d = np.ascontiguousarray(np.asarray(actions).reshape(n), np.int32)
    obs = np.zeros((n, 18), np.float32)
    rew = np.zeros(n, np.float64)
    done = np.zeros(n, np.uint8)
    term = np.zeros((n, 4), np.int64)
    u_used = np.zeros(n, np.int32)
    u_stride = 0
    if uniforms is not None:
        uniforms = np.ascontiguousarray(uniforms, np.float64).reshape(n, -1)
        u_stride = uniforms.shape[1]
    nxt = {}
    if want_next:
        nxt = dict(next_car=np.zeros((n, 7), np.float64), next_lft=np.zeros((n, 3), np.int32),
                   next_obst=np.zeros((n, 3, 4), np.float64))
    rc = lib().hw_oracle_sa_step(
        C.c_int(n), _p(state.car, C.c_double), _p(state.lane, C.c_int32), _p(state.flags, C.c_int32),
        _p(state.tick, C.c_int32), _p(state.obst, C.c_double), _p(state.ncoll, C.c_int32),
        _p(state.spawn_ctr, C.c_uint32), _p(state.ep_ret_milli, C.c_int64), _p(state.ep_len, C.c_int32),
        _p(act_d, C.c_int32), _p(act_c, C.c_double), _p(obs, C.c_float), _p(rew, C.c_double),
        _p(done, C.c_uint8), _p(term, C.c_int64), C.c_double(dt), C.c_int(tps), C.c_int(timeout),
        C.c_uint64(seed), C.c_uint64(env_id0), C.c_int(flags), _p(uniforms, C.c_double), C.c_int(u_stride),
        _p(u_used, C.c_int32), _p(nxt.get("next_car"), C.c_double), _p(nxt.get("next_lft"), C.c_int32),
        _p(nxt.get("next_obst"), C.c_double), C.c_int(nthreads))
    assert rc == 0
    out = dict(obs=obs, rew=rew, done=done.astype(bool), term=term, u_used=u_used)
    out.update(nxt)
    return out


def spawn_uniforms(seed, env_id, spawn_ctr):
    out = (C.c_double * 2)()
    lib().hw_oracle_spawn_uniforms(C.c_uint64(seed), C.c_uint64(env_id), C.c_uint32(spawn_ctr), out)
    return float(out[0]), float(out[1])
