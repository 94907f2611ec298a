"""shared helpers of the parity tests: golden files <-> oracle states <-> packed device tensors"""
import os

import numpy as np

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def sa_state_from_golden(g):
    n = len(g["in_lane"])
    st = orc.SaState(n)
    st.car[:] = g["in_car"]; st.lane[:] = g["in_lane"]; st.flags[:] = g["in_flags"]; st.tick[:] = g["in_tick"]
    st.obst[:] = g["in_obst"]; st.ncoll[:] = g["in_ncoll"]; st.spawn_ctr[:] = g["in_spawn_ctr"]
    return st


def canonical(st):
    return {f: getattr(st, f) for f in st.FIELDS}


def sa_state_from_canonical(c):
    n = len(c["lane"])
    st = orc.SaState(n)
    for f in st.FIELDS:
        getattr(st, f)[...] = c[f]
    return st


def bits_equal(a, b):
    """bit-for-bit equality of two float arrays (distinguishes -0.0 and NaN payloads)"""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype, a.shape, b.shape)
    u = {4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
    return np.array_equal(a.view(u), b.view(u))


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32)


MA_STATE_FIELDS = ("car", "lane", "cflags", "tick", "leader", "nobs", "obst", "olane", "ofresh", "newest",
                   "nc_obs", "nc_agent", "spawn_ctr")


def ma_state_from_golden(g, prefix="in_"):
    from oracle import ma_oracle as mo
    n = len(g["env_id"])
    st = mo.MaState(n, int(g["num_agents"]), int(g["num_slots"]))
    for f in MA_STATE_FIELDS:
        if prefix + f in g:
            getattr(st, f)[...] = g[prefix + f].reshape(getattr(st, f).shape)
    return st
