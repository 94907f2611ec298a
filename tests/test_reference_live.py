"""CPU, build container only (skipped where /root/reference is absent, e.g. on the GPU box): the committed golden files are
what the UNMODIFIED reference produces here, and the C oracle tracks the live reference on fresh trajectories too.

The generators (oracle/gen_golden.py, gen_golden_ma.py) drive the reference's own env.step under the pygame/gym stand-ins;
a recorded transition depends only on (generator seed, transition index), so re-generating a prefix of a file must give the
file's own rows, bit for bit."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref_harness as rh
from helpers import bits_equal, load_golden

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present")


def _same_rows(fresh, committed, n):
    for k, v in fresh.items():
        if k not in committed:   # metadata keys added to the generator after a file was recorded
            continue
        w = committed[k]
        if np.ndim(v) == 0:
            assert v == w, k
            continue
        a, b = np.asarray(v), np.asarray(w)[:n]
        assert a.shape == b.shape, k
        if a.dtype.kind == "f":
            assert bits_equal(np.ascontiguousarray(a), np.ascontiguousarray(b.astype(a.dtype))), k
        else:
            assert np.array_equal(a, b), k


@pytest.mark.parametrize("name,continuous,seed", [("sa_discrete", False, 11), ("sa_continuous", True, 12)])
def test_committed_sa_golden_files_are_the_reference(name, continuous, seed):
    from oracle import gen_golden
    n = 120
    _same_rows(gen_golden.gen_sa(continuous, n, seed), load_golden(name), n)


@pytest.mark.parametrize("name,A,continuous,seed", [("ma3_discrete", 3, False, 103), ("ma5_continuous", 5, True, 205)])
def test_committed_ma_golden_files_are_the_reference(name, A, continuous, seed):
    from oracle import gen_golden_ma
    n = 60
    _same_rows(gen_golden_ma.gen_ma(A, continuous, n, seed), load_golden(name), n)


@pytest.mark.parametrize("continuous", [False, True])
def test_oracle_tracks_the_live_reference_on_a_fresh_trajectory(continuous):
    """new seed, so none of these transitions is in a committed file"""
    from oracle import gen_golden
    g = gen_golden.gen_sa(continuous, 250, 977)
    n = len(g["out_done"])
    st = orc.SaState(n)
    st.car[:] = g["in_car"]; st.lane[:] = g["in_lane"]; st.flags[:] = g["in_flags"]; st.tick[:] = g["in_tick"]
    st.obst[:] = g["in_obst"]; st.ncoll[:] = g["in_ncoll"]; st.spawn_ctr[:] = g["in_spawn_ctr"]
    out = orc.sa_step(st, g["action"], continuous=continuous, seed=977, env_id0=0, autoreset=False, want_next=True)
    assert bits_equal(out["next_car"], g["out_car"]) and bits_equal(out["next_obst"], g["out_obst"])
    assert np.array_equal(out["next_lft"], g["out_lft"]) and bits_equal(out["obs"], g["out_obs"])
    assert bits_equal(out["rew"], g["out_rew"].astype(np.float64)) and np.array_equal(out["done"], g["out_done"])
