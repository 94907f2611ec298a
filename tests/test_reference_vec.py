"""CPU: the C oracle, driven with auto-reset and Monitor-style episode accounting, must reproduce the trajectories the
reference's own VecEnv stack produced (DummyVecEnv / DummyVecMAEnv over Monitor over the unmodified envs;
tests/golden/vec_*.npz written by oracle/gen_golden_vec.py) - observation of a finished env = post-reset observation,
terminal reward / done / info, info['episode'] = {r: round(sum, 6), l}."""
import numpy as np
import pytest

from oracle import ma_oracle as mo
from oracle import oracle as orc
from helpers import check_vec_golden, load_golden

_RT = {}


def _run_time(tick):
    if not _RT:
        t = 0.0
        _RT[0] = 0.0
        for i in range(1, 2400):
            t += 1.0 / 36.0
            _RT[i] = t
    return _RT[int(tick)]


@pytest.mark.parametrize("name", ["vec_sa_discrete", "vec_sa_continuous"])
def test_sa_oracle_matches_reference_vecenv(name):
    g = load_golden(name)
    n, cont, seed = g["done"].shape[1], bool(g["continuous"]), int(g["seed"])
    st = orc.SaState(n)

    def step(t, a):
        o = orc.sa_step(st, a, continuous=cont, seed=seed, env_id0=0, quantise=True, autoreset=True)
        eps = {int(i): (round(o["term"][i, 2] / 1000.0, 6), int(o["term"][i, 3])) for i in np.nonzero(o["done"])[0]}
        rt = [_run_time(o["term"][i, 0] if o["done"][i] else st.tick[i]) for i in range(n)]
        nc = [int(o["term"][i, 1]) if o["done"][i] else int(st.ncoll[i]) for i in range(n)]
        return o["obs"], o["rew"].astype(np.float32), o["done"], eps, (rt, nc, np.zeros(n, int))

    ends, timeouts = check_vec_golden(g, lambda: orc.sa_reset(st), step)
    assert ends >= 10 and timeouts >= 2


@pytest.mark.parametrize("name", ["vec_ma3_discrete", "vec_ma5_discrete", "vec_ma3_continuous", "vec_ma5_shared"])
def test_ma_oracle_matches_reference_vecenv(name):
    g = load_golden(name)
    n, A = g["done"].shape[1], int(g["num_agents"])
    cont, seed, shared = bool(g["continuous"]), int(g["seed"]), bool(g["shared_reward"])
    st = mo.MaState(n, A)

    def step(t, a):
        o = mo.ma_step(st, a, continuous=cont, seed=seed, env_id0=0, quantise=True, autoreset=True, shared_reward=shared)
        ended = o["done"].any(axis=1)
        eps = {int(i): (round(float(o["term"][i, 3]), 6), int(o["term"][i, 4])) for i in np.nonzero(ended)[0]}
        rt = [_run_time(o["term"][i, 0] if ended[i] else st.tick[i]) for i in range(n)]
        no = [int(o["term"][i, 1]) if ended[i] else int(st.nc_obs[i]) for i in range(n)]
        na = [int(o["term"][i, 2]) if ended[i] else int(st.nc_agent[i]) for i in range(n)]
        assert not st.overflow.any()
        return o["obs"], o["rew"].astype(np.float32), o["done"], eps, (rt, no, na)

    ends, timeouts = check_vec_golden(g, lambda: mo.ma_reset(st), step)
    assert ends >= 5
    if name != "vec_ma5_shared":
        assert timeouts >= 1
