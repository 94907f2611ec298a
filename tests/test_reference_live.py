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


def test_soak_summary_is_committed_and_clean():
    """oracle/soak_reference.py: >= 1e5 fresh lock-step transitions (SA + MA, A in {1, 3, 5}), zero mismatching words"""
    import json, os
    from oracle import soak_reference as soak
    s = json.load(open(soak.SUMMARY))
    assert s["total_transitions"] >= 100000 and s["total_mismatching_words"] == 0
    names = {j["name"] for j in s["jobs"]}
    assert {"ma1_discrete", "ma3_continuous", "ma5_discrete", "ma5_continuous", "sa_discrete_0", "sa_continuous_0"} <= names
    assert all(j["mismatching_words"] == 0 and j["episode_ends"] > 0 for j in s["jobs"])


def test_soak_jobs_rerun_clean_on_a_slice():
    """every soak job, re-run here on a short slice with its own seed: still bit-identical to the live reference"""
    from oracle import soak_reference as soak
    for name, kind, kw in soak.job_list(900, 450):
        r = soak.run_job((name, kind, kw))
        assert r["mismatching_words"] == 0, r


def test_edge_and_vec_goldens_are_the_reference():
    """the forced edge-case files and the VecEnv / runner / memory files regenerate bit for bit from the live reference"""
    from oracle import gen_golden_edge as ge, gen_golden_vec as gv
    _same_rows(ge.gen_sa_edge(False, 41), load_golden("sa_discrete_edge"), 10 ** 9)
    _same_rows(ge.gen_ma_edge(5, False, 43), load_golden("ma5_discrete_edge"), 10 ** 9)
    _same_rows(ge.gen_ma_edge(3, True, 44), load_golden("ma3_continuous_edge"), 10 ** 9)
    g, c = gv.record_vec(4, 120, 26, num_agents=5, shared_reward=True), load_golden("vec_ma5_shared")
    for k in ("obs0", "obs", "rew", "done", "ep_l", "run_time", "n_obs_coll", "n_agent_coll", "action"):
        assert np.array_equal(np.asarray(g[k]), c[k]), k
    assert np.array_equal(np.nan_to_num(g["ep_r"]), np.nan_to_num(c["ep_r"]))
    g, c = gv.record_a2c(8, 5, 60, 33), load_golden("runner_a2c_discrete")
    for k in ("obs", "rewards", "masks", "actions", "values"):
        assert np.array_equal(np.asarray(g[k]), c[k]), k
