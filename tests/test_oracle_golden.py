"""CPU: the C oracle must reproduce the reference-generated golden transitions bit for bit."""
import numpy as np
import pytest

from oracle import oracle as orc
from helpers import bits_equal, load_golden, sa_state_from_golden


def test_philox_known_answers_via_spawn_stream():
    # the uniform pair of (seed, env, ctr) must be reproducible and in [0, 1)
    u1, u2 = orc.spawn_uniforms(0, 0, 0)
    v1, v2 = orc.spawn_uniforms(0, 0, 0)
    assert (u1, u2) == (v1, v2) and 0.0 <= u1 < 1.0 and 0.0 <= u2 < 1.0
    assert orc.spawn_uniforms(0, 0, 1) != (u1, u2) and orc.spawn_uniforms(0, 1, 0) != (u1, u2)
    assert orc.spawn_uniforms(1, 0, 0) != (u1, u2)


@pytest.mark.parametrize("name", ["sa_discrete", "sa_continuous", "sa_discrete_realtime", "sa_discrete_noinf", "sa_discrete_edge", "sa_continuous_edge"])
def test_sa_oracle_matches_reference_golden(name):
    g = load_golden(name)
    cont, rt, seed = bool(g["continuous"]), bool(g["real_time"]), int(g["seed"])
    n = len(g["out_done"])
    st = sa_state_from_golden(g)
    assert np.array_equal(g["env_id"], np.arange(n))  # transition i was recorded as environment i
    out_all = orc.sa_step(st, g["action"], continuous=cont, real_time=rt, seed=seed, env_id0=0,
                          autoreset=False, want_next=True, inf_obs=bool(g.get("inf_obs", True)))
    ncoll, ctr = st.ncoll, st.spawn_ctr
    assert bits_equal(out_all["next_car"], g["out_car"])
    assert bits_equal(out_all["next_obst"], g["out_obst"])
    assert np.array_equal(out_all["next_lft"], g["out_lft"])
    assert bits_equal(out_all["obs"], g["out_obs"])
    assert bits_equal(out_all["rew"], g["out_rew"].astype(np.float64))
    assert np.array_equal(out_all["done"], g["out_done"])
    assert np.array_equal(ncoll, g["out_ncoll"])
    assert np.array_equal(ctr, g["out_spawn_ctr"].astype(np.uint32))
    assert g["out_done"].sum() >= 3


def test_sa_reset_observation_anchor():
    # SURVEY.md section 8c sanity anchor (seed 0 reset observation of the reference)
    st = orc.SaState(3)
    obs = orc.sa_reset(st)
    want = np.array([0.5, 0.5, 0.0166667, 0.46875, 0.1666667, 0.34375, 0.125, 0.5, 0.0, 0.65625, 0.5, 0.5,
                     0.675, 0.5, 0.625, 0.5, 0.65, 0.5], np.float32)
    assert np.allclose(obs, want[None], atol=1e-6)
    assert st.lane.tolist() == [2, 2, 2] and st.tick.tolist() == [0, 0, 0]


def test_sa_first_steps_anchor():
    # rewards of actions [2,2,2,0,3,1] from reset, time-out at step 240
    st = orc.SaState(1)
    orc.sa_reset(st)
    rews = []
    for a in [2, 2, 2, 0, 3, 1]:
        o = orc.sa_step(st, [a], seed=0)
        rews.append(float(o["rew"][0]))
    assert rews == [-0.998, -0.993, -0.985, -0.974, -0.912, -0.849]
    st = orc.SaState(1)
    orc.sa_reset(st)
    for t in range(240):
        o = orc.sa_step(st, [3], seed=0, autoreset=False)
        assert bool(o["done"][0]) == (t == 239)
    assert st.tick[0] == 2160


def test_sa_threads_and_quantise_are_consistent():
    rng = np.random.RandomState(0)
    n = 257
    a = orc.SaState(n); orc.sa_reset(a)
    b = a.copy()
    for t in range(60):
        act = rng.randint(0, 4, n)
        oa = orc.sa_step(a, act, seed=3, quantise=True, nthreads=1)
        ob = orc.sa_step(b, act, seed=3, quantise=True, nthreads=4)
        assert bits_equal(oa["obs"], ob["obs"]) and np.array_equal(oa["done"], ob["done"])
    assert bits_equal(a.car, b.car) and bits_equal(a.obst, b.obst)
    assert np.array_equal(a.car, a.car.astype(np.float32).astype(np.float64))


MA_GOLDEN = ["ma1_discrete", "ma1_continuous", "ma3_discrete", "ma3_continuous", "ma5_discrete", "ma5_continuous",
             "ma5_discrete_shared", "ma3_discrete_realtime", "ma3_continuous_noinf",
             "ma5_discrete_edge", "ma3_continuous_edge", "ma2_discrete_edge"]  # *_edge: forced rare branches (oracle/gen_golden_edge.py)


@pytest.mark.parametrize("name", MA_GOLDEN)
def test_ma_oracle_matches_reference_golden(name):
    from oracle import ma_oracle as mo
    from helpers import ma_state_from_golden
    g = load_golden(name)
    n = len(g["env_id"])
    assert np.array_equal(g["env_id"], np.arange(n))
    st = ma_state_from_golden(g)
    out = mo.ma_step(st, g["action"], continuous=bool(g["continuous"]), seed=int(g["seed"]), env_id0=0,
                     autoreset=False, shared_reward=bool(g["shared_reward"]), real_time=bool(g["real_time"]),
                     inf_obs=bool(g.get("inf_obs", True)))
    for f in ("car", "obst"):
        assert bits_equal(getattr(st, f), g["out_" + f].astype(np.float64)), f
    for f in ("lane", "cflags", "tick", "leader", "nobs", "olane", "ofresh", "newest", "nc_obs", "nc_agent"):
        assert np.array_equal(getattr(st, f), g["out_" + f].reshape(getattr(st, f).shape)), f
    assert np.array_equal(st.spawn_ctr, g["out_spawn_ctr"].astype(np.uint32))
    assert bits_equal(out["obs"], g["out_obs"])
    assert bits_equal(out["rew"], g["out_rew"].astype(np.float64))
    assert np.array_equal(out["done"], g["out_done"])
    assert st.overflow.max() == 0 and g["out_done"].any(axis=1).sum() >= 2


def test_ma_reset_and_threads():
    from oracle import ma_oracle as mo
    for A in (1, 3, 5):
        a = mo.MaState(130, A); obs = mo.ma_reset(a)
        assert obs.shape == (130, A, 4 * A + 14) and np.isfinite(obs).all()
        b = a.copy()
        rng = np.random.RandomState(A)
        for t in range(40):
            act = rng.randint(0, 4, (130, A))
            oa = mo.ma_step(a, act, seed=1, quantise=True, nthreads=1)
            ob = mo.ma_step(b, act, seed=1, quantise=True, nthreads=4)
            assert bits_equal(oa["obs"], ob["obs"]) and np.array_equal(oa["done"], ob["done"])
        assert bits_equal(a.car, b.car) and a.overflow.max() == 0


def test_edge_goldens_contain_the_forced_branches():
    """oracle/gen_golden_edge.py: the rare branches the files were built for are really in them (reference outputs)"""
    from helpers import load_golden
    g = load_golden("ma5_discrete_edge")
    spawned = g["out_spawn_ctr"].astype(np.int64) - g["in_spawn_ctr"].astype(np.int64)
    hit_obs = g["out_nc_obs"] - g["in_nc_obs"]
    one_tick = (g["out_tick"] - g["in_tick"]) == 1
    # phantom default-rect collision: an obstacle hit on the very tick of a spawn, by a car nowhere near a live obstacle
    assert int(np.sum((spawned > 0) & (hit_obs > 0) & one_tick)) >= 4
    assert int(np.sum((spawned > 0) & (hit_obs == 0) & ~g["out_done"].any(axis=1))) >= 2      # rect.x = 76: touching only
    assert g["in_nobs"].max() == 10 and int(np.sum(g["in_nobs"] - g["out_nobs"] >= 3)) >= 3     # retire-heavy lists
    assert int(np.sum(g["out_done"].all(axis=1) & (g["out_tick"] == 2160))) >= 3               # time-out finishes every car
    assert int(np.sum(g["out_nc_agent"] - g["in_nc_agent"] > 0)) >= 4                          # car-car contact
    s = load_golden("sa_discrete_edge")
    assert int(s["out_done"].sum()) >= 10 and (s["out_lft"][:, 2] == 2160 + 9).any()
    c = load_golden("sa_continuous_edge")
    assert (c["out_car"][:, 1] == 1.0).any() and (c["out_car"][:, 1] == 7.0).any()


def test_sa_oracle_matches_reference_on_unreachable_states():
    """The y clamps of the discrete update_d (multi_lane_sim.py:219-222) cannot trigger from reset(): discrete steering is
    non-zero only while a lane change is in flight.  The reference still defines them; these rows start from states with
    left-over steering and pin the oracle there (the CUDA kernel's later ticks are specialised to reachable states)."""
    from helpers import load_golden
    g = load_golden("sa_discrete_offmanifold")
    st = sa_state_from_golden(g)
    o = orc.sa_step(st, g["action"], seed=int(g["seed"]), env_id0=0, autoreset=False, want_next=True)
    assert bits_equal(o["next_car"], g["out_car"]) and bits_equal(o["obs"], g["out_obs"]) and np.array_equal(o["next_lft"], g["out_lft"])
    y = g["out_car"][:, 1]
    assert (y == 0.0).any() and (y == 7.0).any()


def test_ma_extension_start_grid_is_collision_free():
    """the invented continuation of the start grid for 6..8 cars (no reference behind it): distinct, non-overlapping rects,
    observation of length 4A + 14, and the world runs"""
    from oracle import ma_oracle as mo
    for A in (6, 7, 8):
        st = mo.MaState(64, A)
        obs = mo.ma_reset(st)
        assert obs.shape == (64, A, 4 * A + 14) and np.isfinite(obs).all()
        xy = {(float(st.car[0, a, 0]), int(st.lane[0, a])) for a in range(A)}
        assert len(xy) == A
        rng = np.random.RandomState(A)
        ends = 0
        for t in range(60):
            o = mo.ma_step(st, rng.randint(0, 4, (64, A)), seed=2, quantise=True)
            ends += int(o["done"].any(axis=1).sum())
            assert t > 0 or not o["done"].any()      # nobody collides on the first step from the start grid
        assert ends > 0 and st.overflow.max() == 0
