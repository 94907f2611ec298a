"""GPU parity: the sm_100a single-agent kernels (called through the C ABI) against the C oracle and the
reference-generated golden transitions.  Bar: every discrete quantity (lane, mode flags, tick,
collision counters, done, reset) and - because the kernel evaluates the tick in fp64 exactly like the
oracle - every float (state, observation, reward) bit-exact; the float comparisons additionally
assert the contract tolerance of 1e-5 relative so that a last-bit libm difference would be reported
as such rather than hidden."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from helpers import bits_equal, canonical, f32, load_golden, sa_state_from_canonical, sa_state_from_golden

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star: positions, velocities and rewards within 1e-5 relative per step


def _env(n, **kw):
    from gym_highway_b200 import HighwayVecEnv
    return HighwayVecEnv(n, **kw)


def _load_state(env, st):
    from gym_highway_b200.layout import pack_sa
    p = pack_sa(canonical(st))
    env.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})


def _read_state(env):
    from gym_highway_b200.layout import unpack_sa
    sd = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
    return sa_state_from_canonical(unpack_sa(sd["car_a"], sd["car_b"], sd["obst"], sd["stat"]))


def _assert_close_and_report(name, got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    denom = np.maximum(np.abs(want), 1e-3)
    rel = np.abs(got - want) / denom
    assert rel.max() <= REL_TOL, "%s: max rel err %.3g" % (name, rel.max())
    nbad = int(np.sum(f32(got).view(np.uint32) != f32(want).view(np.uint32)))
    assert nbad == 0, "%s: %d of %d values differ in the last bits (max rel %.3g)" % (name, nbad, got.size, rel.max())


def test_reset_observation_matches_oracle():
    env = _env(1000)
    obs = env.reset().cpu().numpy()
    st = orc.SaState(1000)
    want = orc.sa_reset(st)
    assert bits_equal(obs, want)
    got = _read_state(env)
    assert bits_equal(got.car, st.car) and bits_equal(got.obst, st.obst)
    assert np.array_equal(got.lane, st.lane) and np.array_equal(got.tick, st.tick)


@pytest.mark.parametrize("name", ["sa_discrete", "sa_continuous", "sa_discrete_realtime", "sa_discrete_noinf", "sa_discrete_edge", "sa_continuous_edge"])
def test_golden_transitions(name):
    """reference-recorded transitions: one env per transition, env_id taken from the file"""
    g = load_golden(name)
    cont, rt, seed = bool(g["continuous"]), bool(g["real_time"]), int(g["seed"])
    n = len(g["out_done"])
    assert np.array_equal(g["env_id"], np.arange(n))  # transition i was recorded as environment i
    env = _env(n, continuous=cont, real_time=rt, seed=seed, env_id0=0, auto_reset=False, exact_steering=True,
               inf_obs=bool(g.get("inf_obs", True)))
    _load_state(env, sa_state_from_golden(g))
    obs, rew, done, _ = env.step(torch.from_numpy(np.ascontiguousarray(g["action"])).cuda())
    got = _read_state(env)
    _assert_close_and_report(name + " car", got.car, f32(g["out_car"]))
    _assert_close_and_report(name + " obst", got.obst, f32(g["out_obst"]))
    assert np.array_equal(got.lane, g["out_lft"][:, 0]) and np.array_equal(got.flags, g["out_lft"][:, 1])
    assert np.array_equal(got.tick, g["out_lft"][:, 2])
    assert np.array_equal(got.ncoll, g["out_ncoll"])
    assert np.array_equal(got.spawn_ctr, g["out_spawn_ctr"].astype(np.uint32))
    assert np.array_equal(done.cpu().numpy(), g["out_done"])
    _assert_close_and_report(name + " obs", obs.cpu().numpy(), g["out_obs"])
    _assert_close_and_report(name + " rew", rew.cpu().numpy(), f32(g["out_rew"]))


@pytest.mark.parametrize("continuous,real_time", [(False, False), (True, False), (False, True)])
def test_trajectory_parity_1000_steps(continuous, real_time):
    """4096 envs x 1000 consecutive steps with auto-reset, kernel vs oracle on the same actions"""
    n, steps, seed = 4096, 1000, 5
    env = _env(n, continuous=continuous, real_time=real_time, seed=seed, env_id0=100, exact_steering=True)
    st = orc.SaState(n)
    orc.sa_reset(st)
    rng = np.random.RandomState(1)
    ndone = 0
    for t in range(steps):
        if continuous:
            act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        else:
            act = rng.randint(0, 4, n).astype(np.int32)
        obs, rew, done, infos = env.step(torch.from_numpy(act).cuda())
        want = orc.sa_step(st, act, continuous=continuous, real_time=real_time, seed=seed, env_id0=100,
                           quantise=True, nthreads=8)
        d = done.cpu().numpy()
        assert np.array_equal(d, want["done"]), "done differs at step %d" % t
        assert bits_equal(obs.cpu().numpy(), want["obs"]), "obs differs at step %d" % t
        assert bits_equal(rew.cpu().numpy(), f32(want["rew"])), "reward differs at step %d" % t
        ndone += int(d.sum())
        if d.any():
            term = env._out[env._cur]["term"].cpu().numpy()
            assert np.array_equal(term[d].astype(np.int64), want["term"][d])
        if t % 50 == 49 or t == steps - 1:
            got = _read_state(env)
            for f in got.FIELDS:
                a, b = getattr(got, f), getattr(st, f)
                assert np.array_equal(a, b), "state field %s differs at step %d" % (f, t)
    assert ndone > n  # crashes and time-outs both happened
    s = env.episode_statistics()
    assert s["episodes"] == ndone


@pytest.mark.parametrize("continuous", [False, True])
def test_first_tick_handles_states_off_the_reference_manifold(continuous):
    """States the reference cannot reach but a caller can inject: the car away from x = 10 on a running episode, obstacles
    placed around it.  The first tick of a step is the generic one (exact rect test at the car's real x); the oracle is
    generic throughout, so one step from such states must agree (the later ticks see the car pinned at 10 again)."""
    n, seed = 8192, 11
    rng = np.random.RandomState(3)
    st = orc.SaState(n)
    orc.sa_reset(st)
    st.tick[:] = rng.randint(1, 2000, n)
    st.car[:, 0] = f32(rng.uniform(4.0, 16.0, n))           # px
    st.car[:, 1] = f32(rng.uniform(1.0, 6.9, n))            # py
    st.car[:, 2] = f32(rng.uniform(20.0, 900.0, n))         # raw x
    st.car[:, 3] = f32(rng.uniform(0.0, 20.0, n))           # vx
    st.car[:, 4] = f32(rng.uniform(-5.0, 5.0, n))           # acc
    st.lane[:] = rng.randint(1, 4, n)
    if continuous:
        st.car[:, 5] = f32(rng.uniform(-30.0, 30.0, n))
    else:
        st.flags[:] = rng.choice([0, 1, 2, 4, 8, 1 | 4, 2 | 8, 1 | 8, 2 | 4], n)
        st.car[:, 5] = np.where(st.flags & 3, f32(rng.uniform(-30.0, 30.0, n)), 0.0)
        st.car[:, 6] = f32(rng.uniform(0.0, 20.0, n))       # cruise
    st.obst[:, :, 0] = f32(st.car[:, 0:1] + rng.uniform(-6.0, 30.0, (n, 3)))   # around the car: many rect overlaps
    st.obst[:, :, 1] = f32(st.car[:, 2:3] + (st.obst[:, :, 0] - st.car[:, 0:1]))
    st.obst[:, :, 3] = f32(rng.uniform(5.0, 7.0, (n, 3)))
    st.obst[:, :, 2] = st.obst[:, :, 3]
    env = _env(n, continuous=continuous, seed=seed, auto_reset=False, exact_steering=True)
    _load_state(env, st)
    act = rng.uniform(-1, 1, (n, 2)).astype(np.float32) if continuous else rng.randint(0, 4, n).astype(np.int32)
    obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
    want = orc.sa_step(st, act, continuous=continuous, seed=seed, quantise=True, autoreset=False, nthreads=4)
    assert int(want["done"].sum()) > n // 20                  # plenty of collisions in the sample
    assert np.array_equal(done.cpu().numpy(), want["done"])
    assert bits_equal(rew.cpu().numpy(), f32(want["rew"]))
    assert bits_equal(obs.cpu().numpy(), want["obs"])
    got = _read_state(env)
    for f in got.FIELDS:
        assert np.array_equal(getattr(got, f), getattr(st, f)), f


def test_tail_block_and_small_batches():
    for n in (1, 2, 31, 127, 129, 1000):
        env = _env(n, seed=9, exact_steering=True)
        st = orc.SaState(n); orc.sa_reset(st)
        rng = np.random.RandomState(n)
        for t in range(30):
            act = rng.randint(0, 4, n).astype(np.int32)
            obs, rew, done, _ = env.step(act)
            want = orc.sa_step(st, act, seed=9, quantise=True)
            assert bits_equal(obs.cpu().numpy(), want["obs"])
            assert np.array_equal(done.cpu().numpy(), want["done"])


def test_vecenv_contract():
    from gym_highway_b200 import AlreadySteppingError, NotSteppingError
    env = _env(8, return_numpy=True)
    assert env.num_envs == 8 and env.observation_space.shape == (18,) and env.action_space.n == 4
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (8, 18) and obs.dtype == np.float32
    with pytest.raises(NotSteppingError):
        env.step_wait()
    env.step_async(np.zeros(8, np.int64))
    with pytest.raises(AlreadySteppingError):
        env.step_async(np.zeros(8, np.int64))
    obs, rew, done, infos = env.step_wait()
    assert rew.dtype == np.float32 and done.dtype == bool and len(infos) == 8
    assert set(infos[0]) == {"run_time", "num_obs_collisions", "num_agent_collisions"}
    assert infos[0]["run_time"] == 0.25000000000000006
    with pytest.raises(AssertionError):
        env.step(np.zeros(7, np.int64))
    # run a full time-out episode: 240 steps of MAINTAIN from reset, done on the last, obs is the reset obs
    env = _env(2, return_numpy=True)
    first = env.reset().copy()
    for t in range(240):
        obs, rew, done, infos = env.step([3, 3])
        assert bool(done[0]) == (t == 239)
    assert np.array_equal(obs, first)
    ep = infos[0]["episode"]
    assert ep["l"] == 240 and infos[0]["run_time"] >= 60.0 and infos.episodes()[0]["l"] == 240
    env.close(); env.close()


@pytest.mark.parametrize("n,steps,continuous", [(513, 20, False), (262144 + 77, 6, False), (300001, 4, True)])
def test_host_buffer_step_matches_device_step(n, steps, continuous):
    """hw_sa_step_host (host buffers; large batches run as two overlapped sub-range launches) against the device step"""
    a, b = _env(n, seed=2, continuous=continuous), _env(n, seed=2, continuous=continuous)
    rng = np.random.RandomState(0)
    a.rollout_random(40); b.rollout_random(40)  # respawns, lane changes and finished episodes in every chunk
    for t in range(steps):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32) if continuous else rng.randint(0, 4, n).astype(np.int32)
        o1, r1, d1, _ = a.step(torch.from_numpy(act).cuda())
        o2, r2, d2 = b.step_host(act)
        assert bits_equal(o1.cpu().numpy(), o2) and bits_equal(r1.cpu().numpy(), r2)
        assert np.array_equal(d1.cpu().numpy(), d2)
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("n,continuous", [(4096 + 37, False), (300001, False), (4096 + 37, True)])
def test_compact_observation_rows_are_the_reference_rows_without_the_constant_columns(n, continuous):
    """HW_FLAG_COMPACT_OBS: [N, 14] rows = the reference's 18 columns minus 11, 13, 15, 17, which are 0.5 in every row (auto-reset rows
    included); everything else the step produces is unchanged, and expand_obs() restores the reference layout bit for bit."""
    from gym_highway_b200 import _lib
    a, b = _env(n, seed=3, continuous=continuous), _env(n, seed=3, continuous=continuous)
    rng = np.random.RandomState(1)
    a.rollout_random(60); b.rollout_random(60)
    keep = list(_lib.SA_COMPACT_COLUMNS)
    ndone = 0
    for t in range(12):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32) if continuous else rng.randint(0, 4, n).astype(np.int32)
        o1, r1, d1 = a.step_host(act)
        o2, r2, d2 = b.step_host(act, compact=True)
        assert o2.shape == (n, _lib.SA_OBS_DIM_COMPACT)
        assert np.all(o1[:, 11:18:2] == 0.5)
        assert bits_equal(o1[:, keep], o2) and bits_equal(r1, r2) and np.array_equal(d1, d2)
        assert bits_equal(type(b).expand_obs(o2), o1)
        ndone += int(d1.sum())
    assert ndone > 0  # rows written by the auto-reset path were compared too
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k


def test_rollout_random_is_deterministic_and_sane():
    a, b = _env(2048, seed=4), _env(2048, seed=4)
    oa, ra, da = a.rollout_random(37, action_ctr0=5)
    ob, rb, db = b.rollout_random(37, action_ctr0=5)
    assert bits_equal(oa.cpu().numpy(), ob.cpu().numpy())
    o = oa.cpu().numpy()
    assert np.isfinite(o).all() and o.min() >= 0.0 and o.max() <= 1.0


def test_gym_per_env_api_matches_reference_anchor():
    """HighwayEnv (N=1 gym API): the survey's seed-0 anchor of the reference"""
    from gym_highway_b200 import HighwayEnv, HighwayEnvContinuous, make_vec_env
    env = HighwayEnv(manual=False, inf_obs=True, save=False, render=False, real_time=False)
    ob = env.reset()
    want = np.array([0.5, 0.5, 0.0166667, 0.46875, 0.1666667, 0.34375, 0.125, 0.5, 0.0, 0.65625, 0.5, 0.5,
                     0.675, 0.5, 0.625, 0.5, 0.65, 0.5], np.float32)
    assert ob.shape == (18,) and ob.dtype == np.float32 and np.allclose(ob, want, atol=1e-6)
    rews = []
    for a in [2, 2, 2, 0, 3, 1]:
        ob, r, d, info = env.step(a)
        rews.append(r)
        assert d is False and set(info) == {"run_time", "num_obs_collisions", "num_agent_collisions"}
    assert rews == [-0.998, -0.993, -0.985, -0.974, -0.912, -0.849]
    envc = HighwayEnvContinuous(render=False)
    envc.reset()
    ob, r, d, info = envc.step([0.5, -0.25])
    assert ob.shape == (18,) and isinstance(r, float)
    v = make_vec_env('Highway-cont-train-v0', num_env=3, seed=1)
    assert v.num_envs == 3 and v.continuous
    with pytest.raises(NotImplementedError):
        HighwayEnv(render=True)


# ---------------------------------------------------------------------------------------------------------
# Default (fast) steering term: vx*tan/4 with one rounding instead of the reference's two divisions.  Parity
# bar of the contract: discrete state bit-exact, floats within 1e-5 relative per step from identical states.
def _float_report(name, got, want, max_bitdiff_frac):
    got, want = f32(got), f32(want)
    rel = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want.astype(np.float64)), 1e-3)
    assert rel.max() <= REL_TOL, "%s: max rel err %.3g" % (name, rel.max())
    frac = float(np.mean(got.view(np.uint32) != want.view(np.uint32)))
    assert frac <= max_bitdiff_frac, "%s: %.3g of the values differ in the last bit" % (name, frac)
    return frac


@pytest.mark.parametrize("name", ["sa_discrete", "sa_continuous", "sa_discrete_realtime", "sa_discrete_noinf", "sa_discrete_edge", "sa_continuous_edge"])
def test_fast_mode_golden_transitions(name):
    g = load_golden(name)
    n = len(g["out_done"])
    env = _env(n, continuous=bool(g["continuous"]), real_time=bool(g["real_time"]), seed=int(g["seed"]), env_id0=0,
               auto_reset=False, inf_obs=bool(g.get("inf_obs", True)))
    _load_state(env, sa_state_from_golden(g))
    obs, rew, done, _ = env.step(torch.from_numpy(np.ascontiguousarray(g["action"])).cuda())
    got = _read_state(env)
    assert np.array_equal(got.lane, g["out_lft"][:, 0]) and np.array_equal(got.flags, g["out_lft"][:, 1])
    assert np.array_equal(got.tick, g["out_lft"][:, 2]) and np.array_equal(got.ncoll, g["out_ncoll"])
    assert np.array_equal(got.spawn_ctr, g["out_spawn_ctr"].astype(np.uint32))
    assert np.array_equal(done.cpu().numpy(), g["out_done"])
    _float_report(name + " car", got.car, g["out_car"], 1e-4)
    _float_report(name + " obst", got.obst, g["out_obst"], 0.0)
    _float_report(name + " obs", obs.cpu().numpy(), g["out_obs"], 1e-4)
    _float_report(name + " rew", rew.cpu().numpy(), g["out_rew"], 0.0)


@pytest.mark.parametrize("continuous", [False, True])
def test_fast_mode_per_step_parity_from_identical_states(continuous):
    """the contract literally: every step starts from the kernel's own (identical) state in the oracle"""
    n, steps, seed = 4096, 300, 11
    env = _env(n, continuous=continuous, seed=seed, env_id0=5)
    rng = np.random.RandomState(3)
    nbits = nvals = 0
    for t in range(steps):
        st = _read_state(env)
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32) if continuous else rng.randint(0, 4, n).astype(np.int32)
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        want = orc.sa_step(st, act, continuous=continuous, seed=seed, env_id0=5, quantise=True, nthreads=8)
        got = _read_state(env)
        assert np.array_equal(done.cpu().numpy(), want["done"]), "done differs at step %d" % t
        for f in ("lane", "flags", "tick", "ncoll", "spawn_ctr", "ep_len", "ep_ret_milli"):
            assert np.array_equal(getattr(got, f), getattr(st, f)), "%s differs at step %d" % (f, t)
        assert bits_equal(rew.cpu().numpy(), f32(want["rew"]))
        for a, b in ((got.car, st.car), (got.obst, st.obst), (obs.cpu().numpy(), want["obs"])):
            a32, b32 = f32(a), f32(b)
            rel = np.abs(a32.astype(np.float64) - b32) / np.maximum(np.abs(b32.astype(np.float64)), 1e-3)
            assert rel.max() <= REL_TOL
            nbits += int(np.sum(a32.view(np.uint32) != b32.view(np.uint32))); nvals += a32.size
    assert nbits <= 1e-6 * nvals, "%d of %d float values differ in the last bit" % (nbits, nvals)


def test_full_size_properties_4m_envs():
    """BASELINE-size batch (4,194,304 envs, the bench workload): properties that do not need the oracle.
    * a slice of the big batch equals a small batch constructed with the slice's global env ids (what sharding relies on:
      an env's trajectory depends only on seed, global id and its actions);
    * two instances agree bit for bit (no race, no dependence on scheduling);
    * observations are normalised into [0, 1]; the device episode counters add up."""
    n, lo, m = 1 << 22, 1234567 & ~1, 4096
    big, twin, part = _env(n, seed=17), _env(n, seed=17), _env(m, seed=17, env_id0=lo)
    for k in range(3):                      # 3 x 90 steps: first time-outs at step 240, plenty of crashes before
        ob, rb, db = big.rollout_random(90, action_ctr0=90 * k)
        ot, rt, dt = twin.rollout_random(90, action_ctr0=90 * k)
        op, rp, dp = part.rollout_random(90, action_ctr0=90 * k)
    assert torch.equal(ob, ot) and torch.equal(rb, rt) and torch.equal(db, dt)
    assert torch.equal(ob[lo:lo + m], op) and torch.equal(rb[lo:lo + m], rp) and torch.equal(db[lo:lo + m], dp)
    sb, sp = big.state_dict(), part.state_dict()
    assert torch.equal(sb["car_a"][lo:lo + m], sp["car_a"]) and torch.equal(sb["obst"][:, lo:lo + m], sp["obst"])
    assert float(ob.min()) >= 0.0 and float(ob.max()) <= 1.0 and bool(torch.isfinite(rb).all())
    # a python step on top: random actions from a tensor, same comparison through the ordinary step() entry point
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    act = torch.randint(0, 4, (n,), generator=g, device="cuda", dtype=torch.int32)
    ob, rb, db, _ = big.step(act)
    op, rp, dp, _ = part.step(act[lo:lo + m].contiguous())
    assert torch.equal(ob[lo:lo + m], op) and torch.equal(db[lo:lo + m], dp)
    st = big.episode_statistics()
    assert st["episodes"] > n and st["timeouts"] > 0 and st["obs_collisions"] >= st["episodes"] - st["timeouts"]
    assert 1.0 <= st["mean_length"] <= 240.0



def test_env_on_another_device_than_the_current_one():
    """the launch context switches to the environment's device only when it is not current: an env on cuda:1 stepped while cuda:0 is
    current gives what the same env gives on cuda:0 (needs two GPUs)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    assert torch.cuda.current_device() == 0
    n = 2048 + 5
    a, b = _env(n, seed=6), _env(n, seed=6, device="cuda:1")
    rng = np.random.RandomState(3)
    for t in range(30):
        act = rng.randint(0, 4, n).astype(np.int32)
        o1, r1, d1, _ = a.step(torch.from_numpy(act).cuda())
        o2, r2, d2, _ = b.step(torch.from_numpy(act).to("cuda:1"))
        assert o2.device.index == 1 and torch.cuda.current_device() == 0
        assert bits_equal(o1.cpu().numpy(), o2.cpu().numpy()) and bits_equal(r1.cpu().numpy(), r2.cpu().numpy())
        assert np.array_equal(d1.cpu().numpy(), d2.cpu().numpy())
