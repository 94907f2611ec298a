"""GPU parity at the sizes BASELINE.json names, through the C ABI, against the C oracle.

configs[1]: ONE batch of 65,536 single-agent envs stepped for 1,000 consecutive env.steps with auto-reset; every
            environment of the batch is compared with the oracle on every step (not just a 4,096-env subset), in the
            bit-exact steering mode; the benchmarked default mode (single-rounding steering term) is checked at the same
            size per step from identical states against the contract (discrete state exact, floats within 1e-5 relative).
configs[2]: 16,384 multi-agent envs, A = 5 cars (+3..K obstacles = "8 vehicles"), i.i.d. UNIFORM random actions (the
            distribution bench.py times), 320 steps, full-batch compare; the last 1,024 environments of the batch drive a
            no-lane-change policy instead so that time-outs (tick 2159 -> 2160) happen inside the run and are asserted.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle import ma_oracle as mo
from helpers import MA_STATE_FIELDS, bits_equal, canonical, f32, sa_state_from_canonical

pytestmark = pytest.mark.gpu

NTHREADS = min(32, os.cpu_count() or 1)


def _read_sa(env):
    from gym_highway_b200.layout import unpack_sa
    sd = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
    return sa_state_from_canonical(unpack_sa(sd["car_a"], sd["car_b"], sd["obst"], sd["stat"]))


def _read_ma(env):
    from gym_highway_b200.layout import unpack_ma
    sd = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
    c = unpack_ma(sd["cars"], sd["obst"], sd["head"], env.num_agents, env.num_slots)
    st = mo.MaState(env.num_envs, env.num_agents, env.num_slots)
    for f in st.FIELDS:
        getattr(st, f)[...] = c[f]
    return st


@pytest.mark.parametrize("continuous", [False, True])
def test_configs1_65536_envs_1000_steps_bit_exact(continuous):
    from gym_highway_b200 import HighwayVecEnv
    n, steps, seed, id0 = 65536, 1000 if not continuous else 400, 31, 1 << 20
    env = HighwayVecEnv(n, continuous=continuous, seed=seed, env_id0=id0, exact_steering=True)
    st = orc.SaState(n)
    orc.sa_reset(st)
    rng = np.random.RandomState(17)
    ndone = ncrash = ntimeout = 0
    for t in range(steps):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32) if continuous else rng.randint(0, 4, n).astype(np.int32)
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        want = orc.sa_step(st, act, continuous=continuous, seed=seed, env_id0=id0, quantise=True, nthreads=NTHREADS)
        d = done.cpu().numpy()
        assert np.array_equal(d, want["done"]), "done differs at step %d" % t
        assert bits_equal(obs.cpu().numpy(), want["obs"]), "obs differs at step %d" % t
        assert bits_equal(rew.cpu().numpy(), f32(want["rew"])), "reward differs at step %d" % t
        if d.any():
            term = env._out[env._cur]["term"].cpu().numpy()
            assert np.array_equal(term[d].astype(np.int64), want["term"][d]), "Monitor record differs at step %d" % t
            ndone += int(d.sum())
            ntimeout += int((want["term"][d][:, 1] == 0).sum())
            ncrash += int((want["term"][d][:, 1] > 0).sum())
        if t % 100 == 99 or t == steps - 1:
            got = _read_sa(env)
            for f in got.FIELDS:
                assert np.array_equal(getattr(got, f), getattr(st, f)), "state field %s differs at step %d" % (f, t)
    assert ncrash > n // 2, "crash-ended episodes: %d" % ncrash
    if not continuous:
        assert ntimeout > 0, "no episode reached the time-out tick"
    s = env.episode_statistics()
    assert s["episodes"] == ndone and s["timeouts"] == ntimeout


def test_configs1_65536_envs_default_mode_within_contract():
    """the mode bench.py times (HW_FLAG_EXACT_STEERING off): per step from identical states, discrete quantities exact and
    floats within 1e-5 relative (north_star), at most 1e-5 of the float values off by their last fp32 bit"""
    from gym_highway_b200 import HighwayVecEnv
    n, steps, seed = 65536, 200, 8
    env = HighwayVecEnv(n, seed=seed, env_id0=77)
    env.rollout_random(37, action_ctr0=5)   # episodes at different phases before the comparison starts
    rng = np.random.RandomState(3)
    nbits = nvals = 0
    for t in range(steps):
        st = _read_sa(env)
        act = rng.randint(0, 4, n).astype(np.int32)
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        want = orc.sa_step(st, act, seed=seed, env_id0=77, quantise=True, nthreads=NTHREADS)
        got = _read_sa(env)
        assert np.array_equal(done.cpu().numpy(), want["done"]), "done differs at step %d" % t
        for f in ("lane", "flags", "tick", "ncoll", "spawn_ctr", "ep_ret_milli", "ep_len"):
            assert np.array_equal(getattr(got, f), getattr(st, f)), "%s differs at step %d" % (f, t)
        for a, b in ((got.car, st.car), (got.obst, st.obst), (obs.cpu().numpy(), want["obs"]), (rew.cpu().numpy(), want["rew"])):
            a32, b32 = f32(a), f32(b)
            rel = np.abs(a32.astype(np.float64) - b32) / np.maximum(np.abs(b32.astype(np.float64)), 1e-3)
            assert rel.max() <= 1e-5, "step %d: max rel err %.3g" % (t, rel.max())
            nbits += int(np.sum(a32.view(np.uint32) != b32.view(np.uint32))); nvals += a32.size
    assert nbits <= 1e-5 * nvals, "%d of %d float values differ in the last bit" % (nbits, nvals)


@pytest.mark.parametrize("kernel", ["group", "thread"])
@pytest.mark.parametrize("continuous", [False, True])
def test_configs2_16384_ma_envs_uniform_random_actions(continuous, kernel):
    from gym_highway_b200 import HighwayMAVecEnv
    n, A, steps, seed, id0, calm = 16384, 5, 320 if not continuous else 150, 12, 5000, 1024
    env = HighwayMAVecEnv(n, num_agents=A, continuous=continuous, seed=seed, env_id0=id0, exact_steering=True, kernel=kernel)
    st = mo.MaState(n, A)
    mo.ma_reset(st)
    rng = np.random.RandomState(23)
    nd = max_live = 0
    for t in range(steps):
        if continuous:
            act = rng.uniform(-1, 1, (n, A, 2)).astype(np.float32)      # i.i.d. uniform, full steering range
            act[n - calm:, :, 1] = 0.0                                   # the calm tail never steers
        else:
            act = rng.randint(0, 4, (n, A)).astype(np.int32)             # i.i.d. uniform over the four actions
            act[n - calm:] = np.where(rng.rand(calm, A) < 0.5, 2, 3)     # the calm tail: ACCELERATE / MAINTAIN only
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        want = mo.ma_step(st, act, continuous=continuous, seed=seed, env_id0=id0, quantise=True, nthreads=NTHREADS)
        d = done.cpu().numpy()
        assert np.array_equal(d, want["done"]), "done differs at step %d" % t
        assert bits_equal(obs.cpu().numpy(), want["obs"]), "obs differs at step %d" % t
        assert bits_equal(rew.cpu().numpy(), f32(want["rew"])), "reward differs at step %d" % t
        anyd = d.any(axis=1)
        nd += int(anyd.sum())
        if anyd.any():
            term = env._out[env._cur]["term"].cpu().numpy()
            assert np.array_equal(term[anyd], want["term"][anyd]), "Monitor record differs at step %d" % t
        max_live = max(max_live, int(st.nobs.max()))
        if t % 40 == 39 or t == steps - 1:
            got = _read_ma(env)
            for f in MA_STATE_FIELDS + ("ep_len", "ep_ret"):
                assert np.array_equal(getattr(got, f), getattr(st, f)), "state field %s differs at step %d" % (f, t)
    s = env.episode_statistics()
    assert s["episodes"] == nd and nd > n
    assert s["dropped_spawns"] == 0 and int(env.diagnostics().max()) == 0 and int(st.overflow.max()) == 0
    assert s["agent_collisions"] > 0 and s["obs_collisions"] > 0
    if not continuous:
        assert s["timeouts"] > 0, "no multi-agent episode reached the time-out tick (2159 -> 2160)"
    assert max_live >= 5


def test_sweep_size_4m_sa_envs_slice_against_the_oracle():
    """configs[4]'s 4 194 304-env point in the bit-exact mode: environments [lo, lo + m) of the big batch against the oracle stepped
    with the slice's global env ids (the full-size tests in test_sa_gpu.py compare the kernel with itself; this one closes the loop)"""
    from gym_highway_b200 import HighwayVecEnv
    n, lo, m, seed, steps = 4194304, 3000001, 2048, 29, 120
    env = HighwayVecEnv(n, seed=seed, exact_steering=True)
    st = orc.SaState(m)
    orc.sa_reset(st)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    nd = 0
    for t in range(steps):
        act = torch.randint(0, 4, (n,), generator=g, device="cuda", dtype=torch.int32)
        obs, rew, done, _ = env.step(act)
        want = orc.sa_step(st, act[lo:lo + m].cpu().numpy(), seed=seed, env_id0=lo, quantise=True, nthreads=NTHREADS)
        assert np.array_equal(done[lo:lo + m].cpu().numpy(), want["done"]), "done differs at step %d" % t
        assert bits_equal(obs[lo:lo + m].cpu().numpy(), want["obs"]), "obs differs at step %d" % t
        assert bits_equal(rew[lo:lo + m].cpu().numpy(), f32(want["rew"])), "reward differs at step %d" % t
        nd += int(want["done"].sum())
    assert nd > m // 20


@pytest.mark.parametrize("continuous", [False, True])
def test_sweep_size_262144_ma_envs_slice_against_the_oracle(continuous):
    """the 262 144-env multi-agent sweep point (thread-per-environment kernel), bit-exact mode: a slice against the oracle"""
    from gym_highway_b200 import HighwayMAVecEnv
    n, A, lo, m, seed, steps = 262144, 5, 200003, 1500, 31, 100
    env = HighwayMAVecEnv(n, num_agents=A, continuous=continuous, seed=seed, exact_steering=True)
    st = mo.MaState(m, A)
    mo.ma_reset(st)
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    nd = 0
    for t in range(steps):
        if continuous:
            act = torch.rand((n, A, 2), generator=g, device="cuda") * 2 - 1
        else:
            act = torch.randint(0, 4, (n, A), generator=g, device="cuda", dtype=torch.int32)
        obs, rew, done, _ = env.step(act)
        want = mo.ma_step(st, act[lo:lo + m].cpu().numpy(), continuous=continuous, seed=seed, env_id0=lo, quantise=True, nthreads=NTHREADS)
        assert np.array_equal(done[lo:lo + m].cpu().numpy(), want["done"]), "done differs at step %d" % t
        assert bits_equal(obs[lo:lo + m].cpu().numpy(), want["obs"]), "obs differs at step %d" % t
        assert bits_equal(rew[lo:lo + m].cpu().numpy(), f32(want["rew"])), "reward differs at step %d" % t
        nd += int(want["done"].any(axis=1).sum())
    assert nd > m and int(env.diagnostics().max()) == 0 and st.overflow.max() == 0
