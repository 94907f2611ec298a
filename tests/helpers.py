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


def check_vec_golden(g, reset_fn, step_fn):
    """replay a tests/golden/vec_*.npz trajectory (recorded from the reference's DummyVec(MA)Env + Monitor stack,
    oracle/gen_golden_vec.py) through `reset_fn() -> obs0` and `step_fn(t, action) -> (obs, rew f32, done, episodes,
    infos)` and require every output to be identical.  episodes: {env: (r, l)} of the episodes that ended on the step;
    infos: (run_time[n], n_obs_coll[n], n_agent_coll[n]) or None."""
    assert bits_equal(np.asarray(reset_fn(), np.float32), g["obs0"])
    T = g["done"].shape[0]
    n_ends = n_timeouts = 0
    for t in range(T):
        obs, rew, done, episodes, infos = step_fn(t, g["action"][t])
        assert np.array_equal(np.asarray(done, bool), g["done"][t]), "done differs at step %d" % t
        assert bits_equal(np.asarray(rew, np.float32), g["rew"][t]), "reward differs at step %d" % t
        assert bits_equal(np.asarray(obs, np.float32), g["obs"][t]), "observation differs at step %d" % t
        ended = np.nonzero(g["ep_l"][t])[0]
        assert sorted(episodes) == ended.tolist(), "episode ends differ at step %d" % t
        for e in ended:
            r, l = episodes[int(e)]
            assert l == int(g["ep_l"][t, e]) and r == float(g["ep_r"][t, e]), (t, e, r, l, g["ep_r"][t, e], g["ep_l"][t, e])
            n_ends += 1
            n_timeouts += int(g["run_time"][t, e] >= 60.0)
        if infos is not None:
            assert np.array_equal(np.asarray(infos[0], np.float64), g["run_time"][t]), "run_time differs at step %d" % t
            assert np.array_equal(np.asarray(infos[1]), g["n_obs_coll"][t]), "num_obs_collisions differs at step %d" % t
            assert np.array_equal(np.asarray(infos[2]), g["n_agent_coll"][t]), "num_agent_collisions differs at step %d" % t
    return n_ends, n_timeouts


# ---- stand-ins used by the runner / collector parity tests ----
class _Episodes(list):
    def episodes(self):
        return self


class OracleVecEnv(object):
    """the C oracle behind the slice of the VecEnv API the host-side runners use, with CPU torch tensors: lets the
    Runner / A2CRunner / collector logic be checked against the reference-recorded goldens without a GPU"""

    def __init__(self, n, continuous=False, seed=0, num_agents=0):
        import torch
        self._torch = torch
        self.num_envs, self.continuous, self.seed_, self.num_agents = n, continuous, seed, num_agents
        self.device = torch.device("cpu")
        if num_agents:
            from oracle import ma_oracle as mo
            self._mo = mo
            self.st = mo.MaState(n, num_agents)
            self.obs_dim = 4 * num_agents + 14
        else:
            self.st = orc.SaState(n)

    def reset(self):
        ob = self._mo.ma_reset(self.st) if self.num_agents else orc.sa_reset(self.st)
        return self._torch.from_numpy(ob.copy())

    def step(self, actions):
        t = self._torch
        a = actions.cpu().numpy()
        if self.num_agents:
            o = self._mo.ma_step(self.st, a, continuous=self.continuous, seed=self.seed_, quantise=True)
            ended = np.nonzero(o["done"].any(axis=1))[0]
            eps = _Episodes(dict(env=int(i), r=round(float(o["term"][i, 3]), 6), l=int(o["term"][i, 4])) for i in ended)
        else:
            o = orc.sa_step(self.st, a, continuous=self.continuous, seed=self.seed_, quantise=True)
            eps = _Episodes(dict(env=int(i), r=round(o["term"][i, 2] / 1000.0, 6), l=int(o["term"][i, 3])) for i in np.nonzero(o["done"])[0])
        return t.from_numpy(o["obs"]), t.from_numpy(o["rew"].astype(np.float32)), t.from_numpy(o["done"]), eps


class HashModel(object):
    """torch twin of oracle/gen_golden_vec.py's NumPy HashModel (the deterministic "policy" the reference runners were
    driven with): every output is an exact function of the observation's bits"""
    bf16_inference = False
    _TABLE = (2, 3, 2, 3, 2, 0, 1, 3)

    def __init__(self, continuous=False):
        self.continuous = continuous

    def _bits(self, obs):
        import torch
        b = obs.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        h = torch.zeros(obs.shape[0], dtype=torch.int64, device=obs.device)
        for k in range(obs.shape[1]):
            h = (h * 1000003 + b[:, k]) & 0xFFFFFFFF
        return h

    def value(self, obs):
        return obs[:, 2] * 2.0 - obs[:, 12]

    def step(self, obs, generator=None):
        import torch
        h = self._bits(obs)
        neglogp = obs[:, 3] + 0.5
        if self.continuous:
            a0 = ((h >> 3) & 1023).to(torch.float32) / 512.0 - 1.0
            a1 = (((h >> 13) & 1023).to(torch.float32) / 512.0 - 1.0) * 0.125
            actions = torch.stack([a0, a1], dim=1)
        else:
            actions = torch.tensor(self._TABLE, dtype=torch.int32, device=obs.device)[(h >> 5) & 7]
        return actions, self.value(obs), neglogp


def check_ppo2_golden(g, make_env, Runner):
    """our Runner over `make_env(n, continuous, seed)` must return what baselines.ppo2.runner.Runner returned over the
    reference's VecEnv stack (tests/golden/runner_ppo2_*.npz): identical obs / masks / actions / values / neglogpacs /
    returns and the same Monitor episode records"""
    n, T, cont = int(g["n"]), int(g["nsteps"]), bool(g["continuous"])
    run = Runner(make_env(n, cont, int(g["seed"])), HashModel(cont), nsteps=T, gamma=float(g["gamma"]), lam=float(g["lam"]))
    n_eps = 0
    for k in range(g["obs"].shape[0]):
        obs, returns, masks, actions, values, neglogp, states, epinfos = run.run()
        assert bits_equal(obs.cpu().numpy(), g["obs"][k]), "obs differ in rollout %d" % k
        assert np.array_equal(actions.cpu().numpy(), g["actions"][k]), "actions differ in rollout %d" % k
        assert np.array_equal(masks.cpu().numpy() > 0.5, g["masks"][k]), "masks differ in rollout %d" % k
        assert bits_equal(values.cpu().numpy(), g["values"][k]) and bits_equal(neglogp.cpu().numpy(), g["neglogpacs"][k])
        assert bits_equal(returns.cpu().numpy(), g["returns"][k]), "returns differ in rollout %d" % k
        want = [(r, int(l)) for r, l in zip(g["ep_r"][k], g["ep_l"][k]) if l]
        assert [(e["r"], e["l"]) for e in epinfos] == want, "episode records differ in rollout %d" % k
        n_eps += len(want)
    return n_eps


def check_a2c_golden(g, make_env, A2CRunner):
    n, T = int(g["n"]), int(g["nsteps"])
    run = A2CRunner(make_env(n, False, int(g["seed"])), HashModel(False), nsteps=T, gamma=float(g["gamma"]))
    for k in range(g["obs"].shape[0]):
        obs, _states, rewards, masks, actions, values = run.run()
        assert bits_equal(obs.cpu().numpy(), g["obs"][k]), "obs differ in rollout %d" % k
        assert np.array_equal(actions.cpu().numpy(), g["actions"][k]) and np.array_equal(masks.cpu().numpy() > 0.5, g["masks"][k])
        assert bits_equal(values.cpu().numpy(), g["values"][k])
        assert bits_equal(rewards.cpu().numpy(), g["rewards"][k]), "n-step returns differ in rollout %d" % k
    return int(g["masks"].sum())


def check_memory_golden(g, mem, rms, ep_records):
    """ring contents / start / length / fixed-index sample / obs_rms sums / episode records against tests/golden/memory_ma3.npz
    (the reference's Memory filled by the MADDPG rollout loop's store_transition calls)"""
    A = int(g["num_agents"])
    assert mem.start == int(g["start"]) and mem.nb_entries == int(g["length"])
    import torch
    idx = torch.as_tensor(g["sample_idx"], device=mem.obs0.device)
    s = mem.sample(len(idx), index=idx)
    for a in range(A):
        for k in ("obs0", "obs1", "rewards", "actions", "terminals1"):
            got = s[k][:, a].cpu().numpy()
            assert bits_equal(got.reshape(g["a%d_%s" % (a, k)].shape), g["a%d_%s" % (a, k)]), (a, k)
        assert bits_equal(mem.obs0[:, a].cpu().numpy(), g["a%d_ring_obs0" % a])
        assert bits_equal(mem.rewards[:, a].cpu().numpy(), g["a%d_ring_rewards" % a])
    if rms is not None:
        # float64 sums in a different (deterministic) order than the reference's sequential updates
        cnt = float(g["rms_count"])
        assert float(rms._count) == 1e-2 + cnt
        assert np.allclose(rms._sum.cpu().numpy(), g["rms_sum"], rtol=1e-12, atol=1e-9)
        assert np.allclose(rms._sumsq.cpu().numpy() - 1e-2, g["rms_sumsq"], rtol=1e-12, atol=1e-9)
    rews, steps = ep_records
    assert len(steps) == int(g["episodes"]) and list(steps) == g["ep_steps"].tolist()
    assert bits_equal(np.asarray(rews, np.float32).reshape(g["ep_rewards"].shape), g["ep_rewards"].astype(np.float32))
