"""VecEnv-level golden trajectories recorded from the UNMODIFIED reference stack (build container only).

    python -m oracle.gen_golden_vec

Where gen_golden.py pins single `env.step` transitions, this pins the layers ABOVE the env that SURVEY.md §8(b)
and §8(f) name, by running the reference's own classes:

  * `baselines.common.vec_env.dummy_vec_env.DummyVecEnv` / `dummy_vec_ma_env.DummyVecMAEnv` (auto-reset: the
    observation of a finished env is the post-reset one, reward / done / info are terminal, :46-57 / :66-79)
    over `baselines.bench.monitor.Monitor` (info['episode'] = {r, l, t}, :58-79) over the reference envs
        -> tests/golden/vec_*.npz  (actions in; obs / rewards / dones / episode records / info counters out)
  * `baselines.ppo2.runner.Runner.run()` (ppo2/runner.py:20-66) and `baselines.a2c.runner.Runner.run()`
    (a2c/runner.py:21-72) driven by a deterministic NumPy model over that VecEnv
        -> tests/golden/runner_*.npz  (obs, returns, masks, actions, values, neglogpacs / n-step returns)
  * `maddpg.trainer.ma_memory.Memory` filled the way `ma_ddpg_learner.store_transition` (:390-396) fills it from the
    MADDPG rollout loop (ma_ddpg.py:225-237) over DummyVecMAEnv
        -> tests/golden/memory_ma3.npz  (ring contents, start / length, a fixed-index sample)

The envs carry the state the kernels carry: between steps every float of the simulator state is rounded to
float32 in place (the kernels keep fp32 state in HBM and do the nine ticks of a step in fp64), and
`random.uniform` is served from the Philox stream keyed by (seed, env index, spawn counter).  Nothing in the
reference is edited; `tensorflow` / `gym.core` are import-only stand-ins under oracle/shim.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np

from . import ref_harness as rh

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _q(x):
    return float(np.float32(x))


def quantise_sa_inplace(sim):
    """round the single-agent simulator state to float32 where the kernels store float32 (rects follow the rounded
    positions, as they do when a kernel reloads its state)"""
    car = sim.reference_car
    car.position.x, car.position.y = _q(car.position.x), _q(car.position.y)
    car.raw_position.x, car.raw_position.y = _q(car.raw_position.x), car.position.y
    car.velocity.x = _q(car.velocity.x)
    car.acceleration, car.steering, car.cruise_vel = _q(car.acceleration), _q(car.steering), _q(car.cruise_vel)
    rh._set_rect(car)
    for o in sim.lane_max_obs:
        o.position.x, o.raw_position.x = _q(o.position.x), _q(o.raw_position.x)
        o.velocity.x, o.init_velocity.x = _q(o.velocity.x), _q(o.init_velocity.x)
        rh._set_rect(o)


def quantise_ma_inplace(world):
    for car in world.agents:
        car.position.x, car.position.y = _q(car.position.x), _q(car.position.y)
        car.raw_position.x, car.raw_position.y = _q(car.raw_position.x), car.position.y
        car.velocity.x, car.velocity.y = _q(car.velocity.x), _q(car.velocity.y)
        car.acceleration, car.steering, car.cruise_vel = _q(car.acceleration), _q(car.steering), _q(car.cruise_vel)
        rh._set_rect(car)
    for o in world.scripted_agents:
        fresh = (o.rect.y == 0)  # pygame's default rect: spawned in the last tick, not yet updated
        o.position.x, o.raw_position.x = _q(o.position.x), _q(o.raw_position.x)
        o.velocity.x, o.init_velocity.x = _q(o.velocity.x), _q(o.init_velocity.x)
        if not fresh:
            rh._set_rect(o)


class KernelStateEnv(object):
    """the reference env with the two harness conditions applied around its own step(): Philox-served spawn draws and
    float32 state between steps.  Sits between Monitor and the env; forwards everything else."""

    def __init__(self, env, seed, env_id, multi_agent):
        self.env = env
        self.multi_agent = multi_agent
        self.provider = rh.PhiloxUniform(seed, env_id, 0)  # the spawn counter runs on across episodes, like the kernels'
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.reward_range = getattr(env, "reward_range", None)
        self.metadata = getattr(env, "metadata", {})
        self.spec = getattr(env, "spec", None)

    def step(self, action):
        if self.multi_agent:
            a = np.asarray(action)
            action = [np.asarray(v, np.float64) for v in a] if a.ndim == 2 and a.shape[1] == 4 else \
                     [[float(np.float32(v[0])), float(np.float32(v[1]))] for v in a]
        elif not isinstance(action, int):
            action = [float(np.float32(action[0])), float(np.float32(action[1]))]
        with rh._patched_uniform(self.provider):
            ob, r, d, info = self.env.step(action)
        done = any(d) if isinstance(d, (list, tuple)) else bool(d)
        if not done:
            if self.multi_agent:
                quantise_ma_inplace(self.env.world)
            else:
                quantise_sa_inplace(self.env.env)
        return ob, r, d, info

    def reset(self, **kw):
        return self.env.reset(**kw)

    def close(self):
        return None

    def seed(self, s=None):
        return []

    @property
    def unwrapped(self):
        return self.env


def _baselines():
    rh.load()
    from baselines.common.vec_env.dummy_vec_env import DummyVecEnv
    from baselines.common.vec_env.dummy_vec_ma_env import DummyVecMAEnv
    from baselines.bench.monitor import Monitor
    return DummyVecEnv, DummyVecMAEnv, Monitor


def make_ref_vec_env(n, seed, continuous=False, num_agents=0, shared_reward=False):
    """the reference's VecEnv stack over n reference envs (num_agents = 0: single-agent)"""
    DummyVecEnv, DummyVecMAEnv, Monitor = _baselines()
    m = rh.load()

    def mk(i):
        def f():
            if num_agents:
                cls = m['ma'].MultiAgentEnvContinuous if continuous else m['ma'].MultiAgentEnv
                env = cls(world_config=dict(rh.TRAIN_KW), num_agents=num_agents, shared_reward=shared_reward)
            else:
                env = (m['sa'].HighwayEnvContinuous if continuous else m['sa'].HighwayEnv)(**rh.TRAIN_KW)
            return Monitor(KernelStateEnv(env, seed, i, bool(num_agents)), None, allow_early_resets=True)
        return f
    return (DummyVecMAEnv if num_agents else DummyVecEnv)([mk(i) for i in range(n)])


# ---- action scripts: per env one exploration policy, so crashes, time-outs (step 240) and lane changes all occur ----
def _script_sa(rng, T, n, continuous):
    if continuous:
        a = rng.uniform(-1, 1, (T, n, 2)).astype(np.float32)
        a[:, 0] = (0.0, 0.0)                       # env 0 never moves: runs into the time-out
        a[:, 1, 1] *= 0.05; a[:, 1, 0] = np.abs(a[:, 1, 0])
        return a
    a = rng.randint(0, 4, (T, n)).astype(np.int32)
    a[:, 0] = 3                                     # MAINTAIN at v = 0: time-out at step 240
    a[:, 1] = np.where(rng.rand(T) < 0.8, 2, a[:, 1])   # accelerate-heavy: rams obstacles
    a[:, 2] = rng.choice([2, 3, 3, 0, 1], T)
    return a


def _script_ma(rng, T, n, A, continuous):
    if continuous:
        a = rng.uniform(-1, 1, (T, n, A, 2)).astype(np.float32)
        a[:, 0] = 0.0
        a[:, 1, :, 1] *= 0.05; a[:, 1, :, 0] = np.abs(a[:, 1, :, 0])
        return a
    a = rng.randint(0, 4, (T, n, A)).astype(np.int32)
    a[:, 0] = 3                                     # all cars hold v = 0: time-out at step 240
    a[:, 1] = rng.choice([2, 3, 3, 3], (T, A))      # accelerate / maintain only: long episodes, obstacle build-up
    a[:, 2] = np.where(rng.rand(T, A) < 0.85, 2, a[:, 2])
    return a


def record_vec(n, T, seed, continuous=False, num_agents=0, shared_reward=False):
    A = num_agents
    venv = make_ref_vec_env(n, seed, continuous, A, shared_reward)
    rng = np.random.RandomState(seed)
    script = _script_ma(rng, T, n, A, continuous) if A else _script_sa(rng, T, n, continuous)
    obs0 = np.array(venv.reset(), np.float32)
    rec = {k: [] for k in ("obs", "rew", "done", "ep_r", "ep_l", "run_time", "n_obs_coll", "n_agent_coll")}
    for t in range(T):
        act = script[t]
        if A and not continuous:
            act = np.eye(4)[act]                    # the discrete MA env takes one-hot vectors and argmaxes them
        ob, rw, dn, infos = venv.step(list(act) if A else act)
        rec["obs"].append(np.array(ob, np.float32)); rec["rew"].append(np.array(rw, np.float32)); rec["done"].append(np.array(dn, bool))
        ep_r, ep_l, rt, no, na = [], [], [], [], []
        for info in infos:
            ep = info.get('episode')
            ep_r.append(ep['r'] if ep else np.nan); ep_l.append(ep['l'] if ep else 0)
            base = info['n'][0] if A else info
            rt.append(base['run_time']); no.append(base['num_obs_collisions']); na.append(base.get('num_agent_collisions', 0))
        rec["ep_r"].append(ep_r); rec["ep_l"].append(ep_l); rec["run_time"].append(rt); rec["n_obs_coll"].append(no); rec["n_agent_coll"].append(na)
    out = {k: np.asarray(v) for k, v in rec.items()}
    out.update(obs0=obs0, action=script, seed=np.int64(seed), continuous=np.bool_(continuous), num_agents=np.int64(A),
               shared_reward=np.bool_(shared_reward))
    return out


# ---- deterministic NumPy "model" for the runners: every output is a bit-level function of the observation, so one
#      differing observation bit anywhere changes the trajectory that follows ----
class HashModel(object):
    initial_state = None

    def __init__(self, continuous=False):
        self.continuous = continuous
        # a2c's runner reads these off the TF model (a2c/runner.py:24-25, :48, :67)
        self.train_model = type("TM", (), {})()
        self.train_model.action = type("T", (), {"shape": type("S", (), {"as_list": staticmethod(lambda: [None, 2] if continuous else [None])})(),
                                                 "dtype": type("D", (), {"name": "float32" if continuous else "int32"})()})()
        self.train_model.X = type("X", (), {"dtype": type("D", (), {"as_numpy_dtype": np.float32})()})()

    @staticmethod
    def _bits(obs):
        b = np.ascontiguousarray(obs, np.float32).view(np.uint32).astype(np.uint64)
        h = np.zeros(obs.shape[0], np.uint64)
        for k in range(obs.shape[1]):
            h = (h * np.uint64(1000003) + b[:, k]) & np.uint64(0xFFFFFFFF)
        return h

    def step(self, obs, S=None, M=None):
        h = self._bits(obs)
        values = (obs[:, 2] * np.float32(2.0) - obs[:, 12]).astype(np.float32)
        neglogp = (obs[:, 3] + np.float32(0.5)).astype(np.float32)
        if self.continuous:
            a0 = ((h >> np.uint64(3)) & np.uint64(1023)).astype(np.float32) / np.float32(512.0) - np.float32(1.0)
            a1 = (((h >> np.uint64(13)) & np.uint64(1023)).astype(np.float32) / np.float32(512.0) - np.float32(1.0)) * np.float32(0.125)
            actions = np.stack([a0, a1], axis=1).astype(np.float32)
        else:
            # mostly ACCELERATE / MAINTAIN with lane changes mixed in: episodes of useful length
            sel = ((h >> np.uint64(5)) & np.uint64(7)).astype(np.int64)
            actions = np.array([2, 3, 2, 3, 2, 0, 1, 3], np.int32)[sel]
        return actions, values, None, neglogp

    def value(self, obs, S=None, M=None):
        return (obs[:, 2] * np.float32(2.0) - obs[:, 12]).astype(np.float32)


def record_ppo2(n, nsteps, nruns, seed, continuous=False, gamma=0.99, lam=0.95):
    rh.load()
    from baselines.ppo2.runner import Runner
    venv = make_ref_vec_env(n, seed, continuous)
    runner = Runner(env=venv, model=HashModel(continuous), nsteps=nsteps, gamma=gamma, lam=lam)
    keys = ("obs", "returns", "masks", "actions", "values", "neglogpacs")
    rec = {k: [] for k in keys}
    ep_r, ep_l = [], []
    for _ in range(nruns):
        out = runner.run()
        for k, v in zip(keys, out[:6]):
            rec[k].append(np.asarray(v))
        ep_r.append([e['r'] for e in out[7]] + [np.nan] * (4 * n - len(out[7])))
        ep_l.append([e['l'] for e in out[7]] + [0] * (4 * n - len(out[7])))
    res = {k: np.asarray(v) for k, v in rec.items()}
    res.update(ep_r=np.asarray(ep_r), ep_l=np.asarray(ep_l), seed=np.int64(seed), continuous=np.bool_(continuous), n=np.int64(n),
               nsteps=np.int64(nsteps), gamma=np.float64(gamma), lam=np.float64(lam))
    return res


def record_a2c(n, nsteps, nruns, seed, gamma=0.99):
    rh.load()
    from baselines.a2c.runner import Runner
    venv = make_ref_vec_env(n, seed, False)
    runner = Runner(venv, HashModel(False), nsteps=nsteps, gamma=gamma)
    keys = ("obs", "rewards", "masks", "actions", "values")
    rec = {k: [] for k in keys}
    for _ in range(nruns):
        obs, _states, rewards, masks, actions, values = runner.run()
        for k, v in zip(keys, (obs, rewards, masks, actions, values)):
            rec[k].append(np.asarray(v))
    res = {k: np.asarray(v) for k, v in rec.items()}
    res.update(seed=np.int64(seed), n=np.int64(n), nsteps=np.int64(nsteps), gamma=np.float64(gamma))
    return res


def record_memory(n, A, T, limit, seed):
    """the MADDPG rollout loop's bookkeeping (ma_ddpg.py:225-262) over the reference VecEnv and the reference Memory"""
    rh.load()
    from maddpg.trainer.ma_memory import Memory
    venv = make_ref_vec_env(n, seed, continuous=True, num_agents=A)
    D = 4 * A + 14
    mems = [Memory(limit=limit, action_shape=(2,), observation_shape=(D,)) for _ in range(A)]
    rng = np.random.RandomState(seed)
    script = _script_ma(rng, T, n, A, True)
    obs_n = venv.reset()
    ep_rew = np.zeros((n, A), np.float32); ep_step = np.zeros(n, np.int64)
    episodes, ep_rewards, ep_steps = 0, [], []
    s_sum, s_sq, s_cnt = np.zeros((A, D), np.float64), np.zeros((A, D), np.float64), 0.0
    for t in range(T):
        actions_n = script[t]
        new_obs_n, rew_n, done_n, _info = venv.step(list(actions_n))
        ep_rew += rew_n; ep_step += 1
        for i in range(A):                         # agent.store_transition(obs_n, actions_n, rew_n, new_obs_n, done_n)
            for b in range(n):
                mems[i].append(obs_n[b][i], actions_n[b][i], rew_n[b][i], new_obs_n[b][i], done_n[b][i])
        for i in range(A):                         # ... which also feeds the ONE obs_rms all agents share (ma_ddpg.py:110-113,
            for b in range(n):                     # ma_ddpg_learner.py:400-404): update(np.array([obs0[b]])), shape (A, D), count 1
                x = np.array([obs_n[b]]).astype('float64')
                s_sum += x.sum(axis=0); s_sq += np.square(x).sum(axis=0); s_cnt += len(x)
        obs_n = new_obs_n
        for d in range(n):
            if any(done_n[d]):
                ep_rewards.append(ep_rew[d].copy()); ep_steps.append(int(ep_step[d]))
                ep_rew[d] = 0; ep_step[d] = 0; episodes += 1
    idx = np.arange(0, mems[0].nb_entries - 2, 7)
    res = dict(action=script, seed=np.int64(seed), n=np.int64(n), num_agents=np.int64(A), limit=np.int64(limit),
               start=np.int64(mems[0].observations0.start), length=np.int64(mems[0].nb_entries), sample_idx=idx,
               episodes=np.int64(episodes), ep_rewards=np.asarray(ep_rewards), ep_steps=np.asarray(ep_steps),
               rms_sum=s_sum, rms_sumsq=s_sq, rms_count=np.float64(s_cnt))
    for i, m in enumerate(mems):
        s = m.sample(len(idx), index=idx)
        for k in ("obs0", "obs1", "rewards", "actions", "terminals1"):
            res["a%d_%s" % (i, k)] = s[k]
        res["a%d_ring_obs0" % i] = m.observations0.data.copy()
        res["a%d_ring_rewards" % i] = m.rewards.data.copy()
    return res


def jobs():
    return [
        ("vec_sa_discrete", lambda: record_vec(8, 520, 21)),
        ("vec_sa_continuous", lambda: record_vec(6, 300, 22, continuous=True)),
        ("vec_ma3_discrete", lambda: record_vec(6, 300, 23, num_agents=3)),
        ("vec_ma5_discrete", lambda: record_vec(6, 300, 24, num_agents=5)),
        ("vec_ma3_continuous", lambda: record_vec(5, 280, 25, continuous=True, num_agents=3)),
        ("vec_ma5_shared", lambda: record_vec(4, 120, 26, num_agents=5, shared_reward=True)),
        ("runner_ppo2_discrete", lambda: record_ppo2(8, 64, 6, 31)),
        ("runner_ppo2_continuous", lambda: record_ppo2(6, 32, 4, 32, continuous=True)),
        ("runner_a2c_discrete", lambda: record_a2c(8, 5, 60, 33)),
        ("memory_ma3", lambda: record_memory(4, 3, 90, 256, 34)),
    ]


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, fn in jobs():
        if only and only not in name:
            continue
        data = fn()
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **data)
        extra = ""
        if "done" in data:
            d = data["done"]
            ends = int(d.sum()) if d.ndim == 2 else int(d.any(axis=-1).sum())
            extra = "%d steps x %d envs, %d episode ends, longest episode %d" % (d.shape[0], d.shape[1], ends, int(data["ep_l"].max()))
        print("%-26s %7.1f KB  %s" % (name, os.path.getsize(path) / 1024, extra))


if __name__ == "__main__":
    main()
