"""Multi-agent golden transitions recorded from the unmodified reference (see gen_golden.py)."""
import numpy as np

from . import ma_oracle as mo
from . import ref_harness as rh

K = 16


def _actions(rng, kind, continuous, A, prev):
    if continuous:
        if kind == 0 or prev is None:
            return rng.uniform(-1, 1, (A, 2)).astype(np.float32)
        if kind == 1:
            a = prev.copy()
            m = rng.rand(A) < 0.2
            a[m] = rng.uniform(-1, 1, (int(m.sum()), 2))
            return a.astype(np.float32)
        a = rng.uniform(-1, 1, (A, 2))
        a[:, 1] *= 0.05  # nearly straight: long episodes, overtakes, obstacle build-up
        a[:, 0] = np.abs(a[:, 0]) * (1.0 if kind == 2 else 0.4)
        return a.astype(np.float32)
    if kind == 0 or prev is None:
        return rng.randint(0, 4, A)
    if kind == 1:
        a = prev.copy()
        m = rng.rand(A) < 0.2
        a[m] = rng.randint(0, 4, int(m.sum()))
        return a
    if kind == 2:
        return np.where(rng.rand(A) < 0.85, 2, rng.randint(0, 4, A))  # accelerate-heavy
    return rng.choice([2, 3, 3, 3], A)  # accelerate / maintain only: no lane changes, time-outs happen


def gen_ma(A, continuous, nsteps, seed, shared_reward=False, real_time=False, inf_obs=True):
    ref = rh.RefMultiAgent(A, continuous=continuous, shared_reward=shared_reward, real_time=real_time, inf_obs=inf_obs)
    rng = np.random.RandomState(seed)
    names = ("in_car", "in_lane", "in_cflags", "in_tick", "in_leader", "in_nobs", "in_obst", "in_olane", "in_ofresh",
             "in_newest", "in_nc_obs", "in_nc_agent", "in_spawn_ctr", "env_id", "action",
             "out_car", "out_lane", "out_cflags", "out_tick", "out_leader", "out_nobs", "out_obst", "out_olane", "out_ofresh",
             "out_newest", "out_nc_obs", "out_nc_agent", "out_spawn_ctr", "out_obs", "out_rew", "out_done")
    rec = {k: [] for k in names}
    st = mo.MaState(1, A, K)
    mo.ma_reset(st)
    kind, prev, episode = 0, None, 0
    for env_id in range(nsteps):
        a = _actions(rng, kind, continuous, A, prev)
        prev = a
        ref.inject(st, 0)
        prov = rh.PhiloxUniform(seed, env_id, int(st.spawn_ctr[0]))
        ob, r, d, _info = ref.step(a, prov)
        nxt = mo.MaState(1, A, K)
        ref.extract(nxt, 0)
        for f in ("car", "lane", "cflags", "tick", "leader", "nobs", "obst", "olane", "ofresh", "newest", "nc_obs", "nc_agent", "spawn_ctr"):
            rec["in_" + f].append(getattr(st, f)[0].copy())
        rec["env_id"].append(env_id)
        rec["action"].append(np.array(a))
        for f in ("car", "lane", "cflags", "tick", "leader", "nobs", "obst", "olane", "ofresh", "newest", "nc_obs", "nc_agent"):
            rec["out_" + f].append(getattr(nxt, f)[0].copy())
        rec["out_spawn_ctr"].append(prov.spawn_ctr)
        rec["out_obs"].append(np.asarray(ob, np.float64).astype(np.float32))  # the float32 VecEnv buffer (dummy_vec_ma_env.py:40)
        rec["out_rew"].append(np.asarray(r, np.float64))
        rec["out_done"].append(np.asarray(d, bool))
        if any(d):
            ctr = prov.spawn_ctr
            st = mo.MaState(1, A, K)
            mo.ma_reset(st)
            episode += 1
            st.spawn_ctr[0] = ctr if episode % 2 else 0
            kind, prev = int(rng.randint(0, 4)), None
        else:
            st = nxt
            st.spawn_ctr[0] = prov.spawn_ctr
            st.quantise_f32()
    out = {k: np.asarray(v) for k, v in rec.items()}
    out.update(seed=np.int64(seed), continuous=np.bool_(continuous), real_time=np.bool_(real_time), inf_obs=np.bool_(inf_obs), num_agents=np.int64(A),
               num_slots=np.int64(K), shared_reward=np.bool_(shared_reward))
    return out


def ma_jobs(nsteps):
    jobs = []
    for A in (1, 3, 5):
        jobs.append(("ma%d_discrete" % A, (lambda A=A: gen_ma(A, False, nsteps, 100 + A))))
        jobs.append(("ma%d_continuous" % A, (lambda A=A: gen_ma(A, True, nsteps, 200 + A))))
    jobs.append(("ma5_discrete_shared", lambda: gen_ma(5, False, max(nsteps // 3, 200), 301, shared_reward=True)))
    jobs.append(("ma3_discrete_realtime", lambda: gen_ma(3, False, max(nsteps // 3, 200), 302, real_time=True)))
    jobs.append(("ma3_continuous_noinf", lambda: gen_ma(3, True, max(nsteps // 3, 200), 303, inf_obs=False)))
    return jobs
