"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python -m oracle.gen_golden [--sa-steps 3000] [--ma-steps 1500]

Each file holds independent transitions (state, action, Philox key) -> (successor state in fp64,
observation, reward, done, counters) recorded from /root/reference's own env.step executed under
oracle/shim.  Input states are fp32-representable (the kernels store fp32) and come from
trajectories driven by a mix of action policies so that lane changes, maintain/accelerate modes,
respawns, crashes and time-outs are all present.  The C oracle must reproduce every file bit for
bit (tests/test_oracle_golden.py); the CUDA kernels must reproduce them after fp32 rounding of the
state (tests/test_sa_gpu.py, tests/test_ma_gpu.py).
"""
import argparse
import os

import numpy as np

from . import oracle as orc
from . import ref_harness as rh

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _policy(rng, kind, continuous, t, prev):
    """action for step t under one of a few exploration policies"""
    if continuous:
        if kind == 0 or prev is None:
            return rng.uniform(-1, 1, 2).astype(np.float32)
        if kind == 1:  # sticky
            return prev if rng.rand() < 0.9 else rng.uniform(-1, 1, 2).astype(np.float32)
        if kind == 2:  # throttle-heavy, gentle steering
            return np.array([rng.uniform(0.2, 1.0), rng.uniform(-0.15, 0.15)], np.float32)
        return np.array([rng.uniform(-1, 1), 0.0 if rng.rand() < 0.5 else rng.uniform(-1, 1)], np.float32)
    if kind == 0 or prev is None:
        return int(rng.randint(0, 4))
    if kind == 1:
        return prev if rng.rand() < 0.85 else int(rng.randint(0, 4))
    if kind == 2:  # mostly accelerate: reaches top speed, rams obstacles
        return 2 if rng.rand() < 0.8 else int(rng.randint(0, 4))
    return int(rng.choice([2, 3, 3, 0, 1]))  # maintain-heavy


def gen_sa(continuous, nsteps, seed, real_time=False, inf_obs=True):
    ref = rh.RefSingleAgent(continuous=continuous, real_time=real_time, inf_obs=inf_obs)
    rng = np.random.RandomState(seed)
    rec = {k: [] for k in ("in_car", "in_lane", "in_flags", "in_tick", "in_obst", "in_ncoll", "in_spawn_ctr",
                           "env_id", "action", "out_car", "out_lft", "out_obst", "out_obs", "out_rew", "out_done",
                           "out_ncoll", "out_spawn_ctr")}
    st = orc.SaState(1)
    orc.sa_reset(st)
    episode, kind, prev, t_ep = 0, 0, None, 0
    for env_id in range(nsteps):  # transition i is stepped as environment i of one batch (env_id0 = 0)
        a = _policy(rng, kind, continuous, t_ep, prev)
        prev = a
        ref.inject(st, 0)
        prov = rh.PhiloxUniform(seed, env_id, int(st.spawn_ctr[0]))
        ob, r, d, _info = ref.step(a, prov)
        nxt = orc.SaState(1)
        ref.extract(nxt, 0)
        rec["in_car"].append(st.car[0].copy()); rec["in_lane"].append(st.lane[0]); rec["in_flags"].append(st.flags[0])
        rec["in_tick"].append(st.tick[0]); rec["in_obst"].append(st.obst[0].copy()); rec["in_ncoll"].append(st.ncoll[0])
        rec["in_spawn_ctr"].append(st.spawn_ctr[0]); rec["env_id"].append(env_id)
        rec["action"].append(np.array(a))
        rec["out_car"].append(nxt.car[0].copy()); rec["out_lft"].append([nxt.lane[0], nxt.flags[0], nxt.tick[0]])
        rec["out_obst"].append(nxt.obst[0].copy()); rec["out_obs"].append(np.asarray(ob, np.float32))
        rec["out_rew"].append(float(r)); rec["out_done"].append(bool(d)); rec["out_ncoll"].append(nxt.ncoll[0])
        rec["out_spawn_ctr"].append(prov.spawn_ctr)
        t_ep += 1
        if d:
            ctr = prov.spawn_ctr
            st = orc.SaState(1)
            orc.sa_reset(st)
            episode += 1
            # keep a running spawn counter on every second episode, restart it otherwise
            st.spawn_ctr[0] = ctr if episode % 2 else 0
            kind, prev, t_ep = int(rng.randint(0, 4)), None, 0
        else:
            st = nxt
            st.spawn_ctr[0] = prov.spawn_ctr
            st.quantise_f32()
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["seed"] = np.int64(seed)
    out["continuous"] = np.bool_(continuous)
    out["real_time"] = np.bool_(real_time)
    out["inf_obs"] = np.bool_(inf_obs)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sa-steps", type=int, default=3000)
    ap.add_argument("--ma-steps", type=int, default=1500)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    jobs = [("sa_discrete", lambda: gen_sa(False, args.sa_steps, 11)),
            ("sa_continuous", lambda: gen_sa(True, args.sa_steps, 12)),
            ("sa_discrete_realtime", lambda: gen_sa(False, max(args.sa_steps // 3, 300), 13, real_time=True)),
            ("sa_discrete_noinf", lambda: gen_sa(False, max(args.sa_steps // 3, 300), 14, inf_obs=False))]
    try:
        from .gen_golden_ma import ma_jobs
        jobs += ma_jobs(args.ma_steps)
    except ImportError:
        pass
    for name, fn in jobs:
        if args.only and args.only not in name:
            continue
        data = fn()
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **data)
        nd = int(np.sum(data["out_done"])) if data["out_done"].ndim == 1 else int(np.sum(np.any(data["out_done"], axis=-1)))
        print("%-28s %6d transitions, %4d terminal, %7.1f KB" % (name, len(data["out_done"]), nd, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
