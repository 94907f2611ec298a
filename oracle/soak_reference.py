"""Soak: the C oracle against the LIVE, unmodified reference on fresh trajectories (build container only).

    python -m oracle.soak_reference [--sa 30000] [--ma 12000] [--procs 8]  ->  tests/golden/soak_summary.json

Every job drives the reference's own env.step (oracle/gen_golden.py / gen_golden_ma.py generators: state injection,
Philox-served spawn draws, a mix of exploration policies) for thousands of transitions with a seed no committed golden
file uses, then hands the recorded (state, action) rows to the C oracle and requires successor state, observation,
reward, done flags, counters and spawn counters to be identical bit for bit.  The summary (transitions, episode ends,
mismatches per job) is committed; tests/test_reference_live.py re-runs a slice of every job and checks the summary's
totals.  TEST INFRASTRUCTURE ONLY.
"""
import argparse
import json
import multiprocessing as mp
import os
import time

import numpy as np

SUMMARY = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "soak_summary.json")


def _bits(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(np.asarray(b).astype(a.dtype))
    u = {4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
    return int(np.count_nonzero(a.view(u) != b.view(u)))


def check_sa(g):
    """mismatching words between the oracle's step and the reference rows of a gen_sa() record"""
    from . import oracle as orc
    n = len(g["out_done"])
    st = orc.SaState(n)
    st.car[:] = g["in_car"]; st.lane[:] = g["in_lane"]; st.flags[:] = g["in_flags"]; st.tick[:] = g["in_tick"]
    st.obst[:] = g["in_obst"]; st.ncoll[:] = g["in_ncoll"]; st.spawn_ctr[:] = g["in_spawn_ctr"]
    o = orc.sa_step(st, g["action"], continuous=bool(g["continuous"]), seed=int(g["seed"]), env_id0=0, autoreset=False,
                    want_next=True, real_time=bool(g["real_time"]), inf_obs=bool(g["inf_obs"]))
    bad = _bits(o["next_car"], g["out_car"]) + _bits(o["next_obst"], g["out_obst"]) + _bits(o["obs"], g["out_obs"])
    bad += _bits(o["rew"], g["out_rew"].astype(np.float64))
    bad += int(np.count_nonzero(o["next_lft"] != g["out_lft"])) + int(np.count_nonzero(o["done"] != g["out_done"]))
    bad += int(np.count_nonzero(st.ncoll != g["out_ncoll"])) + int(np.count_nonzero(st.spawn_ctr != g["out_spawn_ctr"].astype(np.uint32)))
    return bad


def check_ma(g):
    from . import ma_oracle as mo
    n, A, K = len(g["env_id"]), int(g["num_agents"]), int(g["num_slots"])
    st = mo.MaState(n, A, K)
    for f in ("car", "lane", "cflags", "tick", "leader", "nobs", "obst", "olane", "ofresh", "newest", "nc_obs", "nc_agent", "spawn_ctr"):
        getattr(st, f)[...] = g["in_" + f].reshape(getattr(st, f).shape)
    o = mo.ma_step(st, g["action"], continuous=bool(g["continuous"]), seed=int(g["seed"]), env_id0=0, autoreset=False,
                   shared_reward=bool(g["shared_reward"]), real_time=bool(g["real_time"]), inf_obs=bool(g["inf_obs"]))
    bad = _bits(st.car, g["out_car"].astype(np.float64)) + _bits(st.obst, g["out_obst"].astype(np.float64))
    for f in ("lane", "cflags", "tick", "leader", "nobs", "olane", "ofresh", "newest", "nc_obs", "nc_agent"):
        bad += int(np.count_nonzero(getattr(st, f) != g["out_" + f].reshape(getattr(st, f).shape)))
    bad += int(np.count_nonzero(st.spawn_ctr != g["out_spawn_ctr"].astype(np.uint32)))
    bad += _bits(o["obs"], g["out_obs"]) + _bits(o["rew"], g["out_rew"].astype(np.float64)) + int(np.count_nonzero(o["done"] != g["out_done"]))
    bad += int(st.overflow.sum())
    return bad


def job_list(n_sa, n_ma):
    """(name, kind, kwargs): seeds 5000+ are used by no committed golden file"""
    jobs = []
    for i, cont in enumerate((False, True)):
        for part in range(3):
            jobs.append(("sa_%s_%d" % ("continuous" if cont else "discrete", part), "sa", dict(continuous=cont, nsteps=n_sa // 3, seed=5000 + 10 * i + part)))
    jobs.append(("sa_discrete_realtime", "sa", dict(continuous=False, nsteps=n_sa // 6, seed=5031, real_time=True)))
    jobs.append(("sa_continuous_noinf", "sa", dict(continuous=True, nsteps=n_sa // 6, seed=5032, inf_obs=False)))
    for A in (1, 3, 5):
        for cont in (False, True):
            jobs.append(("ma%d_%s" % (A, "continuous" if cont else "discrete"), "ma", dict(A=A, continuous=cont, nsteps=n_ma, seed=6000 + 10 * A + int(cont))))
    jobs.append(("ma5_discrete_shared", "ma", dict(A=5, continuous=False, nsteps=n_ma // 3, seed=6091, shared_reward=True)))
    jobs.append(("ma3_discrete_realtime", "ma", dict(A=3, continuous=False, nsteps=n_ma // 3, seed=6092, real_time=True)))
    return jobs


def run_job(job):
    name, kind, kw = job
    t0 = time.time()
    if kind == "sa":
        from . import gen_golden
        g = gen_golden.gen_sa(kw["continuous"], kw["nsteps"], kw["seed"], real_time=kw.get("real_time", False), inf_obs=kw.get("inf_obs", True))
        bad, ends = check_sa(g), int(np.sum(g["out_done"]))
        extra = dict(timeouts=int(np.sum(g["out_done"] & (g["out_ncoll"] == g["in_ncoll"]))), respawns=int(np.sum(g["out_spawn_ctr"] - g["in_spawn_ctr"])))
    else:
        from . import gen_golden_ma
        g = gen_golden_ma.gen_ma(kw["A"], kw["continuous"], kw["nsteps"], kw["seed"], shared_reward=kw.get("shared_reward", False),
                                 real_time=kw.get("real_time", False))
        bad, ends = check_ma(g), int(np.sum(np.any(g["out_done"], axis=-1)))
        extra = dict(max_live_obstacles=int(g["out_nobs"].max()), agent_collisions=int(np.sum(g["out_nc_agent"] - g["in_nc_agent"])),
                     timeouts=int(np.sum(np.all(g["out_done"], axis=-1) & (g["out_nc_obs"] == g["in_nc_obs"]) & (g["out_nc_agent"] == g["in_nc_agent"]))),
                     retired=int(np.sum(np.maximum(g["in_nobs"] - g["out_nobs"], 0))))
    return dict(name=name, transitions=int(len(g["out_done"])), episode_ends=ends, mismatching_words=int(bad), seed=kw["seed"],
                seconds=round(time.time() - t0, 1), **extra)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sa", type=int, default=30000, help="single-agent transitions per action type")
    ap.add_argument("--ma", type=int, default=12000, help="multi-agent transitions per (A, action type)")
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--out", default=SUMMARY)
    args = ap.parse_args()
    jobs = job_list(args.sa, args.ma)
    with mp.get_context("spawn").Pool(args.procs) as pool:
        res = pool.map(run_job, jobs, chunksize=1)
    tot = sum(r["transitions"] for r in res)
    summary = dict(what="C oracle vs the live unmodified reference, lock-step on fresh trajectories (oracle/soak_reference.py)",
                   args=dict(sa=args.sa, ma=args.ma), total_transitions=tot, total_episode_ends=sum(r["episode_ends"] for r in res),
                   total_mismatching_words=sum(r["mismatching_words"] for r in res), jobs=res)
    with open(args.out, "w") as fh:
        json.dump(summary, fh, indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k != "jobs"}))
    for r in res:
        print(r)


if __name__ == "__main__":
    main()
