"""soak: long trajectories of larger batches against the C oracle, value by value; for the multi-agent env the two step kernels (a
thread per environment / a lane per car) run side by side on the same actions, so a data race in either or a difference between them
shows up as a mismatch.

The oracle evaluates libm's tan (the reference's math.tan); the kernels evaluate a polynomial within 1 ulp of it (DESIGN.md,
deviation 3).  About once per 10^9 car-ticks that ulp reaches a float32 the step returns.  A job with mismatches is therefore re-run
against the oracle with the kernels' own tangent (oracle/tan_small.h): what disappears there is that declared deviation (it must also
be confined to observation floats, within 2 float32 ulps, no `done` flag), what is left is a kernel bug.  The soak FAILS on anything
unexplained.

    python scripts/long_parity.py [summary.json]      (GPU box; a few minutes; HW_SOAK_SCALE multiplies the steps, HW_SOAK_SEED shifts the seeds)"""
import json
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gym_highway_b200 import HighwayVecEnv, HighwayMAVecEnv
from oracle import oracle as orc, ma_oracle as mo

RESULTS = []
SCALE = int(os.environ.get("HW_SOAK_SCALE", "1"))    # multiplies the number of steps of every job
SEED0 = int(os.environ.get("HW_SOAK_SEED", "0"))     # added to the environment and action seeds


def _ulps(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def _compare(counts, obs, want_obs, done, want_done, rew=None, want_rew=None):
    o = obs.cpu().numpy()
    neq = o.view(np.uint32) != want_obs.view(np.uint32)
    if neq.any():
        counts["obs"] += int(neq.sum())
        counts["max_ulps"] = max(counts["max_ulps"], int(_ulps(o[neq], want_obs[neq]).max()))
    counts["done"] += int((done.cpu().numpy() != want_done).sum())
    if rew is not None:
        counts["rew"] += int((rew.cpu().numpy().view(np.uint32) != want_rew.astype(np.float32).view(np.uint32)).sum())


def _run_sa(n, steps, cont, poly, real_time=False, inf_obs=True):
    orc.set_tan(poly)
    env = HighwayVecEnv(n, continuous=cont, seed=21 + SEED0, env_id0=5, exact_steering=True, real_time=real_time, inf_obs=inf_obs)
    st = orc.SaState(n); orc.sa_reset(st)
    rng = np.random.RandomState(7 + SEED0)
    c = dict(obs=0, done=0, rew=0, max_ulps=0, kernels=0)
    for t in range(steps):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32) if cont else rng.randint(0, 4, n).astype(np.int32)
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        want = orc.sa_step(st, act, continuous=cont, seed=21 + SEED0, env_id0=5, quantise=True, nthreads=16, real_time=real_time, inf_obs=inf_obs)
        _compare(c, obs, want["obs"], done, want["done"], rew, want["rew"])
    orc.set_tan(False)
    return c, dict(episodes=int(env.episode_statistics()["episodes"]))


def _run_ma(n, steps, cont, A, uniform, poly, real_time=False, inf_obs=True, shared_reward=False):
    ext = A > 5
    orc.set_tan(poly)
    kw = dict(num_agents=A, continuous=cont, seed=22 + SEED0, env_id0=3, exact_steering=True, extended_agents=ext, shared_reward=shared_reward,
              world_config=dict(real_time=real_time, inf_obs=inf_obs))
    env = HighwayMAVecEnv(n, kernel="thread", **kw)
    env2 = HighwayMAVecEnv(n, kernel="thread" if ext else "group", **kw)
    st = mo.MaState(n, A); mo.ma_reset(st)
    rng = np.random.RandomState(8 + SEED0)
    c = dict(obs=0, done=0, rew=0, max_ulps=0, kernels=0)
    for t in range(steps):
        if cont:
            act = rng.uniform(-1, 1, (n, A, 2)).astype(np.float32)
            if not uniform: act[:, :, 1] *= 0.1
        else:
            act = rng.randint(0, 4, (n, A)).astype(np.int32) if uniform else np.where(rng.rand(n, A) < 0.6, 2, rng.randint(0, 4, (n, A))).astype(np.int32)
        a = torch.from_numpy(act).cuda()
        obs, rew, done, _ = env.step(a)
        obs2, rew2, done2, _ = env2.step(a)
        c["kernels"] += int((obs != obs2).sum()) + int((done != done2).sum())
        want = mo.ma_step(st, act, continuous=cont, seed=22 + SEED0, env_id0=3, quantise=True, nthreads=16, real_time=real_time, inf_obs=inf_obs,
                          shared_reward=shared_reward)
        _compare(c, obs, want["obs"], done, want["done"], rew, want["rew"])
    orc.set_tan(False)
    return c, dict(max_live_obstacles=int(st.nobs.max()), overflow=int(env.diagnostics().max()), episodes=int(env.episode_statistics()["episodes"]))


def _job(label, runner, record):
    """run against the libm oracle; on mismatches, once more against the oracle with the kernels' tangent; returns the unexplained count"""
    c, extra = runner(False)
    bad = c["obs"] + c["done"] + c["rew"]
    record.update(extra)
    record.update(mismatching_values=bad, kernel_vs_kernel_differences=c["kernels"])
    unexplained = c["kernels"] + extra.get("overflow", 0)
    note = ""
    if bad:
        cp, _ = runner(True)
        left = cp["obs"] + cp["done"] + cp["rew"] + cp["kernels"]
        confined = c["done"] == 0 and c["rew"] == 0 and c["max_ulps"] <= 2
        record.update(max_float32_ulps=c["max_ulps"], mismatching_values_with_the_kernels_tangent=left, explained_by_the_tangent_deviation=bool(left == 0 and confined))
        unexplained += 0 if (left == 0 and confined) else bad
        note = " (observation floats only, <= %d float32 ulps; %d left against the oracle with the kernels' tangent%s)" % (
            c["max_ulps"], left, ": the declared 1-ulp tan deviation" if left == 0 and confined else ": UNEXPLAINED")
    print("%s: mismatching values %d%s, differences between the two kernels %d%s" % (
        label, bad, note, c["kernels"], "".join(", %s %s" % (k.replace("_", " "), v) for k, v in extra.items() if k != "episodes")))
    RESULTS.append(record)
    return unexplained


def soak_sa(n, steps, cont, **modes):
    steps *= SCALE
    return _job("SA cont=%d %s %d envs x %d steps" % (cont, modes or "", n, steps), lambda poly: _run_sa(n, steps, cont, poly, **modes),
                dict(env="sa", continuous=bool(cont), envs=n, steps=steps, **modes))


def soak_ma(n, steps, cont, A=5, uniform=False, **modes):
    steps *= SCALE
    return _job("MA cont=%d A=%d %s %d envs x %d steps" % (cont, A, modes or "", n, steps), lambda poly: _run_ma(n, steps, cont, A, uniform, poly, **modes),
                dict(env="ma", agents=A, continuous=bool(cont), uniform_actions=bool(uniform), envs=n, steps=steps, **modes))


if __name__ == "__main__":
    total = (soak_sa(65536, 600, False) + soak_sa(65536, 300, True) + soak_ma(24576, 500, False) + soak_ma(24576, 300, True)
             + soak_ma(9999, 300, False, A=3) + soak_ma(65536, 250, False, uniform=True) + soak_ma(65536, 150, True, uniform=True)
             + soak_ma(20011, 200, False, A=1) + soak_ma(20011, 200, True, A=2) + soak_ma(20011, 200, False, A=4)
             + soak_ma(8191, 200, False, A=8) + soak_ma(8191, 150, True, A=6)
             # the other modes of the path: 60 Hz clock (15 ticks per step), no respawns, shared reward
             + soak_sa(32768, 150, False, real_time=True) + soak_sa(32768, 150, True, inf_obs=False)
             + soak_ma(16384, 120, False, uniform=True, shared_reward=True) + soak_ma(16384, 100, False, uniform=True, real_time=True)
             + soak_ma(16384, 100, True, A=3, uniform=True, inf_obs=False))
    steps = sum(r["envs"] * r["steps"] for r in RESULTS)
    explained = sum(r["mismatching_values"] for r in RESULTS if r.get("explained_by_the_tangent_deviation"))
    print("long parity:", "OK" if total == 0 else "FAILED (%d unexplained)" % total, "- %.2e env-steps compared, %d values explained by the tangent deviation"
          % (steps, explained))
    if len(sys.argv) > 1:
        json.dump(dict(what="CUDA kernels vs the C oracle, value by value (bit-exact mode), scripts/long_parity.py", env_steps_compared=steps,
                       unexplained_mismatches=total, values_explained_by_the_tangent_deviation=explained, seed_offset=SEED0, runs=RESULTS),
                  open(sys.argv[1], "w"), indent=1)
    sys.exit(0 if total == 0 else 1)
