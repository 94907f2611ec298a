"""soak: long trajectories of larger batches against the C oracle; for the multi-agent env the two step kernels (a thread per
environment / a lane per car) run side by side on the same actions, so a data race in either or a difference between them shows up as
a mismatch.  python scripts/long_parity.py [summary.json]   (GPU box; a few minutes)"""
import json
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gym_highway_b200 import HighwayVecEnv, HighwayMAVecEnv
from oracle import oracle as orc, ma_oracle as mo


def soak_sa(n, steps, cont):
    env = HighwayVecEnv(n, continuous=cont, seed=21, env_id0=5, exact_steering=True)
    st = orc.SaState(n); orc.sa_reset(st)
    rng = np.random.RandomState(7)
    bad = 0
    for t in range(steps):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32) if cont else rng.randint(0, 4, n).astype(np.int32)
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        want = orc.sa_step(st, act, continuous=cont, seed=21, env_id0=5, quantise=True, nthreads=16)
        bad += int((done.cpu().numpy() != want["done"]).sum()) + int((obs.cpu().numpy().view(np.uint32) != want["obs"].view(np.uint32)).sum())
        bad += int((rew.cpu().numpy().view(np.uint32) != want["rew"].astype(np.float32).view(np.uint32)).sum())
    print("SA cont=%d  %d envs x %d steps: mismatching values %d" % (cont, n, steps, bad))
    RESULTS.append(dict(env="sa", continuous=bool(cont), envs=n, steps=steps, mismatching_values=bad, episodes=int(env.episode_statistics()["episodes"])))
    return bad


RESULTS = []


def soak_ma(n, steps, cont, A=5, uniform=False):
    ext = A > 5
    env = HighwayMAVecEnv(n, num_agents=A, continuous=cont, seed=22, env_id0=3, exact_steering=True, kernel="thread", extended_agents=ext)
    env2 = HighwayMAVecEnv(n, num_agents=A, continuous=cont, seed=22, env_id0=3, exact_steering=True, kernel="thread" if ext else "group",
                           extended_agents=ext)
    st = mo.MaState(n, A); mo.ma_reset(st)
    rng = np.random.RandomState(8)
    bad = nondet = 0
    for t in range(steps):
        if cont:
            act = rng.uniform(-1, 1, (n, A, 2)).astype(np.float32)
            if not uniform: act[:, :, 1] *= 0.1
        else:
            act = rng.randint(0, 4, (n, A)).astype(np.int32) if uniform else np.where(rng.rand(n, A) < 0.6, 2, rng.randint(0, 4, (n, A))).astype(np.int32)
        a = torch.from_numpy(act).cuda()
        obs, rew, done, _ = env.step(a)
        obs2, rew2, done2, _ = env2.step(a)
        nondet += int((obs != obs2).sum()) + int((done != done2).sum())
        want = mo.ma_step(st, act, continuous=cont, seed=22, env_id0=3, quantise=True, nthreads=16)
        bad += int((done.cpu().numpy() != want["done"]).sum()) + int((obs.cpu().numpy().view(np.uint32) != want["obs"].view(np.uint32)).sum())
    print("MA cont=%d A=%d  %d envs x %d steps: mismatching values %d, differences between the two kernels %d, max live obstacles %d, overflow %d"
          % (cont, A, n, steps, bad, nondet, int(st.nobs.max()), int(env.diagnostics().max())))
    RESULTS.append(dict(env="ma", agents=A, continuous=bool(cont), uniform_actions=bool(uniform), envs=n, steps=steps, mismatching_values=bad,
                        kernel_vs_kernel_differences=nondet, max_live_obstacles=int(st.nobs.max()), episodes=int(env.episode_statistics()["episodes"])))
    return bad + nondet


if __name__ == "__main__":
    total = (soak_sa(65536, 600, False) + soak_sa(65536, 300, True) + soak_ma(24576, 500, False) + soak_ma(24576, 300, True)
             + soak_ma(9999, 300, False, A=3) + soak_ma(65536, 250, False, uniform=True) + soak_ma(65536, 150, True, uniform=True)
             + soak_ma(20011, 200, False, A=1) + soak_ma(20011, 200, True, A=2) + soak_ma(20011, 200, False, A=4)
             + soak_ma(8191, 200, False, A=8) + soak_ma(8191, 150, True, A=6))
    steps = sum(r["envs"] * r["steps"] for r in RESULTS)
    print("long parity:", "OK" if total == 0 else "FAILED (%d)" % total, "- %.2e env-steps compared" % steps)
    if len(sys.argv) > 1:
        json.dump(dict(what="CUDA kernels vs the C oracle, value by value (bit-exact mode), scripts/long_parity.py", env_steps_compared=steps,
                       total_mismatches=total, runs=RESULTS), open(sys.argv[1], "w"), indent=1)
    sys.exit(0 if total == 0 else 1)
