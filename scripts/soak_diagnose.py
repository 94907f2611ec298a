"""A soak mismatch, explained: re-runs one multi-agent soak job against the oracle with libm's tan (the reference's) and against the
oracle with the kernels' own tangent polynomial (oracle/tan_small.h).  Mismatches that exist only against libm are the declared 1-ulp
tangent deviation (deviation 3) reaching a stored float32; anything left against the polynomial oracle is a kernel bug.

    python scripts/soak_diagnose.py [envs] [steps] [seed_offset] [summary.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_highway_b200 import HighwayMAVecEnv
from oracle import oracle as orc, ma_oracle as mo


def run(n, steps, seed0, poly, A=5):
    orc.set_tan(poly)
    env = HighwayMAVecEnv(n, num_agents=A, seed=22 + seed0, env_id0=3, exact_steering=True, kernel="thread")
    st = mo.MaState(n, A); mo.ma_reset(st)
    rng = np.random.RandomState(8 + seed0)
    events = []
    for t in range(steps):
        act = rng.randint(0, 4, (n, A)).astype(np.int32)
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        want = mo.ma_step(st, act, seed=22 + seed0, env_id0=3, quantise=True, nthreads=16)
        o, d = obs.cpu().numpy(), done.cpu().numpy()
        bad = np.argwhere(o.view(np.uint32) != want["obs"].view(np.uint32))
        nd = int((d != want["done"]).sum())
        if len(bad) or nd:
            for e, a, c in bad:
                g, w = o[e, a, c], want["obs"][e, a, c]
                events.append(dict(step=t, env=int(e), car=int(a), col=int(c), gpu=float(g), cpu=float(w),
                                   ulps=int(abs(int(np.float32(g).view(np.int32)) - int(np.float32(w).view(np.int32))))))
            if nd:
                events.append(dict(step=t, done_mismatches=nd))
    orc.set_tan(False)
    return events


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    libm = run(n, steps, seed0, False)
    poly = run(n, steps, seed0, True)
    print("%d envs x %d steps (A = 5, uniform random discrete actions, seed offset %d): %.2e car-ticks" % (n, steps, seed0, n * steps * 5 * 9.0))
    print("against the oracle with libm tan (the reference's): %d mismatching values" % len(libm))
    for ev in libm[:40]:
        print("   ", ev)
    print("against the oracle with the kernels' tangent polynomial: %d mismatching values" % len(poly))
    for ev in poly[:40]:
        print("   ", ev)
    if len(sys.argv) > 4:
        json.dump(dict(what="scripts/soak_diagnose.py", envs=n, steps=steps, seed_offset=seed0, agents=5, libm_oracle_mismatches=libm,
                       polynomial_oracle_mismatches=poly), open(sys.argv[4], "w"), indent=1)
