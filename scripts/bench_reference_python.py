"""BASELINE.json configs[0]: the reference's own CPU path, run where /root/reference exists (the build container).

    python scripts/bench_reference_python.py [--steps 10000] [--procs P]

Each process steps ONE unmodified reference env (gym_highway.envs.HighwayEnv under the pygame/gym stand-ins of oracle/shim, the
reference's own MT19937 respawns) with its own uniform random actions for `steps` env.steps, resetting on done, and is timed
by wall clock; prints per-core and aggregate env-steps/s.  This is the number bench.py's CPU arm cannot produce on the GPU
box (no /root/reference there): it times the C restatement instead and quotes this one.
"""
import argparse, json, multiprocessing as mp, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(args):
    rank, steps, kind = args
    import random
    import numpy as np
    from oracle import ref_harness as rh
    m = rh.load()
    random.seed(rank)
    rng = np.random.RandomState(rank)
    if kind == "ma":
        env = m['ma'].MultiAgentEnv(world_config=dict(rh.TRAIN_KW), num_agents=5)
        env.reset()
        t0 = time.perf_counter()
        for _ in range(steps):
            onehot = np.eye(4)[rng.randint(0, 4, 5)]
            _, _, done, _ = env.step(list(onehot))
            if any(done):
                env.reset()
        return steps / (time.perf_counter() - t0)
    env = m['sa'].HighwayEnv(**rh.TRAIN_KW)
    env.reset()
    t0 = time.perf_counter()
    for _ in range(steps):
        _, _, done, _ = env.step(int(rng.randint(0, 4)))
        if done:
            env.reset()
    return steps / (time.perf_counter() - t0)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--kind", default="sa", choices=["sa", "ma"])
    a = ap.parse_args()
    from oracle import ref_harness as rh
    if not rh.available():
        print(json.dumps({"unavailable": "reference tree not present (this script runs in the build container only)"}))
        sys.exit(0)
    one = worker((0, a.steps, a.kind))
    with mp.Pool(a.procs) as pool:
        t0 = time.perf_counter()
        rates = pool.map(worker, [(r, a.steps, a.kind) for r in range(a.procs)])
        wall = time.perf_counter() - t0
    print(json.dumps({"workload": "reference %s env, 1 env per process, random actions, %d steps" % (a.kind, a.steps),
                      "single_process_env_steps_per_s": one, "procs": a.procs,
                      "aggregate_env_steps_per_s": a.procs * a.steps / wall, "per_core_under_load": sum(rates) / len(rates)}))
