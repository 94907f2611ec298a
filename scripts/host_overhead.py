"""Host cost of one env.step() call: wall time per step of a tight Python loop over a SMALL batch (the GPU is never the bound:
a 4 096-env step kernel takes ~11 us) - what a trainer that steps small batches pays per call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_highway_b200 import HighwayVecEnv, HighwayMAVecEnv


def loop(env, acts, steps):
    for w in range(50):
        env.step_async(acts[w % 8]); env.step_wait()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        env.step_async(acts[k % 8]); env.step_wait()
    t1 = time.perf_counter()       # host time to ISSUE the steps
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return 1e6 * (t1 - t0) / steps, 1e6 * (t2 - t0) / steps


if __name__ == "__main__":
    n, steps = 4096, 3000
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    env = HighwayVecEnv(n, seed=0)
    acts = torch.randint(0, 4, (8, n), generator=g, device='cuda', dtype=torch.int32)
    print("SA  n=%d: %.1f us of host time per step() call, %.1f us per step end to end" % ((n,) + loop(env, acts, steps)))
    obs, rew, done = torch.empty(n, 18, device='cuda'), torch.empty(n, device='cuda'), torch.empty(n, dtype=torch.uint8, device='cuda')
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(steps):
        env.step_into(acts[k % 8], obs, rew, done)
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print("SA  n=%d: %.1f us of host time per step_into() call" % (n, 1e6 * (t1 - t0) / steps))
    env = HighwayMAVecEnv(n, num_agents=5, seed=0)
    acts = torch.randint(0, 4, (8, n, 5), generator=g, device='cuda', dtype=torch.int32)
    print("MA  n=%d: %.1f us of host time per step() call, %.1f us per step end to end" % ((n,) + loop(env, acts, steps)))
