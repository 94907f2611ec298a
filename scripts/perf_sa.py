"""quick device-timed throughput probe of the single-agent step at several batch sizes"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_highway_b200 import HighwayVecEnv

def probe(n, cont=False, steps=60, exact=bool(int(os.environ.get('HW_EXACT', '0')))):
    env = HighwayVecEnv(n, continuous=cont, seed=0, exact_steering=exact)
    env.rollout_random(150)
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    acts = (torch.rand((8, n, 2), generator=g, device='cuda') * 2 - 1) if cont else torch.randint(0, 4, (8, n), generator=g, device='cuda', dtype=torch.int32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for w in range(5):
        env.step_async(acts[w % 8]); env.step_wait()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        if n * 273 < 256 << 20: flush.fill_(k & 255)
        ev[k][0].record(); env.step_async(acts[k % 8]); ev[k][1].record(); env.step_wait()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    med = ms[len(ms) // 2]
    bps = 277 if cont else 273
    print("n=%9d cont=%d  median %.1f us  %.3e env-steps/s  %.0f GB/s algorithmic (%.1f%% of 6554)" % (n, cont, med * 1e3, n / (med * 1e-3), n * bps / (med * 1e-3) / 1e9, 100 * n * bps / (med * 1e-3) / 1e9 / 6553.9))

if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:]] or [65536, 1 << 20, 1 << 22]
    for n in sizes:
        probe(n)
    probe(sizes[-1], cont=True)
