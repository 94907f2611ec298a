"""quick device-timed throughput probe of the multi-agent step"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_highway_b200 import HighwayMAVecEnv

def probe(n, A=5, cont=False, steps=40, kernel=os.environ.get('HW_MA_KERNEL', 'auto')):
    env = HighwayMAVecEnv(n, num_agents=A, continuous=cont, seed=0, kernel=kernel, exact_steering=bool(int(os.environ.get('HW_EXACT', '0'))),
                          num_slots=int(os.environ.get('HW_MA_SLOTS', '16')))
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    acts = (torch.rand((8, n, A, 2), generator=g, device='cuda') * 2 - 1) if cont else torch.randint(0, 4, (8, n, A), generator=g, device='cuda', dtype=torch.int32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for w in range(30):
        env.step_async(acts[w % 8]); env.step_wait()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        if n * 1300 < 256 << 20: flush.fill_(k & 255)
        ev[k][0].record(); env.step_async(acts[k % 8]); ev[k][1].record(); env.step_wait()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    med = ms[len(ms) // 2]
    live = float(((env.head[0] >> 20) & 31).float().mean())
    print("MA n=%8d A=%d cont=%d  median %.1f us  %.3e env-steps/s  %.3e agent-steps/s  live obstacles %.2f" % (n, A, cont, med * 1e3, n / (med * 1e-3), n * A / (med * 1e-3), live))

if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:]] or [16384, 262144]
    for n in sizes:
        probe(n)
    probe(sizes[-1], cont=True)
