"""Where the time of a SMALL-batch step launch goes: launch latency vs. per-tick latency vs. cold caches.

Times the single-agent step at a few batch sizes with ticks_per_step forced to 1 / 5 / 9 (HwSimParams is a per-call
argument of the C ABI), with and without an L2 flush between launches, next to an empty launch pair of CUDA events.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_highway_b200 import HighwayVecEnv, HighwayMAVecEnv


def timed(fn, steps, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        if flush is not None:
            flush.fill_(k & 255)
        ev[k][0].record(); fn(k); ev[k][1].record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2] * 1e3, ms[0] * 1e3


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    x = torch.zeros(32, device='cuda')
    med, mn = timed(lambda k: x.add_(1), 50, None)
    print("tiny torch kernel between two events: median %.1f us, min %.1f us" % (med, mn))
    sizes = [int(a) for a in sys.argv[1:]] or [32, 4096, 65536]
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    for n in sizes:
        env = HighwayVecEnv(n, seed=0)
        env.rollout_random(150)
        acts = torch.randint(0, 4, (8, n), generator=g, device='cuda', dtype=torch.int32)
        for ticks in (1, 5, 9):
            env.params.ticks_per_step = ticks
            def step(k):
                env.step_async(acts[k % 8]); env.step_wait()
            for w in range(5): step(w)
            m_f, mn_f = timed(step, 40, flush)
            m_w, mn_w = timed(step, 40, None)
            print("SA n=%7d ticks=%d  flushed: median %.1f min %.1f us   warm: median %.1f min %.1f us" % (n, ticks, m_f, mn_f, m_w, mn_w))
    for n in (2048, 16384):
        for kern in ("group", "thread"):
            env = HighwayMAVecEnv(n, num_agents=5, seed=0, kernel=kern) if 'kernel' in HighwayMAVecEnv.__init__.__code__.co_varnames else None
            if env is None:
                break
            acts = torch.randint(0, 4, (8, n, 5), generator=g, device='cuda', dtype=torch.int32)
            for ticks in (1, 9):
                env.params.ticks_per_step = ticks
                def step(k):
                    env.step_async(acts[k % 8]); env.step_wait()
                for w in range(5): step(w)
                m_f, mn_f = timed(step, 40, flush)
                m_w, mn_w = timed(step, 40, None)
                print("MA(%s) n=%6d ticks=%d  flushed: median %.1f min %.1f us   warm: median %.1f min %.1f us" % (kern, n, ticks, m_f, mn_f, m_w, mn_w))


if __name__ == "__main__":
    main()
