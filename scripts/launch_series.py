import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_highway_b200 import HighwayVecEnv
n = 4194304
env = HighwayVecEnv(n, seed=0)
env.rollout_random(150)
g = torch.Generator(device='cuda'); g.manual_seed(1)
acts = torch.randint(0, 4, (8, n), generator=g, device='cuda', dtype=torch.int32)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(260)]
for k in range(260):
    ev[k][0].record(); env.step_async(acts[k % 8]); ev[k][1].record(); env.step_wait()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) * 1e3 for a, b in ev]
print(' '.join('%d' % x for x in ms))
tick = (env.car_b[:, 3].contiguous().view(torch.int32) & 0xFFFF)
print('tick histogram (ticks/9 % 240 buckets of 24):', torch.histc((tick // 9).float(), bins=10, min=0, max=240).int().tolist())
