"""a few steps of every kernel on small ragged batches: the target of compute-sanitizer runs (memcheck / racecheck)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_highway_b200 import HighwayVecEnv, HighwayMAVecEnv
g = torch.Generator(device="cuda"); g.manual_seed(0)
for cont in (False, True):
    env = HighwayVecEnv(1000, continuous=cont, seed=1)
    env.rollout_random(60)
    for t in range(12):
        a = (torch.rand((1000, 2), generator=g, device="cuda") * 2 - 1) if cont else torch.randint(0, 4, (1000,), generator=g, device="cuda", dtype=torch.int32)
        env.step(a)
    env.observe(); env.reset()
    for kernel in ("group", "thread"):
        for A in (5, 3):
            ma = HighwayMAVecEnv(333, num_agents=A, continuous=cont, seed=2, kernel=kernel)
            for t in range(40):
                a = (torch.rand((333, A, 2), generator=g, device="cuda") * 2 - 1) if cont else torch.where(torch.rand((333, A), generator=g, device="cuda") < 0.6, 2, torch.randint(0, 4, (333, A), generator=g, device="cuda")).to(torch.int32)
                ma.step(a)
            assert int(ma.diagnostics().max()) == 0
torch.cuda.synchronize()
print("sanity ok")
