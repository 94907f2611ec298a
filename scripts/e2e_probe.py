import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from gym_highway_b200 import HighwayVecEnv
n = 4194304
env = HighwayVecEnv(n, seed=0)
acts = np.random.randint(0, 4, (4, n)).astype(np.int32)
for w in range(3): env.step_host(acts[w])
h = env._host
def t(f, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("full step_host       %.3f ms" % t(lambda: env.step_host(acts[0])))
print("copy into pinned     %.3f ms" % t(lambda: h['act'].copy_(torch.as_tensor(acts[1]))))
print("done astype(bool)    %.3f ms" % t(lambda: h['done'].numpy().astype(bool)))
d_obs = env._out[0]['obs']
print("D2H obs only         %.3f ms" % t(lambda: h['obs'].copy_(d_obs, non_blocking=True)))
print("H2D act only         %.3f ms" % t(lambda: h['d_act'].copy_(h['act'], non_blocking=True)))
