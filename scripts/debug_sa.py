import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests'))
import numpy as np, torch
from oracle import oracle as orc
from gym_highway_b200 import HighwayVecEnv
n, seed = 4096, 5
for env_id0 in (0, 100):
    env = HighwayVecEnv(n, seed=seed, env_id0=env_id0)
    st = orc.SaState(n); orc.sa_reset(st)
    rng = np.random.RandomState(1)
    for t in range(3):
        act = rng.randint(0, 4, n).astype(np.int32)
        obs, rew, done, infos = env.step(torch.from_numpy(act).cuda())
        want = orc.sa_step(st, act, seed=seed, env_id0=env_id0, quantise=True, nthreads=8)
        o = obs.cpu().numpy()
        bad = np.argwhere(o.view(np.uint32) != want['obs'].view(np.uint32))
        print('env_id0', env_id0, 'step', t, 'nbad', len(bad), 'rows', len(np.unique(bad[:,0])) if len(bad) else 0, 'cols', np.unique(bad[:,1]) if len(bad) else [])
        for r, c in bad[:5]:
            print('   row', r, 'col', c, 'act', act[r], 'gpu', repr(o[r, c]), 'cpu', repr(want['obs'][r, c]), 'diff', float(o[r,c]) - float(want['obs'][r,c]))
