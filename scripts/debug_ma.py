import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from oracle import ma_oracle as mo
from gym_highway_b200 import HighwayMAVecEnv
A, continuous = 3, False
n, steps, seed = 2048, 400, 9
env = HighwayMAVecEnv(n, num_agents=A, continuous=continuous, seed=seed, env_id0=7)
st = mo.MaState(n, A); mo.ma_reset(st)
rng = np.random.RandomState(2)
for t in range(steps):
    act = np.where(rng.rand(n, A) < 0.6, 2, rng.randint(0, 4, (n, A))).astype(np.int32)
    obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
    want = mo.ma_step(st, act, continuous=continuous, seed=seed, env_id0=7, quantise=True, nthreads=8)
    o = obs.cpu().numpy()
    bad = np.argwhere(o.view(np.uint32) != want['obs'].view(np.uint32))
    if len(bad):
        print('step', t, 'nbad', len(bad))
        for e, a, c in bad[:6]:
            print('  env', e, 'car', a, 'col', c, 'gpu', repr(o[e, a, c]), 'cpu', repr(want['obs'][e, a, c]), 'done', done[e].cpu().numpy(), want['done'][e])
        from gym_highway_b200.layout import unpack_ma
        sd = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
        c = unpack_ma(sd['cars'], sd['obst'], sd['head'], A, 16)
        e = bad[0][0]
        print('gpu car', c['car'][e]); print('cpu car', st.car[e]); print('diff', c['car'][e] - st.car[e])
        print('nobs', c['nobs'][e], st.nobs[e], 'tick', c['tick'][e], st.tick[e])
        break
