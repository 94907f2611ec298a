"""rows of a golden file the SA kernel does not reproduce (first-mismatch finder for tests/golden/sa_*_edge.npz)"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from helpers import load_golden, sa_state_from_golden, canonical, sa_state_from_canonical, f32
from gym_highway_b200 import HighwayVecEnv
from gym_highway_b200.layout import pack_sa, unpack_sa
for name in sys.argv[1:] or ["sa_discrete_edge", "sa_continuous_edge"]:
    g = load_golden(name)
    n = len(g["out_done"])
    env = HighwayVecEnv(n, continuous=bool(g["continuous"]), seed=int(g["seed"]), env_id0=0, auto_reset=False, exact_steering=True)
    p = pack_sa(canonical(sa_state_from_golden(g)))
    env.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    obs, rew, done, _ = env.step(torch.from_numpy(np.ascontiguousarray(g["action"])).cuda())
    sd = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
    got = sa_state_from_canonical(unpack_sa(sd["car_a"], sd["car_b"], sd["obst"], sd["stat"]))
    for i in range(n):
        bad = []
        if not np.array_equal(f32(got.car[i]).view(np.uint32), f32(g["out_car"][i]).view(np.uint32)): bad.append("car")
        if not np.array_equal(f32(got.obst[i]).view(np.uint32), f32(g["out_obst"][i]).view(np.uint32)): bad.append("obst")
        if (got.lane[i], got.flags[i], got.tick[i]) != tuple(g["out_lft"][i]): bad.append("lft")
        if bool(done[i]) != bool(g["out_done"][i]): bad.append("done")
        if not np.array_equal(obs[i].cpu().numpy().view(np.uint32), g["out_obs"][i].view(np.uint32)): bad.append("obs")
        if float(rew[i]) != float(np.float32(g["out_rew"][i])): bad.append("rew")
        if bad:
            print(name, "row", i, bad, "in car", g["in_car"][i], "lane/flags/tick", g["in_lane"][i], g["in_flags"][i], g["in_tick"][i], "act", g["action"][i])
            print("    got car", got.car[i], (got.lane[i], got.flags[i], got.tick[i]), float(rew[i]), bool(done[i]))
            print("   want car", g["out_car"][i], tuple(g["out_lft"][i]), g["out_rew"][i], g["out_done"][i])
    print(name, "checked", n)
