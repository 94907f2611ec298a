"""eager MADDPG collection steps (for an ncu launch list: where the time of a collection step goes)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_highway_b200 import HighwayMAVecEnv
from gym_highway_b200.replay import MADeviceCollector, RingMemory, RunningMeanStd
n, A = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 5
env = HighwayMAVecEnv(n, num_agents=A, continuous=True, seed=0)
mem = RingMemory(n * 12, A, env.obs_dim, 2, device=env.device)
rms = RunningMeanStd(shape=(A, env.obs_dim), device=env.device)
actor = torch.nn.Sequential(torch.nn.Linear(env.obs_dim, 64), torch.nn.ReLU(), torch.nn.Linear(64, 2), torch.nn.Tanh()).cuda()
col = MADeviceCollector(env, lambda obs: actor(obs.reshape(-1, env.obs_dim)), mem, obs_rms=rms)
col.collect(6)
torch.cuda.synchronize()
