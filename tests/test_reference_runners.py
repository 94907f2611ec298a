"""CPU: the host-side consumers of the env against goldens recorded from the reference's own
classes (oracle/gen_golden_vec.py): ppo2's Runner, a2c's Runner, MADDPG's Memory / store_transition / obs_rms.
The env underneath is the C oracle here; tests/test_reference_gpu.py repeats them over the CUDA env."""
import numpy as np
import torch

from helpers import OracleVecEnv, check_a2c_golden, check_ppo2_golden, load_golden


def _make_env(n, cont, seed):
    return OracleVecEnv(n, continuous=cont, seed=seed)


def test_ppo2_runner_matches_reference_runner():
    from gym_highway_b200.rollout import Runner
    assert check_ppo2_golden(load_golden("runner_ppo2_discrete"), _make_env, Runner) >= 3
    check_ppo2_golden(load_golden("runner_ppo2_continuous"), _make_env, Runner)


def test_a2c_runner_matches_reference_runner():
    from gym_highway_b200.rollout import A2CRunner
    assert check_a2c_golden(load_golden("runner_a2c_discrete"), _make_env, A2CRunner) >= 2


def test_ring_memory_matches_reference_memory():
    """RingMemory / RunningMeanStd (CPU tensors) driven by the MADDPG rollout loop's bookkeeping over the oracle env against
    maddpg.trainer.ma_memory.Memory + the loop of ma_ddpg.py:225-262 over the reference VecEnv (tests/golden/memory_ma3.npz)"""
    from gym_highway_b200.replay import RingMemory, RunningMeanStd
    from helpers import check_memory_golden
    g = load_golden("memory_ma3")
    n, A, limit = int(g["n"]), int(g["num_agents"]), int(g["limit"])
    D = 4 * A + 14
    env = OracleVecEnv(n, continuous=True, seed=int(g["seed"]), num_agents=A)
    mem = RingMemory(limit, A, D, 2)
    rms = RunningMeanStd(shape=(A, D))
    obs = env.reset()
    ep_rew = torch.zeros(n, A); ep_step = torch.zeros(n, dtype=torch.int64)
    rews, steps = [], []
    for t in range(g["action"].shape[0]):
        act = torch.from_numpy(g["action"][t])
        new_obs, rew, done, _ = env.step(act)
        ep_rew += rew; ep_step += 1
        mem.append(obs, act, rew, new_obs, done)
        rms.update(obs, repeat=A)
        obs = new_obs
        for d in range(n):
            if bool(done[d].any()):
                rews.append(ep_rew[d].numpy().copy()); steps.append(int(ep_step[d]))
                ep_rew[d] = 0; ep_step[d] = 0
    assert mem.nb_entries == limit and len(steps) >= 3      # the ring wrapped
    check_memory_golden(g, mem, rms, (rews, steps))
