"""GPU: on-device ppo2-style rollout collection over the batched env."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_runner_collects_a_full_episode_on_device():
    from gym_highway_b200 import HighwayVecEnv
    from gym_highway_b200.rollout import MlpPolicy, Runner
    torch.manual_seed(0)
    n, T = 512, 240
    env = HighwayVecEnv(n, seed=1)
    pol = MlpPolicy(18, 4).cuda()
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    run = Runner(env, pol, nsteps=T, gamma=0.99, lam=0.95, generator=g)
    obs, returns, masks, actions, values, neglogp, states, epinfos = run.run()
    assert obs.shape == (n * T, 18) and obs.is_cuda and returns.shape == (n * T,) and actions.shape == (n * T,)
    assert states is None and torch.isfinite(returns).all() and actions.min() >= 0 and actions.max() <= 3
    # T = 240 steps is exactly one time-out episode: every env finished at least one episode
    assert len(epinfos) >= n and all(1 <= e["l"] <= 240 for e in epinfos)
    s = env.episode_statistics()
    assert s["episodes"] == len(epinfos)
    # sf01 order: rows [e*T, (e+1)*T) belong to env e; the first row is the reset observation
    first = obs.reshape(n, T, 18)[:, 0, :]
    assert torch.equal(first, first[0:1].expand_as(first))
    # the observation stored at step t+1 is what the env returned at step t: replay env 0's actions on a fresh env
    env2 = HighwayVecEnv(n, seed=1)
    a = actions.reshape(n, T)
    o = obs.reshape(n, T, 18)
    for t in range(5):
        ob, _, _, _ = env2.step(a[:, t].contiguous())
        assert torch.equal(ob, o[:, t + 1, :])


def test_vec_monitor_on_the_real_env(tmp_path):
    from gym_highway_b200 import HighwayVecEnv
    from gym_highway_b200.monitor import VecMonitor, load_results
    env = VecMonitor(HighwayVecEnv(64, seed=3), filename=str(tmp_path / "m"), env_id="Highway-train-v0")
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for t in range(250):
        env.step(torch.randint(0, 4, (64,), generator=g, device="cuda", dtype=torch.int32))
    env.close()
    header, rows = load_results(str(tmp_path / "m.monitor.csv"))
    assert header["env_id"] == "Highway-train-v0" and len(rows) >= 64
    assert all(1 <= r["l"] <= 240 for r in rows) and any(r["l"] == 240 for r in rows)
    # a time-out episode of random actions has return in (-240, 0]; a crash adds -250 per hit obstacle
    assert all(r["r"] <= 0 for r in rows)
    assert len(env.get_episode_rewards()) == len(rows)
