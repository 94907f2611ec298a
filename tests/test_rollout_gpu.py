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


def test_graphed_rollout_replays_real_transitions():
    """the whole rollout replayed from one CUDA graph: the stored (obs, action, obs') triples are the env's own"""
    from gym_highway_b200 import HighwayVecEnv
    from gym_highway_b200.rollout import MlpPolicy, Runner
    torch.manual_seed(0)
    n, T = 1024, 16
    pol = MlpPolicy(18, 4).cuda()
    run = Runner(HighwayVecEnv(n, seed=1), pol, nsteps=T, generator=None)
    out1 = [x.clone() for x in run.run_graphed()[:6]]        # warm-up rollout, capture, first replay
    snapshot = {k: v.clone() for k, v in run.env.state_dict().items()}
    start_obs = run.obs.clone()
    obs, returns, masks, actions, values, neglogp = [x.clone() for x in run.run_graphed()[:6]]   # second replay
    assert obs.shape == (n * T, 18) and torch.isfinite(returns).all() and actions.min() >= 0 and actions.max() <= 3
    assert not torch.equal(out1[0], obs)
    o, a = obs.reshape(n, T, 18), actions.reshape(n, T)
    assert torch.equal(o[:, 0, :], start_obs)
    ref = HighwayVecEnv(n, seed=1)
    ref.load_state_dict(snapshot)
    for t in range(T - 1):
        ob, _, _, _ = ref.step(a[:, t].contiguous())
        assert torch.equal(ob, o[:, t + 1, :]), t
    with pytest.raises(ValueError):
        Runner(HighwayVecEnv(8, seed=1), pol, nsteps=3, generator=None).run_graphed()


@pytest.mark.parametrize("graphed", [False, True])
def test_device_runner_zero_copy_rollout(graphed):
    """DeviceRunner: fused sampling kernel + env outputs written straight into the [T, N] buffers; the stored transitions are
    the env's own, the stored values / neglogpacs are the policy's, fresh actions on every rollout (also under graph replay)"""
    from gym_highway_b200 import HighwayVecEnv
    from gym_highway_b200.rollout import MlpPolicy, DeviceRunner, gae, sf01
    torch.manual_seed(0)
    n, T = 2048, 12
    pol = MlpPolicy(18, 4).cuda()
    run = DeviceRunner(HighwayVecEnv(n, seed=4), pol, nsteps=T, seed=9)
    collect = run.run_graphed if graphed else run.run
    first = [x.clone() for x in collect()[:6]]
    snapshot = {k: v.clone() for k, v in run.env.state_dict().items()}
    start_obs = run.mb_obs[T].clone()
    obs, returns, masks, actions, values, neglogp = [x.clone() for x in collect()[:6]]
    assert obs.shape == (n * T, 18) and actions.dtype == torch.int32 and actions.min() >= 0 and actions.max() <= 3
    assert not torch.equal(first[3], actions)           # new random numbers on every rollout
    o, a = obs.reshape(n, T, 18), actions.reshape(n, T)
    assert torch.equal(o[:, 0, :], start_obs)
    ref = HighwayVecEnv(n, seed=4)
    ref.load_state_dict(snapshot)
    rews, dones = [], []
    for t in range(T):
        ob, r, d, _ = ref.step(a[:, t].contiguous())
        rews.append(r.clone()); dones.append(d.clone())
        if t + 1 < T:
            assert torch.equal(ob, o[:, t + 1, :]), t
    # values / neglogpacs: the policy's own, from the same bf16 head
    with torch.no_grad():
        head = pol.head_bf16(obs).float()
    logp = torch.log_softmax(head[:, :4], dim=-1)
    assert torch.allclose(values, head[:, 4], atol=1e-6)
    assert torch.allclose(neglogp, -logp.gather(1, actions.long().unsqueeze(1)).squeeze(1), atol=2e-3)
    # masks: episode ended before step t; returns: GAE over the env's rewards
    m = masks.reshape(n, T)
    for t in range(1, T):
        assert torch.equal(m[:, t] > 0.5, dones[t - 1])
    R = torch.stack(rews)                                   # [T, N]
    with torch.no_grad():
        last_v = pol.value(ref._out[ref._cur]["obs"])
    D = torch.cat([m.t(), dones[-1].float().unsqueeze(0)], 0)
    want, _ = gae(R, values.reshape(n, T).t().contiguous(), D[:T], last_v, D[T], 0.99, 0.95)
    assert torch.allclose(returns, sf01(want), atol=1e-3)
    # the sampler follows the policy: empirical frequency of each action vs the mean probability
    freq = torch.bincount(actions.long(), minlength=4).double() / actions.numel()
    assert (freq - logp.exp().double().mean(0)).abs().max() < 0.02


@pytest.mark.parametrize("runner", ["device", "runner_bf16"])
def test_graph_replay_follows_parameter_updates(runner):
    """a captured rollout must sample from the CURRENT policy.  The bf16 weight copies are persistent
    buffers refreshed in place inside run(), so the refresh is part of the graph: after an in-place parameter update
    (what optimizer.step() does) a replay returns values / neglogpacs of the updated policy, never of the capture-time one."""
    from gym_highway_b200 import HighwayVecEnv
    from gym_highway_b200.rollout import MlpPolicy, DeviceRunner, Runner
    torch.manual_seed(0)
    n, T = 1024, 8
    pol = MlpPolicy(18, 4, bf16_inference=(runner == "runner_bf16")).cuda()
    env = HighwayVecEnv(n, seed=2)
    run = DeviceRunner(env, pol, nsteps=T, seed=1) if runner == "device" else Runner(env, pol, nsteps=T, generator=None)
    run.run_graphed()
    ptrs = [w.data_ptr() for w, _ in pol._bf16_body] + [pol._bf16_head_w.data_ptr()]
    obs, _, _, actions, values, neglogp = [x.clone() for x in run.run_graphed()[:6]]
    with torch.no_grad():
        head = pol.head_bf16(obs).float()
    assert torch.allclose(values, head[:, 4], atol=1e-6)
    with torch.no_grad():   # an "optimizer step": every parameter moves, the value head by a lot
        for p_ in pol.parameters():
            p_.add_(0.05 * torch.randn_like(p_))
        pol.vf.bias.add_(3.0)
    obs2, _, _, actions2, values2, neglogp2 = [x.clone() for x in run.run_graphed()[:6]]
    assert [w.data_ptr() for w, _ in pol._bf16_body] + [pol._bf16_head_w.data_ptr()] == ptrs  # never rebound
    with torch.no_grad():
        head2 = pol.head_bf16(obs2).float()
    assert torch.allclose(values2, head2[:, 4], atol=1e-6), "the replay used stale weights"
    assert float((values2.mean() - values.mean()).abs()) > 1.0
    logp2 = torch.log_softmax(head2[:, :4], dim=-1)
    assert torch.allclose(neglogp2, -logp2.gather(1, actions2.long().unsqueeze(1)).squeeze(1), atol=2e-3)


def test_device_runner_continuous_env():
    """Gaussian sampler + continuous env: stored transitions are the env's own for the clamped actions, neglogpacs follow
    DiagGaussianPd, normals have unit variance"""
    from gym_highway_b200 import HighwayVecEnv
    from gym_highway_b200.rollout import MlpPolicy, DeviceRunner
    torch.manual_seed(0)
    n, T = 4096, 10
    pol = MlpPolicy(18, continuous=True).cuda()
    with torch.no_grad():
        pol.logstd.copy_(torch.tensor([-0.5, -1.0]))
    run = DeviceRunner(HighwayVecEnv(n, continuous=True, seed=6), pol, nsteps=T, seed=3)
    run.run_graphed()
    snapshot = {k: v.clone() for k, v in run.env.state_dict().items()}
    obs, returns, masks, actions, values, neglogp = [x.clone() for x in run.run_graphed()[:6]]
    assert actions.shape == (n * T, 2) and actions.dtype == torch.float32 and torch.isfinite(returns).all()
    o, a = obs.reshape(n, T, 18), actions.reshape(n, T, 2)
    ref = HighwayVecEnv(n, continuous=True, seed=6)
    ref.load_state_dict(snapshot)
    for t in range(T - 1):
        ob, _, _, _ = ref.step(a[:, t].clamp(-1, 1).contiguous())
        assert torch.equal(ob, o[:, t + 1, :]), t
    with torch.no_grad():
        head = pol.head_bf16(obs).float()
    std = pol.logstd.detach().exp()
    z = (actions - head[:, :2]) / std
    want = 0.5 * (z ** 2).sum(-1) + 0.5 * 1.8378770664093453 * 2 + pol.logstd.detach().sum()
    assert torch.allclose(neglogp, want, atol=2e-3) and torch.allclose(values, head[:, 2], atol=1e-6)
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02


def test_gae_and_cast_kernels_match_torch():
    import ctypes as C
    from gym_highway_b200 import _lib
    from gym_highway_b200.rollout import gae
    lib = _lib.load()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    T, n = 37, 5001
    rew = torch.randn(T, n, generator=g, device="cuda")
    val = torch.randn(T, n, generator=g, device="cuda")
    done = (torch.rand(T + 1, n, generator=g, device="cuda") < 0.1).to(torch.uint8)
    last = torch.randn(n, generator=g, device="cuda")
    ret, adv = torch.empty_like(rew), torch.empty_like(rew)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.hw_gae(C.c_void_p(rew.data_ptr()), C.c_void_p(val.data_ptr()), C.c_void_p(done.data_ptr()), C.c_void_p(last.data_ptr()),
                    0.99, 0.95, C.c_void_p(ret.data_ptr()), C.c_void_p(adv.data_ptr()), T, n, stream)
    assert rc == 0
    d = done.float()
    want_ret, want_adv = gae(rew, val, d[:T], last, d[T], 0.99, 0.95)
    assert torch.equal(adv, want_adv) and torch.equal(ret, want_ret)  # same precisions operation by operation
    # cast + pad
    x = torch.randn(777, 18, generator=g, device="cuda")
    y = torch.empty(777, 24, dtype=torch.bfloat16, device="cuda")
    assert lib.hw_cast_pad_bf16(C.c_void_p(x.data_ptr()), 18, C.c_void_p(y.data_ptr()), 24, 777, stream) == 0
    assert torch.equal(y[:, :18], x.to(torch.bfloat16)) and bool((y[:, 18:] == 0).all())
    # the sampler: frequencies of a fixed distribution, float32 and bfloat16 heads
    probs = torch.tensor([0.05, 0.15, 0.3, 0.5], device="cuda")
    for dt in (torch.float32, torch.bfloat16):
        head = torch.zeros(200000, 8, dtype=dt, device="cuda")
        head[:, :4] = probs.log().to(dt)
        head[:, 4] = 1.5
        a = torch.empty(200000, dtype=torch.int32, device="cuda"); nl = torch.empty(200000, device="cuda"); v = torch.empty(200000, device="cuda")
        rc = lib.hw_policy_sample_discrete(C.c_void_p(head.data_ptr()), int(dt == torch.bfloat16), 8, 4, 7, 0, 3, None,
                                           C.c_void_p(a.data_ptr()), C.c_void_p(nl.data_ptr()), C.c_void_p(v.data_ptr()), 200000, stream)
        assert rc == 0
        p = torch.softmax(head[:, :4].float(), -1)[0]
        freq = torch.bincount(a.long(), minlength=4).float() / a.numel()
        assert (freq - p).abs().max() < 5e-3 and bool((v == 1.5).all())
        assert torch.allclose(nl, -p.log()[a.long()], atol=1e-5)


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
