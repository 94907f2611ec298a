"""GPU: the CUDA envs and the on-device consumers against goldens recorded from the reference's OWN upper layers
(oracle/gen_golden_vec.py: DummyVecEnv / DummyVecMAEnv + Monitor, ppo2 / a2c Runner, MADDPG Memory + rollout loop).
Same checks as tests/test_reference_vec.py / test_reference_runners.py, with the kernels underneath instead of the oracle."""
import numpy as np
import pytest
import torch

from helpers import (HashModel, check_a2c_golden, check_memory_golden, check_ppo2_golden, check_vec_golden, load_golden)

pytestmark = pytest.mark.gpu


def _infos_sa(infos):
    return ([i["run_time"] for i in infos], [i["num_obs_collisions"] for i in infos], [i["num_agent_collisions"] for i in infos])


def _infos_ma(infos):
    b = [i["n"][0] for i in infos]
    return ([i["run_time"] for i in b], [i["num_obs_collisions"] for i in b], [i["num_agent_collisions"] for i in b])


@pytest.mark.parametrize("name", ["vec_sa_discrete", "vec_sa_continuous"])
def test_sa_vecenv_matches_reference_vecenv(name):
    from gym_highway_b200 import HighwayVecEnv
    g = load_golden(name)
    env = HighwayVecEnv(g["done"].shape[1], continuous=bool(g["continuous"]), seed=int(g["seed"]), return_numpy=True, exact_steering=True)

    def step(t, a):
        obs, rew, done, infos = env.step(a)
        infos = list(infos)
        eps = {i: (inf["episode"]["r"], inf["episode"]["l"]) for i, inf in enumerate(infos) if "episode" in inf}
        return obs, rew, done, eps, _infos_sa(infos)

    ends, timeouts = check_vec_golden(g, env.reset, step)
    assert ends >= 10 and timeouts >= 2


@pytest.mark.parametrize("kernel", ["group", "thread"])
@pytest.mark.parametrize("name", ["vec_ma3_discrete", "vec_ma5_discrete", "vec_ma3_continuous", "vec_ma5_shared"])
def test_ma_vecenv_matches_reference_vecenv(name, kernel):
    from gym_highway_b200 import HighwayMAVecEnv
    g = load_golden(name)
    env = HighwayMAVecEnv(g["done"].shape[1], num_agents=int(g["num_agents"]), continuous=bool(g["continuous"]), seed=int(g["seed"]),
                          shared_reward=bool(g["shared_reward"]), return_numpy=True, exact_steering=True, kernel=kernel)

    def step(t, a):
        obs, rew, done, infos = env.step(a)
        infos = list(infos)
        eps = {i: (inf["episode"]["r"], inf["episode"]["l"]) for i, inf in enumerate(infos) if "episode" in inf}
        return obs, rew, done, eps, _infos_ma(infos)

    ends, _ = check_vec_golden(g, env.reset, step)
    assert ends >= 5 and int(env.diagnostics().max()) == 0


def _make_env(n, cont, seed):
    from gym_highway_b200 import HighwayVecEnv
    return HighwayVecEnv(n, continuous=cont, seed=seed, exact_steering=True)


def test_ppo2_runner_on_the_cuda_env_matches_reference_runner():
    from gym_highway_b200.rollout import Runner
    assert check_ppo2_golden(load_golden("runner_ppo2_discrete"), _make_env, Runner) >= 3
    check_ppo2_golden(load_golden("runner_ppo2_continuous"), _make_env, Runner)


def test_a2c_runner_on_the_cuda_env_matches_reference_runner():
    from gym_highway_b200.rollout import A2CRunner
    assert check_a2c_golden(load_golden("runner_a2c_discrete"), _make_env, A2CRunner) >= 2


def test_device_collector_matches_reference_memory():
    """MADeviceCollector (env step_into + hw_ring_store + hw_obs_moments) replays the action script the reference's MADDPG
    rollout loop was driven with and must leave the reference Memory's contents, ring position, obs_rms sums and episode
    records behind"""
    from gym_highway_b200 import HighwayMAVecEnv
    from gym_highway_b200.replay import MADeviceCollector, RingMemory, RunningMeanStd
    g = load_golden("memory_ma3")
    n, A, limit = int(g["n"]), int(g["num_agents"]), int(g["limit"])
    env = HighwayMAVecEnv(n, num_agents=A, continuous=True, seed=int(g["seed"]), exact_steering=True)
    mem = RingMemory(limit, A, env.obs_dim, 2, device=env.device)
    rms = RunningMeanStd(shape=(A, env.obs_dim), device=env.device)
    script = torch.from_numpy(g["action"]).cuda()
    col = MADeviceCollector(env, lambda obs: script[col.t], mem, obs_rms=rms)
    rews, steps = [], []
    for t in range(script.shape[0]):
        col.step()
        r, s, _ = col.finished_episodes()
        rews += [x.numpy() for x in r]; steps += [int(x) for x in s]
    assert mem.state.cpu().tolist() == [mem.start, mem.length, n * script.shape[0]]   # the device words and their host mirror agree
    check_memory_golden(g, mem, rms, (rews, steps))
    st = col.epoch_statistics()
    assert st["episodes"] == len(steps) and st["episode_steps"] == sum(steps)
    assert np.allclose(st["epoch_episode_rewards"].numpy(), np.sum(np.asarray(rews, np.float64), axis=0), rtol=1e-12)


def test_device_collector_graph_replay_equals_eager_collection():
    """MADeviceCollector.collect_graphed: the rollout loop replayed from one CUDA graph leaves the ring, its device-side position,
    the episode bookkeeping and obs_rms exactly where the eager loop leaves them"""
    from gym_highway_b200 import HighwayMAVecEnv
    from gym_highway_b200.replay import MADeviceCollector, RingMemory, RunningMeanStd
    torch.manual_seed(0)
    n, A, T = 3000, 5, 6
    actor = torch.nn.Sequential(torch.nn.Linear(34, 32), torch.nn.Tanh(), torch.nn.Linear(32, 2), torch.nn.Tanh()).cuda()

    def make():
        env = HighwayMAVecEnv(n, num_agents=A, continuous=True, seed=8)
        mem = RingMemory(n * 5 * T // 2, A, env.obs_dim, 2, device=env.device)     # wraps during the run
        rms = RunningMeanStd(shape=(A, env.obs_dim), device=env.device)
        return MADeviceCollector(env, lambda obs: actor(obs.reshape(-1, 34)), mem, obs_rms=rms), mem, rms

    (ce, me, re_), (cg, mg, rg) = make(), make()
    for _ in range(4):
        ce.collect(T)
        cg.collect_graphed(T)       # first call: eager warm-up + capture + one replay = two rollouts
    ce.collect(T)
    for k in ("obs0", "obs1", "actions", "rewards", "terminals1"):
        assert torch.equal(getattr(me, k), getattr(mg, k)), k
    assert (me.start, me.length) == (mg.start, mg.length) and mg.state.tolist() == me.state.tolist() == [mg.start, mg.length, 5 * T * n]
    assert torch.equal(ce.ep_reward, cg.ep_reward) and torch.equal(ce.ep_step, cg.ep_step) and torch.equal(ce.epoch_counts, cg.epoch_counts)
    assert torch.equal(re_._sum, rg._sum) and float(re_._count) == float(rg._count) and ce.t == cg.t == 5 * T
    with pytest.raises(ValueError):
        cg.collect_graphed(3)


def test_gae_and_nstep_kernels_are_bit_exact_with_the_reference_precisions():
    import ctypes as C
    from gym_highway_b200 import _lib
    from gym_highway_b200.rollout import discount_with_dones, gae
    lib = _lib.load()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    T, n = 37, 5001
    rew = torch.randn(T, n, generator=g, device="cuda")
    val = torch.randn(T, n, generator=g, device="cuda")
    done = (torch.rand(T + 1, n, generator=g, device="cuda") < 0.1).to(torch.uint8)
    last = torch.randn(n, generator=g, device="cuda")
    ret, adv, nst = torch.empty_like(rew), torch.empty_like(rew), torch.empty_like(rew)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda x: C.c_void_p(x.data_ptr())
    assert lib.hw_gae(p(rew), p(val), p(done), p(last), 0.99, 0.95, p(ret), p(adv), T, n, stream) == 0
    d = done.float()
    want_ret, want_adv = gae(rew, val, d[:T], last, d[T], 0.99, 0.95)
    assert torch.equal(adv, want_adv) and torch.equal(ret, want_ret)
    assert lib.hw_nstep_returns(p(rew), p(done), p(last), 0.99, p(nst), T, n, stream) == 0
    assert torch.equal(nst, discount_with_dones(rew, d[1:], 0.99, bootstrap=last))


@pytest.mark.parametrize("graphed", [False, True])
def test_device_a2c_runner(graphed):
    """DeviceA2CRunner: the stored transitions are the env's own, masks / n-step returns follow a2c/runner.py:48-66"""
    from gym_highway_b200 import HighwayVecEnv
    from gym_highway_b200.rollout import DeviceA2CRunner, MlpPolicy, discount_with_dones, sf01
    torch.manual_seed(0)
    n, T = 2048, 5
    pol = MlpPolicy(18, 4).cuda()
    run = DeviceA2CRunner(HighwayVecEnv(n, seed=4), pol, nsteps=T, gamma=0.99, seed=9)
    collect = run.run_graphed if graphed else run.run
    for _ in range(30):                                   # far enough for crashes to have happened
        collect()
    snapshot = {k: v.clone() for k, v in run.env.state_dict().items()}
    start_obs, start_done = run.mb_obs[T].clone(), run.mb_done_u8[T].clone()
    obs, states, rewards, masks, actions, values = [x.clone() if x is not None else None for x in collect()]
    assert states is None and obs.shape == (n * T, 18) and actions.dtype == torch.int32
    o, a = obs.reshape(n, T, 18), actions.reshape(n, T)
    assert torch.equal(o[:, 0, :], start_obs) and torch.equal(masks.reshape(n, T)[:, 0] > 0.5, start_done.bool())
    ref = HighwayVecEnv(n, seed=4)
    ref.load_state_dict(snapshot)
    rews, dones = [], []
    for t in range(T):
        ob, r, d, _ = ref.step(a[:, t].contiguous())
        rews.append(r.clone()); dones.append(d.clone())
        if t + 1 < T:
            assert torch.equal(ob, o[:, t + 1, :]), t
            assert torch.equal(masks.reshape(n, T)[:, t + 1] > 0.5, d)
    with torch.no_grad():
        last_v = pol.value(ref._out[ref._cur]["obs"])
        head = pol.head_bf16(obs).float()
    assert torch.allclose(values, head[:, 4], atol=1e-6)
    want = discount_with_dones(torch.stack(rews), torch.stack(dones).float(), 0.99, bootstrap=last_v)
    assert torch.equal(rewards, sf01(want))
    assert sum(int(d.sum()) for d in dones) >= 0


@pytest.mark.parametrize("continuous", [False, True])
def test_device_runner_on_the_multi_agent_env(continuous):
    """ppo2-style collection over HighwayMAVecEnv: every car is a row of the shared policy's batch; the stored [T, N, A, ...]
    transitions are the env's own"""
    from gym_highway_b200 import HighwayMAVecEnv
    from gym_highway_b200.rollout import DeviceRunner, MlpPolicy
    torch.manual_seed(0)
    n, A, T = 1024, 5, 8
    D = 4 * A + 14
    pol = MlpPolicy(D, 4, continuous=continuous).cuda()
    run = DeviceRunner(HighwayMAVecEnv(n, num_agents=A, continuous=continuous, seed=4), pol, nsteps=T, seed=9)
    run.run_graphed()
    snapshot = {k: v.clone() for k, v in run.env.state_dict().items()}
    start_obs = run.mb_obs[T].clone()
    obs, returns, masks, actions, values, neglogp = [x.clone() for x in run.run_graphed()[:6]]
    assert obs.shape == (n * A * T, D) and torch.isfinite(returns).all()
    o = obs.reshape(n * A, T, D)
    assert torch.equal(o[:, 0, :], start_obs)
    act = actions.reshape(n, A, T, 2) if continuous else actions.reshape(n, A, T)
    ref = HighwayMAVecEnv(n, num_agents=A, continuous=continuous, seed=4)
    ref.load_state_dict(snapshot)
    ended = 0
    for t in range(T - 1):
        a = act[:, :, t].clamp(-1, 1).contiguous() if continuous else act[:, :, t].contiguous()
        ob, r, d, _ = ref.step(a)
        assert torch.equal(ob.reshape(n * A, D), o[:, t + 1, :]), t
        assert torch.equal(masks.reshape(n * A, T)[:, t + 1] > 0.5, d.reshape(-1))
        ended += int(d.any(dim=1).sum())
    assert ended > 0
    with torch.no_grad():
        head = pol.head_bf16(obs).float()
    assert torch.allclose(values, head[:, 2 if continuous else 4], atol=1e-6)


def test_collector_central_replay_over_nccl():
    """2 ranks over NCCL: all_gathered transition blocks + all_reduced moments leave the single-process ring behind
    (scripts/nccl_collector_check.py); needs two GPUs"""
    import json, os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run scripts/nccl_collector_check.py under `gpurun --gpus 2`)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29531", os.path.join(root, "scripts", "nccl_collector_check.py")],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, HW_N="2048", HW_T="30"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok"] is True
