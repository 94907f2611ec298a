"""CPU: host-side logic that needs no GPU — sharding, gloo collectives (world_size 2), GAE / returns / sf01
against numpy restatements of the baselines runners, spaces and the run-time table."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_highway_b200 import distributed as hd
from gym_highway_b200 import rollout
from gym_highway_b200.spaces import Box, Discrete, highway_obs_bounds


def test_shard_covers_everything_once():
    for total in (1, 7, 64, 1000, 1048576):
        for ws in (1, 2, 3, 4, 8):
            spans = [hd.shard(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (a0, an), (b0, _) in zip(spans, spans[1:]):
                assert a0 + an == b0
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, ws, port, total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    env_id0, n_local = hd.shard(total, rank, ws)
    # every rank "finishes" one episode per local env with return -(global id) milli and length id
    ids = torch.arange(env_id0, env_id0 + n_local)
    stats = torch.zeros(8, dtype=torch.int64)
    stats[0] = n_local; stats[1] = ids.sum(); stats[2] = -ids.sum(); stats[3] = rank
    red = hd.reduce_statistics(stats)
    roll = torch.arange(3 * n_local, dtype=torch.float32).reshape(3, n_local) + 1000 * rank
    gat = hd.gather_env_axis(roll, axis=1)
    mx = hd.max_over_ranks(1.0 + rank)
    q.put((rank, red.tolist(), tuple(gat.shape), gat[0].tolist(), mx, stats.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world_size_2_statistics_and_gather():
    ws, total = 2, 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, total, q)) for r in range(ws)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(ws))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ids = np.arange(total)
    for rank, red, shape, row0, mx, local in res:
        assert red[0] == total and red[1] == ids.sum() and red[2] == -ids.sum() and red[3] == 1
        assert shape == (3, total) and mx == 2.0
        assert row0 == [0.0, 1.0, 2.0, 3.0, 4.0, 1000.0, 1001.0, 1002.0, 1003.0, 1004.0]
        assert local[0] == 5  # the local counters are left untouched
    d = hd.statistics_dict(torch.tensor(res[0][1]))
    assert d["episodes"] == total and abs(d["mean_length"] - 4.5) < 1e-12 and abs(d["mean_return"] + 0.0045) < 1e-12


def _np_gae(rew, val, dones, last_val, last_done, gamma, lam):
    # baselines/ppo2/runner.py:53-65 restated
    T = rew.shape[0]
    adv = np.zeros_like(rew); last = 0
    for t in reversed(range(T)):
        if t == T - 1:
            nnt, nv = 1.0 - last_done, last_val
        else:
            nnt, nv = 1.0 - dones[t + 1], val[t + 1]
        delta = rew[t] + gamma * nv * nnt - val[t]
        adv[t] = last = delta + gamma * lam * nnt * last
    return adv + val


def test_gae_sf01_and_nstep_returns_match_the_runner_formulas():
    rng = np.random.RandomState(0)
    T, N = 17, 6
    rew, val = rng.randn(T, N).astype(np.float32), rng.randn(T, N).astype(np.float32)
    dones = (rng.rand(T, N) < 0.2).astype(np.float32)
    lv, ld = rng.randn(N).astype(np.float32), (rng.rand(N) < 0.3).astype(np.float32)
    ret, adv = rollout.gae(*(torch.from_numpy(x) for x in (rew, val, dones, lv, ld)), 0.99, 0.95)
    assert np.allclose(ret.numpy(), _np_gae(rew, val, dones, lv, ld, 0.99, 0.95), atol=1e-5)
    x = torch.arange(T * N * 3).reshape(T, N, 3)
    f = rollout.sf01(x)
    assert f.shape == (T * N, 3) and torch.equal(f[1], x[1, 0]) and torch.equal(f[T], x[0, 1])
    # a2c n-step returns: per env, discount_with_dones(rewards + [value], dones + [0])[:-1] when the last step is not terminal
    out = rollout.discount_with_dones(torch.from_numpy(rew), torch.from_numpy(dones), 0.9, bootstrap=torch.from_numpy(lv)).numpy()
    for n in range(N):
        r = 0.0 if dones[-1, n] else lv[n]
        want = np.zeros(T, np.float32)
        for t in reversed(range(T)):
            r = rew[t, n] + 0.9 * r * (1.0 - dones[t, n])
            want[t] = r
        assert np.allclose(out[:, n], want, atol=1e-5)


def test_a2c_runner_matches_the_reference_loop_on_a_toy_env():
    """A2CRunner against a literal transcription of baselines/a2c/runner.py:21-72 + a2c/utils.py discount_with_dones,
    over a deterministic toy VecEnv (the runner only needs reset/step/num_envs)"""
    class ToyEnv(object):
        num_envs, continuous = 6, False
        def __init__(self):
            self.t = 0
        def reset(self):
            self.t = 0
            return torch.zeros(self.num_envs, 18)
        def step(self, actions):
            self.t += 1
            g = torch.Generator().manual_seed(self.t)
            obs = torch.rand(self.num_envs, 18, generator=g)
            rew = torch.rand(self.num_envs, generator=g) - 0.5
            done = torch.rand(self.num_envs, generator=g) < 0.3
            return obs, rew, done, None
    torch.manual_seed(1)
    pol = rollout.MlpPolicy(18, 4)
    T, gamma = 7, 0.9
    run = rollout.A2CRunner(ToyEnv(), pol, nsteps=T, gamma=gamma, generator=torch.Generator().manual_seed(3))
    for _ in range(2):   # second rollout starts mid-episode with carried dones
        obs0, dones0 = run.obs.clone(), run.dones.clone()
        env_t0 = run.env.t
        gstate = run.generator.get_state()
        obs, states, rewards, masks, actions, values = run.run()
        # reference loop
        ref_env = ToyEnv(); ref_env.t = env_t0
        g = torch.Generator(); g.set_state(gstate)
        o, d = obs0, dones0
        mb_r, mb_d, mb_v = [], [], []
        with torch.no_grad():
            for n in range(T):
                a, v, _ = pol.step(o, generator=g)
                mb_d.append(d.clone()); mb_v.append(v)
                o, r, dn, _ = ref_env.step(a)
                d = dn.float(); mb_r.append(r)
            mb_d.append(d.clone())
            last = pol.value(o).tolist()
        R = torch.stack(mb_r).t().tolist(); D = torch.stack(mb_d).t()
        m_ref, d_ref = D[:, :-1], D[:, 1:].tolist()
        def dwd(rewards, dones, gamma):
            out, r = [], 0.0
            for rew, dn in zip(rewards[::-1], dones[::-1]):
                r = rew + gamma * r * (1.0 - dn); out.append(r)
            return out[::-1]
        want = []
        for rews, dns, val in zip(R, d_ref, last):
            want.append(dwd(rews + [val], dns + [0], gamma)[:-1] if dns[-1] == 0 else dwd(rews, dns, gamma))
        assert states is None and obs.shape == (6 * T, 18) and actions.shape == (6 * T,)
        assert torch.allclose(rewards.reshape(6, T), torch.tensor(want), atol=1e-5)
        assert torch.equal(masks.reshape(6, T), m_ref)
        assert torch.allclose(values.reshape(6, T), torch.stack(mb_v).t(), atol=1e-6)


def test_policy_shapes_on_cpu():
    pol = rollout.MlpPolicy(18, 4)
    a, v, nl = pol.step(torch.rand(5, 18))
    assert a.shape == (5,) and a.dtype == torch.int32 and v.shape == (5,) and nl.shape == (5,) and (nl > 0).all()
    pc = rollout.MlpPolicy(18, continuous=True)
    a, v, nl = pc.step(torch.rand(5, 18))
    assert a.shape == (5, 2)
    # the continuous sample is returned RAW (neglogp belongs to it); only the env's copy is clamped
    pc.logstd.data.fill_(2.0)
    a, v, nl = pc.step(torch.rand(4000, 18))
    mean, _ = pc.forward(torch.rand(1, 18))
    assert a.abs().max() > 1.0
    assert sum(p.numel() for p in pol.body.parameters()) == 18 * 256 + 256 + 2 * (256 * 256 + 256)


def test_policy_bf16_inference_path_and_sampler():
    """rollout-time forward (bf16 weights, fused bias+ReLU GEMMs, padded widths) tracks the fp32 modules, is refreshed when
    the parameters change, and the inverse-CDF sampler follows the categorical distribution"""
    from gym_highway_b200.rollout import MlpPolicy
    torch.manual_seed(0)
    pol = MlpPolicy(18, 4)
    x = torch.rand(2048, 18)
    with torch.no_grad():
        logits, v = pol.forward(x)
        l16, v16 = pol._forward_bf16(x)
        assert l16.shape == logits.shape and v16.shape == v.shape
        assert (l16 - logits).abs().max() < 5e-3 and (v16 - v).abs().max() < 5e-3
        ptrs = [w.data_ptr() for w, _ in pol._bf16_body] + [pol._bf16_head_w.data_ptr(), pol._bf16_head_b.data_ptr()]
        pol.pi.bias.add_(1.0)                      # in-place update, as an optimiser step would do
        l16b, _ = pol._forward_bf16(x)
        assert ((l16b - l16) - 1.0).abs().max() < 2e-2
        # the bf16 copies are refreshed IN PLACE: a CUDA graph holding their addresses stays valid
        assert [w.data_ptr() for w, _ in pol._bf16_body] + [pol._bf16_head_w.data_ptr(), pol._bf16_head_b.data_ptr()] == ptrs
        # sampler: empirical frequencies of one fixed distribution
        pol.pi.weight.zero_(); pol.pi.bias.copy_(torch.log(torch.tensor([0.1, 0.2, 0.3, 0.4])))
        a, val, nlp = pol.step(torch.rand(200000, 18))
        assert a.dtype == torch.int32 and a.min() >= 0 and a.max() <= 3
        freq = torch.bincount(a.long(), minlength=4).double() / a.numel()
        assert (freq - torch.tensor([0.1, 0.2, 0.3, 0.4], dtype=torch.float64)).abs().max() < 5e-3
        assert torch.allclose(nlp, -torch.log(torch.tensor([0.1, 0.2, 0.3, 0.4]))[a.long()], atol=1e-5)


def test_spaces_and_bounds():
    low, high = highway_obs_bounds(4, np.float32)
    assert low.shape == (18,) and low.dtype == np.float32
    assert low.tolist() == [-5, -30, 0, 0] + [-60, -8] * 3 + [-20, -20] * 4
    assert high.tolist() == [5, 30, 1200, 8] + [60, 8] * 3 + [20, 20] * 4
    low8, _ = highway_obs_bounds(8)
    assert low8.shape == (34,) and low8.dtype == np.float64
    b = Box(low, high, dtype=np.float32)
    assert b.shape == (18,) and b.contains(b.sample()) and Discrete(4).contains(3) and not Discrete(4).contains(4)


def test_run_time_table_accumulates_like_the_reference():
    from gym_highway_b200 import _lib
    from gym_highway_b200.vec_env import _run_time_table
    tab = _run_time_table(_lib.sim_params(False))
    assert tab[9] == 0.25000000000000006 and tab[2160] == 60.00000000000145 and tab[2159] < 60.0


def test_make_vec_env_rejects_unknown_ids():
    from gym_highway_b200 import ENV_IDS
    from gym_highway_b200.cmd_util import make_vec_env
    assert ENV_IDS['Highway-v0'] == ('sa', False) and ENV_IDS['MA-Highway-cont-train-v0'] == ('ma', True)
    with pytest.raises(KeyError):
        make_vec_env('CartPole-v0')


def test_make_vec_env_gives_every_rank_a_disjoint_env_id_range(monkeypatch):
    """with the reference's fixed 10000*rank stride two ranks holding 131072 envs each would share most
    of their Philox streams; the id range of rank r is [start + r*num_env, start + (r+1)*num_env)"""
    from gym_highway_b200 import cmd_util, vec_env, vec_ma_env
    seen = []

    class Fake(object):
        def __init__(self, num_envs, **kw):
            seen.append((num_envs, kw["env_id0"]))
    monkeypatch.setattr(vec_env, "HighwayVecEnv", Fake)
    monkeypatch.setattr(vec_ma_env, "HighwayMAVecEnv", Fake)
    for rank in range(4):
        monkeypatch.setenv("RANK", str(rank))
        cmd_util.make_vec_env('Highway-train-v0', num_env=131072, seed=1, start_index=5)
        cmd_util.make_vec_env('MA-Highway-train-v0', num_env=16384, seed=1)
    sa = sorted((lo, lo + n) for n, lo in seen if n == 131072)
    ma = sorted((lo, lo + n) for n, lo in seen if n == 16384)
    for rng_ in (sa, ma):
        assert all(a[1] == b[0] for a, b in zip(rng_, rng_[1:])), rng_   # contiguous, no overlap
    assert sa[0][0] == 5 and ma[0][0] == 0


class _FakeInfos(object):
    def __init__(self, eps):
        self._eps = eps

    def episodes(self):
        return self._eps


class _FakeVenv(object):
    num_envs = 4

    def __init__(self):
        self.t = 0

    def reset(self):
        return "obs0"

    def step_async(self, a):
        self.t += 1

    def step_wait(self):
        eps = [dict(env=1, r=-12.345678, l=7, t=0.0)] if self.t % 2 == 0 else []
        return "obs", "rew", "done", _FakeInfos(eps)

    def close(self):
        pass


def test_monitor_file_format(tmp_path):
    from gym_highway_b200.monitor import VecMonitor, load_results
    path = str(tmp_path / "run0")
    m = VecMonitor(_FakeVenv(), filename=path, env_id="Highway-v0")
    assert m.reset() == "obs0"
    for _ in range(6):
        m.step([0, 0, 0, 0])
    m.close()
    lines = open(path + ".monitor.csv").read().splitlines()
    assert lines[0].startswith('# {') and '"env_id": "Highway-v0"' in lines[0] and lines[1] == "r,l,t"
    header, rows = load_results(path + ".monitor.csv")
    assert len(rows) == 3 and rows[0]["r"] == -12.345678 and rows[0]["l"] == 7 and header["env_id"] == "Highway-v0"
    assert m.get_total_steps() == 24 and m.means() == (-12.345678, 7.0) and m.get_episode_lengths() == [7, 7, 7]


def test_ring_memory_matches_sequential_appends():
    from gym_highway_b200.replay import RingMemory, RunningMeanStd
    A, od, ad, limit = 3, 5, 2, 10
    mem = RingMemory(limit, A, od, ad)
    ref = []  # what a python list capped at `limit` would hold
    g = torch.Generator().manual_seed(0)
    for it in range(5):
        n = 4
        o0, o1 = torch.rand(n, A, od, generator=g), torch.rand(n, A, od, generator=g)
        a, r, d = torch.rand(n, A, ad, generator=g), torch.rand(n, A, generator=g), (torch.rand(n, A, generator=g) < 0.3).float()
        mem.append(o0, a, r, o1, d)
        for i in range(n):
            ref.append((o0[i], a[i], r[i], o1[i], d[i]))
        ref = ref[-limit:]
        assert mem.nb_entries == len(ref)
        for i in (0, len(ref) // 2, len(ref) - 1):
            e = mem.get(i)
            assert torch.equal(e["obs0"], ref[i][0]) and torch.equal(e["actions"], ref[i][1]) and torch.equal(e["obs1"], ref[i][3])
            assert torch.equal(e["rewards"][:, 0], ref[i][2]) and torch.equal(e["terminals1"][:, 0], ref[i][4])
    b = mem.sample(6, generator=g)
    assert b["obs0"].shape == (6, A, od) and b["rewards"].shape == (6, A, 1)
    with pytest.raises(KeyError):
        mem.get(limit)
    rms = RunningMeanStd(shape=(od,), epsilon=0.0)
    x = torch.randn(50, od, generator=g).double()
    rms.update(x[:20]); rms.update(x[20:])
    assert torch.allclose(rms.mean.double(), x.mean(0), atol=1e-6)
    assert torch.allclose(rms.std.double(), x.std(0, unbiased=False).clamp(min=0.1), atol=1e-6)


def test_packed_rect_filter_of_the_thread_per_env_kernel_is_exact():
    """The per-tick collision filter of csrc/hw_ma_tpe.cu (tpe_filter_word) restated in uint32 arithmetic: the masked word is zero iff
    the two 76 x 38 rects overlap (pygame.Rect.colliderect, highway_core.py:401-422) - for every offset around the boundaries and for
    rect coordinates on either side of zero (the packing wraps modulo 2^16 per field)."""
    MASK, BIAS, SIZE = np.uint32(0xFF00FF80), np.uint32((105 << 16) + 53), np.uint32((75 << 16) + 37)

    def pack(x, y):
        return ((x.astype(np.int64) << 16) + y.astype(np.int64)).astype(np.uint32)

    dx, dy = np.meshgrid(np.arange(-400, 401), np.arange(-200, 201), indexing="ij")
    with np.errstate(over="ignore"):
        for bx0, by0 in ((282, 101), (-1318, 21), (0, 0), (3162, 181), (-30, -19), (40000, 101), (-20000, 21)):
            bx, by = np.full_like(dx, bx0), np.full_like(dy, by0)
            ax, ay = bx + dx, by + dy
            pa, pb = pack(ax + 75, ay + 37), pack(bx, by)
            word = ((pa - pb + BIAS) | (pb - pa + (np.uint32(2) * SIZE + BIAS))) & MASK
            overlap = (ax < bx + 76) & (ay < by + 38) & (ax + 76 > bx) & (ay + 38 > by)
            assert np.array_equal(word == 0, overlap), (bx0, by0)
            # the one-word filter of round 2a is a superset only: it also passes a rect up to 180 px ahead
            one = (pa - pb + BIAS) & MASK
            assert np.all((one == 0) | ~overlap) and np.any((one == 0) & ~overlap)
