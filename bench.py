#!/usr/bin/env python
"""Benchmark of the batched highway step path (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--workload sa_discrete|sa_continuous|ma_discrete|ma_continuous]
    python bench.py --impl reference ...     # the CPU arm: C restatement of the reference on all host cores

One "step" = one env.step() of every environment of the batch (9 simulator ticks, collision tests,
observation, reward, done, auto-reset) = ONE kernel launch.  Prints ONE JSON line on rank 0.

Timing: CUDA events on the launching stream around every step launch; the L2 (126 MB) is flushed
between timed launches by overwriting a 256 MiB buffer (outside the event pairs) unless the batch's
working set is itself larger than L2 (the default: 4,194,304 single-agent envs per GPU move 1.1 GB per
launch), so every launch reads its state from HBM.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SA_BYTES_PER_STEP = {False: 273, True: 277}  # 96 B state in + 96 B out + action + 72 B obs + 4 B rew + 1 B done
L2_BYTES = 126 * 1024 * 1024


def ma_bytes_per_step(A, K):
    return 2 * (36 * A + 16 * K + 32) + 4 * A + 4 * A * (4 * A + 14) + 4 * A + A


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--envs", type=int, default=0, help="environments PER GPU (weak scaling); 0 = 4,194,304 (SA) / 262,144 (MA)")
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements of the other BASELINE configs")
    ap.add_argument("--workload", default="sa_discrete")
    ap.add_argument("--agents", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=0, help="steps of the bounded CPU sample (0 = auto)")
    return ap.parse_args()


def workload_info(args):
    w = args.workload
    cont = w.endswith("continuous")
    if w.startswith("sa_"):
        return dict(kind="sa", continuous=cont, bytes_per_step=SA_BYTES_PER_STEP[cont],
                    name="single-agent 3-lane env, %d envs/GPU, %s random actions" % (args.envs, "continuous" if cont else "discrete"),
                    dtype="f64 tick arithmetic on f32 state")
    return dict(kind="ma", continuous=cont, bytes_per_step=ma_bytes_per_step(args.agents, 16),
                name="multi-agent highway env, %d cars x %d envs/GPU, %s random actions" % (args.agents, args.envs, "continuous" if cont else "discrete"),
                dtype="f64 tick arithmetic on f32 state")


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_arm(args, info, steps_hint=0, nthreads=None, warmup=3, full_batch=False):
    """time the C restatement of the reference step on the host cores: a bounded sample of the workload (65,536 envs per step for
    the cpu_baseline leg of the GPU arm; the workload's own batch for `--impl reference`, where a step of 4,194,304 envs is ~0.25 s)"""
    from oracle import oracle as orc
    nthreads = nthreads or os.cpu_count() or 1
    n = args.envs if full_batch else min(args.envs, 65536)
    rng = np.random.RandomState(1234)
    if info["kind"] == "sa":
        st = orc.SaState(n)
        orc.sa_reset(st)

        def one_step():
            act = rng.uniform(-1, 1, (n, 2)).astype(np.float32) if info["continuous"] else rng.randint(0, 4, n).astype(np.int32)
            t0 = time.perf_counter()
            orc.sa_step(st, act, continuous=info["continuous"], seed=0, quantise=True, nthreads=nthreads)
            return time.perf_counter() - t0
    else:
        from oracle import ma_oracle as mo
        st = mo.MaState(n, args.agents)
        mo.ma_reset(st)

        def one_step():
            A = args.agents
            act = rng.uniform(-1, 1, (n, A, 2)).astype(np.float32) if info["continuous"] else rng.randint(0, 4, (n, A)).astype(np.int32)
            t0 = time.perf_counter()
            mo.ma_step(st, act, continuous=info["continuous"], seed=0, quantise=True, nthreads=nthreads)
            return time.perf_counter() - t0
    for _ in range(max(3, warmup)):
        one_step()
    probe = one_step()
    # bounded: about 12 s of CPU work unless the caller names a step count, and never more than ~60 s
    steps = int(max(10, min(2000, 12.0 / max(probe, 1e-6)))) if not steps_hint else int(max(1, min(steps_hint, 60.0 / max(probe, 1e-6))))
    total = sum(one_step() for _ in range(steps))
    return dict(value=n * steps / total, unit="env-steps/s", cores=nthreads, kind="port",
                sample="%d envs x %d steps of the C restatement (oracle/*.c, fp64, pthreads), action generation excluded; "
                       "the Python reference itself cannot run on the GPU box (no /root/reference, no pygame/gym): "
                       "it measured ~4.9e3 env-steps/s/core single-agent in the build container (BASELINE.md section 2)" % (n, steps),
                ms_per_step=1e3 * total / steps, steps=steps, envs=n)


def reference_main(args, info):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" of this arm is one env.step of the arm's own batch (the same envs per GPU as the GPU arm) on all host cores; the
    # number of steps is the requested one, capped so that the run stays within a minute
    base = cpu_arm(args, info, steps_hint=args.cpu_steps or args.steps, warmup=args.warmup, full_batch=True)
    line = {"impl": "reference", "metric": "env-steps/s", "value": base["value"], "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": base["steps"], "warmup": max(3, args.warmup), "ms_per_step": base["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": info["name"], "envs_per_gpu": args.envs, "envs_per_step_sample": base["envs"], "requested_steps": args.steps},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
def make_env(info, args, n, dev, rank):
    if info["kind"] == "sa":
        from gym_highway_b200 import HighwayVecEnv
        return HighwayVecEnv(n, continuous=info["continuous"], seed=0, device=dev, env_id0=rank * n), 1
    from gym_highway_b200 import HighwayMAVecEnv
    return HighwayMAVecEnv(n, num_agents=args.agents, continuous=info["continuous"], seed=0, device=dev, env_id0=rank * n), args.agents


def make_actions(info, n, A, dev, rank, ring=8):
    """synthetic actions, i.i.d. uniform, generated on device: a ring of `ring` per-step buffers"""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    if info["continuous"]:
        shape = (ring, n, 2) if info["kind"] == "sa" else (ring, n, A, 2)
        return torch.rand(shape, generator=g, device=dev) * 2 - 1
    shape = (ring, n) if info["kind"] == "sa" else (ring, n, A)
    return torch.randint(0, 4, shape, generator=g, device=dev, dtype=torch.int32)


def timed_steps(env, acts, steps, warmup, flush_buf, world, dev, sampler=None):
    """W untimed + K timed env.step launches; CUDA events around every launch on the launching stream;
    returns per-launch device times [ms] and the wall time of the timed region"""
    import torch
    import torch.distributed as dist
    ring = acts.shape[0]
    if hasattr(env, "rollout_random"):
        env.rollout_random(150, action_ctr0=0)  # decorrelate the batch: episodes at different phases
    else:
        for w in range(40):                      # no in-kernel random roll-out for the multi-agent env: ordinary steps do the same
            env.step_async(acts[w % ring]); env.step_wait()
    for w in range(max(warmup, 3)):
        env.step_async(acts[w % ring]); env.step_wait()
    torch.cuda.synchronize(dev)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler is not None:
        sampler.start()
    t0 = time.perf_counter()
    for k in range(steps):
        if flush_buf is not None:
            flush_buf.fill_(k & 0xFF)
        starts[k].record()
        env.step_async(acts[k % ring])
        ends[k].record()
        env.step_wait()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    return np.array([s.elapsed_time(e) for s, e in zip(starts, ends)]), wall


def ncu_traffic(kind, continuous, n):
    """DRAM bytes per launch of the step kernel from the committed ncu --set full capture, if one exists for this size"""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return tab.get("%s_%s_%d" % (kind, "continuous" if continuous else "discrete", n))
    except Exception:
        return None


def side_measure(kind, continuous, n, dev, rank, args, steps=100, warmup=5, agents=5):
    """one extra workload timed like the headline (CUDA events per launch, L2 flushed when the working set is small)"""
    import torch
    info = dict(kind=kind, continuous=continuous)
    a = argparse.Namespace(agents=agents)
    env, A = make_env(info, a, n, dev, rank)
    acts = make_actions(info, n, A, dev, rank)
    bps = SA_BYTES_PER_STEP[continuous] if kind == "sa" else ma_bytes_per_step(A, 16)
    flush = n * bps < 2 * L2_BYTES
    fb = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev) if flush else None
    ms, wall = timed_steps(env, acts, steps, warmup, fb, 1, dev)  # decorrelates the batch first
    peak = measured_peak()[0]
    t = float(ms.mean()) * 1e-3
    out = {"envs": n, "avg_launch_us": t * 1e6, "min_launch_us": float(ms.min()) * 1e3, "env_steps_per_s": n / t,
           "wall_env_steps_per_s": n * steps / wall, "l2": "flushed between launches" if flush else "working set > L2",
           "algorithmic_bytes_per_env_step": bps, "achieved_GBps": n * bps / t / 1e9, "roofline_frac": n * bps / t / 1e9 / peak}
    if kind == "ma":
        live = float(((env.head[0].view(torch.int32) >> 20) & 31).float().mean().item())
        b_live = 2 * (36 * A + 16 * live + 40) + 4 * A + 4 * A * (4 * A + 14) + 4 * A + A
        st = env.episode_statistics()
        out.update({"agents": A, "agent_steps_per_s": A * n / t, "mean_live_obstacles": live,
                    "live_slot_bytes_per_env_step": b_live, "roofline_frac_live_slots": n * b_live / t / 1e9 / peak,
                    "dropped_spawns": st.get("dropped_spawns"), "mean_episode_length": st.get("mean_length"),
                    "kernel": "ma_tpe_step_kernel (thread per environment, cars in registers)" if n >= 16384
                              else "ma_grp_step_kernel (lane per car)"})
    del env, acts, fb
    torch.cuda.empty_cache()
    return out


def stream_measure(kind, continuous, n, dev, rank, agents=5, rounds=4, replays=5):
    """A small batch the way a rollout issues it: launches back to back (one CUDA graph, no host in between), so launch latency
    overlaps the previous kernel.  L2 rule: the graph round-robins over enough independent batches of `n` envs that the combined
    working set is over twice the L2 - every launch finds its batch evicted ("inputs larger than L2", no flush kernel in the timed
    region).  Time = CUDA events around the graph replays / number of launches."""
    import torch
    info = dict(kind=kind, continuous=continuous)
    a = argparse.Namespace(agents=agents)
    bps = SA_BYTES_PER_STEP[continuous] if kind == "sa" else ma_bytes_per_step(agents, 16)
    batches = max(2, int(np.ceil(2.2 * L2_BYTES / (n * bps))))
    envs, A = [], 1
    for b in range(batches):
        if kind == "sa":
            from gym_highway_b200 import HighwayVecEnv
            e = HighwayVecEnv(n, continuous=continuous, seed=0, device=dev, env_id0=(rank * batches + b) * n)
        else:
            from gym_highway_b200 import HighwayMAVecEnv
            e = HighwayMAVecEnv(n, num_agents=agents, continuous=continuous, seed=0, device=dev, env_id0=(rank * batches + b) * n)
            A = agents
        envs.append(e)
    acts = make_actions(info, n, A, dev, rank, ring=rounds)
    for e in envs:  # decorrelate the batches: episodes at different phases
        if hasattr(e, "rollout_random"):
            e.rollout_random(150, action_ctr0=0)
        for w in range(40 if kind == "ma" else 3):
            e.step_async(acts[w % rounds]); e.step_wait()
    torch.cuda.synchronize(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g):
            for r in range(rounds):
                for e in envs:
                    e.step_async(acts[r]); e.step_wait()
        g.replay()  # warm-up replay
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(side)
        for _ in range(replays):
            g.replay()
        t1.record(side)
    torch.cuda.synchronize(dev)
    launches = replays * rounds * batches
    t = t0.elapsed_time(t1) * 1e-3 / launches
    out = {"envs": n, "batches": batches, "launches_timed": launches, "us_per_launch_back_to_back": t * 1e6, "env_steps_per_s": n / t,
           "achieved_GBps": n * bps / t / 1e9, "roofline_frac": n * bps / t / 1e9 / measured_peak()[0],
           "l2": "%d independent batches of %d envs stepped round-robin from one CUDA graph: %.0f MB combined working set vs %d MB L2, no flush "
                 "kernel" % (batches, n, batches * n * bps / 1e6, L2_BYTES // 1000000)}
    del envs, acts, g
    torch.cuda.empty_cache()
    return out


def rollout_measure(dev, rank, world, envs=131072, nsteps=240, rollouts=3):
    """BASELINE.json configs[3]: ppo2 rollout collection (policy forward + sampling + env step + GAE), DeviceRunner under one
    CUDA graph per rollout; every rank collects from its own `envs` environments, statistics all-reduced once per rollout"""
    import torch
    import torch.distributed as dist
    from gym_highway_b200 import HighwayVecEnv
    from gym_highway_b200.rollout import DeviceRunner, MlpPolicy
    torch.manual_seed(0)
    env = HighwayVecEnv(envs, seed=0, device=dev, env_id0=rank * envs)
    pol = MlpPolicy(18, 4, bf16_inference=True).to(dev)
    run = DeviceRunner(env, pol, nsteps=nsteps, gamma=0.99, lam=0.95, seed=0)
    run.run_graphed()  # eager warm-up rollout, capture, first replay
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(rollouts):
        run.run_graphed()
        stats = env.episode_statistics(reset=True)  # one all_reduce of 8 counters per rollout
    t1.record()
    torch.cuda.synchronize(dev)
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out = {"workload": "configs[3]: ppo2 rollout collection, %d envs/GPU x %d GPU(s), T=%d, MLP 18-256-256-256 (bf16 GEMMs) on device, "
                       "fused sampling kernel, env step_into, hw_gae, one CUDA graph per rollout" % (envs, world, nsteps),
           "rollout_env_steps_per_s": world * envs * nsteps * rollouts / (ms * 1e-3), "ms_per_rollout": ms / rollouts,
           "us_per_step": 1e3 * ms / rollouts / nsteps, "episodes_last_rollout": stats["episodes"]}
    del run, env, pol
    torch.cuda.empty_cache()
    return out


def consumers_measure(dev, rank):
    """SURVEY.md section 8(f) consumers on the device, one GPU: a2c's 5-step rollouts (DeviceA2CRunner, one CUDA graph per
    rollout), ppo2-style collection over the multi-agent env (DeviceRunner, every car a row of the shared policy), and the
    MADDPG rollout loop's transition store (MADeviceCollector: env step + hw_ring_store + hw_obs_moments)"""
    import torch
    from gym_highway_b200 import HighwayMAVecEnv, HighwayVecEnv
    from gym_highway_b200.replay import MADeviceCollector, RingMemory, RunningMeanStd
    from gym_highway_b200.rollout import DeviceA2CRunner, DeviceRunner, MlpPolicy
    torch.manual_seed(0)
    out = {}

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize(dev)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            fn()
        t1.record()
        torch.cuda.synchronize(dev)
        return t0.elapsed_time(t1) * 1e-3 / reps

    n, T = 131072, 5
    run = DeviceA2CRunner(HighwayVecEnv(n, seed=0, device=dev, env_id0=rank * n), MlpPolicy(18, 4, bf16_inference=True).to(dev), nsteps=T, seed=0)
    t = timed(run.run_graphed, 100)
    out["a2c_rollout"] = {"workload": "a2c runner: %d single-agent envs, nsteps=%d, MLP 18-256-256-256 (bf16), fused sampling, step_into, "
                                      "hw_nstep_returns, one CUDA graph per rollout" % (n, T),
                          "rollout_env_steps_per_s": n * T / t, "us_per_step": 1e6 * t / T}
    del run
    n, A, T = 32768, 5, 64
    run = DeviceRunner(HighwayMAVecEnv(n, num_agents=A, seed=0, device=dev, env_id0=rank * n), MlpPolicy(4 * A + 14, 4, bf16_inference=True).to(dev),
                       nsteps=T, seed=0)
    t = timed(run.run_graphed, 5)
    out["ma_ppo2_rollout"] = {"workload": "ppo2-style collection over the multi-agent env: %d envs x %d cars (every car a row of the shared "
                                          "MLP 34-256-256-256), T=%d, one CUDA graph per rollout" % (n, A, T),
                              "rollout_env_steps_per_s": n * T / t, "agent_steps_per_s": n * A * T / t, "us_per_step": 1e6 * t / T}
    del run
    n, A, T = 65536, 5, 20
    env = HighwayMAVecEnv(n, num_agents=A, continuous=True, seed=0, device=dev, env_id0=rank * n)
    mem = RingMemory(n * 12, A, env.obs_dim, 2, device=dev)
    rms = RunningMeanStd(shape=(A, env.obs_dim), device=dev)
    actor = torch.nn.Sequential(torch.nn.Linear(env.obs_dim, 64), torch.nn.ReLU(), torch.nn.Linear(64, 2), torch.nn.Tanh()).to(dev).to(torch.bfloat16)
    col = MADeviceCollector(env, lambda obs: actor(obs.reshape(-1, env.obs_dim).to(torch.bfloat16)), mem, obs_rms=rms)
    t_eager = timed(lambda: col.collect(T), 3)
    t = timed(lambda: col.collect_graphed(T), 5)
    out["maddpg_collection"] = {"workload": "MADDPG rollout loop: %d multi-agent envs x %d cars (continuous), torch actor 34-64-2 (bf16), env step_into, "
                                            "hw_ring_store into a %d-row ring, hw_obs_moments; the %d-step loop replayed from one CUDA graph" % (n, A, mem.limit, T),
                                "transitions_per_s": n * T / t, "agent_transitions_per_s": n * A * T / t, "us_per_step": 1e6 * t / T, "us_per_step_eager": 1e6 * t_eager / T,
                                "ring_wrapped": mem.nb_entries == mem.limit}
    del col, mem, env, rms
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    if args.envs <= 0:
        # SA: a point of the scaling sweep (BASELINE.json configs[4]) whose working set (1.1 GB/step) is far larger than
        # L2, so launches are HBM-fed without flushing; configs[1] (65,536 envs) is reported alongside under "configs1".
        args.envs = 4194304 if args.workload.startswith("sa_") else 262144
    info = workload_info(args)
    if args.impl == "reference":
        return reference_main(args, info)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from gym_highway_b200 import numa
    numa_note = numa.bind_to_gpu(local_rank)  # before the first pinned allocation: host buffers land on the GPU's NUMA node
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.envs
    env, A = make_env(info, args, n, dev, rank)
    acts = make_actions(info, n, A, dev, rank)
    working_set = n * info["bytes_per_step"]
    flush = (not args.no_flush) and working_set < 2 * L2_BYTES
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev) if flush else None

    sampler = ClockSampler(local_rank)
    dev_ms, t_wall = timed_steps(env, acts, args.steps, args.warmup, flush_buf, world, dev, sampler)
    clocks = sampler.result()
    total_ms = float(dev_ms.sum())
    if world > 1:
        t = torch.tensor([total_ms, t_wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, t_wall = float(t[0].item()), float(t[1].item())
    value = world * n * args.steps / (total_ms * 1e-3)

    # end to end through the host-buffer C-ABI call: pinned host actions in, host obs/rew/done out
    h_acts = acts.cpu().numpy()
    ring = h_acts.shape[0]
    e2e_steps = max(5, min(args.steps, 50))
    for w in range(3):
        env.step_host(h_acts[w % ring])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        env.step_host(h_acts[k % ring])
    torch.cuda.synchronize(dev)
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    act_bytes = int(h_acts[0].nbytes)
    out_bytes = int(getattr(env, "host_d2h_bytes_per_step", n * A * (4 * A + 14 if info["kind"] == "ma" else 18) * 4 + n * A * 4 + n * A))
    e2e = {"value": world * n * e2e_steps / t_e2e, "unit": "env-steps/s", "h2d_bytes_per_step": act_bytes,
           "d2h_bytes_per_step": out_bytes, "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps,
           "api": ("HighwayVecEnv.step_host -> hw_sa_step_host" if info["kind"] == "sa" else "HighwayMAVecEnv.step_host -> hw_ma_step_host")
                  + " (C ABI, host buffers: pinned-host actions H2D, kernel, obs/rew/done D2H, stream drain, all inside the call)",
           "host_binding": numa_note}
    if info["kind"] == "sa":
        # the same call with HW_FLAG_COMPACT_OBS: observation rows of 14 floats (the four constant columns stay on the device side of PCIe)
        for w in range(2):
            env.step_host(h_acts[w % ring], compact=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            env.step_host(h_acts[k % ring], compact=True)
        torch.cuda.synchronize(dev)
        t_c = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([t_c], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_c = float(t.item())
        e2e["compact_obs"] = {"value": world * n * e2e_steps / t_c, "unit": "env-steps/s", "ms_per_step": 1e3 * t_c / e2e_steps,
                              "d2h_bytes_per_step": n * (14 * 4 + 4 + 1),
                              "note": "step_host(compact=True): [N, 14] observation rows, columns 11/13/15/17 (constant 0.5) dropped; "
                                      "HighwayVecEnv.expand_obs restores the reference layout on the host; not the headline e2e"}
    stats = env.episode_statistics()
    live = None
    if info["kind"] == "ma":
        live = float(((env.head[0].view(torch.int32) >> 20) & 31).float().mean().item())
    del env, acts
    torch.cuda.empty_cache()

    # side measurements of the other BASELINE.json configs (one GPU; the driver's N=1 run records them)
    extras = {}
    if not args.no_extras and world == 1 and info["kind"] == "sa" and not info["continuous"] and n != 65536:
        # configs[1]: 65,536 single-agent envs on one GPU (0.43 waves: latency-bound), L2 flushed between launches
        c1 = side_measure("sa", False, 65536, dev, rank, args, steps=200, warmup=10)
        extras["configs1"] = dict(c1, workload="configs[1]: 65,536 single-agent envs on 1 GPU, discrete random actions",
                                  back_to_back=stream_measure("sa", False, 65536, dev, rank))
        extras["sa_16m"] = dict(side_measure("sa", False, 16777216, dev, rank, args, steps=30, warmup=3),
                                workload="configs[4]: the 16,777,216-env point (largest single-GPU configuration), discrete")
        extras["continuous"] = dict(side_measure("sa", True, 4194304, dev, rank, args, steps=60, warmup=5),
                                    workload="single-agent continuous actions (HighwayEnvContinuous), 4,194,304 envs")
        extras["configs2"] = {
            "workload": "configs[2]: multi-agent env, 5 cars (+3..16 obstacles) per env, uniform random discrete actions, car-car collisions",
            "envs_16384": dict(side_measure("ma", False, 16384, dev, rank, args, steps=200, warmup=10),
                               back_to_back=stream_measure("ma", False, 16384, dev, rank)),
            "envs_262144": side_measure("ma", False, 262144, dev, rank, args, steps=60, warmup=5),
            "continuous_262144": side_measure("ma", True, 262144, dev, rank, args, steps=60, warmup=5)}
    if not args.no_extras and info["kind"] == "sa" and not info["continuous"]:
        extras["configs3"] = rollout_measure(dev, rank, world)
    if not args.no_extras and world == 1 and info["kind"] == "sa" and not info["continuous"] and n != 65536:
        extras["consumers"] = consumers_measure(dev, rank)

    if rank == 0:
        peak, peak_src = measured_peak()
        avg_launch_s = float(dev_ms.mean()) * 1e-3
        bps = info["bytes_per_step"]
        if live is not None:  # only the live obstacle slots are read and written
            bps = 2 * (36 * A + 16 * live + 40) + 4 * A + 4 * A * (4 * A + 14) + 4 * A + A
        achieved = n * bps / avg_launch_s / 1e9
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": info["dtype"], "data": "synthetic",
            "config": {"workload": info["name"], "envs_per_gpu": n, "ticks_per_step": 9,
                       "steering": "fast (vx*tan/4, one rounding: discrete state bit-exact, floats within 1e-5 relative per step; "
                                   "HW_FLAG_EXACT_STEERING, the bit-exact mode of the golden tests, costs ~20 %)",
                       "baseline_config": ("BASELINE.json configs[4] (scaling sweep 4k-16M envs, HBM GB/s vs roofline): the %d-env point, weak-scaled over the GPUs; "
                                           "configs[1] (65,536 envs, launch/latency-bound), the 16M point, continuous actions, configs[2] (multi-agent) and "
                                           "configs[3] (ppo2 rollout) are measured in the same run and reported under their own keys" % n)
                                          if info["kind"] == "sa" else "BASELINE.json configs[2]/[4]: multi-agent env, %d-env point" % n,
                       "l2": "flushed between timed launches (256 MiB overwrite outside the event pairs)" if flush
                             else "inputs larger than L2 (%.0f MB per launch vs 126 MB), no flush" % (working_set / 1e6),
                       "parallelism": "independent env shards, %d rank(s), no data-path collective" % world},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(info["kind"], info["continuous"], n), "peak_source": peak_src,
                         "kernel": "sa_step_kernel" if info["kind"] == "sa" else ("ma_tpe_step_kernel" if n >= 16384 else "ma_grp_step_kernel"),
                         "algorithmic_bytes_per_env_step": bps,
                         "avg_launch_us": avg_launch_s * 1e6, "min_launch_us": float(dev_ms.min()) * 1e3},
            "e2e": e2e, "gpu_launches": args.steps, "clocks": clocks,
            "wall_ms_per_step": 1e3 * t_wall / args.steps, "value_wall": world * n * args.steps / t_wall, "episodes": stats,
        }
        line.update(extras)
        if info["kind"] == "ma":
            line["agent_steps_per_s"] = value * A
            line["mean_live_obstacles"] = live
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_arm(args, info).items() if k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
