"""BASELINE.json configs[3]: ppo2 rollout collection from batched single-agent envs, torch policy on device.

    python scripts/bench_rollout.py [--envs 131072] [--nsteps 240]           # one GPU
    torchrun --nproc-per-node 8 scripts/bench_rollout.py                     # 8 x 131,072 = 1,048,576 envs

Rollout = nsteps x (policy forward 18->256->256->256->(4,1) in bf16 autocast, sample, env.step kernel, buffer writes)
+ GAE on device; episode statistics are all-reduced once per rollout (NCCL), nothing else crosses GPUs.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=131072)
ap.add_argument("--nsteps", type=int, default=240)
ap.add_argument("--rollouts", type=int, default=3)
ap.add_argument("--runner", default="device", choices=["device", "torch"], help="device: fused sampling kernel + zero-copy buffers (DeviceRunner); torch: Runner")
ap.add_argument("--eager", action="store_true", help="issue every launch from the host instead of replaying one CUDA graph per rollout")
args = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from gym_highway_b200 import HighwayVecEnv
from gym_highway_b200.rollout import MlpPolicy, Runner, DeviceRunner
torch.manual_seed(0)
env = HighwayVecEnv(args.envs, seed=0, device=dev, env_id0=rank * args.envs)
pol = MlpPolicy(18, 4, bf16_inference=True).to(dev)
torch.cuda.manual_seed(rank)
if args.runner == "device":
    run = DeviceRunner(env, pol, nsteps=args.nsteps, gamma=0.99, lam=0.95, seed=rank)
    collect = run.run if args.eager else run.run_graphed
else:
    run = Runner(env, pol, nsteps=args.nsteps, gamma=0.99, lam=0.95, generator=None)
    collect = (lambda: run.run(want_epinfos=False)) if args.eager else run.run_graphed
collect()  # warm-up rollout (and graph capture)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(args.rollouts):
    out = collect()
    stats = env.episode_statistics(reset=True)  # one all_reduce of 8 counters per rollout
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
if rank == 0:
    total = world * args.envs * args.nsteps * args.rollouts
    print(json.dumps({"workload": "ppo2 rollout collection, %d envs/GPU x %d GPU(s), T=%d, MLP 18-256-256-256 policy on device" % (args.envs, world, args.nsteps),
                      "runner": args.runner, "mode": "eager" if args.eager else "cuda graph", "rollout_env_steps_per_s": total / (ms * 1e-3), "ms_per_rollout": ms / args.rollouts,
                      "rollout_buffer_GB_per_gpu": args.nsteps * args.envs * 18 * 4 / 1e9, "last_rollout_episodes": stats}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
