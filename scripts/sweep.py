"""BASELINE.json configs[4]: scaling sweep of the step kernels, 4k..16M envs per job at 1/2/4/8 GPUs.

    python scripts/sweep.py [--sizes 4096,...] [--workloads sa_discrete,ma_discrete]         # one GPU
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/sweep.py                      # 8 GPUs: N = total envs of the job

Per point: envs are sharded contiguously over the ranks (no data-path collective), 30 step launches are timed with CUDA
events on the launching stream (L2 flushed between launches when the shard's working set is below 2 x L2), the time is the
max over ranks; one JSON line per point on rank 0: env-steps/s of the whole job and algorithmic HBM GB/s per GPU.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="4096,16384,65536,262144,1048576,4194304,16777216")
ap.add_argument("--workloads", default="sa_discrete,sa_continuous,ma_discrete")
ap.add_argument("--steps", type=int, default=30)
args = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from gym_highway_b200 import HighwayVecEnv, HighwayMAVecEnv
PEAK = 6553.9
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for wl in args.workloads.split(","):
    cont, ma = wl.endswith("continuous"), wl.startswith("ma_")
    for N in [int(x) for x in args.sizes.split(",")]:
        n = N // world
        if n < 1:
            continue
        g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
        if ma:
            A = 5
            env = HighwayMAVecEnv(n, num_agents=A, continuous=cont, seed=0, device=dev, env_id0=rank * n)
            acts = (torch.rand((4, n, A, 2), generator=g, device=dev) * 2 - 1) if cont else torch.randint(0, 4, (4, n, A), generator=g, device=dev, dtype=torch.int32)
            warm = 30
        else:
            env = HighwayVecEnv(n, continuous=cont, seed=0, device=dev, env_id0=rank * n)
            acts = (torch.rand((4, n, 2), generator=g, device=dev) * 2 - 1) if cont else torch.randint(0, 4, (4, n), generator=g, device=dev, dtype=torch.int32)
            env.rollout_random(150)
            warm = 5
        for w in range(warm):
            env.step_async(acts[w % 4]); env.step_wait()
        live = float(((env.head[0].view(torch.int32) >> 20) & 31).float().mean().item()) if ma else 3.0
        bps = (2 * (36 * 5 + 16 * live + 40) + 4 * 5 + 4 * 5 * 34 + 4 * 5 + 5) if ma else (277 if cont else 273)
        do_flush = n * bps < 2 * 126 * 1024 * 1024
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        for k in range(args.steps):
            if do_flush:
                flush.fill_(k & 255)
            ev[k][0].record(); env.step_async(acts[k % 4]); ev[k][1].record(); env.step_wait()
        torch.cuda.synchronize(dev)
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        if rank == 0:
            sps = N // world * world * args.steps / (ms * 1e-3)
            print(json.dumps({"workload": wl, "envs_total": N // world * world, "gpus": world, "envs_per_gpu": n, "us_per_launch": 1e3 * ms / args.steps,
                              "env_steps_per_s": sps, "algorithmic_GBps_per_gpu": sps / world * bps / 1e9,
                              "frac_of_measured_hbm_peak": sps / world * bps / 1e9 / PEAK, "l2_flushed": do_flush}), flush=True)
        del env, acts
        torch.cuda.empty_cache()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
