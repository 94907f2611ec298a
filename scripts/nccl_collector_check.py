"""Multi-rank check of the MADDPG collection path over NCCL (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/nccl_collector_check.py

Every rank steps its shard of the multi-agent envs (global env ids rank * n ...), all_gathers the transition blocks
(`MADeviceCollector(central=True)`: one NCCL all_gather per step) into its ring and all_reduces the observation moments.
Rank 0 then repeats the run in ONE process over all world * n envs and requires the central ring (every rank's), the
episode records and obs_rms to be identical - the sharded collection is the single-process one.  Prints one JSON line
with the timing of the sharded run (transitions/s = env-steps/s through env + store + collectives)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from gym_highway_b200 import HighwayMAVecEnv
from gym_highway_b200.replay import MADeviceCollector, RingMemory, RunningMeanStd


def policy_for(env_id0, A):
    """deterministic continuous actions from (global env id, agent, step): the same for any sharding"""
    def f(obs):
        n = obs.shape[0]
        e = torch.arange(env_id0, env_id0 + n, device=obs.device, dtype=torch.int64).view(n, 1, 1)
        a = torch.arange(A, device=obs.device, dtype=torch.int64).view(1, A, 1)
        k = torch.arange(2, device=obs.device, dtype=torch.int64).view(1, 1, 2)
        h = (e * 1000003 + a * 7919 + k * 104729 + f.t * 15485863) % 2048
        f.t += 1
        return h.to(torch.float32) / 1024.0 - 1.0
    f.t = 0
    return f


def run(n, A, T, limit, env_id0, central, seed=5):
    env = HighwayMAVecEnv(n, num_agents=A, continuous=True, seed=seed, env_id0=env_id0)
    world = dist.get_world_size() if central else 1
    mem = RingMemory(limit, A, env.obs_dim, 2, device=env.device)
    rms = RunningMeanStd(shape=(A, env.obs_dim), device=env.device)
    col = MADeviceCollector(env, policy_for(env_id0, A), mem, obs_rms=rms if central else None, central=central)
    if not central:
        col.obs_rms = rms   # single process: plain local update
        col.group = None
    torch.cuda.synchronize()
    if central:
        dist.barrier()
    t0 = time.perf_counter()
    col.collect(T)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return mem, rms, col, dt, world


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl")
    n, A, T = int(os.environ.get("HW_N", 4096)), 5, int(os.environ.get("HW_T", 40))
    limit = n * world * (T - 3)          # the ring wraps
    mem, rms, col, dt, _ = run(n, A, T, limit, rank * n, central=True)
    ok = True
    # every rank holds the same central ring
    digest = torch.stack([mem.obs0.double().sum(), mem.obs1.double().sum(), mem.actions.double().sum(), mem.rewards.double().sum(),
                          mem.terminals1.double().sum(), rms._sum.sum(), rms._sumsq.sum(), rms._count])
    all_d = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    ok &= all(torch.equal(all_d[0], d) for d in all_d)
    if rank == 0:
        # rank 0 recomputes in one process; the other ranks wait at the final barrier
        dist_world = world
        env = HighwayMAVecEnv(n * world, num_agents=A, continuous=True, seed=5, env_id0=0)
        mem1 = RingMemory(limit, A, env.obs_dim, 2, device=env.device)
        rms1 = RunningMeanStd(shape=(A, env.obs_dim), device=env.device)
        col1 = MADeviceCollector(env, policy_for(0, A), mem1, obs_rms=None, central=False)
        for _ in range(T):
            obs0 = col1.obs[col1.cur]
            # local moments without the collective (the sharded run all_reduced the same rows)
            x = obs0.reshape(-1, A, env.obs_dim).double()
            rms1._sum += x.sum(0) * A; rms1._sumsq += (x * x).sum(0) * A; rms1._count += float(x.shape[0]) * A
            col1.step()
        for k in ("obs0", "obs1", "actions", "rewards", "terminals1"):
            ok &= bool(torch.equal(getattr(mem, k), getattr(mem1, k)))
        ok &= (mem.start, mem.length) == (mem1.start, mem1.length) and mem.state.tolist() == mem1.state.tolist()
        ok &= bool(torch.equal(col.fin_step, col1.fin_step)) and bool(torch.equal(col.ep_reward, col1.ep_reward))
        ok &= bool(torch.allclose(rms.mean, rms1.mean, rtol=0, atol=1e-6)) and bool(torch.allclose(rms.std, rms1.std, rtol=0, atol=1e-6))
        ok &= float(rms._count) == float(rms1._count)
        st = col.epoch_statistics()
        print(json.dumps({"check": "nccl_collector", "ok": bool(ok), "world": dist_world, "envs_per_rank": n, "agents": A, "steps": T,
                          "ring_limit": limit, "ring_wrapped": mem.length == limit, "episodes": st["episodes"],
                          "transitions_per_s": n * world * T / dt, "ms_per_step": 1e3 * dt / T,
                          "collectives_per_step": "1 all_gather (packed [N, %d] block) + 1 all_reduce (obs moments)" % (A * (2 * env.obs_dim + 4))}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
