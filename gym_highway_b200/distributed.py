"""Multi-GPU plumbing: one process per GPU, environments sharded contiguously, no data-path collective.

The environments are independent, so the step itself never communicates.  Collectives exist only for the
consumers of the path: episode statistics (replaces Monitor + the trainers' deque means and the MPI
allreduce at maddpg/trainer/ma_ddpg.py:354 / baselines/ddpg/ddpg.py:273) and optional rollout gathers for a
central learner.  Everything here works on CPU tensors with the gloo backend and on CUDA tensors with NCCL.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(total_envs, rank=None, world_size=None):
    """contiguous partition of `total_envs` environments: returns (env_id0, n_local) of this rank.
    Ranks differ by at most one environment; env ids are global, so a run is reproducible for any world size
    (the Philox stream of env i is keyed by its global id)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(total_envs), int(world_size))
    n_local = base + (1 if rank < rem else 0)
    env_id0 = rank * base + min(rank, rem)
    return env_id0, n_local


def reduce_statistics(stats, group=None):
    """sum the kernels' episode counters over all ranks (one small all_reduce per rollout, not per step)"""
    s = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return s


def statistics_dict(stats):
    v = [int(x) for x in stats.cpu().tolist()]
    eps = v[0]
    return dict(episodes=eps, mean_return=(v[2] / 1000.0 / eps) if eps else float('nan'),
                mean_length=(v[1] / eps) if eps else float('nan'), obs_collisions=v[3], timeouts=v[4])


def gather_env_axis(local, axis=1, group=None):
    """all_gather a rollout block along its environment axis, ranks in order: [T, N_local, ...] -> [T, N_total, ...].
    Needs equal N_local on every rank (pad the batch or use shard() with a divisible total)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    ws = dist.get_world_size(group)
    parts = [torch.empty_like(local) for _ in range(ws)]
    dist.all_gather(parts, local.contiguous(), group=group)
    return torch.cat(parts, dim=axis)


def max_over_ranks(value, device=None, group=None):
    """device-timed durations are reported as the max over ranks"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
