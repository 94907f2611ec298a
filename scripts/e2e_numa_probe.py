"""Where does the multi-GPU host-buffer path lose its scaling?  (run under torchrun, one rank per GPU)

Prints the box's topology (rank 0: `nvidia-smi topo -m`, NUMA nodes, this cpuset) and, for every rank at the same time,
the device->host bandwidth into pinned memory (the 323 MB observation copy of the 4M-env step) with the default placement
and after `numa.bind_to_gpu`.  One JSON line per rank and mode, then an aggregate line from rank 0."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gym_highway_b200 import numa


def d2h_gbs(nbytes, reps=8):
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.copy_(dev, non_blocking=True); torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return nbytes * reps / dt / 1e9


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        for cmd in (["nvidia-smi", "topo", "-m"], ["sh", "-c", "ls /sys/devices/system/node/ | grep node; cat /sys/devices/system/node/node*/cpulist; nproc; cat /proc/self/status | grep -i allowed"]):
            try:
                print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout)
            except Exception as e:
                print("failed:", cmd, e)
    res = {}
    for mode in ("default", "bound"):
        note = numa.bind_to_gpu(local) if mode == "bound" else "default placement, affinity %d cpus" % len(os.sched_getaffinity(0))
        g = d2h_gbs(323 << 20)
        t = torch.tensor([g], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        res[mode] = dict(rank_gbs=g, aggregate_gbs=float(t), note=note, node=numa.gpu_numa_node(local))
        print(json.dumps(dict(rank=rank, mode=mode, **res[mode])), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
