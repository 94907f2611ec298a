"""Per-phase clock64() accounting of the thread-per-environment multi-agent step (needs a library built with
HW_NVCC_EXTRA_HW_MA_TPE="-DHW_TPE_PROF"): where a warp's cycles go."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_highway_b200 import HighwayMAVecEnv, _lib

NAMES = ["-", "before tick (prologue / barrier)", "actions (+maintain search)", "car update: velocity, steering", "leader velocity + tan / angular velocity",
         "position pass", "leader election + rect packing", "obstacle loop", "spawn/retire + car-car filter + exact collide",
         "idle ticks / barriers after the last tick", "rewards + observe obstacles + state store", "observation rows: stage + stream out"]

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    A = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    lib = _lib.load()
    env = HighwayMAVecEnv(n, num_agents=A, seed=0, kernel="thread")
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    acts = torch.randint(0, 4, (8, n, A), generator=g, device='cuda', dtype=torch.int32)
    for w in range(30):
        env.step_async(acts[w % 8]); env.step_wait()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    f = lib.hw_debug_tpe_prof
    f.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    assert f(buf, 1) == 0
    steps = 10
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(steps):
        env.step_async(acts[k % 8]); env.step_wait()
    ev1.record(); torch.cuda.synchronize()
    assert f(buf, 0) == 0
    warps = buf[12]
    tot = sum(buf[1:12])
    print("n=%d A=%d: %.1f us per step (instrumented), %d warp-steps, %.0f cycles per warp-step" % (n, A, ev0.elapsed_time(ev1) * 1e3 / steps, warps, tot / warps))
    for j in range(1, 12):
        print("  %5.1f %%  %8.0f cycles  %s" % (100.0 * buf[j] / tot, buf[j] / warps, NAMES[j]))

if __name__ == "__main__":
    main()
