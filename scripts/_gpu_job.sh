python -m pytest tests -m gpu -x -q > gpurun_out/r3d_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r3d_tests.log
tail -6 gpurun_out/r3d_tests.log
python - <<'PY'
import sys
sys.path.insert(0, 'scripts'); sys.path.insert(0, '.')
import torch
from gym_highway_b200 import HighwayMAVecEnv
import perf_ma
def probe(n, A):
    env = HighwayMAVecEnv(n, num_agents=A, seed=0, extended_agents=True)
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    acts = torch.randint(0, 4, (8, n, A), generator=g, device='cuda', dtype=torch.int32)
    for w in range(30): env.step_async(acts[w % 8]); env.step_wait()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
    for k in range(30):
        ev[k][0].record(); env.step_async(acts[k % 8]); ev[k][1].record(); env.step_wait()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev); med = ms[15]
    print("EXT MA n=%d A=%d median %.1f us %.3e env-steps/s %.3e agent-steps/s" % (n, A, med * 1e3, n / (med * 1e-3), n * A / (med * 1e-3)))
probe(262144, 8); probe(262144, 6); probe(16384, 8)
PY
