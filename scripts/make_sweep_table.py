"""profiles/<tag>_sweep.md from gpurun_out/sweep_{1,2,4,8}gpu.jsonl (the outputs of scripts/sweep.py): python scripts/make_sweep_table.py r1t"""
import collections, json, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "sweep"
tab = collections.defaultdict(dict)
for g in (1, 2, 4, 8):
    p = os.path.join(root, "gpurun_out", "sweep_%dgpu.jsonl" % g)
    if not os.path.exists(p):
        continue
    for l in open(p):
        if l.startswith('{'):
            d = json.loads(l)
            tab[(d["workload"], d["envs_total"])][g] = d
out = ["# Scaling sweep (BASELINE.json configs[4]) - `scripts/sweep.py`", "",
       "N = environments of the whole job, sharded contiguously over the GPUs (no data-path collective). Per cell: job env-steps/s and, in",
       "brackets, algorithmic HBM GB/s per GPU as a fraction of the measured 6 553.9 GB/s. 30 timed launches per point, CUDA events on the",
       "launching stream, max over ranks; the L2 is flushed between launches when a shard's working set is below 2 x L2. Small shards are",
       "launch/latency-bound (a 4 096-env launch is 13 us), which is why the per-GPU fraction only approaches the 1-GPU figure at >= 1 M envs per GPU.", ""]
for wl in ("sa_discrete", "sa_continuous", "ma_discrete"):
    if not any(k[0] == wl for k in tab):
        continue
    out += ["## %s" % wl, "", "| N | 1 GPU | 2 GPUs | 4 GPUs | 8 GPUs |", "|---|---|---|---|---|"]
    for N in (4096, 16384, 65536, 262144, 1048576, 4194304, 16777216):
        row = ["%d" % N]
        for g in (1, 2, 4, 8):
            d = tab.get((wl, N), {}).get(g)
            row.append("%.2e (%.1f %%)" % (d["env_steps_per_s"], 100 * d["frac_of_measured_hbm_peak"]) if d else "-")
        out.append("| " + " | ".join(row) + " |")
    out.append("")
open(os.path.join(root, "profiles", "%s_sweep.md" % tag), "w").write("\n".join(out))
print("\n".join(out))
