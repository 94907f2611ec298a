#!/usr/bin/env python
"""text summary of an .ncu-rep: key raw metrics per captured launch + SASS opcode mix + stall reasons.
usage: ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active', 'sm__cycles_elapsed.max',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum']
print("report:", rep)
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-70s %s" % (w, " | ".join(r[i][:60] for r in rows[1:])))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None:
        if cur["hdr"] is None: cur["hdr"] = r
        else: cur["rows"].append(r)
for b in blocks[:1]:
    h = b["hdr"]; iS, iN, iT = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    tot = sum(int(r[iN]) for r in b["rows"])
    print("\nSASS: %d static instructions, %d executed warp-instructions" % (len(b["rows"]), tot))
    ops = collections.Counter(); thr = collections.Counter()
    for r in b["rows"]:
        s = r[iS].strip()
        if s.startswith("@"): s = s.split(None, 1)[1]
        op = s.split()[0].split(".")[0]
        ops[op] += int(r[iN]); thr[op] += int(r[iT])
    for op, c in ops.most_common(22):
        print("  %-10s %12d %5.1f%%  avg active threads %.1f" % (op, c, 100.0 * c / tot, thr[op] / max(c, 1)))
    st = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    acc = {h[i]: 0 for i in st}
    for r in b["rows"]:
        for i in st:
            try: acc[h[i]] += int(r[i])
            except ValueError: pass
    s = sum(acc.values()) or 1
    print("warp stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / s) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:9]))
