#!/usr/bin/env python
"""executed warp-instructions per CUDA source line of the first kernel in an .ncu-rep (needs -lineinfo and --import-source on)
usage: ncu_lines.py report.ncu-rep [min_percent]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; iN = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr and r[0].isdigit():
        try:
            out.append((fname, int(r[0]), r[1].strip(), int(r[iN]), int(r[iT]), int(r[iS] or 0)))
        except ValueError:
            pass
tot = sum(o[3] for o in out) or 1
tots = sum(o[5] for o in out) or 1
print("total executed warp-instructions %d, samples %d" % (tot, tots))
for f, ln, text, n, t, s in sorted(out, key=lambda o: -o[3]):
    if 100.0 * n / tot < minpct:
        break
    print("%5.2f%% instr %5.2f%% samples  thr %4.1f  %s:%d  %s" % (100.0 * n / tot, 100.0 * s / tots, t / max(n, 1), f, ln, text[:110]))
