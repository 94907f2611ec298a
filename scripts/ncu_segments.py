#!/usr/bin/env python
"""per-basic-block execution profile of the first kernel in an .ncu-rep (source page): executions per warp,
share of executed instructions and of stall samples.  usage: ncu_segments.py report.ncu-rep n_warps [min_share]"""
import csv, io, re, subprocess, sys
rep, nw = sys.argv[1], int(sys.argv[2])
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for i, r in enumerate(rows):
    if 'Source' in r and 'Instructions Executed' in r:
        h = r; start = i + 1; break
iS, iN, iT, iSm = h.index('Source'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
out = []
for r in rows[start:]:
    try: out.append((r[iS].strip(), int(r[iN]), int(r[iT]), int(r[iSm])))
    except (ValueError, IndexError): pass
tot = sum(o[1] for o in out); ts = sum(o[3] for o in out)
print("executed warp-instructions %d (%.0f per warp), samples %d" % (tot, tot / nw, ts))
seg, cur = [], []
for k, o in enumerate(out):
    cur.append((k, o))
    if re.search(r"\b(BRA|CALL|BSYNC|BSSY|BREAK|EXIT|RET)\b", o[0]):
        seg.append(cur); cur = []
if cur: seg.append(cur)
for sg in seg:
    n = sum(o[1] for k, o in sg); smp = sum(o[3] for k, o in sg)
    if 100.0 * n / tot < min_share and 100.0 * smp / ts < min_share: continue
    print("%4d..%4d size %3d exec/warp %6.2f  instr %5.2f%%  samples %5.2f%%  thr %4.1f  end: %s" % (
        sg[0][0], sg[-1][0], len(sg), sg[0][1][1] / nw, 100.0 * n / tot, 100.0 * smp / ts,
        sum(o[2] for k, o in sg) / max(n, 1), sg[-1][1][0][:44]))
