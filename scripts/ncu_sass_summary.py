#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): executed warp instructions by opcode,
and the hottest contiguous SASS regions.  usage: ncu_sass_summary.py src_sass.csv [kernel_index]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# split per kernel: a row starting with "Kernel Name" begins a block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None:
        if cur["hdr"] is None:
            cur["hdr"] = r
        else:
            cur["rows"].append(r)
b = blocks[which]
h = b["hdr"]
iS, iN, iT = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
iSamp = h.index("# Samples")
tot = sum(int(r[iN]) for r in b["rows"])
print(b["name"][:90]); print("SASS instructions: %d, executed warp-instr: %d" % (len(b["rows"]), tot))
ops = collections.Counter(); thr = collections.Counter(); samp = collections.Counter()
for r in b["rows"]:
    s = r[iS].strip()
    if s.startswith("@"):
        s = s.split(None, 1)[1]
    op = s.split()[0].split(".")[0]
    ops[op] += int(r[iN]); thr[op] += int(r[iT]); samp[op] += int(r[iSamp])
tsamp = sum(samp.values())
print("%-10s %12s %7s %9s %8s" % ("opcode", "warp-instr", "share", "thr/instr", "samples%"))
for op, c in ops.most_common(28):
    print("%-10s %12d %6.1f%% %9.1f %7.1f%%" % (op, c, 100.0 * c / tot, thr[op] / max(c, 1), 100.0 * samp[op] / max(tsamp, 1)))
