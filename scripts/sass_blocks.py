#!/usr/bin/env python
"""control-flow skeleton of one loop of a kernel: straight-line segments with their sizes.
usage: scripts/sass_blocks.py lib.so pattern lo hi   (addresses in hex)"""
import re, subprocess, sys
so, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn = None; ins = []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); continue
    if fn and pat in fn:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and lo <= int(m.group(1), 16) <= hi:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
seg = 0; start = None
for a, t in ins:
    if start is None: start = a
    seg += 1
    if re.search(r"\b(BRA|CALL|BSYNC|BSSY|BREAK|EXIT)\b", t):
        print("%04x..%04x %3d  %s" % (start, a, seg, t))
        seg = 0; start = None
if seg: print("%04x..     %3d" % (start, seg))
