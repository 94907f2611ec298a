#!/usr/bin/env python
"""Static issue schedule of a SASS address range: the stall count ptxas encoded per instruction (bits 105..108 of the 128-bit word)
and its sum = the cycles ONE warp needs to issue the range when nothing else runs and no scoreboard wait is hit.
usage: sass_stalls.py lib.so kernel_substring [lo_hex hi_hex]"""
import re, subprocess, sys, collections
so, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
fn, cur, rows = None, None, []
for line in txt:
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); continue
    if not fn or pat not in fn:
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", line)
    if m:
        cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]; rows.append(cur); continue
    m = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", line)
    if m and cur is not None and cur[3] is None:
        cur[3] = int(m.group(1), 16)
tot = 0; n = 0; byop = collections.Counter(); waits = 0
for a, t, w0, w1 in rows:
    if not (lo <= a <= hi) or w1 is None:
        continue
    stall = (w1 >> 41) & 0xF
    wmask = (w1 >> 52) & 0x3F
    tot += max(stall, 1); n += 1
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
    byop[op] += max(stall, 1)
    waits += 1 if wmask else 0
    if len(sys.argv) > 5:
        print("%05x  stall %2d  wait %02x  %s" % (a, stall, wmask, t[:90]))
print("%d instructions, %d stall cycles (%.1f per instruction), %d instructions wait on a scoreboard" % (n, tot, tot / max(n, 1), waits))
print(", ".join("%s %d" % kv for kv in byop.most_common(14)))
