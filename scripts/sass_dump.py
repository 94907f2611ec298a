#!/usr/bin/env python
"""print the SASS of one kernel between two addresses (hex), stripped of encodings: scripts/sass_dump.py lib.so pattern [lo hi]"""
import re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); continue
    if fn and pat in fn:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and lo <= int(m.group(1), 16) <= hi:
            print("%04x  %s" % (int(m.group(1), 16), m.group(2).strip()))
