#!/usr/bin/env python
"""static SASS report of one kernel: size, opcode histogram of the hottest backward-branch loop"""
import re, subprocess, sys, collections
so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn = None; ins = []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); continue
    if fn and pat in fn:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
print("kernel", pat, "instructions", len(ins))
# loops: backward branches
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?([0-9a-fx]+)\)?", t)
    m2 = re.search(r"BRA.*0x([0-9a-f]+)", t)
    if m2:
        tgt = int(m2.group(1), 16)
        if tgt < a:
            loops.append((tgt, a))
for tgt, a in loops:
    body = [t for ad, t in ins if tgt <= ad <= a]
    ops = collections.Counter()
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0].split(".")[0]] += 1
    print("loop 0x%x..0x%x: %d instructions; %s" % (tgt, a, len(body), ", ".join("%s %d" % kv for kv in ops.most_common(16))))
