#!/usr/bin/env python
"""Stall samples and executed instructions of one kernel per OUTERMOST-file source line, inlined helpers charged to their call site.
usage: ncu_phase_lines.py report.ncu-rep cubin mangled_kernel_name source_file_basename [top [first_non_helper_line]]
(ncu's source page charges an inlined helper to the helper's own line; nvdisasm -gi knows the inlining chain)"""
import csv, io, re, subprocess, sys, collections
rep, cubin, kern, fbase = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
minline = int(sys.argv[6]) if len(sys.argv) > 6 else 0   # lines below this one are small helpers: charge them to their caller
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kern + ":"))
line_of = {}
chain = []
for l in dis[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        chain.append((m.group(1), int(m.group(2))))
        if m.group(3):
            chain.append((m.group(3), int(m.group(4))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m:
        off = int(m.group(1), 16)
        if chain:
            own = [ln for f, ln in chain if f.endswith(fbase) and ln >= minline]
            cur = own[0] if own else -1
        line_of[off] = cur
        chain = []
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if r and r[0] == "Address")
iS, iN, iT, iA = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
ins = []
for r in rows:
    if r and r[0].startswith("0x"):
        ins.append((int(r[0], 16), r[iA].strip(), int(r[iS] or 0), int(r[iN] or 0), int(r[iT] or 0)))
base = min(a for a, *_ in ins)
per = collections.defaultdict(lambda: [0, 0, 0])
for a, t, s, n, th in ins:
    ln = line_of.get(a - base, -1)
    per[ln][0] += s; per[ln][1] += n; per[ln][2] += th
ts, tn = sum(v[0] for v in per.values()) or 1, sum(v[1] for v in per.values()) or 1
text = open(next(f for f, _ in [(c[0], 0) for c in [(p, 0) for p in set(f for l in dis[start:start + 4000] for f in re.findall(r'File "([^"]+)"', l))]] if f.endswith(fbase))).read().splitlines()
print("kernel %s: %d samples, %d executed warp-instructions" % (kern[:50], ts, tn))
print("-- by line (top %d by samples)" % top)
for ln, (s, n, th) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% samples %5.2f%% instr  thr %4.1f  :%d  %s" % (100.0 * s / ts, 100.0 * n / tn, th / max(n, 1), ln, text[ln - 1].strip()[:100] if 0 < ln <= len(text) else ""))
print("-- cumulative by line order")
acc_s = acc_n = 0
for ln in sorted(per):
    s, n, th = per[ln]
    acc_s += s; acc_n += n
    if s * 200 >= ts or n * 200 >= tn:
        print(":%-4d %5.2f%% samples (cum %5.1f%%) %5.2f%% instr (cum %5.1f%%)  %s" % (ln, 100.0 * s / ts, 100.0 * acc_s / ts, 100.0 * n / tn, 100.0 * acc_n / tn, text[ln - 1].strip()[:90] if 0 < ln <= len(text) else ""))
