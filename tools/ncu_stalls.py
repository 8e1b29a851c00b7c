#!/usr/bin/env python3
"""Stall-reason totals of one kernel restricted to a source-line range: tools/ncu_stalls.py <rep> <cubin> <kernel substr> <file> <lo> <hi>"""
import csv, re, subprocess, sys
rep, cubin, fun, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi_ = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi_]
si, ie = h.index("# Samples"), h.index("Instructions Executed")
stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
sass = [r for r in rows[hi_ + 1:] if len(r) > ie and r[si].isdigit()]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
keep, lines, cur = False, [], ("?", 0)
for l in dis:
    m = re.match(r"\s*\.section\s+(\S+)", l)
    if m: keep = m.group(1).startswith(".text.") and fun in m.group(1)
    if not keep: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l): lines.append(cur)
tot = {h[i]: 0 for i in stall}; ns = ni = 0; alls = 0
for r, ln in zip(sass, lines):
    alls += int(r[si])
    if ln[0] == fname and lo <= ln[1] <= hi:
        ns += int(r[si]); ni += int(r[ie])
        for i in stall: tot[h[i]] += int(r[i] or 0)
print(f"{fname}:{lo}-{hi}: {ns} samples ({100*ns/alls:.1f}% of kernel), {ni} warp instructions, {ns/max(ni,1)*1e3:.2f} samples per 1k instr")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]: print(f"  {k:28s} {100*v/max(ns,1):5.1f}%")
