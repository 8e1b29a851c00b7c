#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel: joins `ncu --page source --csv` (SASS rows with stall samples and
instruction counts) with `nvdisasm -g` line info of the same cubin.  Usage:
  tools/ncu_lines.py <report.ncu-rep> <cubin> [top_n] [mangled-name substring of the kernel, when the cubin holds several]"""
import csv, re, subprocess, sys
rep, cubin = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fun = sys.argv[4] if len(sys.argv) > 4 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
si, ie, te = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
sass = [(int(r[si]), int(r[ie]), int(r[te]), r[h.index("Source")].strip()) for r in rows[hi + 1:] if len(r) > te and r[si].isdigit()]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
if fun:      # keep only the .text section of the requested kernel
    keep, out_l = False, []
    for l in dis:
        m = re.match(r"\s*\.section\s+(\S+)", l)
        if m:
            keep = m.group(1).startswith(".text.") and fun in m.group(1)
        if keep:
            out_l.append(l)
    dis = out_l
lines, cur = [], ("?", 0)
for l in dis:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l):
        lines.append(cur)
if len(lines) != len(sass):
    print(f"warning: {len(lines)} disassembled instructions vs {len(sass)} profiled; aligning by index", file=sys.stderr)
agg = {}
for (s, i, t, _), ln in zip(sass, lines):
    a = agg.setdefault(ln, [0, 0, 0]); a[0] += s; a[1] += i; a[2] += t
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"total samples {ts}, warp instructions {ti}")
src_cache = {}
def src(ln):
    import glob
    f = src_cache.get(ln[0])
    if f is None:
        c = glob.glob(f"longcalld_b200/csrc/{ln[0]}")
        f = src_cache[ln[0]] = open(c[0]).read().splitlines() if c else []
    return f[ln[1] - 1].strip()[:100] if 0 < ln[1] <= len(f) else ""
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/ts:5.1f}% samples {100*a[1]/ti:5.1f}% inst  lanes {a[2]/max(a[1],1):4.1f}  {ln[0]}:{ln[1]}  {src(ln)}")
