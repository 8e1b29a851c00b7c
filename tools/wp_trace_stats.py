"""Reads a LCD_DROPIN_TRACE timeline (tools/wp_trace.sh) and prints where the worker threads' time went."""
import collections, sys
for fn in sys.argv[1:]:
    ev = []
    for l in open(fn):
        r = l.rstrip("\n").split("\t"); ev.append((float(r[0]), r[1], r[2], int(r[3]), int(r[4])))
    span = ev[-1][0]
    tids = sorted({e[1] for e in ev if e[2] in ("chunk_begin",)})
    in_chunk = collections.defaultdict(float); wait = collections.defaultdict(float); cb = {}; wb = {}
    first = {}; last = {}
    for t, tid, what, a, b in ev:
        if what == "chunk_begin": cb[tid] = t; first.setdefault(tid, t)
        elif what == "chunk_end": in_chunk[tid] += t - cb[tid]; last[tid] = t
        elif what == "wait_begin": wb[tid] = t
        elif what == "wait_end": wait[tid] += t - wb[tid]
    n = len(tids)
    bb = {}; bt = collections.defaultdict(list)
    for t, tid, what, a, b in ev:
        if what == "batch_begin": bb[(tid, a)] = t
        elif what == "batch_end": bt[a].append((bb[(tid, a)], t, b))
    iv = sorted((a, b) for k in bt for a, b, _ in bt[k]); tot = 0.0; ca, cb_ = iv[0]
    for a, b in iv[1:]:
        if a > cb_: tot += cb_ - ca; ca, cb_ = a, b
        else: cb_ = max(cb_, b)
    tot += cb_ - ca
    print(f"{fn}: span {span:.2f} s, {n} worker threads; per thread: in collect_var_main {sum(in_chunk.values())/n:.2f} s (of which blocked on the GPU {sum(wait[t] for t in tids)/n:.2f} s), "
          f"outside (BAM load, idle) {span - sum(in_chunk.values())/n:.2f} s; some engine batch in flight for {tot:.2f} s")
    for k in sorted(bt):
        d = [b - a for a, b, _ in bt[k]]; ns = [m for _, _, m in bt[k]]
        print(f"   engine {k}: {len(d)} batches, {sum(ns)} problems, library time {sum(d):.2f} s, mean {1e3*sum(d)/len(d):.1f} ms, max {1e3*max(d):.1f} ms")
