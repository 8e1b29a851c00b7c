"""Comparison helpers shared by the tests, `__graft_entry__.smoke()` and the parity leg of `bench.py`: layout-independent views of the
C-ABI outputs (so results of the library, the oracle and the reference shim can be compared record by record) and input transforms.
Pure numpy; nothing here touches the oracle or the GPU library."""
import numpy as np


def digar_view(d, o):
    """A lcd_digar_output_t dict in the layout-independent form T.collect_digar returns."""
    reads = {}
    for i in range(d["n_reads"]):
        r = int(d["ordered_read_ids"][i])
        if d["is_skipped"][r]: continue
        f, n = int(o["digar_first"][r]), int(o["n_digar"][r])
        ev = []
        for k in range(f, f + n):
            t, ln = int(o["digar_type"][k]), int(o["digar_len"][k])
            a0 = int(o["digar_alt_off"][k])
            ev.append((int(o["digar_pos"][k]), t, ln, int(o["digar_qi"][k]), int(o["digar_low_qual"][k]), bytes(o["digar_alt"][a0:a0 + ln]) if t in (1, 8) else b""))
        nf, nn = int(o["nreg_first"][r]), int(o["n_nreg"][r])
        iv = [(int(o["nreg_beg"][k]), int(o["nreg_end"][k]), int(o["nreg_label"][k])) for k in range(nf, nf + nn)]
        reads[r] = (int(o["skip"][r]), int(o["read_beg"][r]), int(o["read_end"][r]), ev, iv)
    civ = [(int(o["cnreg_beg"][k]), int(o["cnreg_end"][k]), int(o["cnreg_label"][k])) for k in range(o["n_cnreg"])]
    return dict(reads=reads, chunk_noisy=civ, qual_counts=o["qual_counts"].tolist(), totals=(o["n_digar_total"], o["n_alt_total"], o["n_nreg_total"]))


def digar_same(a, b, tag):
    for r in b["reads"]:
        assert a["reads"][r] == b["reads"][r], (tag, r, [x for x, y in zip(a["reads"][r], b["reads"][r]) if x != y][:1])
    assert a["qual_counts"] == b["qual_counts"], tag
    assert a["chunk_noisy"] == b["chunk_noisy"] and a["totals"] == b["totals"], tag


def to_md(d, rng):
    """A chunk with =/X CIGARs -> the same chunk with plain-M CIGARs + one standard MD tag per read ([0-9]+(([A-Z]|\\^[A-Z]+)[0-9]+)*)."""
    cig = np.asarray(d["cigar"], np.uint32)
    new_cig, new_off, new_n, mds, md_off = [], [], [], bytearray(), []
    for r in range(d["n_reads"]):
        ops = cig[int(d["cigar_off"][r]):int(d["cigar_off"][r]) + int(d["n_cigar"][r])]
        out, md, run, m = [], "", 0, 0
        for w in ops.tolist():
            op, ln = w & 15, w >> 4
            if op == 7: run += ln; m += ln
            elif op == 8:
                for _ in range(ln):
                    md += str(run) + "ACGTN"[int(rng.integers(0, 5))]; run = 0
                m += ln
            else:
                if m: out.append((m << 4) | 0); m = 0
                if op == 2:
                    md += str(run) + "^" + "".join("ACGT"[int(x)] for x in rng.integers(0, 4, ln)); run = 0
                out.append(w)
        if m: out.append((m << 4) | 0)
        md += str(run)
        new_off.append(len(new_cig)); new_n.append(len(out)); new_cig.extend(out)
        md_off.append(len(mds)); mds += md.encode() + b"\0"
    e = dict(d, cigar=np.array(new_cig + [0], np.uint32), cigar_off=np.array(new_off + [0], np.int64), n_cigar=np.array(new_n + [0], np.int32))
    return e, np.array(md_off + [0], np.int64), np.frombuffer(bytes(mds), np.uint8).copy()


