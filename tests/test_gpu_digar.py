"""GPU: liblcd_gpu.so's K1 kernels (count / scan / fill one thread per read, warp-per-read quality histogram; through the
C-ABI) against the oracle and the golden fixtures, bit-exact: every digar1_t record, read span, skip decision, per-read noisy
interval, the chunk's noisy list and base-quality histogram; then K1's device layout handed to K2 unchanged."""
import numpy as np
import pytest

import lcd_testlib as T
from longcalld_b200 import synth
from test_oracle_digar import digar_cases

pytestmark = pytest.mark.gpu


from longcalld_b200.check import digar_view as view, digar_same as same


def test_gpu_vs_reference_fixtures(gpu):
    g = T.load_golden("digar_lcd")
    chunks = [T.digar_case_from_json(c["in"]) for c in g["cases"]]
    for c, d, o in zip(g["cases"], chunks, gpu.digar_batch(chunks)):
        a = view(d, o)
        assert T.digar_digest(a) == c["digest"] and a["chunk_noisy"] == [tuple(x) for x in c["chunk_noisy"]]


def test_gpu_vs_oracle_random(gpu, oracle):
    cases = list(digar_cases(33, 160))
    for i, (d, o) in enumerate(zip(cases, gpu.digar_batch(cases))):          # one batch of 160 chunks
        same(view(d, o), T.collect_digar(oracle, "lcd_oracle_collect_digar_eqx", d), i)
    assert gpu.digar_batch([]) == []


def test_gpu_many_intervals_and_edges(gpu, oracle):
    """Reads with more than 64 noisy intervals (cgranges' unstable radix order on the host side of the plan), empty chunks,
    chunks whose reads are all skipped, a single one-base read."""
    rng = np.random.default_rng(35)
    big = synth.make_digar_chunk(rng, n_reads=6, read_len=(100000, 140000), err_every=300, tech="ont", low_qual_frac=0.0)
    none = synth.make_digar_chunk(rng, n_reads=5); none["is_skipped"][:] = 1
    one = synth.make_digar_chunk(rng, n_reads=1, read_len=(1, 2), err_every=5)
    empty = synth.make_digar_chunk(rng, n_reads=3); empty["n_reads"] = 0
    cases = [big, none, one, empty, big]
    res = gpu.digar_batch(cases)
    want = [T.collect_digar(oracle, "lcd_oracle_collect_digar_eqx", d) for d in cases]
    assert max(len(v[4]) for v in want[0]["reads"].values()) > 64
    for i, (d, o) in enumerate(zip(cases, res)):
        same(view(d, o), want[i], i)


def test_gpu_rejects_m_cigar(gpu):
    rng = np.random.default_rng(36)
    d = synth.make_digar_chunk(rng, n_reads=4)
    d["cigar"] = d["cigar"].copy(); d["cigar"][int(d["cigar_off"][int(d["ordered_read_ids"][0])]) + 1] &= ~np.uint32(15)      # an 'M' op
    d["is_skipped"][:] = 0
    with pytest.raises(gpu.LcdGpuError, match="'M' op"):
        gpu.digar_batch([d])


def test_gpu_chunk_shaped_plan_and_chain_into_pileup(gpu, oracle):
    """Chunks shaped like 500 kb at 30x (scaled: 120 kb windows, 15 kb reads), resident plan re-run; then the records feed K2
    (lcd_pileup_batch) with candidate sites drawn from them: same per-site coverage as the oracle chain."""
    rng = np.random.default_rng(37)
    cases = [synth.make_digar_chunk(rng, n_reads=240, read_len=(10000, 20000), err_every=400, ref_len=120000) for _ in range(4)]
    plan = gpu.DigarPlan(cases)
    for _ in range(2):
        plan.run(); plan.sync()
    assert plan.work_units() > 4 * 240 * 10000
    res = plan.fetch()
    for i, (d, o) in enumerate(zip(cases, res)):
        same(view(d, o), T.collect_digar(oracle, "lcd_oracle_collect_digar_eqx", d), i)
    d, o = cases[0], res[0]
    nd = o["n_digar_total"]
    ev = [k for k in range(nd) if o["digar_type"][k] in (1, 2, 8) and not o["digar_low_qual"][k]]
    pick = sorted(set(rng.choice(ev, size=400, replace=False).tolist()))
    key = lambda k: (int(o["digar_pos"][k]) - (0 if o["digar_type"][k] == 8 else 1), int(o["digar_type"][k]), int(o["digar_len"][k]),
                     bytes(o["digar_alt"][int(o["digar_alt_off"][k]):int(o["digar_alt_off"][k]) + int(o["digar_len"][k])]) if o["digar_type"][k] != 2 else b"")
    sites = sorted({key(k): k for k in pick}.values(), key=key)
    alt, aoff = [], []
    for k in sites:
        aoff.append(len(alt))
        if o["digar_type"][k] != 2: alt.extend(o["digar_alt"][int(o["digar_alt_off"][k]):int(o["digar_alt_off"][k]) + int(o["digar_len"][k])].tolist())
    t = o["digar_type"]
    pile = dict(n_reads=d["n_reads"], n_sites=len(sites), min_bq=d["min_bq"], min_sv_len=50, ordered_read_ids=d["ordered_read_ids"],
                is_skipped=np.maximum(d["is_skipped"], o["skip"][:d["n_reads"]]), read_beg=o["read_beg"], read_end=o["read_end"], read_is_rev=d["read_is_rev"],
                digar_first=o["digar_first"], n_digar=o["n_digar"], qual_off=d["qual_off"], qual=d["qual"], digar_pos=o["digar_pos"], digar_type=t,
                digar_len=o["digar_len"], digar_qi=o["digar_qi"], digar_low_qual=o["digar_low_qual"], digar_alt_off=o["digar_alt_off"], digar_alt=o["digar_alt"],
                site_pos=np.array([o["digar_pos"][k] for k in sites] + [0], np.int64), site_type=np.array([t[k] for k in sites] + [0], np.int32),
                site_ref_len=np.array([(0 if t[k] == 1 else (o["digar_len"][k] if t[k] == 2 else 1)) for k in sites] + [0], np.int32),
                site_alt_len=np.array([(0 if t[k] == 2 else o["digar_len"][k]) for k in sites] + [0], np.int32),
                site_alt_off=np.array(aoff + [0], np.int64), site_alt=np.array(alt + [0], np.uint8))
    got = gpu.pileup_batch([pile])[0]
    want = T.pileup(oracle, "lcd_oracle_collect_cand_vars", pile)
    assert np.array_equal(got, want)
    # the same, in place: K2 / K3 on the difference lists the digar plan left in HBM (only site lists are uploaded)
    site_keys = ("site_pos", "site_type", "site_ref_len", "site_alt_len", "site_alt_off", "site_alt")
    empty = dict(n_sites=0, min_sv_len=50, var_cate=np.zeros(1, np.int32), **{k: np.zeros(1, pile[k].dtype) for k in site_keys})
    sl = dict(n_sites=len(sites), min_sv_len=50, **{k: pile[k] for k in site_keys})
    sl["var_cate"] = rng.choice(np.array([0x004, 0x008, 0x080, 0x100, 0x800], np.int32), size=len(sites) + 1, p=[0.5, 0.15, 0.15, 0.1, 0.1])
    k2 = gpu.PileupOnDigarPlan(plan, [sl, empty, empty, empty])
    k2.run(); k2.sync()
    assert np.array_equal(k2.fetch()[0], want)
    k3 = gpu.ProfileOnDigarPlan(plan, [sl, empty, empty, empty], [c["n_reads"] for c in cases])
    k3.run(); k3.sync()
    po = k3.fetch()[0]
    prof_in = dict(pile, var_cate=sl["var_cate"], nreg_first=o["nreg_first"], n_nreg=o["n_nreg"], nreg_beg=o["nreg_beg"], nreg_end=o["nreg_end"])
    rows = []
    for r in range(d["n_reads"]):
        ps, pe = int(po["prof_start"][r]), int(po["prof_end"][r]); n = max(0, pe - ps + 1) if ps >= 0 else 0; a0 = int(po["allele_off"][r])
        rows.append((ps, pe, tuple(po["alleles"][a0:a0 + n].tolist()), tuple(po["alt_qi"][a0:a0 + n].tolist())))
    assert rows == T.read_var_profile(oracle, "lcd_oracle_read_var_profile", prof_in)
    assert sum(1 for row in rows for x in row[2] if x == 1) > len(sites) // 2
    assert int(got[:, 3].sum()) > len(sites) // 2          # (events of reads K1 dropped are not counted)


def test_gpu_md_tagged_reads(gpu, oracle):
    """Reads with plain-M CIGARs + MD tags (collect_digar_from_MD_tag): the device-side MD walk feeds the same kernels; results identical to the
    =/X oracle on the chunk's own =/X CIGARs (which tests/test_oracle_md.py pins to the unmodified reference's MD path).  Mixed chunks (some reads
    =/X already), a tag that does not match its CIGAR."""
    from test_oracle_md import to_md
    rng = np.random.default_rng(79)
    cases = list(digar_cases(81, 60))
    conv = [to_md(d, rng) for d in cases]
    mixed = []
    for d, (e, md_off, md) in zip(cases[:10], conv[:10]):            # every other read keeps its =/X CIGAR: md_off < 0
        cig, off, cnt, mo = [], [], [], md_off.copy()
        for r in range(d["n_reads"]):
            src = d if r % 2 else e
            ops = src["cigar"][int(src["cigar_off"][r]):int(src["cigar_off"][r]) + int(src["n_cigar"][r])]
            off.append(len(cig)); cnt.append(len(ops)); cig.extend(ops.tolist())
            if r % 2: mo[r] = -1
        mixed.append((dict(d, cigar=np.array(cig + [0], np.uint32), cigar_off=np.array(off + [0], np.int64), n_cigar=np.array(cnt + [0], np.int32)), mo, md))
    chunks = [e for e, _, _ in conv] + [m for m, _, _ in mixed]
    tags = [(o, m) for _, o, m in conv] + [(o, m) for _, o, m in mixed]
    res = gpu.digar_md_batch(chunks, tags)
    for i, (d, o) in enumerate(zip(cases + cases[:10], res)):
        same(view(d, o), T.collect_digar(oracle, "lcd_oracle_collect_digar_eqx", d), i)
    e, md_off, md = conv[0]
    bad = md.copy(); first = int(md_off[int(e["ordered_read_ids"][0])]); bad[first] = ord("#")
    e2 = dict(e, is_skipped=np.zeros_like(e["is_skipped"]))
    with pytest.raises(gpu.LcdGpuError, match="MD tag and CIGAR do not match"):
        gpu.digar_md_batch([e2], [(md_off, bad)])


def test_gpu_cs_tagged_and_untagged_reads(gpu, oracle):
    """The two other variants of the reference's driver (src/collect_var.c:1072-1080) through lcd_digar_tags_batch: plain-M reads with a cs tag
    (collect_digar_from_cs_tag) and without any tag (collect_digar_from_ref_seq, bases against the chunk's reference window, reads hanging over
    its ends included), one batch with both kinds of chunks, against the oracle restatements that tests/test_oracle_digar_cs.py and
    tests/test_oracle_digar_refseq.py pin to the unmodified reference.  A cs tag whose letters differ from SEQ is refused."""
    from test_oracle_digar_cs import to_cs
    from test_oracle_digar_refseq import to_refseq, refseq_args
    import ctypes as C
    from longcalld_b200.capi import TAG_CS, TAG_REFSEQ
    rng = np.random.default_rng(113)
    cases = list(digar_cases(115, 80))
    chunks, tags, want = [], [], []
    for n, d in enumerate(cases):
        if n % 2 == 0:
            e, off, cs = to_cs(d, rng)
            chunks.append(e); tags.append(dict(kind=np.full(e["n_reads"] + 1, TAG_CS, np.int8), off=off, text=cs))
            want.append(T.collect_digar(oracle, "lcd_oracle_collect_digar_cs", e, mid_args=(off.ctypes.data_as(C.c_void_p), cs.ctypes.data_as(C.c_void_p)), cap_like=d, slack=64))
        else:
            e, ref, rb, re_ = to_refseq(d, rng, trim=n % 3 != 0)
            chunks.append(e); tags.append(dict(kind=np.full(e["n_reads"] + 1, TAG_REFSEQ, np.int8), ref_seq=ref, ref_beg=rb, ref_end=re_))
            want.append(T.collect_digar(oracle, "lcd_oracle_collect_digar_refseq", e, mid_args=refseq_args(ref, rb, re_), cap_like=d, slack=2000))
    res = gpu.digar_tags_batch(chunks, tags)
    for i, (e, o, w) in enumerate(zip(chunks, res, want)):
        same(view(e, o), w, i)
    e, off, cs = to_cs(cases[0], rng)
    bad = cs.copy(); i = [k for k in range(len(bad) - 2) if bad[k] == ord("*")][0]; bad[i + 2] = ord("a") if bad[i + 2] != ord("a") else ord("c")
    with pytest.raises(gpu.LcdGpuError, match="differ from its SEQ"):
        gpu.digar_tags_batch([dict(e, is_skipped=np.zeros_like(e["is_skipped"]))], [dict(kind=np.full(e["n_reads"] + 1, TAG_CS, np.int8), off=off, text=bad)])
