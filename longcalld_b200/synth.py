"""Seeded synthetic workloads shaped like the noisy regions `longcallD call` re-aligns on 30x
long-read data (BASELINE.json configs 2-4).  There is no read simulator or aligner in the image, so
the generator emits directly what the re-alignment stack consumes: per noisy region a reference
window, two haplotype sequences carrying planted variants, and ~15 reads per haplotype with a
HiFi- or ONT-like error model.

Shape parameters come from the reference's behaviour on its bundled data:
~375 re-aligned regions per Mb (HiFi), median window ~60-70 bp with a kilobase tail (SV / VNTR
regions, up to 3.3 kb observed), ~31 seq->graph POA alignments and ~2.5 ref-vs-consensus WFA
alignments per region.
"""
import numpy as np

REGIONS_PER_MB = {"hifi": 375, "ont": 1725}       # ONT: 345 regions / 0.2 Mb on the bundled data


def _low_complexity(rng, n):
    kind = rng.integers(0, 3)
    if kind == 0:                                  # homopolymer
        return np.full(n, rng.integers(0, 4), dtype=np.uint8)
    unit = rng.integers(0, 4, int(rng.integers(2, 7))).astype(np.uint8)
    return np.resize(unit, n)


def sdust_window(rng, n, lc_every=2000, n_frac=0.0005):
    """ASCII reference window of n bases for K0: random sequence with homopolymers, short tandem repeats and AT-rich stretches planted every ~lc_every
    bases, a few N runs and soft-masked (lower-case) bases.  (At the reference's T = 5 random sequence alone gives an interval every ~65 bases.)"""
    s = rng.integers(0, 4, n).astype(np.uint8)
    p = int(rng.integers(0, lc_every))
    while p < n - 80:
        kind = int(rng.integers(0, 3))
        if kind == 0: s[p:p + int(rng.integers(4, 40))] = rng.integers(0, 4)
        elif kind == 1:
            u, c = int(rng.integers(2, 7)), int(rng.integers(3, 15))
            s[p:p + u * c] = np.tile(rng.integers(0, 4, u).astype(np.uint8), c)
        else:
            L = int(rng.integers(10, 60)); s[p:p + L] = rng.choice(np.array([0, 3], np.uint8), L)
        p += int(rng.integers(5, 2 * lc_every))
    a = np.frombuffer(b"ACGT", np.uint8)[s].copy()
    for q in rng.integers(0, n, max(1, int(n * n_frac))): a[q:q + int(rng.integers(1, 30))] = ord("N")
    low = rng.random(n) < 0.05; a[low & (a != ord("N"))] |= 0x20
    return a


def _region_ref(rng, L):
    ref = rng.integers(0, 4, L).astype(np.uint8)
    if rng.random() < 0.6 and L >= 24:             # most noisy regions sit on a repeat tract
        n = int(min(L - 8, rng.integers(8, 40)))
        s = int(rng.integers(4, L - n - 3))
        ref[s:s + n] = _low_complexity(rng, n)
    return ref


def _apply_variants(rng, ref, p_snp, p_indel, sv=None):
    """Haplotype = reference window + SNPs + 1-6 bp indels (+ one SV)."""
    L = len(ref)
    out = ref.copy()
    snp = np.nonzero(rng.random(L) < p_snp)[0]
    out[snp] = (out[snp] + rng.integers(1, 4, len(snp))) % 4
    pieces, last = [], 0
    events = sorted(int(x) for x in np.nonzero(rng.random(L) < p_indel)[0])
    if sv is not None:
        events = sorted(events + [sv[0]])
    for pos in events:
        if pos < last:
            continue
        pieces.append(out[last:pos])
        if sv is not None and pos == sv[0]:
            if sv[1] == "ins":
                pieces.append(rng.integers(0, 4, sv[2]).astype(np.uint8))
                last = pos
            else:
                last = min(L, pos + sv[2])
        elif rng.random() < 0.5:
            k = int(rng.integers(1, 7))
            pieces.append(np.resize(out[max(0, pos - 1):pos + 1], k).astype(np.uint8))   # repeat-like insertion
            last = pos
        else:
            last = min(L, pos + int(rng.integers(1, 7)))
    pieces.append(out[last:])
    return np.concatenate(pieces) if pieces else out


def _read_errors(rng, seq, tech):
    """HiFi: 0.2 % errors, 80 % homopolymer-style indels.  ONT R10: 1.5 %, subs 40 / ins 25 / del 35."""
    L = len(seq)
    if L == 0:
        return seq
    rate = 0.002 if tech == "hifi" else 0.015
    ev = np.nonzero(rng.random(L) < rate)[0]
    if len(ev) == 0:
        return seq
    out = seq.copy()
    kind = rng.random(len(ev))
    if tech == "hifi":
        sub, ins = ev[kind < 0.2], ev[(kind >= 0.2) & (kind < 0.6)]
        dele = ev[kind >= 0.6]
    else:
        sub, ins = ev[kind < 0.4], ev[(kind >= 0.4) & (kind < 0.65)]
        dele = ev[kind >= 0.65]
    out[sub] = (out[sub] + rng.integers(1, 4, len(sub))) % 4
    if len(ins):
        out = np.insert(out, ins, out[ins])        # duplicate the base: homopolymer-style insertion
    if len(dele):
        shift = np.searchsorted(ins, dele)         # positions moved by the insertions before them
        out = np.delete(out, dele + shift)
    return out


def region_length(rng):
    u = rng.random()
    if u < 0.72:
        return int(np.clip(rng.lognormal(np.log(60), 0.7), 20, 1200))
    if u < 0.96:
        return int(rng.integers(100, 1000))
    return int(rng.integers(500, 3500))


class Region:
    __slots__ = ("ref", "haps", "reads", "read_hap")


def make_regions(mbp, tech="hifi", seed=11, with_reads=True, coverage=30):
    """Noisy regions for `mbp` megabases of reference at `coverage`x."""
    rng = np.random.default_rng(seed)
    n_regions = max(1, int(round(mbp * REGIONS_PER_MB[tech])))
    regions = []
    for _ in range(n_regions):
        L = region_length(rng)
        ref = _region_ref(rng, L)
        sv = None
        if L >= 500 and rng.random() < 0.5:
            sv = (int(rng.integers(L // 4, L // 2)), "ins" if rng.random() < 0.5 else "del", int(rng.integers(50, min(3000, 2 * L))))
        h1 = _apply_variants(rng, ref, 0.004, 0.004, sv)
        h2 = _apply_variants(rng, ref, 0.004, 0.004, None) if rng.random() < 0.7 else h1.copy()
        r = Region()
        r.ref, r.haps, r.reads, r.read_hap = ref, (h1, h2), [], []
        if with_reads:
            for h in (0, 1):
                n = int(np.clip(rng.poisson(coverage / 2), 3, 40))
                for _ in range(n):
                    r.reads.append(_read_errors(rng, r.haps[h], tech))
                    r.read_hap.append(h + 1)
        regions.append(r)
    return regions


def wfa_problems(regions):
    """The ref-vs-consensus gap-affine-2p / no-heuristic alignments of wfa_collect_aln_str
    (reference src/align.c:565): one per (region, haplotype)."""
    return [(r.ref, h) for r in regions for h in r.haps]


# ----------------------------------------------------------------------------- K4 workload: read x variant profiles of region chunks
CATE = {"CLEAN_HET_SNP": 0x004, "CLEAN_HET_INDEL": 0x008, "CLEAN_HOM_VAR": 0x080, "NOISY_CAND_HET_VAR": 0x100,
        "NOISY_CAND_HOM_VAR": 0x200, "LOW_COV_VAR": 0x001, "CAND_SOMATIC_VAR": 0x040}
CATE_CLEAN = 0x004 | 0x008 | 0x080
CATE_GERMLINE = CATE_CLEAN | 0x100 | 0x200


def make_phase_chunk(rng, n_vars=60, n_reads=200, err=0.02, tech="hifi", shuffle_order=False):
    """A synthetic chunk for the read -> haplotype assignment: a diploid truth over sorted candidate variants of mixed
    categories, reads sampled from the two haplotypes spanning contiguous variant ranges (some skipped, some empty,
    some carrying noise / low-quality calls), with coverage counts derived from the reads."""
    pos = np.sort(rng.choice(np.arange(1000, 1000 + 400 * max(n_vars, 1)), size=n_vars, replace=False)).astype(np.int64)
    cats = np.array([CATE["CLEAN_HET_SNP"], CATE["CLEAN_HET_INDEL"], CATE["CLEAN_HOM_VAR"], CATE["NOISY_CAND_HET_VAR"],
                     CATE["NOISY_CAND_HOM_VAR"], CATE["LOW_COV_VAR"], CATE["CAND_SOMATIC_VAR"]], dtype=np.int32)
    var_cate = rng.choice(cats, size=n_vars, p=[0.45, 0.12, 0.12, 0.12, 0.05, 0.08, 0.06]).astype(np.int32)
    var_type = np.where(var_cate == CATE["CLEAN_HET_SNP"], 8, rng.choice([8, 1, 2], size=n_vars)).astype(np.int32)
    var_type[var_cate == CATE["CLEAN_HET_INDEL"]] = rng.choice([1, 2], size=int((var_cate == CATE["CLEAN_HET_INDEL"]).sum()))
    is_hp = ((var_type != 8) & (rng.random(n_vars) < 0.3)).astype(np.int32)
    is_hom = (var_cate == CATE["CLEAN_HOM_VAR"]) | (var_cate == CATE["NOISY_CAND_HOM_VAR"])
    alt_hap = rng.integers(1, 3, n_vars)                                   # which haplotype carries the alt allele of a het variant
    n_uniq = np.where(rng.random(n_vars) < 0.1, 3, 2).astype(np.int32)
    starts, ends, haps_true = [], [], []
    for _ in range(n_reads):
        if n_vars == 0 or rng.random() < 0.05:
            starts.append(-1); ends.append(-2)
        else:
            s = int(rng.integers(0, n_vars)); L = int(np.clip(rng.poisson(12 if tech == "hifi" else 25), 1, n_vars))
            starts.append(s); ends.append(min(n_vars - 1, s + L - 1))
        haps_true.append(int(rng.integers(1, 3)))
    order = np.argsort(np.array(starts), kind="stable")                    # reads arrive position-sorted
    starts = np.array(starts, dtype=np.int32)[order]; ends = np.array(ends, dtype=np.int32)[order]; haps_true = np.array(haps_true)[order]
    allele_off = np.zeros(n_reads, dtype=np.int64); alle = []
    alle_covs = np.zeros((n_vars, 4), dtype=np.int32)
    is_skipped = (rng.random(n_reads) < 0.04).astype(np.uint8)
    for r in range(n_reads):
        allele_off[r] = len(alle)
        for v in range(starts[r], ends[r] + 1) if starts[r] >= 0 else ():
            a = 1 if (is_hom[v] or alt_hap[v] == haps_true[r]) else 0
            u = rng.random()
            if u < err: a = 1 - a
            elif u < err + 0.02: a = -1
            elif u < err + 0.03: a = -2
            elif u < err + 0.035 and n_uniq[v] == 3: a = 2
            alle.append(a)
            if a >= 0 and not is_skipped[r]: alle_covs[v, a] += 1
    ordered = np.arange(n_reads, dtype=np.int32)
    if shuffle_order: rng.shuffle(ordered)
    d = dict(n_reads=n_reads, n_vars=n_vars, ordered_read_ids=ordered, is_skipped=is_skipped, prof_start=starts, prof_end=ends,
             allele_off=allele_off, alleles=np.array(alle + [0], dtype=np.int8), var_cate=var_cate, var_type=var_type, is_hp_indel=is_hp,
             n_uniq_alles=n_uniq, alle_covs=np.ascontiguousarray(alle_covs), total_cov=alle_covs.sum(axis=1).astype(np.int32), pos=pos)
    return d



def phase_chunks(mbp, tech="hifi", seed=11):
    """The 500 kb region chunks of `mbp` megabases (LONGCALLD_BAM_CHUNK_REG_SIZE, reference src/bam_utils.h:10) as inputs of
    the read -> haplotype assignment: ~1 050 reads (30x of 15 kb reads incl. overlap reads) and ~6 candidate variants per kb."""
    rng = np.random.default_rng(seed + 7)
    n = max(1, int(round(mbp / 0.5)))
    return [make_phase_chunk(rng, n_vars=int(rng.integers(2500, 3500)), n_reads=int(rng.integers(950, 1150)), err=0.01, tech=tech)
            for _ in range(n)]


def edlib_pairs(regions, per_mb=210, seed=11, mbp=1.0):
    """(read window, first read) pairs of the sampling filter (edlib_xgaps in collect_partial_aln_beg_end, reference
    src/align.c:722-733): ~42 calls per 0.2 Mb on the bundled HiFi data."""
    rng = np.random.default_rng(seed + 13)
    want = max(1, int(round(per_mb * mbp)))
    # the calls observed on the bundled data are 42 per 0.2 Mb with a median of 202 x 202 and a maximum of 941 x 941
    # (only deep regions are sampled), so pairs come from windows of 100-1000 bp
    cand = [(ri, k) for ri, r in enumerate(regions) if 100 <= len(r.reads[0]) <= 1000
            for k in range(1, len(r.reads)) if 100 <= len(r.reads[k]) <= 1000]
    pick = rng.choice(len(cand), size=min(want, len(cand)), replace=False)
    return [(regions[cand[i][0]].reads[cand[i][1]], regions[cand[i][0]].reads[0]) for i in np.sort(pick)]


# ----------------------------------------------------------------------------- K2 workload: difference lists vs candidate sites
def make_pileup_chunk(rng, ref_len=20000, n_reads=200, read_len=(3000, 8000), var_every=150, err_every=800, min_sv_len=50):
    """A synthetic region chunk for the per-site coverage pass: sorted candidate sites (SNPs, small and large insertions,
    deletions) and, per read, the difference list the pileup scan derives from an =/X CIGAR (digar1_t: '=' runs, X bases,
    I / D events with read offset, low-quality flag) plus base qualities.  Reads carry the variants of their haplotype, random
    errors (some promoted to sites), and length-perturbed copies of large insertions (the reference matches those fuzzily)."""
    INS, DEL, EQ, DIFF = 1, 2, 7, 8
    ref0 = 100000
    truth = []
    p = ref0 + int(rng.integers(20, var_every))
    while p < ref0 + ref_len - 200:
        u = rng.random()
        if u < 0.7: v = (p, DIFF, 1, 1, rng.integers(0, 4, 1).astype(np.uint8))
        elif u < 0.85:
            n = int(rng.integers(min_sv_len, 130)) if rng.random() < 0.15 else int(rng.integers(1, 11))
            v = (p, INS, 0, n, rng.integers(0, 4, n).astype(np.uint8))
        else:
            n = int(rng.integers(1, 11)); v = (p, DEL, n, 0, np.zeros(0, np.uint8))
        truth.append((v, int(rng.integers(1, 4))))                     # carried by hap 1, hap 2 or both (3)
        p += int(rng.integers(12, 2 * var_every)) + v[2]
    reads, extra_sites = [], []
    for _ in range(n_reads):
        L = int(rng.integers(*read_len)); beg = ref0 + int(rng.integers(0, max(1, ref_len - L))); end = min(ref0 + ref_len - 1, beg + L - 1)
        hap = int(rng.integers(1, 3)); ev = []
        for v, h in truth:
            if beg < v[0] and v[0] + v[2] + 2 < end and (h & hap) and rng.random() < 0.95:
                alt = v[4]
                if v[1] == INS and v[3] >= min_sv_len and rng.random() < 0.6:       # a noisy copy of a large insertion
                    alt = rng.integers(0, 4, max(1, int(v[3] * rng.uniform(0.7, 1.3)))).astype(np.uint8)
                ev.append((v[0], v[1], v[2], len(alt), alt, int(rng.random() < 0.05)))
        q = beg + int(rng.integers(1, err_every))
        while q + 12 < end:
            t = [DIFF, INS, DEL][int(rng.integers(0, 3))]; n = 1 if t == DIFF else int(rng.integers(1, 4))
            alt = rng.integers(0, 4, n).astype(np.uint8) if t != DEL else np.zeros(0, np.uint8)
            ev.append((q, t, 0 if t == INS else (n if t == DEL else 1), len(alt), alt, int(rng.random() < 0.3)))
            q += int(rng.integers(12, 2 * err_every))
        ev.sort(key=lambda e: (e[0], e[1]))
        keep, last = [], beg
        for e in ev:                                                    # events at least 12 reference bases apart
            if e[0] >= last + 12 and e[0] + e[2] + 12 < end: keep.append(e); last = e[0] + e[2]
        for e in keep:
            if rng.random() < 0.02: extra_sites.append((e[0], e[1], e[2], e[3], e[4]))
        reads.append((beg, end, int(rng.random() < 0.5), keep))
    key = lambda s: (s[0] if s[1] == DIFF else s[0] - 1, s[1], s[2], s[3], bytes(s[4]))
    sites, seen = [], set()
    for s_ in sorted([t[0] for t in truth] + extra_sites, key=key):
        if key(s_) not in seen: seen.add(key(s_)); sites.append(s_)
    d = dict(n_reads=n_reads, n_sites=len(sites), min_bq=10, min_sv_len=min_sv_len)
    dpos, dtype, dlen, dqi, dlow, daoff, dalt, dfirst, ndig, qoff, quals = [], [], [], [], [], [], [], [], [], [], []
    for beg, end, rev, ev in reads:
        dfirst.append(len(dpos)); cur, qi = beg, 0
        for pos, t, rl, al, alt, low in ev:
            if pos > cur: dpos.append(cur); dtype.append(EQ); dlen.append(pos - cur); dqi.append(qi); dlow.append(0); daoff.append(len(dalt)); qi += pos - cur
            dpos.append(pos); dtype.append(t); dlen.append(rl if t == DEL else al); dqi.append(qi); dlow.append(low); daoff.append(len(dalt)); dalt.extend(alt.tolist())
            if t == DIFF: qi += 1; cur = pos + 1
            elif t == INS: qi += al; cur = pos
            else: cur = pos + rl
        if end >= cur: dpos.append(cur); dtype.append(EQ); dlen.append(end - cur + 1); dqi.append(qi); dlow.append(0); daoff.append(len(dalt)); qi += end - cur + 1
        ndig.append(len(dpos) - dfirst[-1]); qoff.append(len(quals)); quals.extend(rng.integers(3, 41, qi + 1).tolist())
    d.update(ordered_read_ids=np.arange(n_reads, dtype=np.int32), is_skipped=(rng.random(n_reads) < 0.04).astype(np.uint8),
             read_beg=np.array([r[0] for r in reads], np.int64), read_end=np.array([r[1] for r in reads], np.int64),
             read_is_rev=np.array([r[2] for r in reads], np.uint8), digar_first=np.array(dfirst, np.int64), n_digar=np.array(ndig, np.int32),
             qual_off=np.array(qoff, np.int64), qual=np.array(quals + [0], np.uint8), digar_pos=np.array(dpos + [0], np.int64),
             digar_type=np.array(dtype + [0], np.int8), digar_len=np.array(dlen + [0], np.int32), digar_qi=np.array(dqi + [0], np.int32),
             digar_low_qual=np.array(dlow + [0], np.uint8), digar_alt_off=np.array(daoff + [0], np.int64), digar_alt=np.array(dalt + [0], np.uint8),
             site_pos=np.array([s_[0] for s_ in sites] + [0], np.int64), site_type=np.array([s_[1] for s_ in sites] + [0], np.int32),
             site_ref_len=np.array([s_[2] for s_ in sites] + [0], np.int32), site_alt_len=np.array([s_[3] for s_ in sites] + [0], np.int32))
    aoff, flat = [], []
    for s_ in sites: aoff.append(len(flat)); flat.extend(s_[4].tolist())
    d.update(site_alt_off=np.array(aoff + [0], np.int64), site_alt=np.array(flat + [0], np.uint8))
    return d


def add_profile_inputs(rng, d):
    """Extend a pileup chunk with what the read x variant profile pass reads besides it: a category per candidate variant
    (a few LONGCALLD_NON_VAR entries are skipped by the reference) and per-read noisy intervals (cgranges [beg, end))."""
    cats = np.array([CATE["CLEAN_HET_SNP"], CATE["CLEAN_HET_INDEL"], CATE["CLEAN_HOM_VAR"], CATE["NOISY_CAND_HET_VAR"], CATE["LOW_COV_VAR"], 0x800], dtype=np.int32)
    d = dict(d)
    d["var_cate"] = rng.choice(cats, size=d["n_sites"] + 1, p=[0.5, 0.15, 0.1, 0.1, 0.05, 0.1]).astype(np.int32)
    first, cnt, nb, ne = [], [], [], []
    for r in range(d["n_reads"]):
        first.append(len(nb)); k = int(rng.integers(0, 3)) if rng.random() < 0.4 else 0
        for _ in range(k):
            b = int(rng.integers(d["read_beg"][r], d["read_end"][r] + 1)); nb.append(b); ne.append(b + int(rng.integers(1, 400)))
        cnt.append(k)
    d.update(nreg_first=np.array(first + [0], np.int64), n_nreg=np.array(cnt + [0], np.int32), nreg_beg=np.array(nb + [0], np.int64), nreg_end=np.array(ne + [0], np.int64))
    return d


# ----------------------------------------------------------------------------- K1 workload: =/X CIGAR reads of a chunk
def make_digar_chunk(rng, n_reads=100, read_len=(800, 4000), err_every=300, tech="hifi", low_qual_frac=0.05, ref0=100000, ref_len=30000):
    """A synthetic region chunk as the loader hands it to the pileup scan: per read the BAM fields the =/X path reads
    (0-based pos, strand, CIGAR words, 4-bit packed SEQ, QUAL).  Reads alternate '=' runs with X runs, insertions, deletions and
    the odd N skip; some carry dense clusters of differences (they become noisy intervals), long soft / hard clips (clip
    intervals; a few flagged as ONT palindromes), sit at the very start / end of the contig, or are dense enough to be dropped."""
    EQ, X, I, D, N, S, H = 7, 8, 1, 2, 3, 4, 5
    ont = tech == "ont"
    whole = ref0 + ref_len + (5 if rng.random() < 0.3 else 5000)
    cig, cig_off, n_cig, pos0, rev, pal, lq, seq_off, qual_off, bseq, qual = [], [], [], [], [], [], [], [], [], [], []
    for r in range(n_reads):
        L = int(rng.integers(*read_len)); ops = []
        start = 0 if rng.random() < 0.03 else ref0 + int(rng.integers(0, max(1, ref_len - L)))
        if rng.random() < 0.25: ops.append((S if rng.random() < 0.8 else H, int(rng.choice([3, 25, 31, 120, 700]))))
        dense_read = rng.random() < 0.04
        every = max(2, err_every // 40) if dense_read else err_every
        left = L
        while left > 0:
            cluster = rng.random() < 0.1
            k = int(rng.integers(3, 12)) if cluster else 1
            run = min(left, int(rng.integers(1, 2 * every)))
            ops.append((EQ, run)); left -= run
            for _ in range(k):
                if left <= 0: break
                u = rng.random()
                if u < 0.5: n = int(rng.integers(1, 4)); ops.append((X, n)); left -= n
                elif u < 0.72: ops.append((I, int(rng.integers(1, 8)) if rng.random() < 0.9 else int(rng.integers(20, 300))))
                elif u < 0.97: n = int(rng.integers(1, 8)) if rng.random() < 0.9 else int(rng.integers(20, 300)); ops.append((D, n)); left -= n
                else: n = int(rng.integers(50, 500)); ops.append((N, n)); left -= n
                if cluster and left > 0: g = min(left, int(rng.integers(1, 30 if not ont else 12))); ops.append((EQ, g)); left -= g
        if ops[-1][0] not in (EQ, X): ops.append((EQ, int(rng.integers(1, 50))))
        if rng.random() < 0.25: ops.append((S if rng.random() < 0.8 else H, int(rng.choice([3, 25, 31, 120, 700]))))
        merged = []
        for o in ops:                                                  # no two equal neighbouring ops
            if merged and merged[-1][0] == o[0]: merged[-1] = (o[0], merged[-1][1] + o[1])
            else: merged.append(o)
        ql = sum(n for t, n in merged if t in (EQ, X, I, S))
        cig_off.append(len(cig)); n_cig.append(len(merged)); cig.extend((n << 4) | t for t, n in merged)
        pos0.append(start); rev.append(int(rng.random() < 0.5)); pal.append(int(ont and rng.random() < 0.2)); lq.append(ql)
        codes = rng.choice(np.array([1, 2, 4, 8, 15], np.uint8), size=ql + (ql & 1), p=[0.245, 0.245, 0.245, 0.245, 0.02])
        seq_off.append(len(bseq)); bseq.extend(((codes[0::2] << 4) | codes[1::2]).tolist())
        q = rng.integers(12, 45, ql); q[rng.random(ql) < low_qual_frac] = rng.integers(0, 10)
        qual_off.append(len(qual)); qual.extend(q.tolist())
    hifi_win, ont_win = 100, 25
    return dict(n_reads=n_reads, min_bq=10, noisy_reg_max_xgaps=5, noisy_reg_slide_win=ont_win if ont else hifi_win, end_clip_reg=30,
                end_clip_reg_flank_win=100, max_noisy_frac_per_read=0.5, max_var_ratio_per_read=0.05, whole_ref_len=whole,
                reg_beg=ref0 + ref_len // 4, reg_end=ref0 + 3 * ref_len // 4,
                ordered_read_ids=rng.permutation(n_reads).astype(np.int32), is_skipped=(rng.random(n_reads) < 0.05).astype(np.uint8),
                read_pos0=np.array(pos0, np.int64), read_is_rev=np.array(rev, np.uint8), is_palindrome=np.array(pal, np.uint8),
                n_cigar=np.array(n_cig, np.int32), cigar_off=np.array(cig_off, np.int64), cigar=np.array(cig + [0], np.uint32),
                l_qseq=np.array(lq, np.int32), seq_off=np.array(seq_off, np.int64), bseq=np.array(bseq + [0], np.uint8),
                qual_off=np.array(qual_off, np.int64), qual=np.array(qual + [0], np.uint8))


def digar_chunks_30x(n_chunks, tech="hifi", seed=11, chunk_len=500000, read_mean=15000, coverage=30):
    """K1 workload at BASELINE shape: `n_chunks` 500 kb region chunks at 30x, ~1 050 reads of ~15 kb each with =/X CIGARs.
    Two haplotypes carry het / hom SNPs (1 per kb) and small indels (1 per 8 kb); reads add sequencing errors (HiFi 0.2 %,
    mostly 1-bp indels; ONT 1.5 %) and 3 % of them a >= 100 bp soft clip.  Built with numpy per read (no simulator in the image)."""
    EQ, X, I, D, S = 7, 8, 1, 2, 4
    rng = np.random.default_rng(seed)
    ont = tech == "ont"
    err = 0.015 if ont else 0.002
    out = []
    for c in range(n_chunks):
        reg_beg = 1000000 + c * chunk_len; reg_end = reg_beg + chunk_len - 1
        n_var = chunk_len // 1000 + chunk_len // 8000
        vpos = np.sort(rng.choice(np.arange(reg_beg - read_mean, reg_end + read_mean, 3), size=n_var * 2, replace=False))[: n_var * 2]
        vtype = rng.choice(np.array([X, I, D]), size=len(vpos), p=[0.89, 0.055, 0.055])
        vlen = np.where(vtype == X, 1, rng.integers(1, 9, len(vpos)))
        vhap = rng.integers(1, 4, len(vpos))                         # carried by hap 1, hap 2, or both
        valt = rng.integers(0, 4, (len(vpos), 8)).astype(np.uint8)
        # sequencing errors: 80 % recur at homopolymer-like hot spots (a fixed 1-bp insertion or deletion per spot, one per 250 bp),
        # 20 % fall anywhere -- so a chunk has a few thousand distinct candidate sites, as real data does
        hpos = np.sort(rng.choice(np.arange(reg_beg - read_mean, reg_end + read_mean, 5), size=(chunk_len + 2 * read_mean) // 250, replace=False))
        htype = rng.choice(np.array([I, D]), size=len(hpos)); halt = rng.integers(0, 4, (len(hpos), 8)).astype(np.uint8)
        n_reads = int(coverage * (chunk_len + read_mean) / read_mean)
        starts = np.sort(rng.integers(reg_beg - read_mean, reg_end, n_reads))
        lens = np.clip(rng.normal(read_mean, read_mean / 5, n_reads).astype(np.int64), 2000, 3 * read_mean) if not ont else \
            np.clip(rng.lognormal(np.log(read_mean), 0.6, n_reads).astype(np.int64), 2000, 6 * read_mean)
        cig, cig_off, n_cig, lq, seq_off, qual_off, seqs, quals = [], [], [], [], [], [], [], []
        c_top = s_top = q_top = 0
        code = np.array([1, 2, 4, 8], np.uint8)
        for r in range(n_reads):
            beg, L = int(starts[r]), int(lens[r]); hap = int(rng.integers(1, 3))
            lo, hi = np.searchsorted(vpos, [beg + 20, beg + L - 20])
            m = (vhap[lo:hi] & hap) != 0
            p1, t1, l1 = vpos[lo:hi][m], vtype[lo:hi][m], vlen[lo:hi][m]; a1 = valt[lo:hi][m]
            ne_all = rng.poisson(err * L); nh = int(rng.binomial(ne_all, 0.8)); ne = ne_all - nh
            p2 = rng.integers(beg + 20, beg + L - 20, ne); u = rng.random(ne)
            t2 = np.where(u < (0.6 if ont else 0.5), X, np.where(u < (0.8 if ont else 0.75), I, D)); l2 = np.where(t2 == X, 1, rng.integers(1, 3, ne))
            a2 = rng.integers(0, 4, (ne, 8)).astype(np.uint8)
            hl, hh = np.searchsorted(hpos, [beg + 20, beg + L - 20])
            hi_ = np.unique(rng.integers(hl, max(hl + 1, hh), nh)) if hh > hl else np.zeros(0, np.int64)
            p, t, l, a = (np.concatenate([p1, p2, hpos[hi_]]), np.concatenate([t1, t2, htype[hi_]]), np.concatenate([l1, l2, np.ones(len(hi_), np.int64)]),
                          np.concatenate([a1, a2, halt[hi_]]))
            o = np.argsort(p, kind="stable"); p, t, l, a = p[o], t[o], l[o], a[o]
            ref_span = np.where(t == I, 0, l)
            keep = np.ones(len(p), bool)
            if len(p) > 1:
                keep[1:] = p[1:] >= (p[:-1] + ref_span[:-1] + 2)          # an '=' run of >= 2 between neighbouring events
            p, t, l, a, ref_span = p[keep], t[keep], l[keep], a[keep], ref_span[keep]
            if len(p) > 1:                                                # second pass: dropping can only widen gaps, but re-check chained overlaps
                ok = np.ones(len(p), bool); ok[1:] = p[1:] >= p[:-1] + ref_span[:-1] + 2
                p, t, l, a, ref_span = p[ok], t[ok], l[ok], a[ok], ref_span[ok]
            k = len(p)
            eq = np.empty(k + 1, np.int64)
            eq[0] = (p[0] - beg) if k else L
            if k: eq[1:k] = p[1:] - (p[:-1] + ref_span[:-1]); eq[k] = beg + L - (p[-1] + ref_span[-1])
            ops = np.empty(2 * k + 1, np.uint32); ops[0::2] = (eq.astype(np.uint32) << 4) | EQ
            if k: ops[1::2] = (l.astype(np.uint32) << 4) | t.astype(np.uint32)
            clip = int(rng.integers(100, 2000)) if rng.random() < 0.03 else 0
            if clip: ops = np.concatenate([np.array([(clip << 4) | S], np.uint32), ops])
            q_adv = np.where(t == D, 0, l)                                 # read bases an event consumes
            qlen = clip + int(eq.sum()) + int(q_adv.sum())
            bases = rng.integers(0, 4, qlen + (qlen & 1)).astype(np.uint8)
            if k:                                                          # planted alt bases at the events' read offsets
                qi = clip + np.cumsum(eq[:k]) + np.concatenate([[0], np.cumsum(q_adv[:-1])])
                for j in np.nonzero(t != D)[0]: bases[qi[j]:qi[j] + l[j]] = a[j, :l[j]]
            packed = (code[bases[0::2]] << 4) | code[bases[1::2]]
            qv = rng.choice(np.array([93, 60, 40, 30, 20, 8], np.uint8), size=qlen, p=[0.6, 0.12, 0.1, 0.08, 0.06, 0.04]) if not ont else \
                rng.integers(5, 40, qlen).astype(np.uint8)
            cig_off.append(c_top); n_cig.append(len(ops)); cig.append(ops); c_top += len(ops)
            lq.append(qlen); seq_off.append(s_top); seqs.append(packed); s_top += len(packed)
            qual_off.append(q_top); quals.append(qv); q_top += qlen
        out.append(dict(n_reads=n_reads, min_bq=10, noisy_reg_max_xgaps=5, noisy_reg_slide_win=25 if ont else 100, end_clip_reg=30,
                        end_clip_reg_flank_win=100, max_noisy_frac_per_read=0.5, max_var_ratio_per_read=0.05, whole_ref_len=250000000,
                        reg_beg=reg_beg, reg_end=reg_end, ordered_read_ids=np.arange(n_reads, dtype=np.int32), is_skipped=np.zeros(n_reads, np.uint8),
                        read_pos0=(starts - 1).astype(np.int64), read_is_rev=rng.integers(0, 2, n_reads).astype(np.uint8), is_palindrome=np.zeros(n_reads, np.uint8),
                        n_cigar=np.array(n_cig, np.int32), cigar_off=np.array(cig_off, np.int64), cigar=np.concatenate(cig + [np.zeros(1, np.uint32)]),
                        l_qseq=np.array(lq, np.int32), seq_off=np.array(seq_off, np.int64), bseq=np.concatenate(seqs + [np.zeros(1, np.uint8)]),
                        qual_off=np.array(qual_off, np.int64), qual=np.concatenate(quals + [np.zeros(1, np.uint8)])))
    return out


def sites_from_digar_output(d, o, min_sv_len=50):
    """Workload preparation (the reference's step 1.2, collect_all_cand_var_sites src/collect_var.c:1209, restated with numpy for
    the bench set-up; large insertions are matched exactly here): the sorted unique candidate sites of a chunk from K1's output
    `o` -- every non-low-quality X / I / D record of a kept read that starts inside [reg_beg, reg_end]."""
    nr = d["n_reads"]; nd = int(o["n_digar_total"])
    t = o["digar_type"][:nd].astype(np.int32); pos = o["digar_pos"][:nd]; ln = o["digar_len"][:nd]; aoff = o["digar_alt_off"][:nd]
    read_of = np.repeat(np.arange(nr), o["n_digar"][:nr])
    dropped = (np.asarray(d["is_skipped"][:nr]) != 0) | (o["skip"][:nr] != 0)
    ok = ((t == 8) | (t == 1) | (t == 2)) & (o["digar_low_qual"][:nd] == 0) & (pos >= d["reg_beg"]) & (pos <= d["reg_end"]) & ~dropped[read_of]
    idx = np.nonzero(ok)[0]
    alt = o["digar_alt"]
    keys = {}
    for k in idx.tolist():
        tk, lk = int(t[k]), int(ln[k])
        a = bytes(alt[aoff[k]:aoff[k] + lk]) if tk != 2 else b""
        keys.setdefault((int(pos[k]) - (0 if tk == 8 else 1), tk, lk if tk == 2 else (0 if tk == 1 else 1), 0 if tk == 2 else lk, a), int(pos[k]))
    order = sorted(keys)
    site_alt, site_alt_off = [], []
    for key in order:
        site_alt_off.append(len(site_alt)); site_alt.extend(key[4])
    return dict(n_sites=len(order), min_sv_len=min_sv_len,
                site_pos=np.array([keys[k] for k in order] + [0], np.int64), site_type=np.array([k[1] for k in order] + [0], np.int32),
                site_ref_len=np.array([k[2] for k in order] + [0], np.int32), site_alt_len=np.array([k[3] for k in order] + [0], np.int32),
                site_alt_off=np.array(site_alt_off + [0], np.int64), site_alt=np.array(site_alt + [0], np.uint8))


def site_list_from_sites(o, st, min_sv_len=50, src_is_offset=False):
    """lcd_site_list_t arrays of a chunk from the candidate-site list `st` (K1b's output: site_pos / type / ref_len / alt_len / site_src) and
    K1's output `o`: a site's alt bases are those of the record it stands for (site_src = record index, or the offset of the bases in
    digar_alt when src_is_offset)."""
    n = int(st["n_sites"])
    src = np.asarray(st["site_src"][:n], np.int64); typ = np.asarray(st["site_type"][:n], np.int32)
    al = np.where(typ == 2, 0, np.asarray(st["site_alt_len"][:n], np.int64))
    a0 = src if src_is_offset else np.asarray(o["digar_alt_off"], np.int64)[src] if n else src
    off = np.zeros(n + 1, np.int64); np.cumsum(al, out=off[1:])
    idx = np.repeat(a0 - off[:-1], al) + np.arange(int(off[-1]), dtype=np.int64)
    return dict(n_sites=n, min_sv_len=min_sv_len, site_pos=np.append(np.asarray(st["site_pos"][:n], np.int64), 0), site_type=np.append(typ, 0).astype(np.int32),
                site_ref_len=np.append(np.asarray(st["site_ref_len"][:n], np.int32), 0).astype(np.int32), site_alt_len=np.append(np.asarray(st["site_alt_len"][:n], np.int32), 0).astype(np.int32),
                site_alt_off=off.copy(), site_alt=np.append(np.asarray(o["digar_alt"], np.uint8)[idx], 0).astype(np.uint8))


def empty_site_list(min_sv_len=50):
    return dict(n_sites=0, min_sv_len=min_sv_len, site_pos=np.zeros(1, np.int64), site_type=np.zeros(1, np.int32), site_ref_len=np.zeros(1, np.int32),
                site_alt_len=np.zeros(1, np.int32), site_alt_off=np.zeros(1, np.int64), site_alt=np.zeros(1, np.uint8))


def pileup_input_from_digar(d, o, sites):
    """lcd_pileup_input_t (+ lcd_profile_extra_t when `sites` carries var_cate) for a chunk from K1's host-side output."""
    nr = d["n_reads"]
    p = dict(n_reads=nr, n_sites=sites["n_sites"], min_bq=d["min_bq"], min_sv_len=sites["min_sv_len"], ordered_read_ids=d["ordered_read_ids"],
             is_skipped=np.maximum(np.asarray(d["is_skipped"][:nr]), o["skip"][:nr]), read_beg=o["read_beg"], read_end=o["read_end"], read_is_rev=d["read_is_rev"],
             digar_first=o["digar_first"], n_digar=o["n_digar"], qual_off=d["qual_off"], qual=d["qual"], digar_pos=o["digar_pos"], digar_type=o["digar_type"],
             digar_len=o["digar_len"], digar_qi=o["digar_qi"], digar_low_qual=o["digar_low_qual"], digar_alt_off=o["digar_alt_off"], digar_alt=o["digar_alt"],
             **{k: sites[k] for k in ("site_pos", "site_type", "site_ref_len", "site_alt_len", "site_alt_off", "site_alt")})
    if "var_cate" in sites:
        p.update(var_cate=sites["var_cate"], nreg_first=o["nreg_first"], n_nreg=o["n_nreg"], nreg_beg=o["nreg_beg"], nreg_end=o["nreg_end"])
    return p


def make_classify_chunk(rng, ref_len=4000, n_sites=400, max_xgaps=5, ref0=100000, is_ont=0):
    """Candidate sites with K2-style counters on a reference window with planted homopolymers / short tandem repeats (and a few N /
    lower-case bases): the input of the per-site category (classify_var_cate, src/collect_var.c:413).  Small indels often repeat the
    reference's own bases, so that both context tests (var_is_homopolymer, var_is_repeat_region) fire."""
    ref = rng.integers(0, 4, ref_len).astype(np.uint8)
    p = 40
    while p < ref_len - 80:
        kind = rng.integers(0, 3)
        if kind == 0:
            n = int(rng.integers(3, 13)); ref[p:p + n] = rng.integers(0, 4)
        elif kind == 1:
            u = int(rng.integers(2, 7)); c = int(rng.integers(3, 9)); unit = rng.integers(0, 4, u).astype(np.uint8)
            ref[p:p + u * c] = np.tile(unit, c)[:max(0, min(u * c, ref_len - p))]
        p += int(rng.integers(15, 120))
    letters = np.frombuffer(b"ACGT", np.uint8)[ref].copy()
    for q in rng.integers(0, ref_len, 6): letters[q] = ord("N")
    for q in rng.integers(0, ref_len, 12): letters[q] = letters[q] | 0x20                     # lower case
    pos = np.sort(rng.choice(np.arange(ref0 + 40, ref0 + ref_len - 60), size=n_sites, replace=False)).astype(np.int64)
    typ = rng.choice(np.array([8, 1, 2], np.int32), size=n_sites, p=[0.5, 0.25, 0.25])
    ln = np.where(rng.random(n_sites) < 0.85, rng.integers(1, max_xgaps + 1, n_sites), rng.integers(max_xgaps + 1, 12, n_sites)).astype(np.int32)
    ref_l = np.where(typ == 1, 0, np.where(typ == 2, ln, 1)).astype(np.int32); alt_l = np.where(typ == 2, 0, np.where(typ == 1, ln, 1)).astype(np.int32)
    alt, off = [], []
    for i in range(n_sites):
        off.append(len(alt))
        if typ[i] == 2: continue
        k = int(alt_l[i]); q = int(pos[i] - ref0)
        alt.extend((ref[q:q + k] if (typ[i] == 1 and rng.random() < 0.6) else rng.integers(0, 4, k)).tolist())
    total = rng.integers(0, 60, n_sites).astype(np.int32); low = rng.integers(0, 12, n_sites).astype(np.int32)
    frac = rng.choice(np.array([0.0, 0.1, 0.19, 0.2, 0.5, 0.8, 0.81, 1.0]), size=n_sites)
    altc = np.minimum(total, np.round(total * frac)).astype(np.int32)
    counts = np.zeros((n_sites + 1, 8), np.int32)
    counts[:n_sites, 0] = total; counts[:n_sites, 1] = low; counts[:n_sites, 2] = total - altc; counts[:n_sites, 3] = altc
    counts[:n_sites, 5] = altc // 2; counts[:n_sites, 7] = altc - altc // 2
    if is_ont:       # strand-skewed alternative-allele counts (ONT's strand-bias test, var_is_strand_bias src/collect_var.c:270): balanced, mildly and fully skewed
        skew = rng.choice(np.array([0.5, 0.5, 0.35, 0.2, 0.1, 0.0, 1.0]), size=n_sites)
        fwd = np.minimum(altc, np.round(altc * skew)).astype(np.int32)
        counts[:n_sites, 5] = fwd; counts[:n_sites, 7] = altc - fwd
        refc = total - altc; counts[:n_sites, 4] = refc // 2; counts[:n_sites, 6] = refc - refc // 2
    return dict(n_sites=n_sites, min_dp=5, min_alt_dp=2, max_xgaps=max_xgaps, is_ont=int(is_ont), min_af=0.20, max_af=0.80, ref_beg=ref0, ref_end=ref0 + ref_len - 1,
                ref_seq=letters, site_pos=np.append(pos, 0), site_type=np.append(typ, 0).astype(np.int32), site_ref_len=np.append(ref_l, 0).astype(np.int32),
                site_alt_len=np.append(alt_l, 0).astype(np.int32), site_alt_off=np.array(off + [0], np.int64), site_alt=np.array(alt + [0], np.uint8), site_counts=counts)


def classify_input_from_sites(d, sites, counts, seed, flank=64, max_xgaps=5, is_ont=0):
    """lcd_classify_input_t of a bench chunk: its candidate sites and K2's counters on a synthetic reference window (uniform ACGT with a
    homopolymer or short tandem repeat every ~200 bases; the bench's reads are CIGARs without a reference of their own)."""
    n = int(sites["n_sites"]); pos = np.asarray(sites["site_pos"][:n], np.int64)
    lo = int(min(int(d["reg_beg"]), int(pos.min()) if n else int(d["reg_beg"]))) - flank
    hi = int(max(int(d["reg_end"]), int((pos + np.asarray(sites["site_ref_len"][:n])).max()) if n else int(d["reg_end"]))) + flank
    rng = np.random.default_rng(seed)
    L = hi - lo + 1
    ref = rng.integers(0, 4, L).astype(np.uint8)
    starts = np.arange(100, L - 100, 200) + rng.integers(0, 60, len(np.arange(100, L - 100, 200)))
    unit_len = rng.integers(1, 7, len(starts)); copies = rng.integers(3, 9, len(starts))
    for s0, u, c in zip(starts.tolist(), unit_len.tolist(), copies.tolist()):
        ref[s0:s0 + u * c] = np.tile(ref[s0:s0 + u], c)
    return dict(n_sites=n, min_dp=5, min_alt_dp=2, max_xgaps=max_xgaps, is_ont=int(is_ont), min_af=0.20, max_af=0.80, ref_beg=lo, ref_end=hi,
                ref_seq=np.frombuffer(b"ACGT", np.uint8)[ref].copy(), site_pos=sites["site_pos"], site_type=sites["site_type"], site_ref_len=sites["site_ref_len"],
                site_alt_len=sites["site_alt_len"], site_alt_off=sites["site_alt_off"], site_alt=sites["site_alt"],
                site_counts=np.ascontiguousarray(np.vstack([counts[:n], np.zeros((1, 8), np.int32)]), dtype=np.int32))


def noisyreg_input_from(d, o, sites, cate, seed, low_every=250, is_ont=0):
    """lcd_noisyreg_input_t of a chunk: K1's host-side output `o`, the candidate sites `sites` (K1b) with K2b's categories `cate`, and
    low-complexity intervals at sdust-like density (one 8 - 40 bp interval per ~250 bp; the bench's reads are CIGARs without a reference of
    their own, so the intervals are drawn, a third of them on candidate sites / noisy intervals as homopolymer runs are in real data)."""
    rng = np.random.default_rng(seed)
    n = int(sites["n_sites"]); nr = int(d["n_reads"]); ncn = int(o["n_cnreg"])
    span = int(d["reg_end"]) - int(d["reg_beg"])
    k = max(1, span // low_every)
    lb = rng.integers(int(d["reg_beg"]) - 200, int(d["reg_end"]) + 200, k)
    if n: lb[::3] = np.asarray(sites["site_pos"][:n])[rng.integers(0, n, len(lb[::3]))] - rng.integers(0, 6, len(lb[::3]))
    if ncn: lb[1::7] = np.asarray(o["cnreg_beg"][:ncn])[rng.integers(0, ncn, len(lb[1::7]))] - rng.integers(-3, 12, len(lb[1::7]))
    lb = np.sort(lb).astype(np.int64); le = lb + rng.integers(8, 40, k)
    return dict(reg_beg=int(d["reg_beg"]), reg_end=int(d["reg_end"]), min_alt_dp=2, noisy_reg_flank_len=10, is_ont=int(is_ont), min_af=0.20, n_sites=n, n_reads=nr,
                site_pos=sites["site_pos"], site_type=sites["site_type"], site_ref_len=sites["site_ref_len"], var_cate=np.append(np.asarray(cate[:n], np.int32), 0).astype(np.int32),
                n_cnreg=ncn, cnreg_beg=np.asarray(o["cnreg_beg"][:ncn], np.int64), cnreg_end=np.asarray(o["cnreg_end"][:ncn], np.int64), cnreg_label=np.asarray(o["cnreg_label"][:ncn], np.int32),
                n_low=k, low_beg=lb, low_end=le, is_skipped=np.maximum(np.asarray(d["is_skipped"][:nr]), np.asarray(o["skip"][:nr])),
                **{f: o[f] for f in ("read_beg", "read_end", "digar_first", "n_digar", "digar_pos", "digar_type", "digar_len", "nreg_first", "n_nreg", "nreg_beg", "nreg_end")})


def kept_site_list(sites, keep, cate):
    """The compacted candidate list classify_cand_vars leaves (chunk->cand_vars + var_i_to_cate): the sites K2c kept, with its categories."""
    n = int(sites["n_sites"])
    idx = np.nonzero(np.asarray(keep[:n]))[0]
    alt_len = np.asarray(sites["site_alt_len"])[idx]; off = np.asarray(sites["site_alt_off"])[idx]; t = np.asarray(sites["site_type"])[idx]
    al = np.where(t == 2, 0, alt_len).astype(np.int64)
    new_off = np.zeros(len(idx) + 1, np.int64); np.cumsum(al, out=new_off[1:])
    src = np.repeat(off - new_off[:-1], al) + np.arange(int(new_off[-1]), dtype=np.int64)
    pad = lambda a, dt: np.concatenate([np.asarray(a), np.zeros(1, dt)]).astype(dt)
    return dict(n_sites=len(idx), min_sv_len=sites["min_sv_len"], site_pos=pad(np.asarray(sites["site_pos"])[idx], np.int64), site_type=pad(t, np.int32),
                site_ref_len=pad(np.asarray(sites["site_ref_len"])[idx], np.int32), site_alt_len=pad(alt_len, np.int32), site_alt_off=new_off.copy(),
                site_alt=pad(np.asarray(sites["site_alt"], np.uint8)[src], np.uint8), var_cate=pad(np.asarray(cate[:n])[idx], np.int32))
