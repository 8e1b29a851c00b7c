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
