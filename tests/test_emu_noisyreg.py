"""CPU: the product's K2c device logic (longcalld_b200/csrc/noisyreg_device.cuh: the chunk's noisy-region set, one CTA per chunk) compiled for the
host with a one-thread CTA (tests/emu) against the oracle: kept sites, working categories and the final region list."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lcd_testlib as T
from test_oracle_noisyreg import noisyreg_cases, noisyreg_fixture_cases

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libnoisyreg_emu.so"])
    return C.CDLL(os.path.join(EMU_DIR, "libnoisyreg_emu.so"))


def test_emu_vs_oracle(emu, oracle):
    n = 0
    for case, _ in noisyreg_cases(oracle):
        for variant in (case, dict(case, n_low=0), dict(case, n_cnreg=0), dict(case, n_sites=0), dict(case, min_alt_dp=1, noisy_reg_flank_len=0)):
            a = T.noisy_regs(oracle, "lcd_oracle_noisy_regs", variant)
            b = T.noisy_regs(emu, "emu_noisy_regs", variant)
            assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]), n
        n += 1
    assert n >= 20


def fabricated_case(rng, n_sites=160, span=900, n_reads=14):
    """Sites packed into a short window in collect_all_cand_var_sites' order (same and neighbouring anchors, long deletions over many sites), arbitrary
    categories, reads whose records tile their span: stresses the site-overlap test, the flank sweep and the containment rule."""
    base = 200000
    pos = base + rng.integers(0, span, n_sites); typ = rng.choice(np.array([8, 1, 2]), n_sites, p=[0.4, 0.3, 0.3])
    rl = np.where(typ == 1, 0, np.where(typ == 2, np.where(rng.random(n_sites) < 0.9, rng.integers(1, 6, n_sites), rng.integers(20, 200, n_sites)), 1))
    anchor = np.where(typ == 8, pos, pos - 1)
    o = np.lexsort((rl, typ, anchor)); pos, typ, rl = pos[o], typ[o], rl[o]
    keep = np.ones(n_sites, bool); keep[1:] = (pos[1:] != pos[:-1]) | (typ[1:] != typ[:-1]) | (rl[1:] != rl[:-1])
    pos, typ, rl = pos[keep], typ[keep], rl[keep]; n = len(pos)
    cate = rng.choice(np.array([0x800, 0x001, 0x002, 0x400, 0x004, 0x008, 0x010, 0x080]), n, p=[0.05, 0.2, 0.05, 0.1, 0.25, 0.15, 0.1, 0.1]).astype(np.int32)
    rb, re_, df, nd, dp, dt, dl, nf, nn, nb, ne = [], [], [], [], [], [], [], [], [], [], []
    for r in range(n_reads):
        b = base - 300 + int(rng.integers(0, 500)); p = b; first = len(dp)
        stop = max(b + 60, base + span + int(rng.integers(-200, 300)))
        while p < stop:
            run = int(rng.integers(1, 40)); dp.append(p); dt.append(7); dl.append(run); p += run
            t = int(rng.choice([8, 1, 2])); ln = int(rng.integers(1, 5)) if rng.random() < 0.9 else int(rng.integers(20, 120))
            dp.append(p); dt.append(t); dl.append(ln); p += 0 if t == 1 else ln
        rb.append(b); re_.append(p - 1); df.append(first); nd.append(len(dp) - first)
        k = int(rng.integers(0, 3)); nf.append(len(nb)); nn.append(k)
        st = np.sort(rng.integers(b, p, k))
        for s_ in st.tolist(): nb.append(s_); ne.append(s_ + int(rng.integers(5, 80)))
    ncn = int(rng.integers(0, 12)); cb = base - 100 + rng.integers(0, span + 200, ncn); ce = cb + rng.integers(3, 120, ncn); cl = rng.integers(1, 60, ncn)
    nl = int(rng.integers(0, 30)); lb = np.sort(base - 50 + rng.integers(0, span + 100, nl)); le = lb + rng.integers(3, 40, nl)
    return dict(reg_beg=base + int(rng.integers(0, 50)), reg_end=base + span - int(rng.integers(0, 50)), min_alt_dp=int(rng.integers(1, 4)), noisy_reg_flank_len=int(rng.choice([0, 10, 25])),
                is_ont=int(rng.integers(0, 2)), min_af=float(rng.choice([0.1, 0.2, 0.5])), n_sites=n, n_reads=n_reads, site_pos=np.append(pos, 0), site_type=np.append(typ, 0),
                site_ref_len=np.append(rl, 0), var_cate=np.append(cate, 0), n_cnreg=ncn, cnreg_beg=cb, cnreg_end=ce, cnreg_label=cl, n_low=nl, low_beg=lb, low_end=le,
                is_skipped=(rng.random(n_reads) < 0.1).astype(np.uint8), read_beg=np.array(rb), read_end=np.array(re_), digar_first=np.array(df), n_digar=np.array(nd),
                digar_pos=np.array(dp + [0]), digar_type=np.array(dt + [0]), digar_len=np.array(dl + [0]), nreg_first=np.array(nf), n_nreg=np.array(nn),
                nreg_beg=np.array(nb + [0]), nreg_end=np.array(ne + [0]))


def test_emu_vs_oracle_fabricated(emu, oracle):
    rng = np.random.default_rng(86)
    adds = 0
    for n in range(900):
        case = fabricated_case(rng, n_sites=int(rng.integers(2, 220)), span=int(rng.choice([60, 300, 900])))
        a = T.noisy_regs(oracle, "lcd_oracle_noisy_regs", case)
        b = T.noisy_regs(emu, "emu_noisy_regs", case)
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]), n
        adds += len(a[1])
    assert adds > 300


def test_emu_vs_fixtures(emu):
    for n, (case, kept, regs) in enumerate(noisyreg_fixture_cases()):
        got = T.noisy_regs(emu, "emu_noisy_regs", case)
        assert got[0] == kept and got[1] == regs, n


def test_emu_chained_to_sdust(emu, oracle):
    """K2c chained to a K0 plan (lcd_noisyreg_plan_create_on_sdust): the number of low-complexity intervals and K0's status are read through pointers on the
    device; the answers are those of the plain form, and a window K0 failed on fails the chunk (status -7) instead of being read"""
    rng = np.random.default_rng(87)
    for n in range(120):
        case = fabricated_case(rng, n_sites=int(rng.integers(2, 220)), span=int(rng.choice([60, 300, 900])))
        a = T.noisy_regs(oracle, "lcd_oracle_noisy_regs", case)
        b = T.noisy_regs(emu, "emu_noisy_regs_chained", case)
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]), n
    with pytest.raises(AssertionError, match="-7"):
        T.noisy_regs(emu, "emu_noisy_regs_k0_failed", case)
