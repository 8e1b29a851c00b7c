"""GPU: the DP engines in two windows of the workspace pool (lcd_gpu_split_pool) with CTA slots reserved (lcd_gpu_reserve_sms): a POA
batch on the main thread / library stream while a WFA + edlib batch runs from a second host thread on its own stream and the phasing
kernel from a third on the auxiliary stream -- every result identical to the serial, single-window run."""
import threading

import numpy as np
import pytest

import lcd_testlib as T
from longcalld_b200 import synth

pytestmark = pytest.mark.gpu


def test_concurrent_engines_split_pool(gpu):
    import torch
    regions = synth.make_regions(1.0, "hifi", seed=5, with_reads=True)
    pairs = synth.wfa_problems(regions)
    problems = [[s for s, h in zip(r.reads, r.read_hap) if h == hap and len(s)] for r in regions for hap in (1, 2)]
    problems = [p for p in problems if p]
    epairs = synth.edlib_pairs(regions, seed=5, mbp=1.0)
    rng = np.random.default_rng(6)
    chunks = [(synth.make_phase_chunk(rng, 600, 400), mask, 0) for mask in (synth.CATE_CLEAN, synth.CATE_GERMLINE) for _ in range(4)]
    # serial, one window
    want_wfa = gpu.wfa_batch(pairs, gpu.wfa_params())
    want_poa = gpu.poa_batch(problems, gpu.poa_params())
    want_ed = gpu.edlib_batch(epairs, gpu.MODE_NW, 1)
    want_ph = gpu.phase_batch(chunks)
    free_b, _ = torch.cuda.mem_get_info()
    pool = min(free_b // 4, 32 << 30)               # what lcd_gpu_init(0, 0) reserved
    gpu.split_pool(pool // 2); gpu.reserve_sms(12)
    try:
        side = torch.cuda.Stream()
        got, err = {}, []

        def dp2():
            try:
                gpu.set_thread_stream(side.cuda_stream)
                for _ in range(3):
                    got["wfa"] = gpu.wfa_batch(pairs, gpu.wfa_params()); got["ed"] = gpu.edlib_batch(epairs, gpu.MODE_NW, 1)
            except Exception as e:
                err.append(e)

        def ph():
            try:
                gpu.set_thread_stream(gpu.aux_stream())
                for _ in range(3):
                    got["ph"] = gpu.phase_batch(chunks)
            except Exception as e:
                err.append(e)
        ts = [threading.Thread(target=dp2), threading.Thread(target=ph)]
        for t in ts: t.start()
        for _ in range(2):
            got["poa"] = gpu.poa_batch(problems, gpu.poa_params())
        for t in ts: t.join()
        assert not err, err
    finally:
        gpu.split_pool(0); gpu.reserve_sms(0)
    assert got["wfa"] == want_wfa and got["ed"] == want_ed
    for (st, cons, msa), (st0, cons0, msa0) in zip(got["poa"], want_poa):
        assert st == st0 == 0 and cons == cons0 and (msa == msa0).all()
    for g, w in zip(got["ph"], want_ph):
        assert all(np.array_equal(g[k], w[k]) for k in w)
    # and back to one window
    assert gpu.wfa_batch(pairs[:50], gpu.wfa_params()) == want_wfa[:50]
