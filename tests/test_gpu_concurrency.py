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


def test_worker_threads_follow_the_library_device():
    """A host thread other than the one that called lcd_gpu_init starts on CUDA device 0; the entry points bind it to the library's
    device.  Run in a child process that initialises the library on the LAST device of the box (device 0 on a one-GPU box)."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, threading
        sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
        import numpy as np, torch
        import longcalld_b200 as lcd
        from longcalld_b200 import synth
        dev = torch.cuda.device_count() - 1
        torch.cuda.set_device(dev); lcd.init(dev, 1 << 30)
        rng = np.random.default_rng(7)
        chunks = [(synth.make_phase_chunk(rng, 300, 200), synth.CATE_CLEAN, 0)]
        want = lcd.phase_batch(chunks); got = {}
        def work():
            lcd.set_thread_stream(lcd.aux_stream()); got["r"] = lcd.phase_batch(chunks)
        t = threading.Thread(target=work); t.start(); t.join()
        assert all(np.array_equal(got["r"][0][k], want[0][k]) for k in want[0])
        print("ok", dev)
    """ % (T.ROOT, T.ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
