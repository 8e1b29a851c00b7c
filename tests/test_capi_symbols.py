"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/lcd_gpu.h declares,
and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import longcalld_b200 as lcd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "lcd_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lcd_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(lcd.lib_path()):
        lcd.build_library()
    L = C.CDLL(lcd.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert L.lcd_gpu_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lcd.LcdGpuError):
        lcd.init(0, 0)
    import numpy as np
    a = np.zeros(10, dtype=np.uint8)
    with pytest.raises(lcd.LcdGpuError):
        lcd.wfa_batch([(a, a)], lcd.wfa_params())


def test_product_does_not_import_oracle():
    """The product path (package + csrc) must never reference oracle/."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "longcalld_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")) or fn == "Makefile":
                s = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"oracle/|lcd_oracle|liblcd_oracle|libref_shim|lcd_testlib", s):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
