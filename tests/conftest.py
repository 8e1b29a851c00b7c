import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import lcd_testlib as T
    return T.oracle_lib()


@pytest.fixture(scope="session")
def ref():
    """Unmodified reference behind oracle/_ref/libref_shim.so; tests that need it skip when absent."""
    import lcd_testlib as T
    lib = T.ref_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libref_shim.so not built (no /root/reference here)")
    return lib


@pytest.fixture(scope="session")
def gpu():
    import longcalld_b200 as lcd
    lcd.init(0, 0)
    yield lcd
