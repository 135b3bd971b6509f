import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def emu_lib():
    """Host-emulation build of the kernel sources (test harness only, see zokrates_b200/csrc/rt.cuh)."""
    import __graft_entry__ as g
    from zokrates_b200._lib import Library
    return Library(g.build_emu())


@pytest.fixture(scope="session")
def oracle_c():
    import __graft_entry__ as g
    from tests.oracle_c import OracleC
    path = g.build_oracle()
    return OracleC(path)


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a real GPU.  Never falls back: a missing .so or device is a failure."""
    from zokrates_b200._lib import Library, DEFAULT_LIB
    assert os.path.exists(DEFAULT_LIB), "zokrates_b200/libzkb200.so missing: run __graft_entry__.build()"
    lib = Library(DEFAULT_LIB)
    assert lib.dll.zkb_device_count() > 0, "no CUDA device visible"
    return lib
