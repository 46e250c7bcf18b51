import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emu_lib(tmp_path_factory):
    """libbellman_b200_emu.so: the product sources compiled for host fibers (tests/native/), built
    once per session.  BB_EMU_LIB points at a pre-built variant (e.g. with sanitizers) instead."""
    lib = os.environ.get("BB_EMU_LIB")
    if lib:
        return lib
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tests", "native", "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib, launches = mod.build(str(tmp_path_factory.mktemp("bb_emu")))
    assert launches >= 25
    return lib
