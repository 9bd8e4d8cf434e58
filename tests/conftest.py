import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "progressive-x_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def _have_gpu() -> bool:
    try:
        from pyprogressivex import _native
        lib = _native.load_library()
        import ctypes as C
        h = C.c_void_p()
        if lib.pxb_ctx_create(0, C.byref(h)) != 0:
            return False
        lib.pxb_ctx_destroy(h)
        return True
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    """A libpxb200 context on cuda:0. GPU tests FAIL (not skip) if the native library cannot run."""
    from pyprogressivex import _native
    c = _native.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O
