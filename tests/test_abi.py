"""The C-ABI library loads and exports every symbol include/pxb200.h declares (no compute: runs without a GPU)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "pxb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pxb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("pxb_ctx_create", "pxb_residual_matrix", "pxb_score_compound", "pxb_solve_minimal",
                 "pxb_pearl_datacost", "pxb_pearl_label", "pxb_find_homographies", "pxb_find_two_view_motions",
                 "pxb_find_6d_poses"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib_path = ROOT / "progressive-x_b200" / "libpxb200.so"
    assert lib_path.exists(), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(str(lib_path))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in pxb200.h but not exported: {missing}"


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the context must refuse to be created (and say why)."""
    from pyprogressivex import _native
    lib = _native.load_library()
    h = ctypes.c_void_p()
    rc = lib.pxb_ctx_create(0, ctypes.byref(h))
    if rc == 0:  # a GPU is present: nothing to assert here
        lib.pxb_ctx_destroy(h)
        return
    assert rc < 0 and b"no CPU fallback" in lib.pxb_last_error() or b"sm_" in lib.pxb_last_error()
