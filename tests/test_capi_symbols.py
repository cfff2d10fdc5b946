"""The C-ABI library loads on a CPU-only box and exports every symbol include/freud_b200.h declares; the product
path fails loudly (no fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "freud_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("fgpu_points_create", "fgpu_ball_query", "fgpu_knn_query", "fgpu_nlist_copy", "fgpu_rdf_accumulate",
                 "fgpu_rdf_read", "fgpu_rdf_allreduce", "fgpu_steinhardt_compute", "fgpu_comm_create"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from freud_b200 import _capi

    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [name for name in declared_symbols() if not hasattr(lib, name)]
    assert not missing, f"libfreud_b200.so lacks {missing}"
    # and the ctypes binding table covers the header exactly
    assert sorted(_capi.SIGNATURES) == declared_symbols()


def test_no_torch_and_no_oracle_in_the_product_library():
    """The shipped library links neither PyTorch nor anything under oracle/."""
    from freud_b200 import _capi

    with open(_capi.LIB_PATH, "rb") as f:
        blob = f.read()
    for needle in (b"libtorch", b"libc10", b"libfreud_ref", b"libfreud_port", b"fport_", b"fref_"):
        assert needle not in blob, needle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "freud_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports oracle"
                assert "oracle/" not in src or fn == "__init__.py", f"{fn} references oracle/"


def test_fails_loudly_without_a_gpu():
    from freud_b200 import _capi

    lib = _capi.lib()
    assert b"freud_b200" in lib.fgpu_version()
    if lib.fgpu_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(_capi.GpuError, match="no CPU fallback"):
        _capi.Context(0)
    # NULL handles are rejected with an error code, not a crash
    assert lib.fgpu_rdf_reset(None) != 0 and lib.fgpu_ctx_synchronize(None) != 0
    counts = np.zeros(4, np.uint32)
    assert lib.fgpu_rdf_read(None, counts.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))) != 0
