"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/wavesim.h declares, and refuses to
create a solver without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from wsharness import PRODUCT_SO, ROOT, make_desc

HEADER = os.path.join(ROOT, "include", "wavesim.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ws_[a-z_0-9]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(PRODUCT_SO):
        import __graft_entry__ as g
        g.build()
    return C.CDLL(PRODUCT_SO)


def test_all_declared_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for name in syms:
        assert hasattr(lib, name), name


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = make_desc(2, "acoustic", 32, 32, nt=4)
    h = C.c_void_p()
    rc = lib.ws_create(C.byref(d), C.byref(h))
    assert rc != 0
    lib.ws_last_error.restype = C.c_char_p
    assert b"no CUDA device" in lib.ws_last_error()


def test_estimate_memory(lib):
    lib.ws_estimate_memory.restype = C.c_size_t
    d = make_desc(3, "elastic", 1024, 1024, 1024, fd_order=8, damping=2, boundary_width=20, free_surface=1)
    b = lib.ws_estimate_memory(C.byref(d))
    # 9 wavefields + 11 model vectors on the padded 1088 x 1036 x 1036 box, plus CPML slabs
    assert 90e9 < b < 110e9
