"""CPU-only checks of the product library: it loads, exports every symbol the
public header declares, validates arguments, and its geometry matches the
oracle's restatement of state.c:424-671.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import support as S
import theora_b200 as T
from theora_b200 import abi


def header_symbols():
    txt = open(os.path.join(S.ROOT, "include", "theora_b200.h")).read()
    return sorted(set(re.findall(r"OCG_API[^;(]*?\b(ocg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = abi.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libtheora_b200.so does not export %s" % s
    assert set(syms) == set(abi.EXPORTED_SYMBOLS), set(syms) ^ set(abi.EXPORTED_SYMBOLS)


def test_version_string():
    assert b"sm_100a" in abi.lib().ocg_version()


@pytest.mark.parametrize("dims", [(16, 16, 0), (64, 64, 0), (352, 288, 0), (1920, 1088, 0), (3840, 2160, 0),
                                  (640, 480, 2), (320, 240, 3)])
def test_geometry_matches_oracle(dims):
    fw, fh, fmt = dims
    for nrefs in (3, 6):
        g = T.geometry(fw, fh, fmt, nrefs)
        go = S.make_geometry(fw, fh, fmt, nrefs)
        assert bytes(g) == bytes(go)
        a = T.frag_buf_offs(g)
        b = np.empty(go.nfrags, np.int32)
        S.oracle().oco_geometry_frag_buf_offs(C.byref(go), S.ptr(b, S.i32p))
        assert np.array_equal(a, b)


def test_geometry_1080p_numbers():
    """48 960 fragments, 3 279 360-byte buffers."""
    g = T.geometry(1920, 1088)
    assert g.nfrags == 48960
    assert g.ref_frame_sz == 1952 * 1120 + 2 * (976 * 560) + 16
    assert g.planes[0].ystride == -1952 and g.planes[1].ystride == -976


@pytest.mark.parametrize("args", [(0, 16, 0, 3), (24, 16, 0, 3), (16, 16, 1, 3), (16, 16, 0, 2), (16, 16, 0, 7)])
def test_geometry_rejects_bad_arguments(args):
    g = abi.Geometry()
    assert abi.lib().ocg_geometry_init(C.byref(g), *args) == -10  # OCG_EINVAL == TH_EINVAL


def test_null_arguments_fail_cleanly():
    L = abi.lib()
    assert L.ocg_geometry_init(None, 16, 16, 0, 3) == -1
    assert L.ocg_ctx_sync(None) == -1
    assert L.ocg_dec_submit(None, None, None) == -1
    h = C.c_void_p()
    assert L.ocg_ctx_create(C.byref(h), None, 0) == -1


def test_no_cpu_fallback_without_device():
    """On a box without a GPU, creating a context must fail loudly."""
    L = abi.lib()
    if L.ocg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    g = T.geometry(64, 64)
    with pytest.raises(T.OcgError):
        T.Context(g)
    assert b"no CUDA device" in L.ocg_last_error() or b"CUDA" in L.ocg_last_error()
