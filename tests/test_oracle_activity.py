"""Pins the oracle's oco_block_activity (and the OCG_MET_ACTIVITY batch form)
against the reference's file-static oc_mb_activity (analyze.c:1152), called for
real through oracle/_ref/libth_c_analyze.so (analyze.c compiled as part of the
harness translation unit)."""
import ctypes as C
import os

import numpy as np
import pytest

import support as S

LIBA = os.path.join(S.REF_DIR, "libth_c_analyze.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIBA), reason="oracle/_ref/libth_c_analyze.so not built")

W, H, PAD = 192, 96, 16


def frames(rng, kind):
    st = W + 2 * PAD
    shape = (H + 2 * PAD, st)
    if kind == "noise":
        a = rng.integers(0, 256, size=shape)
    elif kind == "flat":
        a = 100 + rng.integers(-2, 3, size=shape)
    elif kind == "edges":  # diagonal ramps/steps: one directional energy dominates -> the "edge block" branch
        yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
        a = np.zeros(shape, np.int64)
        for by in range(0, shape[0], 32):
            for bx in range(0, shape[1], 32):
                sgn = 1 if rng.integers(0, 2) else -1
                g = int(rng.integers(2, 4))
                d = (xx[by:by + 32, bx:bx + 32] - bx - 16) + sgn * (yy[by:by + 32, bx:bx + 32] - by - 16)
                if rng.integers(0, 2):
                    a[by:by + 32, bx:bx + 32] = 128 + g * d + rng.integers(-1, 2, size=d.shape)
                else:
                    a[by:by + 32, bx:bx + 32] = 128 + 20 * g * np.sign(d + 0.5) + rng.integers(-1, 2, size=d.shape)
    else:  # texture
        a = rng.integers(0, 256, size=(shape[0] // 2 + 1, shape[1] // 2 + 1))
        a = np.kron(a, np.ones((2, 2), np.int64))[:shape[0], :shape[1]] // 2 + rng.integers(0, 64, size=shape)
    return np.clip(a, 0, 255).astype(np.uint8), st


@pytest.mark.parametrize("kind", ["noise", "flat", "edges", "texture"])
def test_activity_matches_reference(kind):
    rng = np.random.default_rng(len(kind))
    a, st = frames(rng, kind)
    ystride = -st
    base = (PAD + H - 1) * st + PAD
    A = C.CDLL(LIBA)
    A.refh_mb_activity.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_uint)]
    A.refh_mb_activity.restype = C.c_uint
    O = S.oracle()
    O.oco_block_activity.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    O.oco_block_activity.restype = C.c_uint
    nedge = 0
    for my in range(0, H, 16):
        for mx in range(0, W, 16):
            offs = [(my + by) * ystride + mx + bx for by in (0, 8) for bx in (0, 8)]
            act = (C.c_uint * 4)()
            luma = A.refh_mb_activity(a.ctypes.data + base, ystride, (C.c_long * 4)(*offs), act)
            tot = 0
            for i, o in enumerate(offs):
                sm = C.c_int()
                v = O.oco_block_activity(a.ctypes.data + base + o, ystride, C.byref(sm))
                assert v == act[i], (kind, mx, my, i)
                tot += sm.value
                blk = a[PAD + H - 1 - (my + (i >> 1) * 8) - 7:PAD + H - (my + (i >> 1) * 8), PAD + mx + (i & 1) * 8:PAD + mx + (i & 1) * 8 + 8].astype(np.int64)
                raw = int((blk * blk).sum() * 64 - blk.sum() ** 2)
                nedge += int(raw >= (8 << 12) and v != raw)
            assert tot == luma
    if kind == "edges":
        assert nedge > 20  # the log/exp branch really ran
