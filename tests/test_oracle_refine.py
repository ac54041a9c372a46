"""Pins the oracle's half-pel refinement (oco_mcenc_refine_batch) against the
real oc_mcenc_refine1mv / oc_mcenc_refine4mv of the compiled reference
(oracle/ref_internal_harness.c builds the minimal encoder context)."""
import ctypes as C

import numpy as np
import pytest

import mcgen as M
import support as S

pytestmark = pytest.mark.skipif(not S.ref_available("c"), reason="oracle/_ref not built")

OC_SP_LEVEL_NOSATD = 3  # encint.h: speed level from which the 1MV refinement uses SAD


def bind():
    O = S.oracle()
    O.oco_mcenc_refine_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    R = S.ref("c")
    R.refh_mcenc_refine.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_long), C.c_int, C.c_int, C.c_uint,
                                    C.POINTER(C.c_int), C.POINTER(C.c_uint), C.c_int, C.c_int, C.POINTER(C.c_int)]
    return O, R


@pytest.mark.parametrize("seed,shift,entry,use_sad", [(1, (5, -3), "mixed", False), (2, (-2, 7), "max", False),
                                                       (3, (0, 0), 0, False), (4, (9, 4), "mixed", True),
                                                       (5, (-6, -6), "max", True)])
def test_refine_matches_reference(seed, shift, entry, use_sad):
    rng = np.random.default_rng(seed)
    src, _, ref, bl, ystride = M.make_scene(rng, shift=shift, noise=5)
    n = 48
    mb = M.make_refine_cases(rng, n, ystride=ystride, entry=entry)
    out = np.zeros(n, M.REF_OUT)
    O, R = bind()
    O.oco_mcenc_refine_batch(src.ctypes.data + bl, ref.ctypes.data + bl, ystride, mb.ctypes.data, out.ctypes.data, n,
                             3 | (4 if use_sad else 0))
    changed = 0
    for i in range(n):
        m = mb[i]
        ro = (C.c_int * 10)()
        offs = (C.c_long * 4)(*[int(v) for v in m["frag_off"]])
        bmv = (C.c_int * 4)(*[M.mv_pack(2 * int(v[0]), 2 * int(v[1])) for v in m["block_vec"]])
        bsatd = (C.c_uint * 4)(*[int(v) for v in m["block_satd"]])
        R.refh_mcenc_refine(src.ctypes.data + bl, ref.ctypes.data + bl, ystride, offs, 1,
                            M.mv_pack(2 * int(m["vec"][0]), 2 * int(m["vec"][1])), int(m["satd"]), bmv, bsatd,
                            OC_SP_LEVEL_NOSATD if use_sad else 1, 1, ro)
        o = out[i]
        assert M.mv_pack(int(o["mv"][0]), int(o["mv"][1])) == ro[0], i
        assert int(o["satd"]) == ro[1], i
        for b in range(4):
            assert M.mv_pack(int(o["ref_mv"][b][0]), int(o["ref_mv"][b][1])) == ro[2 + b], (i, b)
            assert int(o["block_satd"][b]) == ro[6 + b], (i, b)
        changed += int(o["mv"][0] & 1 or o["mv"][1] & 1)
    if entry == "max":
        assert changed == n  # every macro block moved to a half-pel site
    if entry == 0:
        assert changed == 0  # nothing beats a zero entry score
