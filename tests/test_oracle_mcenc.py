"""Pins the oracle's motion search (oco_mcenc_search_batch) against the real
oc_mcenc_search_frame of the compiled reference, through a minimal encoder
context (oracle/ref_internal_harness.c)."""
import ctypes as C

import numpy as np
import pytest

import mcgen as M
import support as S

pytestmark = pytest.mark.skipif(not S.ref_available("c"), reason="oracle/_ref not built")


@pytest.mark.parametrize("seed,shift,noise,smooth", [(1, (5, -3), 6, True), (2, (-12, 9), 10, True),
                                                      (3, (0, 0), 2, True), (4, (14, 15), 20, False),
                                                      (5, (-15, -15), 4, True)])
def test_search_matches_reference(seed, shift, noise, smooth):
    rng = np.random.default_rng(seed)
    src, rfull, rsatd, bl, ystride = M.make_scene(rng, shift=shift, noise=noise, smooth=smooth)
    n = 60
    mb_in, cases = M.make_cases(rng, n, ystride=ystride)
    out = np.zeros(n, M.MB_OUT)
    O = S.oracle()
    O.oco_mcenc_search_batch.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    O.oco_mcenc_search_batch(src.ctypes.data + bl, rfull.ctypes.data + bl, rsatd.ctypes.data + bl, ystride,
                             mb_in.ctypes.data, out.ctypes.data, n)
    R = S.ref("c")
    R.refh_mcenc_search_frame.argtypes = [C.c_void_p] * 3 + [C.c_int, C.POINTER(C.c_long), C.c_int, C.c_int, C.c_int,
                                                              C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int,
                                                              C.c_int, C.c_int, C.POINTER(C.c_int)]
    moved = 0
    for i, cs in enumerate(cases):
        ro = (C.c_int * 12)()
        offs = (C.c_long * 4)(*cs["offs"])
        nbm = (C.c_int * max(cs["ncn"], 1))(*(cs["nb_mvs"] or [0]))
        nbe = (C.c_int * max(cs["ncn"], 1))(*(cs["nb_err"] or [0]))
        R.refh_mcenc_search_frame(src.ctypes.data + bl, rfull.ctypes.data + bl, rsatd.ctypes.data + bl, ystride, offs,
                                  cs["frame"], cs["accum"], cs["ncn"], nbm, nbe, cs["mv1"], cs["mv2"], cs["own_err"],
                                  1, ro)
        o = out[i]
        assert M.mv_pack(2 * int(o["best_vec"][0]), 2 * int(o["best_vec"][1])) == ro[0], i
        assert int(o["error"]) == ro[1], i
        assert int(o["satd"]) == ro[2], i
        if cs["frame"] == 1:
            for b in range(4):
                assert M.mv_pack(2 * int(o["block_vec"][b][0]), 2 * int(o["block_vec"][b][1])) == ro[3 + b], (i, b)
                assert int(o["block_satd"][b]) == ro[7 + b], (i, b)
        moved += int(o["best_vec"][0] != 0 or o["best_vec"][1] != 0)
    if shift != (0, 0):
        assert moved > n // 2  # the search really found the displacement
