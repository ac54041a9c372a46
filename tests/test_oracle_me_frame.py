"""Pins the oracle's whole-frame motion analysis (oco_me_frame: history rotation, candidate sets from
neighbours, thresholds, search, refinements in the analysis loop's order) against the REAL reference:
oc_mcenc_search / oc_mcenc_refine1mv / oc_mcenc_refine4mv run by oracle/_ref over every macro block of
the same frames inside a th_encode_alloc context.  Also checks the committed golden vector."""
import ctypes as C
import os

import numpy as np
import pytest

import megen
import support as S
from theora_b200 import abi

P, F4, NS, FA, DR = (abi.OCG_ME_REFINE_PREV, abi.OCG_ME_REFINE_4MV, abi.OCG_ME_NOSATD, abi.OCG_ME_FAST,
                     abi.OCG_ME_DROPPED)
GOLDEN = os.path.join(S.GOLDEN_DIR, "me_frame.npz")
GOLDEN_CASE = dict(fw=176, fh=144, nframes=3, motion=(3, 1), flags=[P | F4, P | F4, P | F4 | DR], density=0.3, seed=77)


def oracle_me(g, topo, state, frames, flags, mask):
    O = S.oracle()
    O.oco_me_frame.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    base = g.base_off
    p = [f.ctypes.data + base for f in frames]  # IO, PREV_ORIG, GOLD_ORIG, PREV, GOLD
    O.oco_me_frame(p[0], p[2], p[1], p[4], p[3], g.planes[0].ystride, topo.ctypes.data, state.ctypes.data, len(state),
                   flags, mask.ctypes.data if mask is not None else None)


def run_sequence(fw, fh, nframes, motion, flags, density, seed, ref=None):
    """Yields (t, flags, oracle_state_copy, reference_state_or_None)."""
    rng = np.random.default_rng(seed)
    g = S.make_geometry(fw, fh, 0, 6)
    orig, recon = megen.scene_buffers(g, rng, nframes + 1, motion=motion)
    L = abi.lib()
    n = L.ocg_me_nmbs(C.byref(g))
    topo = np.zeros(n, abi.ME_TOPO_DTYPE)
    assert L.ocg_me_topology(C.byref(g), topo.ctypes.data) == 0
    state = np.zeros(n, abi.ME_MB_DTYPE)
    h = ref.refh_me_open(fw, fh, 0) if ref is not None else None
    want = np.zeros(n, abi.ME_MB_DTYPE)
    try:
        for t in range(1, nframes + 1):
            gold_t = 0 if t < 3 else 1
            frames = [orig[t], orig[t - 1], orig[gold_t], recon[t - 1], recon[gold_t]]
            mask = (rng.random(n) < density).astype(np.uint8) if density > 0 else None
            fl = flags[t - 1]
            oracle_me(g, topo, state, frames, fl, mask)
            if h is not None:
                ptrs = (C.c_void_p * 5)(*[f.ctypes.data for f in frames])
                ref.refh_me_frame(h, ptrs, fl, mask.ctypes.data if mask is not None else None, want.ctypes.data)
            yield t, fl, topo, state.copy(), (want.copy() if h is not None else None)
    finally:
        if h is not None:
            ref.refh_me_close(h)


CASES = [
    (64, 64, 3, (3, 1), [P | F4] * 3, 0.0, 1),
    (176, 144, 4, (2, -1), [0, P | F4, P | F4, P | F4 | DR], 0.3, 2),
    (352, 288, 3, (5, 2), [P | F4, P, P | F4], 1.0, 3),
    (208, 112, 3, (-3, 1), [P | F4 | FA] * 3, 0.2, 4),
    (320, 240, 3, (1, 3), [P | NS | FA] * 3, 0.5, 5),
]


@pytest.mark.skipif(not S.ref_available("c"), reason="needs oracle/_ref")
@pytest.mark.parametrize("case", CASES)
def test_oracle_frame_analysis_matches_reference(case):
    R = megen.bind_ref_me(S.ref("c"))
    for t, fl, topo, got, want in run_sequence(*case, ref=R):
        megen.assert_me_equal(got, want, topo["valid"], fl, "frame %d" % t)


def test_oracle_matches_committed_golden():
    """tests/golden/me_frame.npz was produced by the reference itself (tests/golden/make_golden.py)."""
    G = np.load(GOLDEN)
    for t, fl, topo, got, _ in run_sequence(**GOLDEN_CASE):
        want = G["frame%d" % t].view(abi.ME_MB_DTYPE).reshape(-1)
        megen.assert_me_equal(got, want, topo["valid"], fl, "golden frame %d" % t)
