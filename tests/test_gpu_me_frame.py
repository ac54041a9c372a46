"""Whole-frame motion analysis on the device (ocg_me_frame: candidates from
already-searched neighbours, wave-front over super-block rows) against the REAL
reference: oc_mcenc_search / oc_mcenc_refine1mv / oc_mcenc_refine4mv run by
oracle/_ref over every macro block of the same frames in coding order, inside a
context made by th_encode_alloc.  Sequences of frames so the per-macro-block
history (analysis_mv[1..2], error) is exercised; bit-exact on every field."""
import ctypes as C

import numpy as np
import pytest

import megen
import support as S
import theora_b200 as T
from theora_b200 import abi

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not S.ref_available("c"), reason="needs oracle/_ref")]

P, F4, NS, FA, DR = (abi.OCG_ME_REFINE_PREV, abi.OCG_ME_REFINE_4MV, abi.OCG_ME_NOSATD, abi.OCG_ME_FAST,
                     abi.OCG_ME_DROPPED)
# (fw, fh, fmt, nframes, motion, per-frame flags, gold refine density)
CASES = [
    (64, 64, 0, 3, (3, 1), [P | F4] * 3, 0.0),
    (176, 144, 0, 4, (2, -1), [0, P | F4, P | F4, P | F4 | DR], 0.3),       # partial super blocks, keyframe-style first
    (352, 288, 0, 4, (5, 2), [P | F4, P, P | F4, P | F4], 1.0),              # every GOLD vector refined
    (208, 112, 2, 3, (-3, 1), [P | F4 | FA] * 3, 0.2),                      # speed level 2
    (320, 240, 0, 3, (1, 3), [P | NS | FA] * 3, 0.5),                       # speed level 3 (SAD scores)
    (1920, 1088, 0, 2, (3, 1), [P | F4] * 2, 0.1),                           # BASELINE configs[3] frame size
]


def run_case(fw, fh, fmt, nframes, motion, flags_seq, gold_density, seed):
    rng = np.random.default_rng(seed)
    R = megen.bind_ref_me(S.ref("c"))
    L = abi.lib()
    g = S.make_geometry(fw, fh, fmt, 6)
    orig, recon = megen.scene_buffers(g, rng, nframes + 1, motion=motion)
    h = R.refh_me_open(fw, fh, fmt)
    assert h
    ctx = T.Context(g, 0)
    me = C.c_void_p()
    abi.check(L.ocg_me_create(C.byref(me), ctx.h, None), "ocg_me_create")
    n = L.ocg_me_nmbs(C.byref(g))
    topo = np.zeros(n, abi.ME_TOPO_DTYPE)
    R.refh_me_topology(h, topo.ctypes.data)
    want = np.zeros(n, abi.ME_MB_DTYPE)
    got = np.zeros(n, abi.ME_MB_DTYPE)
    bufs = (C.c_int * 5)(0, 1, 2, 3, 4)
    try:
        for t in range(1, nframes + 1):
            gold_t = 0 if t < 3 else 1
            frames = [orig[t], orig[t - 1], orig[gold_t], recon[t - 1], recon[gold_t]]
            mask = (rng.random(n) < gold_density).astype(np.uint8)
            fl = flags_seq[t - 1]
            ptrs = (C.c_void_p * 5)(*[f.ctypes.data for f in frames])
            R.refh_me_frame(h, ptrs, fl, mask.ctypes.data if gold_density > 0 else None, want.ctypes.data)
            for i, f in enumerate(frames):
                ctx.upload_frame(i, f)
            abi.check(L.ocg_me_frame(me, bufs, fl, mask.ctypes.data if gold_density > 0 else None), "ocg_me_frame")
            abi.check(L.ocg_me_read(me, got.ctypes.data), "ocg_me_read")
            megen.assert_me_equal(got, want, topo["valid"], fl, "frame %d" % t)
        # the scene really moves: most macro blocks found a non-zero vector
        assert np.mean(want["unref_mv"][topo["valid"] == 1][:, 1] != 0) > 0.5
    finally:
        L.ocg_me_destroy(me)
        ctx.close()
        R.refh_me_close(h)


@pytest.mark.parametrize("case", CASES)
def test_whole_frame_motion_analysis_matches_reference(case):
    run_case(*case, seed=hash(case[:2]) & 0xFFFF)


def test_batch_of_streams_matches_single_calls():
    """ocg_me_frame_batch: 3 independent streams in one launch set == 3 single calls."""
    fw, fh = 176, 144
    rng = np.random.default_rng(5)
    L = abi.lib()
    g = S.make_geometry(fw, fh, 0, 6)
    n = L.ocg_me_nmbs(C.byref(g))
    ctxs, mes, singles = [], [], []
    scenes = [megen.scene_buffers(g, rng, 3, motion=m) for m in ((3, 1), (-2, 2), (0, 4))]
    for s in range(3):
        ctx = T.Context(g, 0)
        ctxs.append(ctx)
        for role, buf in enumerate([scenes[s][0][2], scenes[s][0][1], scenes[s][0][0], scenes[s][1][1], scenes[s][1][0]]):
            ctx.upload_frame(role, buf)
        ctx.sync()
        for k in range(2):
            me = C.c_void_p()
            abi.check(L.ocg_me_create(C.byref(me), ctx.h, None), "ocg_me_create")
            (mes if k == 0 else singles).append(me)
    flags = P | F4
    bufs = (C.c_int * 15)(*([0, 1, 2, 3, 4] * 3))
    arr = (C.c_void_p * 3)(*[m.value for m in mes])
    abi.check(L.ocg_me_frame_batch(arr, bufs, 3, flags, ctxs[0].stream), "ocg_me_frame_batch")
    ctxs[0].sync()
    for s in range(3):
        abi.check(L.ocg_me_frame(singles[s], (C.c_int * 5)(0, 1, 2, 3, 4), flags, None), "ocg_me_frame")
        a = np.zeros(n, abi.ME_MB_DTYPE)
        b = np.zeros(n, abi.ME_MB_DTYPE)
        abi.check(L.ocg_me_read(mes[s], a.ctypes.data), "read")
        abi.check(L.ocg_me_read(singles[s], b.ctypes.data), "read")
        assert a.tobytes() == b.tobytes()
    for m in mes + singles:
        L.ocg_me_destroy(m)
    for c in ctxs:
        c.close()


def test_me_api_rejects_bad_arguments():
    """Error behaviour of the ocg_me_* entry points: negative codes, never a crash."""
    L = abi.lib()
    g3 = S.make_geometry(64, 64, 0, 3)
    g6 = S.make_geometry(64, 64, 0, 6)
    me = C.c_void_p()
    assert L.ocg_me_create(C.byref(me), None, None) == -1            # OCG_EFAULT
    ctx3 = T.Context(g3, 0)
    assert L.ocg_me_create(C.byref(me), ctx3.h, None) == -10         # needs IO + 2 originals + 2 reconstructions
    ctx3.close()
    ctx = T.Context(g6, 0)
    n = L.ocg_me_nmbs(C.byref(g6))
    topo = np.zeros(n, abi.ME_TOPO_DTYPE)
    assert L.ocg_me_topology(C.byref(g6), topo.ctypes.data) == 0
    bad = topo.copy()
    last = int(np.nonzero(bad["valid"])[0][-1])
    first = int(np.nonzero(bad["valid"])[0][0])
    bad["cn"][first][0] = last                                       # a neighbour that comes LATER in coding order
    bad["ncn"][first] = 1
    assert L.ocg_me_create(C.byref(me), ctx.h, bad.ctypes.data) == -10
    abi.check(L.ocg_me_create(C.byref(me), ctx.h, topo.ctypes.data), "ocg_me_create")
    assert L.ocg_me_frame(me, None, 0, None) == -1
    assert L.ocg_me_frame(me, (C.c_int * 5)(0, 1, 2, 3, 4), 1 << 9, None) == -10   # unknown flag
    assert L.ocg_me_frame(me, (C.c_int * 5)(0, 1, 2, 3, 9), 0, None) == -10        # buffer index out of range
    assert L.ocg_me_read(me, None) == -1
    L.ocg_me_destroy(me)
    L.ocg_me_destroy(None)
    ctx.close()


def test_device_matches_oracle_and_golden():
    """Same sequence as tests/golden/me_frame.npz (made by the reference): device == CPU oracle == golden,
    so the check also holds on a box without oracle/_ref."""
    import test_oracle_me_frame as TO
    G = np.load(TO.GOLDEN)
    c = TO.GOLDEN_CASE
    g = S.make_geometry(c["fw"], c["fh"], 0, 6)
    L = abi.lib()
    ctx = T.Context(g, 0)
    me = C.c_void_p()
    abi.check(L.ocg_me_create(C.byref(me), ctx.h, None), "ocg_me_create")
    n = L.ocg_me_nmbs(C.byref(g))
    got = np.zeros(n, abi.ME_MB_DTYPE)
    # regenerate the inputs exactly as run_sequence does
    rng = np.random.default_rng(c["seed"])
    orig, recon = megen.scene_buffers(g, rng, c["nframes"] + 1, motion=c["motion"])
    seq = TO.run_sequence(**c)
    try:
        for t, fl, topo, oracle_state, _ in seq:
            gold_t = 0 if t < 3 else 1
            frames = [orig[t], orig[t - 1], orig[gold_t], recon[t - 1], recon[gold_t]]
            mask = (rng.random(n) < c["density"]).astype(np.uint8)
            for i, f in enumerate(frames):
                ctx.upload_frame(i, f)
            abi.check(L.ocg_me_frame(me, (C.c_int * 5)(0, 1, 2, 3, 4), fl, mask.ctypes.data), "ocg_me_frame")
            abi.check(L.ocg_me_read(me, got.ctypes.data), "ocg_me_read")
            megen.assert_me_equal(got, oracle_state, topo["valid"], fl, "vs oracle, frame %d" % t)
            megen.assert_me_equal(got, G["frame%d" % t].view(abi.ME_MB_DTYPE).reshape(-1), topo["valid"], fl,
                                  "vs golden, frame %d" % t)
    finally:
        L.ocg_me_destroy(me)
        ctx.close()
