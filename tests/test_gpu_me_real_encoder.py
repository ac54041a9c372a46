"""Device motion analysis against a REAL reference encode (BASELINE configs[3]
conditions): the unmodified encoder runs on the host (integrated build,
OCG_ENC_HOST) with the analysis-pass spy installed; every pass hands out the
frame buffers and the per-macro-block state the previous pass left.  For each
frame whose first pass runs oc_mcenc_search (analyze.c:1725 / 2402) the device
starts from the captured state, gets the same five frames, the GOLD-refinement
set the encoder's mode decision actually chose (oc_mb_enc_info.refined & 0x40),
and must reproduce the encoder's own results: vectors incl. history, errors,
SATDs, 4MV vectors, and the 4MV refinements where the encoder ran them
(refined & 0x80).  Covers key frames with search, the first-inter-frame dry
run, speed levels, and scene content that triggers GOLD refinement."""
import ctypes as C

import numpy as np
import pytest

import megen
import support as S
import theora_b200 as T
from theora_b200 import abi
import th_streams as streams

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not streams.available(), reason="needs the integrated build")]


class SpyFrame(C.Structure):
    _fields_ = [("frame_type", C.c_int32), ("prevframe_dropped", C.c_int32), ("sp_level", C.c_int32),
                ("keyframe_frequency_force", C.c_int32), ("nmbs", C.c_int32), ("reserved", C.c_int32),
                ("curframe_num", C.c_int64), ("ref_frame_sz", C.c_int64), ("frames", C.c_void_p * 5),
                ("state", C.c_void_p), ("refined", C.c_void_p)]


SPY_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(SpyFrame))


def capture(w, h, n, q, kf, speed, noise):
    G = streams.lib()
    snaps = []

    def on_pass(user, fp):
        f = fp.contents
        fr = [None if not f.frames[i] else np.frombuffer((C.c_uint8 * f.ref_frame_sz).from_address(f.frames[i]),
                                                         np.uint8).copy() for i in range(5)]
        st = np.frombuffer((C.c_uint8 * (f.nmbs * abi.ME_MB_DTYPE.itemsize)).from_address(f.state),
                           abi.ME_MB_DTYPE).copy()
        rf = np.frombuffer((C.c_uint8 * f.nmbs).from_address(f.refined), np.uint8).copy()
        snaps.append(dict(frame_type=f.frame_type, dropped=f.prevframe_dropped, sp=f.sp_level,
                          kff=f.keyframe_frequency_force, num=f.curframe_num, nmbs=f.nmbs, frames=fr, state=st,
                          refined=rf))
    cb = SPY_FN(on_pass)
    G.ocg_backend_set_enc_spy.argtypes = [SPY_FN, C.c_void_p]
    G.ocg_backend_set_enc_mode(streams.ENC_HOST)
    G.ocg_backend_set_enc_spy(cb, None)
    try:
        hnd = G.refh_encode_synth(w, h, 0, n, q, kf, speed, noise, 12345)
        assert hnd
        G.refh_stream_free(hnd)
    finally:
        G.ocg_backend_set_enc_spy(SPY_FN(), None)
        G.ocg_backend_set_enc_mode(streams.ENC_AUTO)
    return snaps


CASES = [
    # w, h, frames, quality, kf, speed, noise_shift
    (176, 144, 8, 32, 4, 1, 28),    # key frames every 4: searches inside intra analysis too
    (352, 288, 7, 50, 64, 1, 26),   # noisy, fine quantiser: GOLD refinements happen
    (352, 288, 6, 10, 64, 0, 28),   # speed 0
    (208, 112, 6, 32, 64, 2, 28),   # speed 2: no 4MV
    (640, 368, 5, 32, 64, 1, 30),
]


@pytest.mark.parametrize("case", CASES)
def test_device_reproduces_the_encoders_motion_analysis(case):
    w, h, n, q, kf, speed, noise = case
    snaps = capture(w, h, n, q, kf, speed, noise)
    fw, fh = (w + 15) & ~15, (h + 15) & ~15
    g = S.make_geometry(fw, fh, 0, 6)
    L = abi.lib()
    nmbs = L.ocg_me_nmbs(C.byref(g))
    topo = np.zeros(nmbs, abi.ME_TOPO_DTYPE)
    L.ocg_me_topology(C.byref(g), topo.ctypes.data)
    valid = topo["valid"].astype(bool)
    ctx = T.Context(g, 0)
    me = C.c_void_p()
    abi.check(L.ocg_me_create(C.byref(me), ctx.h, None), "ocg_me_create")
    bufs = (C.c_int * 5)(0, 1, 2, 3, 4)
    got = np.zeros(nmbs, abi.ME_MB_DTYPE)
    checked = inter_checked = gold_refined = four_refined = 0
    seen = set()
    try:
        for a, b in zip(snaps[:-1], snaps[1:]):
            first_pass = a["num"] not in seen
            seen.add(a["num"])
            inter = a["frame_type"] == 1
            searches = first_pass and a["num"] > 0 and a["sp"] < 4 and (inter or a["kff"] > 1)
            if not searches:
                # nothing may have touched the vectors' history (refinements of a recode pass aside)
                assert np.array_equal(a["state"]["analysis_mv"][:, 1:], b["state"]["analysis_mv"][:, 1:])
                continue
            assert a["nmbs"] == nmbs and all(f is not None and f.size == g.ref_frame_sz for f in a["frames"])
            flags = (abi.OCG_ME_DROPPED if a["dropped"] else 0)
            flags |= abi.OCG_ME_NOSATD if a["sp"] >= 3 else 0
            flags |= abi.OCG_ME_FAST if a["sp"] >= 2 else 0
            mask = None
            if inter:
                flags |= abi.OCG_ME_REFINE_PREV | abi.OCG_ME_REFINE_4MV
                mask = ((b["refined"] & 0x40) != 0).astype(np.uint8)
                gold_refined += int(mask[valid].sum())
            for i in range(5):
                ctx.upload_frame(i, a["frames"][i])
            abi.check(L.ocg_me_write(me, a["state"].ctypes.data), "ocg_me_write")
            abi.check(L.ocg_me_frame(me, bufs, flags, mask.ctypes.data if mask is not None else None), "ocg_me_frame")
            abi.check(L.ocg_me_read(me, got.ctypes.data), "ocg_me_read")
            want = b["state"]
            where = "frame %d (%s)" % (a["num"], "inter" if inter else "intra")
            for f in ("analysis_mv", "error", "satd"):
                assert np.array_equal(got[f][valid], want[f][valid]), "%s: %s differs" % (where, f)
            if a["sp"] < 2:
                assert np.array_equal(got["block_mv"][valid], want["block_mv"][valid]), where + ": block_mv"
                r4 = valid & ((b["refined"] & 0x80) != 0) if inter else np.zeros(nmbs, bool)
                four_refined += int(r4.sum())
                assert np.array_equal(got["ref_mv"][r4], want["ref_mv"][r4]), where + ": ref_mv"
                assert np.array_equal(got["ref_block_satd"][r4], want["ref_block_satd"][r4]), where + ": refined block_satd"
                nr = valid & ~r4
                assert np.array_equal(got["block_satd"][nr], want["block_satd"][nr]), where + ": block_satd"
            if inter:
                assert np.array_equal(got["unref_mv"][valid], want["unref_mv"][valid]), where + ": unref_mv"
                inter_checked += 1
            checked += 1
    finally:
        L.ocg_me_destroy(me)
        ctx.close()
    assert checked >= n - 3 and inter_checked >= 1
    print("frames checked %d (inter %d), GOLD refinements %d, 4MV refinements %d" % (checked, inter_checked, gold_refined,
                                                                                    four_refined))
    if case[3] == 50:
        assert gold_refined > 0, "this case is meant to exercise GOLD refinement"
