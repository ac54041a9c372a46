"""The vtable back-end's recorder + the oracle's whole-frame executor against
the plain reference decoder on real streams, CPU only.

The integrated library (reference host code + back-end, record mode) produces
the per-frame lists of include/theora_b200.h; replaying them through
oco_dec_frame must reproduce, bit for bit, what the unmodified reference decodes
from the same packets.  This pins the data format, the class binning, the row
packing and the executor's ordering on real bitstreams."""
import ctypes as C

import numpy as np
import pytest

import support as S
import th_streams as streams

pytestmark = pytest.mark.skipif(not (S.ref_available("c") and streams.available()),
                                reason="needs oracle/_ref and the integrated build")

CASES = [
    # w, h, frames, quality, kf, speed, noise_shift
    (64, 64, 2, 48, 64, 1, 30),      # BASELINE configs[0] size, loop filter off
    (64, 64, 6, 32, 4, 1, 28),       # loop filter on, several keyframes
    (176, 144, 8, 20, 64, 1, 28),    # strong loop filter, inter frames
    (350, 270, 5, 40, 64, 1, 30),    # cropped picture (frame 352x272)
    (320, 240, 6, 10, 3, 0, 28),     # low quality, speed 0
    (96, 80, 10, 60, 64, 2, 26),     # high quality / heavy noise: dense blocks
    (208, 112, 6, 24, 4, 1, 28, 2),  # TH_PF_422
    (208, 112, 6, 24, 4, 1, 28, 3),  # TH_PF_444
]


def replay_and_compare(case, dc_mode=streams.DC_HOST):
    w, h, n, q, kf, sp, ns = case[:7]
    fmt = case[7] if len(case) > 7 else 0
    R = S.ref("c")
    st = S.Stream.encode(R, w, h, n, quality=q, kf=kf, speed=sp, noise_shift=ns, fmt=fmt)
    blob = st.to_bytes()
    g, works, _ = streams.capture_stream_work(blob, streams.BACKEND_RECORD, dc_mode=dc_mode)
    assert all(wk is None or wk.dc_residual == (dc_mode == streams.DC_DEVICE) for wk in works)
    dec = S.Decoder(R, st)
    frames = np.full(g.nrefs * g.ref_frame_sz, 0x80, np.uint8)
    assert len(works) == n
    stats = {"coded": 0, "uncoded": 0, "rows": 0, "cls": [0, 0, 0, 0]}
    for i, wk in enumerate(works):
        assert dec.next() >= 0
        want = dec.frame()
        if wk is not None:
            f = wk.as_struct()
            S.oracle().oco_dec_frame(C.byref(g), S.ptr(frames, S.u8p), C.byref(f), 7)
            cur = wk.ref_idx[2]
            stats["coded"] += wk.ncoded
            stats["uncoded"] += wk.nuncoded
            stats["rows"] += len(wk.rows)
            for k in range(4):
                stats["cls"][k] += wk.ncls[k]
        planes = S.planes_from_buffer(g, frames[cur * g.ref_frame_sz:(cur + 1) * g.ref_frame_sz])
        got = np.concatenate([p.ravel() for p in planes])
        assert np.array_equal(got, want), "frame %d differs from the reference decoder" % i
    dec.close()
    st.free()
    return stats


@pytest.mark.parametrize("dc_mode", [streams.DC_DEVICE, streams.DC_HOST])
@pytest.mark.parametrize("case", CASES)
def test_recorded_lists_replay_to_reference_frames(case, dc_mode):
    """dc_mode DC_DEVICE: the records carry DC residuals and the executor undoes the
    prediction (oco_dc_unpredict_plane); DC_HOST: the reference's own routine ran in the hook."""
    stats = replay_and_compare(case, dc_mode)
    assert stats["coded"] > 0


def test_every_fragment_is_accounted_for():
    R = S.ref("c")
    st = S.Stream.encode(R, 176, 144, 5, quality=32, kf=64, speed=1, noise_shift=28)
    g, works, _ = streams.capture_stream_work(st.to_bytes(), streams.BACKEND_RECORD)
    offs = np.empty(g.nfrags, np.int32)
    S.oracle().oco_geometry_frag_buf_offs(C.byref(g), S.ptr(offs, S.i32p))
    for i, wk in enumerate(works):
        assert len(wk.recs) == g.nfrags
        assert np.array_equal(wk.recs["buf_off"], offs)
        for pli in range(3):
            p = g.planes[pli]
            assert np.all((wk.recs["pli_qti"][p.froffset:p.froffset + p.nfrags] & 3) == pli)
        if i == 0:
            assert wk.nuncoded == 0  # keyframe: everything coded, all intra
            assert np.all(wk.recs["refi"] == 2)
        # stored rows are exactly the rows the masks announce
        assert int(sum(bin(int(m)).count("1") for m in wk.recs["rowmask"][wk.coded_mask])) == len(wk.rows)
    st.free()
