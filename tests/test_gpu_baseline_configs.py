"""The BASELINE.json configurations at their full size, pinned frame by frame (-m gpu):
  configs[1]  1080p 4:2:0 decode, 300 synthetic inter/intra frames (key frame every 64), quality 32 (loop filter
              on) and quality 48 (loop filter off): every frame's hash through th_decode_packetin on the B200
              back-end == the compiled reference decoder's;
  configs[2]  1080p intra-only encode, 300 frames: packets byte-identical to the reference encoder's;
  configs[3]  1080p encode with the motion search, speed level 1, key frame every 64: packets byte-identical."""
import ctypes as C

import numpy as np
import pytest

import support as S
import th_streams as streams
import th_workload as wl

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (S.ref_available("c") and streams.available()),
                                 reason="needs oracle/_ref and the integrated build")]


def frame_hashes(lib, blob, nframes):
    buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
    sh = lib.refh_stream_from_blob(buf, len(blob))
    assert sh
    d = lib.refh_dec_open(sh)
    assert d, "decoder failed to open"
    out = []
    h = (C.c_uint64 * 3)()
    for _ in range(nframes):
        assert lib.refh_dec_next(d) >= 0
        lib.refh_dec_hash(d, h)
        out.append((int(h[0]), int(h[1]), int(h[2])))
    assert lib.refh_dec_next(d) == 1000
    lib.refh_dec_close(d)
    lib.refh_stream_free(sh)
    return out


@pytest.mark.parametrize("quality", [32, 48])
@pytest.mark.parametrize("dc_mode", [streams.DC_DEVICE, streams.DC_HOST])
def test_config1_1080p_300_frames_every_frame(quality, dc_mode):
    R = S.ref("asm" if S.ref_available("asm") else "c")
    blob = wl.synth_stream(1920, 1080, 300, quality, 64, lib=R)
    want = frame_hashes(R, blob, 300)
    G = streams.lib()
    G.ocg_backend_set_mode(streams.BACKEND_GPU)
    G.ocg_backend_set_dc_mode(dc_mode)
    G.ocg_backend_set_expand_mode(streams.EXPAND_DEVICE)
    try:
        got = frame_hashes(G, blob, 300)
    finally:
        G.ocg_backend_set_dc_mode(streams.DC_HOST)
    bad = [i for i in range(300) if got[i] != want[i]]
    assert not bad, "frames that differ from the reference decoder: %s" % bad[:10]


def encode_hash(lib, frames, kf, speed, quality=32):
    h, b = C.c_uint64(), C.c_long()
    secs = lib.refh_encode_time_mt(1920, 1080, frames, quality, kf, speed, 30, 12345, 1, C.byref(h), C.byref(b))
    assert secs > 0, "encode failed"
    return h.value, b.value


def test_config2_1080p_intra_only_300_frames():
    R = S.ref("asm" if S.ref_available("asm") else "c")
    G = streams.lib()
    G.ocg_backend_get_enc_stats(None, 1)
    got = encode_hash(G, 300, 1, 1)
    st = streams.EncBackendStats()
    G.ocg_backend_get_enc_stats(C.byref(st), 0)
    assert st.frames >= 300
    assert got == encode_hash(R, 300, 1, 1), "packets differ from the reference encoder's"


@pytest.mark.parametrize("speed", [0, 1, 2])
def test_config3_1080p_inter_encode(speed):
    """kf=64, 72 frames: a key frame, 63 inter frames, a key frame, 7 inter frames."""
    R = S.ref("asm" if S.ref_available("asm") else "c")
    G = streams.lib()
    G.ocg_backend_get_enc_stats(None, 1)
    got = encode_hash(G, 72, 64, speed)
    st = streams.EncBackendStats()
    G.ocg_backend_get_enc_stats(C.byref(st), 0)
    assert st.frames >= 72 and st.me_frames >= 70
    assert got == encode_hash(R, 72, 64, speed), "packets differ from the reference encoder's"
