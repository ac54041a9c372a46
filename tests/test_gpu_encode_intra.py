"""Intra-only encode (BASELINE configs[2]) through the reference's public API:
th_encode_ycbcr_in / th_encode_packetout of the unmodified host code with the
B200 encoder back-end (device pre-pass look-ups + recorded reconstruction) must
produce the SAME PACKETS, byte for byte, as the unmodified reference C encoder,
and the reconstruction the device leaves in the encoder's SELF buffer must equal
the reference encoder's own reconstruction (host kernels, OCG_ENC_HOST) -- and,
where the reference itself is drift-free, what the reference decoder makes of
those packets (closed loop).  (At speed level >= 2 the reference encoder's
reconstruction differs from its decoder's: the fast tokeniser can zero every AC
coefficient of a block that analyze.c:803-806 still reconstructs with the full
iDCT, while the decoder takes the DC-only shortcut of state.c:967 -- the TODO at
analyze.c:787.  The device reproduces the ENCODER there, as a drop-in must.)"""
import ctypes as C

import numpy as np
import pytest

import support as S
import th_streams as streams

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (S.ref_available("c") and streams.available()),
                                 reason="needs oracle/_ref and the integrated build")]

# (w, h, frames, quality, speed, noise_shift)
CASES = [
    (64, 64, 2, 48, 1, 30),        # BASELINE configs[0] size
    (64, 64, 3, 63, 0, 28),        # finest quantiser, loop filter off
    (176, 144, 3, 20, 1, 28),      # loop filter on, adaptive quantisers (nqis up to 3)
    (350, 270, 2, 40, 0, 30),      # picture not a multiple of 16
    (320, 240, 3, 5, 1, 26),       # coarse quantiser: many DC-only blocks
    (96, 80, 3, 32, 2, 26),        # speed level 2 (fast tokeniser, oc_mb_intra_satd)
    (640, 360, 2, 0, 1, 30),       # quality 0
    (1920, 1080, 2, 32, 1, 30),    # BASELINE configs[2] frame size, loop filter on
    (1920, 1080, 2, 48, 1, 30),    # loop filter off
]


def frame_bytes(w, h):
    fw, fh = (w + 15) & ~15, (h + 15) & ~15
    return fw * fh + 2 * (fw // 2) * (fh // 2)


@pytest.mark.parametrize("case", CASES)
def test_intra_only_encode_is_bit_identical_and_closed_loop(case):
    w, h, n, q, sp, ns = case
    R = S.ref("c")
    G = streams.lib()
    want = S.Stream.encode(R, w, h, n, quality=q, kf=1, speed=sp, noise_shift=ns)
    st = streams.EncBackendStats()
    G.ocg_backend_get_enc_stats(None, 1)
    recon = np.zeros(frame_bytes(w, h), np.uint8)
    hnd = G.refh_encode_synth_recon(w, h, 0, n, q, 1, sp, ns, 12345, recon.ctypes.data)
    assert hnd, "device encoder failed to allocate"
    got = S.Stream(G, hnd)
    G.ocg_backend_get_enc_stats(C.byref(st), 0)
    # every frame went through the device (frame 0 is analysed twice, encode.c:1283-1290)
    assert st.frames == n + 1 and st.prepass_frames == n + 1
    assert got.packet_sizes() == want.packet_sizes()
    assert got.to_bytes() == want.to_bytes(), "packets differ from the reference encoder's"
    # the reference encoder's own reconstruction of the last frame (host kernels)
    host_recon = np.zeros_like(recon)
    G.ocg_backend_set_enc_mode(streams.ENC_HOST)
    try:
        hh = G.refh_encode_synth_recon(w, h, 0, n, q, 1, sp, ns, 12345, host_recon.ctypes.data)
    finally:
        G.ocg_backend_set_enc_mode(streams.ENC_AUTO)
    assert hh
    G.refh_stream_free(hh)
    assert np.array_equal(host_recon, recon), "device reconstruction differs from the reference encoder's"
    if sp < 2:
        # closed loop: reference decoder output of the last frame == device reconstruction
        dec = S.Decoder(R, want)
        for _ in range(n):
            assert dec.next() >= 0
        assert np.array_equal(dec.frame(), recon), "device reconstruction differs from the decoded frame"
        dec.close()
    want.free()
    got.free()


@pytest.mark.parametrize("fmt", [2, 3])
def test_intra_only_encode_422_444(fmt):
    R = S.ref("c")
    G = streams.lib()
    want = S.Stream.encode(R, 144, 96, 3, quality=24, kf=1, speed=1, noise_shift=28, fmt=fmt)
    G.ocg_backend_get_enc_stats(None, 1)
    got = S.Stream.encode(G, 144, 96, 3, quality=24, kf=1, speed=1, noise_shift=28, fmt=fmt)
    st = streams.EncBackendStats()
    G.ocg_backend_get_enc_stats(C.byref(st), 0)
    assert st.frames == 4
    assert got.to_bytes() == want.to_bytes()
    want.free()
    got.free()


def test_multithreaded_encoders_agree_with_the_reference():
    R = S.ref("c")
    G = streams.lib()
    hr, hg = C.c_uint64(), C.c_uint64()
    br, bg = C.c_long(), C.c_long()
    assert R.refh_encode_time_mt(320, 240, 4, 32, 1, 1, 28, 777, 1, C.byref(hr), C.byref(br)) > 0
    assert G.refh_encode_time_mt(320, 240, 4, 32, 1, 1, 28, 777, 3, C.byref(hg), C.byref(bg)) > 0
    assert (hr.value, br.value) == (hg.value, bg.value)
