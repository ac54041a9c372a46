"""Inter-frame encode (BASELINE configs[3]) through the reference's public API: an encoder that can emit
inter frames (keyframe_granule_shift > 0) runs on the B200 back-end -- every frame's reconstruction, uncoded
copies, loop filter and borders on the device (recorded from the analysis loop and flushed as one graph), the
analysis loop itself fed by the device tables where its inputs are known ahead of it -- and must produce
the SAME PACKETS, byte for byte, as the unmodified reference C encoder, and leave the reference encoder's own
reconstruction in its SELF buffer (so the next frame predicts from the right pixels)."""
import ctypes as C

import numpy as np
import pytest

import support as S
import th_streams as streams

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (S.ref_available("c") and streams.available()),
                                 reason="needs oracle/_ref and the integrated build")]

# (w, h, frames, quality, kf, speed, noise_shift)
CASES = [
    (64, 64, 4, 48, 64, 1, 30),
    (96, 80, 6, 32, 4, 1, 28),        # key frames inside the run
    (176, 144, 8, 20, 64, 0, 28),     # speed 0, loop filter on, adaptive quantisers
    (176, 144, 8, 40, 64, 2, 27),     # speed 2 (fast analysis: no 4MV)
    (350, 270, 5, 32, 64, 1, 28),     # picture not a multiple of 16 (border SSD masks)
    (320, 240, 6, 5, 64, 1, 26),      # coarse quantiser: early skips, many uncoded blocks
    (640, 360, 5, 63, 64, 1, 29),     # finest quantiser, loop filter off
    (1920, 1080, 3, 32, 64, 1, 30),   # BASELINE configs[3]
]


def frame_bytes(w, h):
    fw, fh = (w + 15) & ~15, (h + 15) & ~15
    return fw * fh + 2 * (fw // 2) * (fh // 2)


@pytest.mark.parametrize("case", CASES)
def test_inter_encode_is_bit_identical(case):
    w, h, n, q, kf, sp, ns = case
    R = S.ref("c")
    G = streams.lib()
    want = S.Stream.encode(R, w, h, n, quality=q, kf=kf, speed=sp, noise_shift=ns)
    st = streams.EncBackendStats()
    G.ocg_backend_get_enc_stats(None, 1)
    recon = np.zeros(frame_bytes(w, h), np.uint8)
    hnd = G.refh_encode_synth_recon(w, h, 0, n, q, kf, sp, ns, 12345, recon.ctypes.data)
    assert hnd, "device encoder failed to allocate"
    got = S.Stream(G, hnd)
    G.ocg_backend_get_enc_stats(C.byref(st), 0)
    # every packed frame was reconstructed on the device (the first key frame and the first inter frame are
    # packed twice, encode.c:1283-1290, 1304-1317)
    assert st.frames >= n
    assert got.packet_sizes() == want.packet_sizes()
    assert got.to_bytes() == want.to_bytes(), "packets differ from the reference encoder's"
    # closed loop (speed < 2, see test_gpu_encode_intra.py): the reference decoder's last frame == the
    # reconstruction the device left in the encoder's SELF buffer
    if sp < 2:
        dec = S.Decoder(R, want)
        for _ in range(n):
            assert dec.next() >= 0
        assert np.array_equal(dec.frame(), recon), "device reconstruction differs from the decoded frame"
        dec.close()
    want.free()
    got.free()


@pytest.mark.parametrize("fmt", [2, 3])
def test_inter_encode_422_444(fmt):
    R = S.ref("c")
    G = streams.lib()
    want = S.Stream.encode(R, 144, 96, 6, quality=24, kf=64, speed=1, noise_shift=28, fmt=fmt)
    got = S.Stream.encode(G, 144, 96, 6, quality=24, kf=64, speed=1, noise_shift=28, fmt=fmt)
    assert got.to_bytes() == want.to_bytes()
    want.free()
    got.free()


def test_inter_capable_encoder_runs_on_the_device():
    """keyframe_granule_shift > 0: device frames == packed frames."""
    R = S.ref("c")
    G = streams.lib()
    G.ocg_backend_get_enc_stats(None, 1)
    want = S.Stream.encode(R, 96, 80, 4, quality=32, kf=4, speed=1, noise_shift=28)
    got = S.Stream.encode(G, 96, 80, 4, quality=32, kf=4, speed=1, noise_shift=28)
    st = streams.EncBackendStats()
    G.ocg_backend_get_enc_stats(C.byref(st), 0)
    assert st.frames >= 4
    assert got.to_bytes() == want.to_bytes()
    want.free()
    got.free()
