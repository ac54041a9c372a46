"""Device-side token expansion (ocg_dec_flush_tokens: SURVEY 8(f)1, reference decode.c:1511-1586): streams
with dense coefficient lists, long EOB runs, every quantiser count, large values and all pixel formats must
decode to exactly the reference decoder's frames with the host recording nothing per fragment."""
import numpy as np
import pytest

import support as S
import th_streams as streams

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (S.ref_available("c") and streams.available()),
                                 reason="needs oracle/_ref and the integrated build")]

# (w, h, frames, quality, kf, speed, noise_shift, fmt)
CASES = [
    (64, 64, 3, 63, 64, 1, 24, 0),      # finest quantiser + heavy noise: long token lists, large values
    (176, 144, 6, 63, 3, 0, 25, 0),     # key frames inside the run, speed 0 (up to 3 quantisers per frame)
    (176, 144, 5, 0, 64, 1, 30, 0),     # coarsest quantiser: almost everything ends at index 0/1, long EOB runs
    (320, 240, 5, 40, 64, 1, 26, 0),
    (352, 288, 4, 56, 64, 2, 24, 0),
    (208, 112, 5, 50, 4, 1, 25, 2),     # 4:2:2
    (208, 112, 5, 50, 4, 1, 25, 3),     # 4:4:4
    (1920, 1080, 3, 48, 64, 1, 26, 0),  # BASELINE frame size, noisy
    (1920, 1080, 3, 20, 64, 1, 30, 0),
]


@pytest.mark.parametrize("dc_mode", [streams.DC_HOST, streams.DC_DEVICE])
@pytest.mark.parametrize("case", CASES)
def test_token_path_matches_reference(case, dc_mode):
    w, h, n, q, kf, sp, ns, fmt = case
    R = S.ref("c")
    st = S.Stream.encode(R, w, h, n, quality=q, kf=kf, speed=sp, noise_shift=ns, fmt=fmt)
    _, works, outs = streams.capture_stream_work(st.to_bytes(), streams.BACKEND_GPU, dc_mode=dc_mode,
                                                 expand=streams.EXPAND_DEVICE)
    assert all(wk is None for wk in works), "the token path must not hand out host lists"
    dec = S.Decoder(R, st)
    assert len(outs) == n
    for i in range(n):
        assert dec.next() >= 0
        assert np.array_equal(outs[i], dec.frame()), "frame %d differs" % i
    dec.close()
    st.free()


def test_token_path_long_stream_reuses_its_graph():
    """More frames than the kernel-by-kernel warm-up: the flush graph replays from frame 16 on."""
    R = S.ref("c")
    st = S.Stream.encode(R, 96, 80, 40, quality=36, kf=16, speed=1, noise_shift=27)
    _, _, outs = streams.capture_stream_work(st.to_bytes(), streams.BACKEND_GPU)
    dec = S.Decoder(R, st)
    for i in range(40):
        assert dec.next() >= 0
        assert np.array_equal(outs[i], dec.frame()), "frame %d differs" % i
    dec.close()
    st.free()
