"""End-to-end on the GPU through the reference's public API: real packets ->
th_decode_packetin (reference host code, B200 vtable back-end) -> frame in host
memory, compared bit-for-bit with the unmodified reference decoder."""
import numpy as np
import pytest

import support as S
import th_streams as streams

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (S.ref_available("c") and streams.available()),
                                 reason="needs oracle/_ref and the integrated build")]

CASES = [
    (64, 64, 2, 48, 64, 1, 30),
    (64, 64, 6, 32, 4, 1, 28),
    (176, 144, 8, 20, 64, 1, 28),
    (350, 270, 5, 40, 64, 1, 30),
    (320, 240, 6, 10, 3, 0, 28),
    (96, 80, 10, 60, 64, 2, 26),
    (1920, 1080, 4, 32, 64, 1, 30),
    (3840, 2160, 3, 32, 64, 2, 30),   # BASELINE configs[4] frame size
]


@pytest.mark.parametrize("expand", [streams.EXPAND_DEVICE, streams.EXPAND_REFERENCE])
@pytest.mark.parametrize("dc_mode", [streams.DC_DEVICE, streams.DC_HOST, streams.DC_DEVICE_AHEAD])
@pytest.mark.parametrize("case", CASES)
def test_public_api_decode_matches_reference(case, dc_mode, expand):
    """expand EXPAND_DEVICE (the default): the device walks the decoder's token lists (ocg_dec_flush_tokens);
    EXPAND_REFERENCE: the reference's expansion loop with the per-fragment recorder hook.
    DC_DEVICE: DC un-prediction by the wave-front kernel inside the flush; DC_DEVICE_AHEAD: the same kernel
    started ahead of the lists (recorder path only); DC_HOST: on the host, in the hook."""
    if expand == streams.EXPAND_DEVICE and dc_mode == streams.DC_DEVICE_AHEAD:
        pytest.skip("the token path undoes the DC prediction inside its flush")
    w, h, n, q, kf, sp, ns = case
    R = S.ref("c")
    st = S.Stream.encode(R, w, h, n, quality=q, kf=kf, speed=sp, noise_shift=ns)
    g, works, outs = streams.capture_stream_work(st.to_bytes(), streams.BACKEND_GPU, dc_mode=dc_mode, expand=expand)
    if expand == streams.EXPAND_REFERENCE:
        want = {streams.DC_DEVICE: 1, streams.DC_HOST: 0, streams.DC_DEVICE_AHEAD: 2}[dc_mode]
        assert all(wk is None or wk.dc_residual == want for wk in works)
    dec = S.Decoder(R, st)
    assert len(outs) == n
    for i in range(n):
        assert dec.next() >= 0
        want = dec.frame()
        assert np.array_equal(outs[i], want), "frame %d differs" % i
    dec.close()
    st.free()


@pytest.mark.parametrize("fmt", [2, 3])
@pytest.mark.parametrize("q", [20, 48])
def test_422_and_444_streams_match_reference(fmt, q):
    """TH_PF_422 / TH_PF_444: full-resolution chroma planes, half-pel chroma vectors
    (state.c:888-957), luma-sized strides for every plane in 4:4:4."""
    R = S.ref("c")
    st = S.Stream.encode(R, 208, 112, 7, quality=q, kf=5, speed=1, noise_shift=28, fmt=fmt)
    g, works, outs = streams.capture_stream_work(st.to_bytes(), streams.BACKEND_GPU)
    assert g.pixel_fmt == fmt
    dec = S.Decoder(R, st)
    assert len(outs) == 7
    for i in range(7):
        assert dec.next() >= 0
        assert np.array_equal(outs[i], dec.frame()), "frame %d differs" % i
    dec.close()
    st.free()


def test_stream_starting_on_inter_frame_uses_grey_reference():
    """decode.c:2053 oc_dec_init_dummy_frame: drop the keyframe, decode the rest."""
    import ctypes as C
    R = S.ref("c")
    st = S.Stream.encode(R, 96, 80, 4, quality=32, kf=64, speed=1, noise_shift=28)
    # rebuild a stream without the first data packet
    blob = st.to_bytes()
    hdr = np.frombuffer(blob[:16], np.uint32)
    npk = int(hdr[1])
    sizes = np.frombuffer(blob[16:16 + 4 * npk], np.uint32).copy()
    data_off = 16 + 4 * npk
    offs = (np.concatenate([[0], np.cumsum(sizes)]) + data_off).astype(np.int64).tolist()
    keep = [0, 1, 2] + list(range(4, npk))
    new = np.array([hdr[0], len(keep), hdr[2], hdr[3]], np.uint32).tobytes() + sizes[keep].tobytes() + \
        b"".join(blob[offs[i]:offs[i + 1]] for i in keep)
    st2 = S.Stream.from_bytes(R, new)
    g, works, outs = streams.capture_stream_work(new, streams.BACKEND_GPU)
    dec = S.Decoder(R, st2)
    for i in range(len(outs)):
        assert dec.next() >= 0
        assert np.array_equal(outs[i], dec.frame()), i
    dec.close()
    st.free()
    st2.free()
