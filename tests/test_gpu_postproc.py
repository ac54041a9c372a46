"""TH_DECCTL_SET_PPLEVEL through the B200 back-end: the reconstruction and the (non-normative)
de-blocking / de-ringing filters of decode.c:1609-1957 run on the device (ocg_dec_postproc.cu); the
back-end takes the level away from the reference's MCU loop, so the host filters never run.  Output
must equal the unmodified reference decoder's at the same level, every frame."""
import ctypes as C

import numpy as np
import pytest

import support as S
import th_streams as streams

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (S.ref_available("c") and streams.available()),
                                 reason="needs oracle/_ref and the integrated build")]


def decode_all(lib, st_blob, level, nframes):
    buf = (C.c_uint8 * len(st_blob)).from_buffer_copy(st_blob)
    sh = lib.refh_stream_from_blob(buf, len(st_blob))
    stream = S.Stream(lib, sh)
    dec = S.Decoder(lib, stream)
    assert dec.set_pplevel(level) == 0
    out = []
    for _ in range(nframes):
        assert dec.next() >= 0
        out.append(dec.frame())
    dec.close()
    stream.free()
    return out


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("case", [(176, 144, 6, 10, 4, 28), (352, 288, 5, 24, 64, 30), (208, 112, 5, 5, 3, 26),
                                  (1920, 1088, 3, 16, 64, 27), (64, 8, 4, 8, 2, 26), (8, 64, 4, 8, 2, 26)])
def test_postprocessed_output_matches_reference(case, level):
    w, h, n, q, kf, ns = case
    R = S.ref("c")
    st = S.Stream.encode(R, w, h, n, quality=q, kf=kf, speed=1, noise_shift=ns)
    blob = st.to_bytes()
    st.free()
    want = decode_all(R, blob, level, n)
    G = streams.lib()
    G.ocg_backend_set_mode(streams.BACKEND_GPU)
    got = decode_all(G, blob, level, n)
    for i in range(n):
        assert np.array_equal(got[i], want[i]), "pp level %d: frame %d differs" % (level, i)
    if level >= 2:
        plain = decode_all(R, blob, 0, n)
        assert any(not np.array_equal(plain[i], want[i]) for i in range(n)), "the filters changed nothing"


@pytest.mark.parametrize("fmt", [2, 3])  # TH_PF_422, TH_PF_444
def test_postprocessing_other_pixel_formats(fmt):
    R = S.ref("c")
    st = S.Stream.encode(R, 176, 144, 4, quality=8, kf=3, speed=1, noise_shift=27, fmt=fmt)
    blob = st.to_bytes()
    st.free()
    G = streams.lib()
    G.ocg_backend_set_mode(streams.BACKEND_GPU)
    for level in (2, 4, 5, 7):
        want = decode_all(R, blob, level, 4)
        got = decode_all(G, blob, level, 4)
        for i in range(4):
            assert np.array_equal(got[i], want[i]), "format %d, pp level %d: frame %d differs" % (fmt, level, i)


def test_level_changes_between_frames():
    """The level may change with every packet (TH_DECCTL_SET_PPLEVEL between th_decode_packetin calls)."""
    R = S.ref("c")
    st = S.Stream.encode(R, 352, 288, 8, quality=10, kf=64, speed=1, noise_shift=27)
    blob = st.to_bytes()
    st.free()
    levels = [0, 7, 3, 0, 5, 1, 6, 2]

    def run(lib):
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        stream = S.Stream(lib, lib.refh_stream_from_blob(buf, len(blob)))
        dec = S.Decoder(lib, stream)
        out = []
        for lv in levels:
            assert dec.set_pplevel(lv) == 0
            assert dec.next() >= 0
            out.append(dec.frame())
        dec.close()
        stream.free()
        return out

    want = run(R)
    G = streams.lib()
    G.ocg_backend_set_mode(streams.BACKEND_GPU)
    got = run(G)
    for i in range(len(levels)):
        assert np.array_equal(got[i], want[i]), "frame %d (level %d) differs" % (i, levels[i])
