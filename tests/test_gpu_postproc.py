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


def _smoothish_frame(g, rng):
    """A padded buffer whose picture has regions of every variance class: smooth ramps, mild and heavy noise."""
    import workgen as W
    buf = W.random_frames(g, rng)[:g.ref_frame_sz].copy()
    for pli in range(3):
        p = g.planes[pli]
        w, h, stride = p.width, p.height, -p.ystride
        yy, xx = np.mgrid[0:h, 0:w]
        base = (xx * 3 + yy * 2) % 256
        amp = np.choose(((xx // 32) + (yy // 24)) % 4, [0, 3, 12, 60])
        img = np.clip(base + rng.integers(-1, 2, size=(h, w)) * amp + ((xx // 8 + yy // 8) % 2) * (amp // 2), 0, 255).astype(np.uint8)
        top = g.base_off + p.plane_off + (h - 1) * p.ystride  # top-left pixel of the picture
        for y in range(h):  # memory row y (top-down) = internal row h-1-y
            buf[top + y * stride: top + y * stride + w] = img[h - 1 - y]
    return buf


@pytest.mark.parametrize("level", [2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("dims", [(176, 144, 0), (320, 64, 3), (64, 256, 2), (1920, 1088, 0)])
def test_device_filters_match_the_oracle_on_synthetic_frames(dims, level):
    """ocg_pp_run through the C ABI against the oracle's restatement (itself pinned to the reference by
    tests/test_oracle_postproc.py): random quantiser indices and tables, frames with every variance class."""
    import theora_b200 as T
    from theora_b200 import abi
    rng = np.random.default_rng(1000 * level + dims[0])
    g = S.make_geometry(dims[0], dims[1], dims[2], 3)
    buf = _smoothish_frame(g, rng)
    nf = g.nfrags
    dc_qis = rng.integers(0, 64, nf).astype(np.uint8)
    qis = rng.integers(0, 64, nf).astype(np.uint8)
    dc_scale = rng.integers(1, 120, 64).astype(np.int32)
    sharp_mod = (-rng.integers(0, 64, 64)).astype(np.int32)
    O = S.oracle()
    O.oco_pp_deblock_plane.restype = None
    O.oco_pp_dering_plane.restype = None
    O.oco_pp_deblock_plane.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    O.oco_pp_dering_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    want, want_var = [], []
    for pli in range(3):
        p = g.planes[pli]
        w, h, stride = p.width, p.height, -p.ystride
        top = g.base_off + p.plane_off + (h - 1) * p.ystride
        img = np.stack([buf[top + y * stride: top + y * stride + w] for y in range(h)])  # top-down
        off = 3 * (pli != 0)
        if level >= 2 + off:
            s = np.ascontiguousarray(img[::-1])
            d = np.empty_like(s)
            var = np.zeros(p.nfrags, np.int32)
            dq = np.ascontiguousarray(dc_qis[p.froffset:p.froffset + p.nfrags])
            qq = np.ascontiguousarray(qis[p.froffset:p.froffset + p.nfrags])
            O.oco_pp_deblock_plane(d.ctypes.data, w, s.ctypes.data, w, w, h, dq.ctypes.data, dc_scale.ctypes.data, var.ctypes.data)
            if level >= 3 + off:
                O.oco_pp_dering_plane(d.ctypes.data, w, w, h, pli, int(level >= 4 + off), qq.ctypes.data, dc_scale.ctypes.data,
                                      sharp_mod.ctypes.data, var.ctypes.data)
            want.append(np.ascontiguousarray(d[::-1]).ravel())
            want_var.append(var)
    L = abi.lib()
    ctx = T.Context(g, 0)
    ctx.upload_frame(0, buf)
    pp = C.c_void_p()
    abi.check(L.ocg_pp_create(C.byref(pp), ctx.h))
    try:
        abi.check(L.ocg_pp_run(pp, 0, level, dc_scale.ctypes.data, sharp_mod.ctypes.data, dc_qis.ctypes.data, qis.ctypes.data))
        total = sum(g.planes[i].width * g.planes[i].height for i in range(3))
        out = np.zeros(total, np.uint8)
        abi.check(L.ocg_pp_download(pp, out.ctypes.data))
        var = np.zeros(nf, np.int32)
        abi.check(L.ocg_pp_download_variances(pp, var.ctypes.data))
    finally:
        L.ocg_pp_destroy(pp)
        ctx.close()
    at = 0
    for k, wnt in enumerate(want):
        got = out[at:at + wnt.size]
        assert np.array_equal(got, wnt), "plane %d differs at %s" % (k, np.flatnonzero(got != wnt)[:5])
        p = g.planes[k]
        assert np.array_equal(var[p.froffset:p.froffset + p.nfrags], want_var[k]), "variances of plane %d differ" % k
        at += wnt.size
    classes = np.concatenate(want_var)
    assert (classes > 3840).any() and (classes <= 384).any(), "the frame does not exercise the variance classes"
