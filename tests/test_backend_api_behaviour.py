"""API behaviour of the integrated library that differs from / must match the
reference: post-processing levels are accepted, the stripe
callback is delivered once per frame, dropped (0-byte) packets are
TH_DUPFRAME, a missing device fails th_decode_alloc.  CPU-only parts use the
recorder mode; the GPU parts are marked."""
import ctypes as C
import os

import numpy as np
import pytest

import support as S
from theora_b200 import abi
import th_streams as streams

G = np.load(os.path.join(S.GOLDEN_DIR, "streams.npz"))
pytestmark = pytest.mark.skipif(not streams.available(), reason="integrated build not present")

TH_DECCTL_GET_PPLEVEL_MAX, TH_DECCTL_SET_PPLEVEL, TH_DECCTL_SET_STRIPE_CB = 1, 3, 7
TH_EIMPL, TH_DUPFRAME = -23, 1


def open_decoder(blob, mode):
    L = streams.lib()
    L.ocg_backend_set_mode(mode)
    buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
    sh = L.refh_stream_from_blob(buf, len(blob))
    d = L.refh_dec_open(sh)
    L.ocg_backend_set_mode(streams.BACKEND_GPU)
    return L, sh, d, buf


def test_postprocessing_levels_are_accepted_like_the_reference():
    """TH_DECCTL_SET_PPLEVEL / GET_PPLEVEL_MAX behave as in decode.c:1980-2000 (the filters themselves run
    on the host after the flush: tests/test_gpu_postproc.py)."""
    L, sh, d, _ = open_decoder(G["s64_q48_blob"].tobytes(), streams.BACKEND_RECORD)
    assert d
    L.refh_dec_ctx.restype = C.c_void_p
    L.refh_dec_ctx.argtypes = [C.c_void_p]
    L.th_decode_ctl.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    ctx = L.refh_dec_ctx(d)
    v = C.c_int(-1)
    assert L.th_decode_ctl(ctx, TH_DECCTL_GET_PPLEVEL_MAX, C.byref(v), C.sizeof(v)) == 0 and v.value == 7
    for lvl in (2, 7, 0):
        v = C.c_int(lvl)
        assert L.th_decode_ctl(ctx, TH_DECCTL_SET_PPLEVEL, C.byref(v), C.sizeof(v)) == 0
    v = C.c_int(8)
    assert L.th_decode_ctl(ctx, TH_DECCTL_SET_PPLEVEL, C.byref(v), C.sizeof(v)) == -10  # TH_EINVAL: out of range
    v = C.c_int(0)
    assert L.th_decode_ctl(ctx, TH_DECCTL_SET_PPLEVEL, C.byref(v), 1) == -10  # TH_EINVAL, as decode.c:1985
    L.refh_dec_close(d)
    L.refh_stream_free(sh)


def test_alloc_fails_without_a_device():
    if abi.lib().ocg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    L, sh, d, _ = open_decoder(G["s64_q48_blob"].tobytes(), streams.BACKEND_GPU)
    assert not d, "th_decode_alloc must fail (no CPU fallback) when no device is present"
    L.refh_stream_free(sh)


def blob_with_empty_packet(blob, after):
    hdr = np.frombuffer(blob[:16], np.uint32)
    npk = int(hdr[1])
    sizes = np.frombuffer(blob[16:16 + 4 * npk], np.uint32).copy()
    offs = (np.concatenate([[0], np.cumsum(sizes)]) + 16 + 4 * npk).astype(np.int64).tolist()
    parts = [blob[offs[i]:offs[i + 1]] for i in range(npk)]
    parts.insert(after, b"")
    new_sizes = np.array([len(p) for p in parts], np.uint32)
    return np.array([hdr[0], len(parts), hdr[2], hdr[3]], np.uint32).tobytes() + new_sizes.tobytes() + b"".join(parts)


@pytest.mark.gpu
def test_dropped_packet_is_dupframe_and_keeps_the_picture():
    blob = blob_with_empty_packet(G["s64_q32_kf4_blob"].tobytes(), 5)  # after two data packets
    want = G["s64_q32_kf4_hashes"]
    L, sh, d, _ = open_decoder(blob, streams.BACKEND_GPU)
    assert d
    hashes, rets = [], []
    for _ in range(7):
        rets.append(L.refh_dec_next(d))
        h = (C.c_uint64 * 3)()
        L.refh_dec_hash(d, h)
        hashes.append([int(x) for x in h])
    assert rets[2] == TH_DUPFRAME and all(r == 0 for i, r in enumerate(rets) if i != 2)
    assert hashes[2] == hashes[1]  # picture unchanged by the dropped frame
    got = [hashes[i] for i in range(7) if i != 2]
    assert got == [[int(x) for x in row] for row in want]
    L.refh_dec_close(d)
    L.refh_stream_free(sh)


@pytest.mark.gpu
def test_stripe_callback_gets_the_whole_final_frame_once_per_frame():
    blob = G["qcif_q20_blob"].tobytes()
    want = G["qcif_q20_hashes"]
    L, sh, d, _ = open_decoder(blob, streams.BACKEND_GPU)
    assert d
    L.refh_dec_ctx.restype = C.c_void_p
    L.refh_dec_ctx.argtypes = [C.c_void_p]
    L.th_decode_ctl.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]

    class ImgPlane(C.Structure):
        _fields_ = [("width", C.c_int), ("height", C.c_int), ("stride", C.c_int), ("data", C.POINTER(C.c_uint8))]

    CB = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(ImgPlane), C.c_int, C.c_int)
    calls = []

    def on_stripe(ctx, planes, y0, yend):
        p = planes[0]
        rows = np.ctypeslib.as_array(p.data, shape=(p.height * p.stride,))[:p.height * p.stride]
        luma = rows.reshape(p.height, p.stride)[:, :p.width]
        calls.append((y0, yend, S.fnv1a64(luma)))

    cb = CB(on_stripe)

    class StripeCb(C.Structure):
        _fields_ = [("ctx", C.c_void_p), ("stripe_decoded", CB)]

    scb = StripeCb(None, cb)
    assert L.th_decode_ctl(L.refh_dec_ctx(d), TH_DECCTL_SET_STRIPE_CB, C.byref(scb), C.sizeof(scb)) == 0
    n = len(want)
    for _ in range(n):
        assert L.refh_dec_next(d) >= 0
    assert len(calls) == n
    for i, (y0, yend, h) in enumerate(calls):
        assert (y0, yend) == (0, 144 // 8)
        assert h == int(want[i][0]), i  # the luma handed to the callback is the final frame
    L.refh_dec_close(d)
    L.refh_stream_free(sh)


def test_intra_only_encoder_alloc_fails_without_a_device():
    """keyframe_granule_shift == 0 selects the device encoder: no device, no encoder."""
    if abi.lib().ocg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    L = streams.lib()
    assert not L.refh_encode_synth(64, 64, 0, 2, 48, 1, 1, 30, 12345)


def test_host_mode_encoder_matches_the_reference_bitstream():
    """Tooling mode (OCG_ENC_HOST): the integrated build's encoder is the reference's."""
    if not S.ref_available("c"):
        pytest.skip("needs oracle/_ref")
    L = streams.lib()
    R = S.ref("c")
    L.ocg_backend_set_enc_mode(streams.ENC_HOST)
    try:
        got = S.Stream.encode(L, 64, 64, 3, quality=32, kf=1, speed=1, noise_shift=28)
    finally:
        L.ocg_backend_set_enc_mode(streams.ENC_AUTO)
    want = S.Stream.encode(R, 64, 64, 3, quality=32, kf=1, speed=1, noise_shift=28)
    assert got.to_bytes() == want.to_bytes()
    got.free()
    want.free()
