"""Pins the CPU restatement (oracle/theora_oracle.c) bit-for-bit against the
compiled, unmodified reference (oracle/_ref/libth_c.so) on randomised and
adversarial inputs.  Skipped where the reference build is absent; the committed
golden vectors (test_oracle_golden.py) cover that case."""
import ctypes as C

import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.skipif(not S.ref_available("c"), reason="oracle/_ref not built")


def rnd_coeffs(rng, n, last_zzi, big=False):
    fz = FZ
    x = np.zeros((n, 64), np.int16)
    nz = 64 if last_zzi > 10 else last_zzi
    amp = 32767 if big else 600
    for i in range(n):
        k = rng.integers(1, nz + 1)
        pos = fz[rng.choice(nz, size=min(k, nz), replace=False)]
        x[i, pos] = rng.integers(-amp, amp + 1, size=len(pos))
    return x


FZ = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,
               14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39,
               46, 53, 60, 61, 54, 47, 55, 62, 63])


@pytest.mark.parametrize("last_zzi", [0, 1, 2, 3, 4, 10, 11, 30, 64])
@pytest.mark.parametrize("big", [False, True])
def test_idct(last_zzi, big):
    rng = np.random.default_rng(last_zzi * 2 + big)
    O, R = S.oracle(), S.ref("c")
    x = rnd_coeffs(rng, 400, max(last_zzi, 1), big)
    if big:
        # garbage outside the class footprint must be ignored identically
        x[::3] = rng.integers(-32768, 32768, size=x[::3].shape)
    for i in range(len(x)):
        a, b = x[i].copy(), x[i].copy()
        ya, yb = np.zeros(64, np.int16), np.zeros(64, np.int16)
        O.oco_idct8x8(S.ptr(ya, S.i16p), S.ptr(a, S.i16p), last_zzi)
        R.oc_idct8x8_c(S.ptr(yb, S.i16p), S.ptr(b, S.i16p), last_zzi)
        assert np.array_equal(ya, yb)
        assert np.array_equal(a, b)  # input-clearing side effect


@pytest.mark.parametrize("fmt", [0, 2, 3])
def test_mv_offsets(fmt):
    O, R = S.oracle(), S.ref("c")
    for pli in range(3):
        for dy in range(-31, 32):
            for dx in range(-31, 32):
                mv = np.int16(((dy & 0xFF) << 8 | (dx & 0xFF)) - (65536 if dy < 0 else 0))
                oa, ob = (C.c_int * 2)(0, 0), (C.c_int * 2)(0, 0)
                na = O.oco_mv_offsets(oa, -1952, pli, fmt, int(mv))
                nb = R.refh_mv_offsets(fmt, -1952, pli, int(mv), ob)
                assert na == nb and oa[0] == ob[0] and (na == 1 or oa[1] == ob[1]), (pli, dx, dy)


def test_state_frag_recon():
    rng = np.random.default_rng(7)
    O, R = S.oracle(), S.ref("c")
    stride = 64
    for it in range(600):
        last_zzi = int(rng.choice([0, 1, 2, 3, 5, 10, 11, 40, 64]))
        x = rnd_coeffs(rng, 1, max(last_zzi, 1), big=bool(it % 5 == 0))[0]
        x[0] = rng.integers(-2000, 2000)
        dcq = int(rng.integers(1, 4000))
        intra = int(rng.integers(0, 2))
        pli = int(rng.integers(0, 3))
        dx, dy = int(rng.integers(-31, 32)), int(rng.integers(-31, 32))
        mv = ((dy & 0xFF) << 8 | (dx & 0xFF))
        mv = mv - 65536 if mv >= 32768 else mv
        ref = rng.integers(0, 256, size=(64, stride), dtype=np.uint8)
        outs = []
        for fn, lib in ((O.oco_state_frag_recon, "o"), (R.refh_state_frag_recon, "r")):
            dst = np.full((64, stride), 77, np.uint8)
            co = np.zeros(128, np.int16)
            co[:64] = x
            # bottom-left pixel = last row; fragment in the middle
            bl = (63 * stride)
            off = -28 * stride + 24
            fn(C.cast(dst.ctypes.data + bl, S.u8p) if lib == "o" else dst.ctypes.data + bl,
               C.cast(ref.ctypes.data + bl, S.u8p) if lib == "o" else ref.ctypes.data + bl,
               off, -stride, pli, 0, intra, mv, S.ptr(co, S.i16p), last_zzi, dcq)
            outs.append(dst)
        assert np.array_equal(outs[0], outs[1]), (it, last_zzi, intra, dx, dy)


@pytest.mark.parametrize("limit", [1, 2, 5, 13, 40, 127])
def test_loop_filter_table(limit):
    O, R = S.oracle(), S.ref("c")
    a, b = (C.c_byte * 256)(), (C.c_byte * 256)()
    O.oco_loop_filter_init(a, limit)
    R.refh_loop_filter_table(b, limit)
    assert list(a) == list(b)


@pytest.mark.parametrize("density", [0.0, 0.2, 0.5, 0.8, 1.0])
def test_loop_filter_plane(density):
    rng = np.random.default_rng(int(density * 10))
    O, R = S.oracle(), S.ref("c")
    for it in range(40):
        nh, nv = int(rng.integers(1, 12)), int(rng.integers(1, 9))
        limit = int(rng.integers(1, 41))
        stride = nh * 8 + 16
        if it % 2:
            img = rng.integers(0, 256, size=(nv * 8, stride), dtype=np.uint8)
        else:
            img = (rng.integers(0, 2, size=(nv * 8, stride)) * 255).astype(np.uint8)
        coded = (rng.random(nh * nv) < density).astype(np.uint8)
        res = []
        for fn in (R.refh_loop_filter_plane, O.oco_loop_filter_plane_seq, O.oco_loop_filter_plane_cells):
            p = img.copy()
            fn(p.ctypes.data + (nv * 8 - 1) * stride + 8, -stride, nh, nv, S.ptr(coded, S.u8p), limit)
            res.append(p)
        assert np.array_equal(res[0], res[1]), "sequential oracle != reference"
        assert np.array_equal(res[0], res[2]), "cell decomposition != reference"


def test_borders_fill():
    rng = np.random.default_rng(3)
    O, R = S.oracle(), S.ref("c")
    for pli, fmt, hp, vp in ((0, 0, 16, 16), (1, 0, 8, 8), (2, 2, 8, 16), (1, 3, 16, 16)):
        w, h = 48, 32
        stride = w + 2 * hp
        a = rng.integers(0, 256, size=(h + 2 * vp, stride), dtype=np.uint8)
        b = a.copy()
        bl = (vp + h - 1) * stride + hp
        O.oco_borders_fill_plane(a.ctypes.data + bl, -stride, w, h, hp, vp)
        R.refh_borders_fill(b.ctypes.data + bl, -stride, w, h, pli, fmt)
        assert np.array_equal(a, b)


def test_fdct_quantize():
    rng = np.random.default_rng(11)
    O, R = S.oracle(), S.ref("c")
    for it in range(1500):
        amp = int(rng.choice([3, 40, 255]))
        x = rng.integers(-amp, amp + 1, size=64).astype(np.int16)
        if it % 50 == 0:
            x[:] = 0
        ya, yb = np.zeros(64, np.int16), np.zeros(64, np.int16)
        O.oco_fdct8x8(S.ptr(ya, S.i16p), S.ptr(x, S.i16p))
        R.oc_enc_fdct8x8_c(S.ptr(yb, S.i16p), S.ptr(x, S.i16p))
        assert np.array_equal(ya, yb)
        deq = rng.integers(2, 2000, size=64).astype(np.uint16)
        if it % 7 == 0:
            deq = rng.integers(1, 65536 // 2, size=64).astype(np.uint16)
        ea, eb = np.zeros(128, np.int16), np.zeros(128, np.int16)
        O.oco_enquant_init(S.ptr(ea, S.i16p), S.ptr(deq, S.u16p))
        R.oc_enc_enquant_table_init_c(eb.ctypes.data, S.ptr(deq, S.u16p))
        assert np.array_equal(ea, eb)
        qa, qb = np.zeros(64, np.int16), np.zeros(64, np.int16)
        na = O.oco_quantize(S.ptr(qa, S.i16p), S.ptr(ya, S.i16p), S.ptr(deq, S.u16p), S.ptr(ea, S.i16p))
        nb = R.oc_enc_quantize_c(S.ptr(qb, S.i16p), S.ptr(yb, S.i16p), S.ptr(deq, S.u16p), eb.ctypes.data)
        assert na == nb and np.array_equal(qa, qb)


def test_block_metrics():
    rng = np.random.default_rng(5)
    O, R = S.oracle(), S.ref("c")
    for name in ("sad", "ssd"):
        getattr(R, "oc_enc_frag_%s_c" % name).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        getattr(R, "oc_enc_frag_%s_c" % name).restype = C.c_uint
    R.oc_enc_frag_sad_thresh_c.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
    R.oc_enc_frag_sad2_thresh_c.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
    R.oc_enc_frag_intra_sad_c.argtypes = [C.c_void_p, C.c_int]
    R.oc_enc_frag_satd_c.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_int]
    R.oc_enc_frag_satd2_c.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    R.oc_enc_frag_intra_satd_c.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_int]
    R.oc_enc_frag_border_ssd_c.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64]
    R.oc_enc_frag_copy2_c.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    R.oc_enc_frag_sub_c.argtypes = [S.i16p, C.c_void_p, C.c_void_p, C.c_int]
    R.oc_enc_frag_sub_128_c.argtypes = [S.i16p, C.c_void_p, C.c_int]
    for fn in (R.oc_enc_frag_sad_thresh_c, R.oc_enc_frag_sad2_thresh_c, R.oc_enc_frag_intra_sad_c,
               R.oc_enc_frag_satd_c, R.oc_enc_frag_satd2_c, R.oc_enc_frag_intra_satd_c, R.oc_enc_frag_border_ssd_c):
        fn.restype = C.c_uint
    st = 40
    for it in range(500):
        mode = it % 3
        if mode == 0:
            imgs = [rng.integers(0, 256, size=(8, st), dtype=np.uint8) for _ in range(3)]
        elif mode == 1:
            base = rng.integers(0, 256, size=(8, st))
            imgs = [np.clip(base + rng.integers(-6, 7, size=(8, st)), 0, 255).astype(np.uint8) for _ in range(3)]
        else:
            imgs = [(rng.integers(0, 2, size=(8, st)) * 255).astype(np.uint8) for _ in range(3)]
        s, r1, r2 = (a.ctypes.data + 7 * st + 3 for a in imgs)  # bottom-up, negative stride
        ys = -st
        assert O.oco_frag_sad(s, r1, ys) == R.oc_enc_frag_sad_c(s, r1, ys)
        assert O.oco_frag_ssd(s, r1, ys) == R.oc_enc_frag_ssd_c(s, r1, ys)
        th = int(rng.integers(0, 3000))
        assert O.oco_frag_sad_thresh(s, r1, ys, th) == R.oc_enc_frag_sad_thresh_c(s, r1, ys, th)
        assert O.oco_frag_sad2_thresh(s, r1, r2, ys, th) == R.oc_enc_frag_sad2_thresh_c(s, r1, r2, ys, th)
        assert O.oco_frag_intra_sad(s, ys) == R.oc_enc_frag_intra_sad_c(s, ys)
        da, db = C.c_int(0), C.c_int(0)
        assert O.oco_frag_satd(C.byref(da), s, r1, ys) == R.oc_enc_frag_satd_c(C.byref(db), s, r1, ys)
        assert da.value == db.value
        assert O.oco_frag_satd2(C.byref(da), s, r1, r2, ys) == R.oc_enc_frag_satd2_c(C.byref(db), s, r1, r2, ys)
        assert da.value == db.value
        assert O.oco_frag_intra_satd(C.byref(da), s, ys) == R.oc_enc_frag_intra_satd_c(C.byref(db), s, ys)
        assert da.value == db.value
        mask = int(rng.integers(0, 2 ** 63))
        assert O.oco_frag_border_ssd(s, r1, ys, mask) == R.oc_enc_frag_border_ssd_c(s, r1, ys, mask)
        a, b = np.zeros((8, st), np.uint8), np.zeros((8, st), np.uint8)
        O.oco_frag_copy2(a.ctypes.data + 7 * st, r1, r2, ys)
        R.oc_enc_frag_copy2_c(b.ctypes.data + 7 * st, r1, r2, ys)
        assert np.array_equal(a, b)
        xa, xb = np.zeros(64, np.int16), np.zeros(64, np.int16)
        O.oco_frag_sub(S.ptr(xa, S.i16p), s, r1, ys)
        R.oc_enc_frag_sub_c(S.ptr(xb, S.i16p), s, r1, ys)
        assert np.array_equal(xa, xb)
        O.oco_frag_sub_128(S.ptr(xa, S.i16p), s, ys)
        R.oc_enc_frag_sub_128_c(S.ptr(xb, S.i16p), s, ys)
        assert np.array_equal(xa, xb)
