"""Generates tests/golden/*.npz from the compiled, UNMODIFIED reference
(oracle/_ref/libth_c.so, built from /root/reference by oracle/Makefile).

Run in the build container (needs oracle/_ref):  python tests/golden/make_golden.py
The vectors pin the oracle (and through it the CUDA path) on boxes where the
reference build is not present.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import support as S  # noqa: E402

FZ = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,
               14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39,
               46, 53, 60, 61, 54, 47, 55, 62, 63])


def units():
    R = S.ref("c")
    rng = np.random.default_rng(20261017)
    out = {}
    # --- iDCT (idct.c:301) -------------------------------------------------
    n = 96
    x = np.zeros((n, 64), np.int16)
    lz = rng.choice([0, 1, 2, 3, 4, 10, 11, 30, 64], size=n).astype(np.int32)
    for i in range(n):
        nz = 64 if lz[i] > 10 else max(int(lz[i]), 1)
        pos = FZ[rng.choice(nz, size=rng.integers(1, nz + 1), replace=False)]
        x[i, pos] = rng.integers(-2000, 2001, size=len(pos))
        if i % 4 == 0:
            x[i] = rng.integers(-32768, 32768, size=64)
    y = np.zeros((n, 64), np.int16)
    xc = x.copy()
    for i in range(n):
        R.oc_idct8x8_c(S.ptr(y[i], S.i16p), S.ptr(xc[i], S.i16p), int(lz[i]))
    out.update(idct_x=x, idct_lz=lz, idct_y=y, idct_x_after=xc)
    # --- MV offsets (state.c:846), all vectors, both plane kinds, 3 formats --
    mvs = []
    for fmt in (0, 2, 3):
        for pli in (0, 1):
            for dy in range(-31, 32):
                for dx in range(-31, 32):
                    mv = ((dy & 0xFF) << 8) | (dx & 0xFF)
                    mv = mv - 65536 if mv >= 32768 else mv
                    o = (C.c_int * 2)(0, 0)
                    k = R.refh_mv_offsets(fmt, -976, pli, mv, o)
                    mvs.append((fmt, pli, mv, k, o[0], o[1] if k == 2 else 0))
    out["mv_table"] = np.array(mvs, np.int32)
    # --- loop filter planes (state.c:1055) ---------------------------------
    lf_in, lf_out, lf_meta, lf_coded = [], [], [], []
    for it in range(24):
        nh, nv = int(rng.integers(1, 9)), int(rng.integers(1, 7))
        limit = int(rng.choice([1, 2, 5, 13, 40, 127]))
        stride = nh * 8 + 16
        img = rng.integers(0, 256, size=(nv * 8, stride), dtype=np.uint8) if it % 2 else \
            (rng.integers(0, 2, size=(nv * 8, stride)) * 255).astype(np.uint8)
        coded = (rng.random(nh * nv) < rng.random()).astype(np.uint8)
        o = img.copy()
        R.refh_loop_filter_plane(o.ctypes.data + (nv * 8 - 1) * stride + 8, -stride, nh, nv, S.ptr(coded, S.u8p), limit)
        lf_in.append(img.ravel()); lf_out.append(o.ravel()); lf_coded.append(coded)
        lf_meta.append((nh, nv, limit, stride))
    out["lf_in"] = np.concatenate(lf_in); out["lf_out"] = np.concatenate(lf_out)
    out["lf_coded"] = np.concatenate(lf_coded); out["lf_meta"] = np.array(lf_meta, np.int32)
    # --- fDCT + quantise (fdct.c:128, enquant.c:184-249) ---------------------
    n = 64
    fx = rng.integers(-255, 256, size=(n, 64)).astype(np.int16)
    fx[::8] //= 16
    fy = np.zeros((n, 64), np.int16)
    deq = rng.integers(2, 1500, size=(n, 64)).astype(np.uint16)
    enq = np.zeros((n, 128), np.int16)
    q = np.zeros((n, 64), np.int16)
    nzz = np.zeros(n, np.int32)
    for i in range(n):
        R.oc_enc_fdct8x8_c(S.ptr(fy[i], S.i16p), S.ptr(fx[i], S.i16p))
        R.oc_enc_enquant_table_init_c(enq[i].ctypes.data, S.ptr(deq[i], S.u16p))
        nzz[i] = R.oc_enc_quantize_c(S.ptr(q[i], S.i16p), S.ptr(fy[i], S.i16p), S.ptr(deq[i], S.u16p), enq[i].ctypes.data)
    out.update(fdct_x=fx, fdct_y=fy, q_deq=deq, q_enq=enq, q_out=q, q_last=nzz)
    # --- block metrics (encfrag.c) -------------------------------------------
    R.oc_enc_frag_sad_c.restype = C.c_uint
    R.oc_enc_frag_ssd_c.restype = C.c_uint
    R.oc_enc_frag_satd_c.restype = C.c_uint
    R.oc_enc_frag_satd2_c.restype = C.c_uint
    R.oc_enc_frag_intra_satd_c.restype = C.c_uint
    R.oc_enc_frag_intra_sad_c.restype = C.c_uint
    R.oc_enc_frag_sad2_thresh_c.restype = C.c_uint
    for f in (R.oc_enc_frag_sad_c, R.oc_enc_frag_ssd_c):
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    R.oc_enc_frag_satd_c.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_int]
    R.oc_enc_frag_satd2_c.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    R.oc_enc_frag_intra_satd_c.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_int]
    R.oc_enc_frag_intra_sad_c.argtypes = [C.c_void_p, C.c_int]
    R.oc_enc_frag_sad2_thresh_c.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
    n = 64
    blocks = rng.integers(0, 256, size=(n, 3, 8, 8), dtype=np.uint8)
    blocks[n // 2:, 1:] = np.clip(blocks[n // 2:, :1].astype(int) + rng.integers(-4, 5, size=(n - n // 2, 2, 8, 8)), 0, 255)
    met = np.zeros((n, 10), np.int64)
    for i in range(n):
        s, r1, r2 = (blocks[i, k].ctypes.data for k in range(3))
        dc = C.c_int(0)
        met[i, 0] = R.oc_enc_frag_sad_c(s, r1, 8)
        met[i, 1] = R.oc_enc_frag_sad2_thresh_c(s, r1, r2, 8, 0xFFFFFFFF)
        met[i, 2] = R.oc_enc_frag_satd_c(C.byref(dc), s, r1, 8); met[i, 3] = dc.value
        met[i, 4] = R.oc_enc_frag_satd2_c(C.byref(dc), s, r1, r2, 8); met[i, 5] = dc.value
        met[i, 6] = R.oc_enc_frag_intra_satd_c(C.byref(dc), s, 8); met[i, 7] = dc.value
        met[i, 8] = R.oc_enc_frag_ssd_c(s, r1, 8)
        met[i, 9] = R.oc_enc_frag_intra_sad_c(s, 8)
    out.update(met_blocks=blocks, met_out=met)
    np.savez_compressed(os.path.join(HERE, "units.npz"), **out)
    print("units.npz:", {k: v.shape for k, v in out.items()})


STREAMS = [
    ("s64_q48", (64, 64, 2, 48, 64, 1, 30)),       # BASELINE configs[0] shape: 64x64, 2 frames
    ("s64_q32_kf4", (64, 64, 6, 32, 4, 1, 28)),
    ("qcif_q20", (176, 144, 6, 20, 64, 1, 28)),
    ("crop_350x270", (350, 270, 4, 40, 64, 1, 30)),
]


def streams():
    R = S.ref("c")
    out = {}
    for name, (w, h, n, q, kf, sp, ns) in STREAMS:
        st = S.Stream.encode(R, w, h, n, quality=q, kf=kf, speed=sp, noise_shift=ns)
        blob = st.to_bytes()
        dec = S.Decoder(R, st)
        hashes = []
        for i in range(n):
            assert dec.next() >= 0
            hashes.append(dec.hashes())
        dec.close()
        st.free()
        out[name + "_blob"] = np.frombuffer(blob, np.uint8)
        out[name + "_hashes"] = np.array(hashes, np.uint64)
    np.savez_compressed(os.path.join(HERE, "streams.npz"), **out)
    print("streams.npz:", {k: v.shape for k, v in out.items()})


def me_frame():
    """Whole-frame motion analysis by the reference (oc_mcenc_search + refinements over every macro block,
    oracle/ref_internal_harness.c refh_me_frame) on the deterministic scene of tests/test_oracle_me_frame.py;
    also oc_mb_activity on one of its frames through oracle/_ref/libth_c_analyze.so."""
    import megen
    import test_oracle_me_frame as T
    R = megen.bind_ref_me(S.ref("c"))
    out = {}
    for t, fl, topo, got, want in T.run_sequence(**T.GOLDEN_CASE, ref=R):
        out["frame%d" % t] = want.view(np.uint8).copy()
    np.savez_compressed(os.path.join(HERE, "me_frame.npz"), **out)
    print("me_frame.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    units()
    streams()
    me_frame()
