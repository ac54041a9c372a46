"""The oracle's restatement of the out-of-loop post-processing filters (oco_pp_deblock_plane,
oco_pp_dering_plane; decode.c:1609-1957) pinned against the compiled reference: a stream is decoded by the
unmodified reference twice, with and without TH_DECCTL_SET_PPLEVEL; the oracle filters applied to the
unprocessed frames -- with the per-fragment inputs the reference decoder itself holds (tracked DC quantiser
indices, quantiser indices, tables) -- must reproduce the post-processed frames, every frame, every level.
CPU only."""
import ctypes as C

import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.skipif(not S.ref_available("c"), reason="needs oracle/_ref (built from /root/reference)")


def planes(frame, fw, fh, fmt):
    cw = fw >> (0 if fmt & 1 else 1)
    ch = fh >> (0 if fmt & 2 else 1)
    y = frame[:fw * fh].reshape(fh, fw)
    cb = frame[fw * fh:fw * fh + cw * ch].reshape(ch, cw)
    cr = frame[fw * fh + cw * ch:].reshape(ch, cw)
    return [y, cb, cr]


def oracle_pp(O, frame, fw, fh, fmt, level, dc_qis, qis, dc_scale, sharp_mod):
    out = frame.copy()
    src = planes(frame, fw, fh, fmt)
    dst = planes(out, fw, fh, fmt)
    froff = 0
    all_var = []
    for pli in range(3):
        h, w = src[pli].shape
        nfr = (w >> 3) * (h >> 3)
        off = 3 * (pli != 0)
        if level >= 2 + off:
            s = np.ascontiguousarray(src[pli][::-1])  # internal orientation: row 0 = bottom row
            d = np.empty_like(s)
            var = np.zeros(nfr, np.int32)
            dq = np.ascontiguousarray(dc_qis[froff:froff + nfr])
            qq = np.ascontiguousarray(qis[froff:froff + nfr])
            O.oco_pp_deblock_plane(d.ctypes.data, w, s.ctypes.data, w, w, h, dq.ctypes.data, dc_scale.ctypes.data, var.ctypes.data)
            if level >= 3 + off:
                O.oco_pp_dering_plane(d.ctypes.data, w, w, h, pli, int(level >= 4 + off), qq.ctypes.data, dc_scale.ctypes.data,
                                      sharp_mod.ctypes.data, var.ctypes.data)
            dst[pli][:] = d[::-1]
            all_var.append(var)
        froff += nfr
    return out, all_var


@pytest.mark.parametrize("level", [2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("case", [(176, 144, 5, 10, 4, 28, 0), (352, 288, 4, 24, 64, 30, 0), (208, 112, 4, 5, 3, 26, 0),
                                  (176, 144, 3, 8, 64, 27, 3), (176, 144, 3, 8, 64, 27, 2)])
def test_oracle_postprocessing_matches_reference(case, level):
    w, h, n, q, kf, ns, fmt = case
    R = S.ref("c")
    O = S.oracle()
    for f in (O.oco_pp_deblock_plane, O.oco_pp_dering_plane):
        f.restype = None
    O.oco_pp_deblock_plane.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    O.oco_pp_dering_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    R.refh_dec_ctx.restype = C.c_void_p
    R.refh_dec_ctx.argtypes = [C.c_void_p]
    R.refh_dec_pp_state.restype = C.c_long
    R.refh_dec_pp_state.argtypes = [C.c_void_p] * 7
    st = S.Stream.encode(R, w, h, n, quality=q, kf=kf, speed=1, noise_shift=ns, fmt=fmt)
    plain, post = S.Decoder(R, st), S.Decoder(R, st)
    assert post.set_pplevel(level) == 0
    changed = 0
    for i in range(n):
        assert plain.next() >= 0 and post.next() >= 0
        f0, fl = plain.frame(), post.frame()
        nfr_max = (post.fw >> 3) * (post.fh >> 3) * 3
        dc_qis, qis = np.zeros(nfr_max, np.uint8), np.zeros(nfr_max, np.uint8)
        dc_scale, sharp_mod = np.zeros(64, np.int32), np.zeros(64, np.int32)
        ref_var = np.zeros(nfr_max, np.int32)
        used = C.c_int(0)
        nfr = R.refh_dec_pp_state(R.refh_dec_ctx(post.d), dc_qis.ctypes.data, qis.ctypes.data, dc_scale.ctypes.data,
                                  sharp_mod.ctypes.data, ref_var.ctypes.data, C.byref(used))
        if nfr < 0 or used.value < 2:
            assert np.array_equal(f0, fl)
            continue
        assert used.value == level
        want, var = oracle_pp(O, f0, post.fw, post.fh, post.fmt, level, dc_qis[:nfr], qis[:nfr], dc_scale, sharp_mod)
        assert np.array_equal(want, fl), "level %d, frame %d: oracle filters differ from the reference's" % (level, i)
        # the variances the reference accumulated (luma; chroma too from level 5)
        nl = (post.fw >> 3) * (post.fh >> 3)
        assert np.array_equal(var[0], ref_var[:nl])
        changed += int(not np.array_equal(f0, fl))
    assert changed > 0, "the filters changed nothing"
    plain.close()
    post.close()
    st.free()
