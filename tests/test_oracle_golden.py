"""The CPU oracle against the committed golden vectors (tests/golden/*.npz,
generated from the unmodified reference by tests/golden/make_golden.py).  Runs
everywhere, including boxes without the reference build."""
import ctypes as C
import os

import numpy as np
import pytest

import support as S
import th_streams as streams

U = np.load(os.path.join(S.GOLDEN_DIR, "units.npz"))
G = np.load(os.path.join(S.GOLDEN_DIR, "streams.npz"))
STREAM_NAMES = sorted(k[:-5] for k in G.files if k.endswith("_blob"))


def test_idct_golden():
    O = S.oracle()
    for i in range(len(U["idct_lz"])):
        x = U["idct_x"][i].copy()
        y = np.zeros(64, np.int16)
        O.oco_idct8x8(S.ptr(y, S.i16p), S.ptr(x, S.i16p), int(U["idct_lz"][i]))
        assert np.array_equal(y, U["idct_y"][i]), i
        assert np.array_equal(x, U["idct_x_after"][i]), i


def test_mv_offsets_golden():
    O = S.oracle()
    for fmt, pli, mv, k, o0, o1 in U["mv_table"][::7].tolist():
        o = (C.c_int * 2)(0, 0)
        assert O.oco_mv_offsets(o, -976, pli, fmt, mv) == k
        assert o[0] == o0 and (k == 1 or o[1] == o1)


def test_loop_filter_golden():
    O = S.oracle()
    pos = cpos = 0
    for nh, nv, limit, stride in U["lf_meta"].tolist():
        n = nv * 8 * stride
        img = U["lf_in"][pos:pos + n].reshape(nv * 8, stride)
        want = U["lf_out"][pos:pos + n].reshape(nv * 8, stride)
        coded = np.ascontiguousarray(U["lf_coded"][cpos:cpos + nh * nv])
        for fn in (O.oco_loop_filter_plane_seq, O.oco_loop_filter_plane_cells):
            p = img.copy()
            fn(p.ctypes.data + (nv * 8 - 1) * stride + 8, -stride, nh, nv, S.ptr(coded, S.u8p), limit)
            assert np.array_equal(p, want)
        pos += n
        cpos += nh * nv


def test_fdct_quantize_golden():
    O = S.oracle()
    for i in range(len(U["fdct_x"])):
        x = U["fdct_x"][i].copy()
        y = np.zeros(64, np.int16)
        O.oco_fdct8x8(S.ptr(y, S.i16p), S.ptr(x, S.i16p))
        assert np.array_equal(y, U["fdct_y"][i])
        deq = U["q_deq"][i].copy()
        enq = np.zeros(128, np.int16)
        O.oco_enquant_init(S.ptr(enq, S.i16p), S.ptr(deq, S.u16p))
        assert np.array_equal(enq, U["q_enq"][i])
        q = np.zeros(64, np.int16)
        last = O.oco_quantize(S.ptr(q, S.i16p), S.ptr(y, S.i16p), S.ptr(deq, S.u16p), S.ptr(enq, S.i16p))
        assert last == U["q_last"][i] and np.array_equal(q, U["q_out"][i])


def test_block_metrics_golden():
    O = S.oracle()
    for i in range(len(U["met_blocks"])):
        b = np.ascontiguousarray(U["met_blocks"][i])
        s, r1, r2 = (b[k].ctypes.data for k in range(3))
        m = U["met_out"][i].tolist()
        dc = C.c_int(0)
        assert O.oco_frag_sad(s, r1, 8) == m[0]
        assert O.oco_frag_sad2_thresh(s, r1, r2, 8, 0xFFFFFFFF) == m[1]
        assert O.oco_frag_satd(C.byref(dc), s, r1, 8) == m[2] and dc.value == m[3]
        assert O.oco_frag_satd2(C.byref(dc), s, r1, r2, 8) == m[4] and dc.value == m[5]
        assert O.oco_frag_intra_satd(C.byref(dc), s, 8) == m[6] and dc.value == m[7]
        assert O.oco_frag_ssd(s, r1, 8) == m[8]
        assert O.oco_frag_intra_sad(s, 8) == m[9]


@pytest.mark.skipif(not streams.available(), reason="integrated build (recorder) not present")
@pytest.mark.parametrize("name", STREAM_NAMES)
def test_golden_streams_through_recorder_and_oracle(name):
    """Golden packets -> reference host parser + recorder (record mode) -> oracle
    frame executor -> plane hashes equal to what the reference decoder produced."""
    blob = G[name + "_blob"].tobytes()
    want = G[name + "_hashes"]
    g, works, _ = streams.capture_stream_work(blob, streams.BACKEND_RECORD)
    frames = np.full(g.nrefs * g.ref_frame_sz, 0x80, np.uint8)
    assert len(works) == len(want)
    cur = 0
    for i, wk in enumerate(works):
        if wk is not None:
            f = wk.as_struct()
            S.oracle().oco_dec_frame(C.byref(g), S.ptr(frames, S.u8p), C.byref(f), 7)
            cur = wk.ref_idx[2]
        planes = S.planes_from_buffer(g, frames[cur * g.ref_frame_sz:(cur + 1) * g.ref_frame_sz])
        got = [S.fnv1a64(p) for p in planes]
        assert got == [int(x) for x in want[i]], (name, i)
