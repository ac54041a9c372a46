"""BASELINE configs[2] at full size as a parity case: the intra-only encode block
pipeline of one 1080p 4:2:0 frame (every fragment: sub_128 -> fDCT -> quantise,
then the decoder-side dequant + iDCT + intra recon of the quantised blocks) on the
GPU against the CPU oracle; plus inter residual transform and SATD/SAD over all
fragments with motion-displaced predictors (configs[3] metrics)."""
import ctypes as C

import numpy as np
import pytest
import torch

import support as S
import theora_b200 as T
from theora_b200 import abi

pytestmark = pytest.mark.gpu

FW, FH = 1920, 1088


def synth_padded_frame(g, f, seed):
    """One padded reference-layout buffer holding synthetic frame `f` (aprons replicated)."""
    R = None
    buf = np.zeros(g.ref_frame_sz, np.uint8)
    ys, xs = np.mgrid[0:FH, 0:FW]
    rng = np.random.default_rng(seed + f)
    y = ((2 * (xs + 3 * f) + (ys + f)) & 255) + 60 * ((((xs + 3 * f) >> 5) ^ ((ys + f) >> 5)) & 1) + rng.integers(0, 4, size=xs.shape)
    planes = [np.clip(y, 0, 255).astype(np.uint8)]
    cys, cxs = np.mgrid[0:FH // 2, 0:FW // 2]
    planes.append((128 + (((cxs + f) >> 3) & 15)).astype(np.uint8))
    planes.append((128 - (((cys + 2 * f) >> 3) & 15)).astype(np.uint8))
    for pli in range(3):
        p = g.planes[pli]
        stride = -p.ystride
        top_left = g.base_off + p.plane_off + (p.height - 1) * p.ystride
        view = np.lib.stride_tricks.as_strided(buf[top_left:], shape=(p.height, p.width), strides=(stride, 1))
        view[:] = planes[pli]
        S.oracle().oco_borders_fill_plane(buf.ctypes.data + g.base_off + p.plane_off, p.ystride, p.width, p.height,
                                          p.hpad, p.vpad)
    return buf


def quant_tables(rng):
    deq = rng.integers(8, 300, size=(18, 64)).astype(np.uint16)
    deq[:, 0] = rng.integers(8, 60, size=18)
    enq = np.zeros((18, 128), np.int16)
    for t in range(18):
        S.oracle().oco_enquant_init(S.ptr(enq[t], S.i16p), S.ptr(deq[t], S.u16p))
    return deq, enq


def run_fdct_quant(dsrc, dref, ystride, fr, deq, enq):
    n = len(fr)
    dfr = torch.from_numpy(fr.view(np.int32).reshape(n, 4)).cuda()
    ddeq = torch.from_numpy(deq.view(np.int16)).cuda()
    denq = torch.from_numpy(enq).cuda()
    od = torch.zeros((n, 64), dtype=torch.int16, device="cuda")
    oq = torch.zeros((n, 64), dtype=torch.int16, device="cuda")
    onz = torch.zeros(n, dtype=torch.int32, device="cuda")
    abi.check(abi.lib().ocg_enc_fdct_quant_batch(dsrc, dref, ystride, dfr.data_ptr(), n, ddeq.data_ptr(),
                                                 denq.data_ptr(), od.data_ptr(), oq.data_ptr(), onz.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return od.cpu().numpy(), oq.cpu().numpy(), onz.cpu().numpy()


def test_intra_only_encode_pipeline_1080p():
    rng = np.random.default_rng(3)
    g = S.make_geometry(FW, FH, 0, 3)
    offs = np.empty(g.nfrags, np.int32)
    S.oracle().oco_geometry_frag_buf_offs(C.byref(g), S.ptr(offs, S.i32p))
    src = synth_padded_frame(g, 0, 100)
    deq, enq = quant_tables(rng)
    n = g.nfrags
    planes = np.concatenate([np.full(g.planes[p].nfrags, p, np.int32) for p in range(3)])
    fr = np.zeros(n, S.ENC_FRAG_DTYPE)
    fr["src_off"] = offs
    fr["ref_off0"] = S.INT32_MIN
    fr["ref_off1"] = S.INT32_MIN
    fr["aux"] = planes  # pli, qti=0, qii=0
    # the luma and chroma strides differ: run the planes separately, like analyze.c does
    dsrc = torch.from_numpy(src).cuda()
    got_d, got_q, got_nz = np.zeros((n, 64), np.int16), np.zeros((n, 64), np.int16), np.zeros(n, np.int32)
    want_d, want_q, want_nz = np.zeros((n, 64), np.int16), np.zeros((n, 64), np.int16), np.zeros(n, np.int32)
    for pli in range(3):
        p = g.planes[pli]
        sl = slice(p.froffset, p.froffset + p.nfrags)
        sub = np.ascontiguousarray(fr[sl])
        d, q, nz = run_fdct_quant(dsrc.data_ptr() + g.base_off, dsrc.data_ptr() + g.base_off, p.ystride, sub, deq, enq)
        got_d[sl], got_q[sl], got_nz[sl] = d, q, nz
        S.oracle().oco_enc_fdct_quant_batch(src.ctypes.data + g.base_off, src.ctypes.data + g.base_off, p.ystride,
                                            sub.ctypes.data, p.nfrags, S.ptr(deq, S.u16p), S.ptr(enq, S.i16p),
                                            S.ptr(want_d[sl], S.i16p), S.ptr(want_q[sl], S.i16p), S.ptr(want_nz[sl], S.i32p))
    assert np.array_equal(got_d, want_d)
    assert np.array_equal(got_q, want_q)
    assert np.array_equal(got_nz, want_nz)
    # reconstruction leg of the encoder (analyze.c:793-823): dequantise, iDCT, recon_intra -- through the
    # decoder-side kernel, all fragments intra, last_zzi = nonzero+1 as the encoder passes it
    fz = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20,
                   13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52,
                   45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])
    tab = planes * 6  # (pli*2+qti)*3+qii with qti=qii=0
    coeffs = np.zeros((n, 64), np.int16)
    coeffs[:, fz] = (got_q.astype(np.int32) * deq[tab].astype(np.int32)).astype(np.int16)
    dc = got_q[:, 0].copy()
    coeffs[:, 0] = 0
    last_zzi = np.where((got_nz == 0) & (got_q[:, 0] == 0), 0, got_nz + 1).astype(np.uint8)
    recs = np.zeros(n, S.REC_DTYPE)
    recs["buf_off"] = offs
    recs["dc"] = dc
    recs["last_zzi"] = last_zzi
    recs["refi"] = 2
    recs["pli_qti"] = planes
    rows = coeffs.reshape(n, 8, 8)
    cls = S.cls_of_last_zzi(last_zzi)
    nrow = np.array([0, 2, 4, 8])[cls]
    keep = (np.abs(rows).sum(axis=2) > 0) & (np.arange(8)[None, :] < nrow[:, None])
    recs["rowmask"] = (keep * (1 << np.arange(8))[None, :]).sum(axis=1).astype(np.uint8)
    cnt = keep.sum(axis=1)
    recs["coeff_row"] = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.uint32)
    rowpool = rows[keep]
    dcq = np.array([[deq[0, 0], deq[1, 0]], [deq[6, 0], deq[7, 0]], [deq[12, 0], deq[13, 0]]], np.uint16)
    work = T.FrameWork((1, 2, 0), 0, dcq, recs, rowpool)
    frames = np.zeros(3 * g.ref_frame_sz, np.uint8)
    import workgen as W
    want = W.oracle_decode(g, frames, work, 5)
    ctx = T.Context(g)
    for b in range(3):
        ctx.upload_frame(b, frames[b * g.ref_frame_sz:(b + 1) * g.ref_frame_sz])
    T.lib().ocg_set_stage_mask(5)
    out = np.empty(g.ref_frame_sz, np.uint8)
    ctx.submit(work, out)
    ctx.sync()
    T.lib().ocg_set_stage_mask(7)
    ctx.close()
    assert np.array_equal(out, want[:g.ref_frame_sz])
    # sanity: the reconstruction is close to the source (it is a real encode round trip)
    pl = S.planes_from_buffer(g, out)
    sp = S.planes_from_buffer(g, src)
    mse = float(np.mean((pl[0].astype(np.float64) - sp[0]) ** 2))
    assert mse < 200.0, mse


def test_inter_metrics_and_residual_1080p():
    rng = np.random.default_rng(4)
    g = S.make_geometry(FW, FH, 0, 3)
    offs = np.empty(g.nfrags, np.int32)
    S.oracle().oco_geometry_frag_buf_offs(C.byref(g), S.ptr(offs, S.i32p))
    src = synth_padded_frame(g, 1, 100)
    ref = synth_padded_frame(g, 0, 100)
    p = g.planes[0]
    n = p.nfrags
    fr = np.zeros(n, S.ENC_FRAG_DTYPE)
    fr["src_off"] = offs[:n]
    dx = rng.integers(-15, 16, size=n)
    dy = rng.integers(-15, 16, size=n)
    fr["ref_off0"] = offs[:n] + dy * p.ystride + dx
    two = rng.random(n) < 0.5
    fr["ref_off1"] = np.where(two, fr["ref_off0"] + rng.integers(-1, 2, size=n) * p.ystride + rng.integers(-1, 2, size=n),
                              S.INT32_MIN)
    fr["aux"] = 4  # luma, inter tables
    dsrc, dref = torch.from_numpy(src).cuda(), torch.from_numpy(ref).cuda()
    dfr = torch.from_numpy(fr.view(np.int32).reshape(n, 4)).cuda()
    for metric in (0, 1, 3):
        sub = fr.copy()
        if metric == 3:
            sub["ref_off1"] = S.INT32_MIN
        dsub = torch.from_numpy(sub.view(np.int32).reshape(n, 4)).cuda()
        ov = torch.zeros(n, dtype=torch.int32, device="cuda")
        odc = torch.zeros(n, dtype=torch.int32, device="cuda")
        abi.check(abi.lib().ocg_enc_metrics_batch(metric, dsrc.data_ptr() + g.base_off, dref.data_ptr() + g.base_off,
                                                  p.ystride, dsub.data_ptr(), n, ov.data_ptr(), odc.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        wv, wdc = np.zeros(n, np.uint32), np.zeros(n, np.int32)
        S.oracle().oco_enc_metrics_batch(metric, src.ctypes.data + g.base_off, ref.ctypes.data + g.base_off, p.ystride,
                                         sub.ctypes.data, n, S.ptr(wv, S.u32p), S.ptr(wdc, S.i32p))
        assert np.array_equal(ov.cpu().numpy().view(np.uint32), wv), metric
        assert np.array_equal(odc.cpu().numpy(), wdc), metric
    deq, enq = quant_tables(rng)
    d, q, nz = run_fdct_quant(dsrc.data_ptr() + g.base_off, dref.data_ptr() + g.base_off, p.ystride, fr, deq, enq)
    wd, wq, wnz = np.zeros((n, 64), np.int16), np.zeros((n, 64), np.int16), np.zeros(n, np.int32)
    S.oracle().oco_enc_fdct_quant_batch(src.ctypes.data + g.base_off, ref.ctypes.data + g.base_off, p.ystride,
                                        fr.ctypes.data, n, S.ptr(deq, S.u16p), S.ptr(enq, S.i16p), S.ptr(wd, S.i16p),
                                        S.ptr(wq, S.i16p), S.ptr(wnz, S.i32p))
    assert np.array_equal(d, wd) and np.array_equal(q, wq) and np.array_equal(nz, wnz)
