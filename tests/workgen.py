"""Seeded synthetic decoder work lists (coded/uncoded fragments, motion vectors,
coefficient rows) in the C-ABI format, for parity tests at sizes where no real
stream is at hand, plus helpers to run a frame through the CPU oracle."""
import ctypes as C

import numpy as np

import support as S
from theora_b200.abi import FrameWork, REC_DTYPE, cls_of_last_zzi

FZ = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,
               14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39,
               46, 53, 60, 61, 54, 47, 55, 62, 63])


def frag_offsets(g):
    offs = np.empty(g.nfrags, np.int32)
    S.oracle().oco_geometry_frag_buf_offs(C.byref(g), S.ptr(offs, S.i32p))
    return offs


def frag_planes(g):
    pl = np.empty(g.nfrags, np.int32)
    for pli in range(3):
        p = g.planes[pli]
        pl[p.froffset:p.froffset + p.nfrags] = pli
    return pl


def random_work(g, rng, density=0.7, intra_only=False, lf_limit=0, big=False, dense_rows=False,
                ref_idx=(0, 1, 2), cls_probs=(0.3, 0.2, 0.2, 0.3), mv_range=31):
    """Builds one FrameWork with random content; every field is exercised."""
    n = g.nfrags
    offs = frag_offsets(g)
    planes = frag_planes(g)
    coded = rng.random(n) < density
    if intra_only:
        coded[:] = True
    recs = np.zeros(n, REC_DTYPE)
    recs["buf_off"] = offs
    recs["refi"] = 3
    recs["pli_qti"] = planes
    idx = np.nonzero(coded)[0]
    nc = len(idx)
    lz_choices = {0: [0, 1], 1: [2, 3], 2: [4, 7, 10], 3: [11, 20, 40, 63, 64]}
    cls = rng.choice(4, size=nc, p=cls_probs)
    last_zzi = np.array([rng.choice(lz_choices[int(c)]) for c in cls], np.int32)
    assert np.array_equal(cls_of_last_zzi(last_zzi), cls)
    # coefficient rows are appended in a shuffled ("coded") order, not raster order
    order = rng.permutation(nc)
    rows = []
    nrows = 0
    amp = 32767 if big else 500
    coeff_row = np.zeros(nc, np.uint32)
    rowmask = np.zeros(nc, np.uint8)
    for i in order:
        blk = np.zeros(64, np.int16)
        lz = int(last_zzi[i])
        nzmax = 64 if lz > 10 else lz
        if nzmax > 1:
            k = int(rng.integers(0, nzmax))
            pos = FZ[rng.choice(np.arange(1, nzmax), size=min(k, nzmax - 1), replace=False)] if k else []
            blk[pos] = rng.integers(-amp, amp + 1, size=len(pos))
        if big and i % 3 == 0:
            blk[1:] = rng.integers(-32768, 32768, size=63)  # garbage outside the class footprint too
        b = blk.reshape(8, 8)
        nrow_cls = (0, 2, 4, 8)[int(cls[i])]
        mask = 0
        for r in range(nrow_cls):
            keep = dense_rows or b[r].any() or rng.random() < 0.05
            if keep:
                mask |= 1 << r
                rows.append(b[r].copy())
        coeff_row[i] = nrows
        nrows += bin(mask).count("1")
        rowmask[i] = mask
    crec = np.zeros(nc, REC_DTYPE)
    crec["buf_off"] = offs[idx]
    crec["coeff_row"] = coeff_row
    crec["rowmask"] = rowmask
    dx = rng.integers(-mv_range, mv_range + 1, size=nc)
    dy = rng.integers(-mv_range, mv_range + 1, size=nc)
    crec["mv"] = (((dy & 0xFF) << 8) | (dx & 0xFF)).astype(np.uint16).view(np.int16)
    crec["dc"] = rng.integers(-32768 if big else -1500, 32768 if big else 1500, size=nc)
    crec["last_zzi"] = last_zzi
    if intra_only:
        crec["refi"] = 2
    else:
        crec["refi"] = rng.choice([0, 1, 2], size=nc, p=[0.2, 0.6, 0.2])
    qti = (crec["refi"] != 2).astype(np.uint8)
    crec["pli_qti"] = planes[idx].astype(np.uint8) | (qti << 2)
    recs[idx] = crec
    dcq = rng.integers(8, 65535 if big else 400, size=(3, 2)).astype(np.uint16)
    rows_arr = np.array(rows, np.int16).reshape(-1, 8) if rows else np.zeros((0, 8), np.int16)
    return FrameWork(ref_idx, lf_limit, dcq, recs, rows_arr)


def random_frames(g, rng):
    """nrefs padded buffers (aprons included) of random pixels."""
    return rng.integers(0, 256, size=g.nrefs * g.ref_frame_sz, dtype=np.uint8)


def oracle_decode(g, frames, work, stage_mask=7):
    out = frames.copy()
    f = work.as_struct()
    S.oracle().oco_dec_frame(C.byref(g), S.ptr(out, S.u8p), C.byref(f), stage_mask)
    return out
