"""GPU parity: the CUDA decode path (through the C ABI) against the CPU oracle
on seeded synthetic work lists -- bit-exact over the whole padded buffer."""
import numpy as np
import pytest

import support as S
import theora_b200 as T
import workgen as W

pytestmark = pytest.mark.gpu


def run_gpu(g, frames, work, stage_mask=7):
    T.lib().ocg_set_stage_mask(stage_mask)
    try:
        ctx = T.Context(g)
        for b in range(g.nrefs):
            ctx.upload_frame(b, frames[b * g.ref_frame_sz:(b + 1) * g.ref_frame_sz])
        out = np.empty(g.ref_frame_sz, np.uint8)
        ctx.submit(work, out)
        ctx.sync()
        full = frames.copy()
        s = work.ref_idx[2]
        full[s * g.ref_frame_sz:(s + 1) * g.ref_frame_sz] = out
        # the other buffers must be untouched
        for b in range(g.nrefs):
            if b != s:
                assert np.array_equal(ctx.download_frame(b), frames[b * g.ref_frame_sz:(b + 1) * g.ref_frame_sz])
        ctx.close()
        return full
    finally:
        T.lib().ocg_set_stage_mask(7)


def first_diff(g, a, b):
    d = np.nonzero(a != b)[0]
    return "first mismatch at byte %d of %d (%d differ)" % (d[0], a.size, d.size) if d.size else "equal"


CASES = [
    # fw, fh, fmt, density, intra_only, lf, big, dense_rows
    (64, 64, 0, 0.7, False, 0, False, False),
    (64, 64, 0, 1.0, True, 0, False, True),
    (64, 48, 0, 0.5, False, 9, False, False),
    (352, 288, 0, 0.6, False, 4, False, False),
    (352, 288, 0, 0.9, False, 30, True, False),
    (176, 144, 2, 0.7, False, 5, False, False),
    (176, 144, 3, 0.7, False, 5, True, True),
    (16, 16, 0, 1.0, False, 3, False, False),
    (32, 16, 0, 0.0, False, 3, False, False),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("stage_mask", [1, 7])
def test_decode_frame_matches_oracle(case, stage_mask):
    fw, fh, fmt, density, intra, lf, big, dense = case
    rng = np.random.default_rng(hash(case) & 0xFFFF)
    g = S.make_geometry(fw, fh, fmt, 3)
    frames = W.random_frames(g, rng)
    work = W.random_work(g, rng, density=density, intra_only=intra, lf_limit=lf, big=big, dense_rows=dense,
                         ref_idx=(1, 2, 0))
    want = W.oracle_decode(g, frames, work, stage_mask)
    got = run_gpu(g, frames, work, stage_mask)
    assert np.array_equal(want, got), first_diff(g, want, got)


@pytest.mark.parametrize("cls", [0, 1, 2, 3])
def test_each_sparsity_class(cls):
    rng = np.random.default_rng(100 + cls)
    g = S.make_geometry(128, 64, 0, 3)
    probs = [0.0] * 4
    probs[cls] = 1.0
    frames = W.random_frames(g, rng)
    work = W.random_work(g, rng, density=1.0, cls_probs=probs, big=True)
    want = W.oracle_decode(g, frames, work, 1)
    got = run_gpu(g, frames, work, 1)
    assert np.array_equal(want, got), first_diff(g, want, got)


@pytest.mark.parametrize("tma", [0, 1, 2])
@pytest.mark.parametrize("dims", [(96, 80, 0), (1040, 48, 0), (1024, 32, 2), (560, 64, 3), (80, 96, 3)])
def test_loop_filter_only_all_limits(tma, dims):
    """Loop filter + borders on random pixels, every coded pattern, many limits,
    every kernel variant (0: strip kernel, the default; 1: TMA tiles through shared memory; 2: one cell per
    thread)."""
    rng = np.random.default_rng(5)
    g = S.make_geometry(dims[0], dims[1], dims[2], 3)
    T.lib().ocg_set_lf_tma(tma)  # before the contexts are created: they build the tensor maps
    try:
        for lim in (1, 2, 3, 7, 16, 31, 63, 127):
            frames = W.random_frames(g, rng)
            work = W.random_work(g, rng, density=float(rng.random()), lf_limit=lim)
            want = W.oracle_decode(g, frames, work, 6)
            got = run_gpu(g, frames, work, 6)
            assert np.array_equal(want, got), (lim, first_diff(g, want, got))
    finally:
        T.lib().ocg_set_lf_tma(0)


def test_1080p_frame():
    rng = np.random.default_rng(1080)
    g = S.make_geometry(1920, 1088, 0, 3)
    frames = W.random_frames(g, rng)
    work = W.random_work(g, rng, density=0.8, lf_limit=6)
    want = W.oracle_decode(g, frames, work)
    got = run_gpu(g, frames, work)
    assert np.array_equal(want, got), first_diff(g, want, got)


def test_batch_of_streams_matches_single_submits():
    """ocg_dec_run_batch over resident packs == per-frame submits == oracle."""
    rng = np.random.default_rng(77)
    g = S.make_geometry(176, 144, 0, 3)
    nstreams, nframes = 5, 3
    ctxs, packs, wants = [], [], []
    for s in range(nstreams):
        frames = W.random_frames(g, rng)
        works = []
        cur = frames
        for f in range(nframes):
            # rotate buffers the way decode.c:2947-2962 does for inter frames
            refs = ((2 + f) % 3, (1 + f) % 3, (0 + f) % 3) if f else (1, 2, 0)
            w = W.random_work(g, rng, density=0.7, lf_limit=int(rng.integers(0, 12)), ref_idx=refs)
            works.append(w)
            cur = W.oracle_decode(g, cur, w)
        wants.append(cur)
        ctx = T.Context(g)
        for b in range(3):
            ctx.upload_frame(b, frames[b * g.ref_frame_sz:(b + 1) * g.ref_frame_sz])
        ctxs.append(ctx)
        packs.append(T.Pack(works, g.nfrags))
    for f in range(nframes):
        T.run_batch(ctxs, packs, [f] * nstreams)
    for s in range(nstreams):
        ctxs[s].sync()
    ctxs[0].sync()
    import ctypes
    for s in range(nstreams):
        got = np.concatenate([ctxs[s].download_frame(b) for b in range(3)])
        assert np.array_equal(wants[s], got), (s, first_diff(g, wants[s], got))


DC_CASES = [
    # fw, fh, fmt, density, intra_only
    (64, 64, 0, 0.7, False),
    (64, 64, 0, 1.0, True),      # every pattern is 15/7: the longest chains
    (352, 288, 0, 0.15, False),  # sparse: pred_last reaches far back in raster order
    (352, 288, 0, 0.6, False),
    (176, 144, 2, 0.5, False),
    (176, 144, 3, 0.9, False),
    (16, 16, 0, 1.0, False),     # a single super block
    (1920, 1088, 0, 0.4, False),
    (32, 2048, 0, 0.5, False),   # tall: 256 fragment rows in one CTA
    (3840, 2160, 0, 0.5, False), # luma DC values do not fit shared memory: global scratch variant
]


@pytest.mark.parametrize("case", DC_CASES)
def test_dc_unprediction_on_device_matches_oracle(case):
    """dc_residual=1: the device undoes the DC prediction (decode.c:1392-1500) before reconstructing;
    the records' dc fields are treated as residuals by both sides."""
    from theora_b200 import abi
    fw, fh, fmt, density, intra = case
    rng = np.random.default_rng(hash(case) & 0xFFFF)
    g = S.make_geometry(fw, fh, fmt, 3)
    assert abi.lib().ocg_dc_unpredict_supported(g) == 1
    frames = W.random_frames(g, rng)
    work = W.random_work(g, rng, density=density, intra_only=intra, lf_limit=0, ref_idx=(1, 2, 0))
    # large residuals too, so the 16-bit wrap of frags[].dc is reached now and then
    coded = work.recs["refi"] != 3
    big = rng.random(len(work.recs)) < 0.02
    work.recs["dc"][coded & big] = rng.integers(-32768, 32768, size=int((coded & big).sum()))
    w2 = abi.FrameWork(work.ref_idx, work.lf_limit, work.dc_quant, work.recs, work.rows, dc_residual=1)
    want = W.oracle_decode(g, frames, w2, 1)
    got = run_gpu(g, frames, w2, 1)
    assert np.array_equal(want, got), first_diff(g, want, got)
    # and it is not a no-op: with the flag off the same records reconstruct differently
    if density > 0.3:
        assert not np.array_equal(want, W.oracle_decode(g, frames, work, 1))


def test_dc_residual_frames_are_refused_in_resident_packs():
    from theora_b200 import abi
    rng = np.random.default_rng(3)
    g = S.make_geometry(64, 64, 0, 3)
    work = W.random_work(g, rng, density=0.5)
    w2 = abi.FrameWork(work.ref_idx, work.lf_limit, work.dc_quant, work.recs, work.rows, dc_residual=1)
    ctx = T.Context(g)
    pack = T.Pack([w2], g.nfrags)
    with pytest.raises(abi.OcgError):
        T.run_batch([ctx], [pack], [0])
    ctx.close()
