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


@pytest.mark.parametrize("tma", [1, 0])
@pytest.mark.parametrize("dims", [(96, 80, 0), (1040, 48, 0), (1024, 32, 2), (560, 64, 3), (80, 96, 3)])
def test_loop_filter_only_all_limits(tma, dims):
    """Loop filter + borders on random pixels, every coded pattern, many limits,
    both kernel variants (TMA tiles through shared memory / per-thread accesses)."""
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
