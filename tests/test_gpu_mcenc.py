"""GPU parity of the motion-search kernel (ocg_mcenc_search_batch) against the
CPU oracle (itself pinned against oc_mcenc_search_frame, test_oracle_mcenc.py),
at unit scale and over every macro block of a 1080p frame pair (configs[3])."""
import ctypes as C

import numpy as np
import pytest
import torch

import mcgen as M
import support as S
from theora_b200 import abi

pytestmark = pytest.mark.gpu


def run_both(src, rfull, rsatd, bl, ystride, mb_in):
    n = len(mb_in)
    want = np.zeros(n, M.MB_OUT)
    O = S.oracle()
    O.oco_mcenc_search_batch.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    O.oco_mcenc_search_batch(src.ctypes.data + bl, rfull.ctypes.data + bl, rsatd.ctypes.data + bl, ystride,
                             mb_in.ctypes.data, want.ctypes.data, n)
    ds, df, dt = (torch.from_numpy(a).cuda() for a in (src, rfull, rsatd))
    din = torch.from_numpy(mb_in.view(np.uint8).reshape(n, 48)).cuda()
    dout = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    abi.check(abi.lib().ocg_mcenc_search_batch(ds.data_ptr() + bl, df.data_ptr() + bl, dt.data_ptr() + bl, ystride,
                                               din.data_ptr(), dout.data_ptr(), n,
                                               torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = dout.cpu().numpy().view(M.MB_OUT).reshape(n)
    return want, got


@pytest.mark.parametrize("seed,shift,noise,smooth", [(1, (5, -3), 6, True), (2, (-12, 9), 10, True),
                                                      (3, (0, 0), 2, True), (4, (14, 15), 20, False),
                                                      (5, (-15, -15), 4, True)])
def test_search_kernel_matches_oracle(seed, shift, noise, smooth):
    rng = np.random.default_rng(seed)
    src, rfull, rsatd, bl, ystride = M.make_scene(rng, shift=shift, noise=noise, smooth=smooth)
    mb_in, _ = M.make_cases(rng, 501, ystride=ystride)
    want, got = run_both(src, rfull, rsatd, bl, ystride, mb_in)
    for f in M.MB_OUT.names:
        assert np.array_equal(want[f], got[f]), f


def test_search_every_macro_block_of_a_1080p_frame():
    rng = np.random.default_rng(9)
    w, h = 1920, 1088
    src, rfull, rsatd, bl, ystride = M.make_scene(rng, w=w, h=h, pad=16, shift=(3, 1), noise=3)
    nmb = (w // 16) * (h // 16)
    mb_in = np.zeros(nmb, M.MB_IN)
    i = 0
    for my in range(0, h, 16):
        for mx in range(0, w, 16):
            mb_in[i]["frag_off"] = [(my + by) * ystride + mx + bx for by in (0, 8) for bx in (0, 8)]
            c, setb0, ncand = M.candidates([M.mv_pack(int(rng.integers(-8, 9)), int(rng.integers(-8, 9)))
                                            for _ in range(int(rng.integers(0, 5)))], 0,
                                           M.mv_pack(int(rng.integers(-8, 9)), int(rng.integers(-8, 9))),
                                           M.mv_pack(int(rng.integers(-8, 9)), int(rng.integers(-8, 9))))
            for k, (x, y) in enumerate(c):
                mb_in[i]["cand"][k] = (x, y)
            mb_in[i]["setb0"], mb_in[i]["ncand"] = setb0, ncand
            mb_in[i]["t2_base"] = int(rng.integers(0, 1500))
            mb_in[i]["is_prev"] = i & 1
            i += 1
    assert nmb == 8160
    want, got = run_both(src, rfull, rsatd, bl, ystride, mb_in)
    for f in M.MB_OUT.names:
        assert np.array_equal(want[f], got[f]), f
    # the planted displacement is recovered (rows are addressed bottom-up, so only |dy| is checked)
    found = (got["best_vec"][:, 0] == 3) & (np.abs(got["best_vec"][:, 1]) == 1)
    assert found.mean() > 0.5, found.mean()


def run_refine(src, ref, bl, ystride, mb, flags):
    n = len(mb)
    want = np.zeros(n, M.REF_OUT)
    O = S.oracle()
    O.oco_mcenc_refine_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    O.oco_mcenc_refine_batch(src.ctypes.data + bl, ref.ctypes.data + bl, ystride, mb.ctypes.data, want.ctypes.data, n,
                             flags)
    ds, dr = torch.from_numpy(src).cuda(), torch.from_numpy(ref).cuda()
    din = torch.from_numpy(mb.view(np.uint8).reshape(n, 48)).cuda()
    dout = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    abi.check(abi.lib().ocg_mcenc_refine_batch(ds.data_ptr() + bl, dr.data_ptr() + bl, ystride, din.data_ptr(),
                                               dout.data_ptr(), n, flags, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return want, dout.cpu().numpy().view(M.REF_OUT).reshape(n)


@pytest.mark.parametrize("seed,shift,entry,flags", [(1, (5, -3), "mixed", 3), (2, (-2, 7), "max", 3), (3, (0, 0), 0, 3),
                                                     (4, (9, 4), "mixed", 7), (5, (-6, -6), "max", 5),
                                                     (6, (1, 1), "mixed", 1), (7, (1, 1), "mixed", 2)])
def test_refine_kernel_matches_oracle(seed, shift, entry, flags):
    """ocg_mcenc_refine_batch (oc_mcenc_refine1mv / refine4mv, mcenc.c:606-791)."""
    rng = np.random.default_rng(seed)
    src, _, ref, bl, ystride = M.make_scene(rng, shift=shift, noise=5)
    mb = M.make_refine_cases(rng, 777, ystride=ystride, entry=entry)
    want, got = run_refine(src, ref, bl, ystride, mb, flags)
    fields = (["mv", "satd"] if flags & 1 else []) + (["ref_mv", "block_satd"] if flags & 2 else [])
    for f in fields:
        assert np.array_equal(want[f], got[f]), f


def test_refine_every_macro_block_of_a_1080p_frame():
    rng = np.random.default_rng(19)
    w, h = 1920, 1088
    src, _, ref, bl, ystride = M.make_scene(rng, w=w, h=h, pad=16, shift=(3, 1), noise=3)
    mb = M.make_refine_cases(rng, 8160, w=w, h=h, ystride=ystride, vmax=6, entry="max")
    i = 0
    for my in range(0, h, 16):
        for mx in range(0, w, 16):
            mb[i]["frag_off"] = [(my + by) * ystride + mx + bx for by in (0, 8) for bx in (0, 8)]
            i += 1
    want, got = run_refine(src, ref, bl, ystride, mb, 3)
    for f in ("mv", "satd", "ref_mv", "block_satd"):
        assert np.array_equal(want[f], got[f]), f
