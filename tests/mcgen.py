"""Motion-search test cases: frames with known displacement + the candidate
sets of lib/mcenc.c:90-164 restated, in the C-ABI layout of ocg_mb_search_in."""
import numpy as np

MB_IN = np.dtype([("frag_off", "<i4", (4,)), ("cand", "i1", (13, 2)), ("setb0", "u1"), ("ncand", "u1"),
                  ("t2_base", "<u2"), ("is_prev", "u1"), ("pad", "u1")])
MB_OUT = np.dtype([("best_vec", "i1", (2,)), ("error", "<u2"), ("satd", "<u4"), ("block_vec", "i1", (4, 2)),
                   ("block_satd", "<u4", (4,))])
assert MB_IN.itemsize == 48 and MB_OUT.itemsize == 32


def mv_pack(x, y):
    v = ((y & 0xFF) << 8) | (x & 0xFF)
    return v - 65536 if v >= 32768 else v


def mv_x(mv):
    v = mv & 0xFF
    return v - 256 if v >= 128 else v


def mv_y(mv):
    return mv >> 8


def clamp31(v):
    return max(-31, min(31, v))


def candidates(nb_mvs, accum, mv1, mv2):
    """mcenc.c:90-164 -> (list of (x,y), setb0, ncand)."""
    ax, ay = mv_x(accum), mv_y(accum)
    c = [None]
    for m in nb_mvs:
        c.append((mv_x(m), mv_y(m)))
    c.append((ax, ay))
    c.append((clamp31(mv_x(mv1) + ax), clamp31(mv_y(mv1) + ay)))
    c.append((0, 0))
    a = c[1:4]
    c[0] = (sorted(v[0] for v in a)[1], sorted(v[1] for v in a)[1])
    setb0 = len(c)
    c.append((clamp31(2 * mv_x(mv1) - mv_x(mv2) + ax), clamp31(2 * mv_y(mv1) - mv_y(mv2) + ay)))
    return c, setb0, len(c)


def make_scene(rng, w=192, h=128, pad=32, shift=(5, -3), noise=6, smooth=True):
    """src = ref shifted by `shift` + noise; padded buffers, bottom-up addressing.
    Returns src, ref_full, ref_satd (uint8 2-D), base offset of the bottom-left picture pixel, ystride."""
    st = w + 2 * pad
    hh = h + 2 * pad
    base_img = rng.integers(0, 256, size=(hh + 64, st + 64)).astype(np.float64)
    if smooth:
        k = np.ones(5) / 5.0
        base_img = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, base_img)
        base_img = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 0, base_img)
        base_img = (base_img - base_img.min()) / (base_img.max() - base_img.min()) * 255
    ref = base_img[32:32 + hh, 32:32 + st]
    src = base_img[32 + shift[1]:32 + shift[1] + hh, 32 + shift[0]:32 + shift[0] + st]
    src = np.clip(src + rng.integers(-noise, noise + 1, size=src.shape), 0, 255).astype(np.uint8)
    ref_full = np.clip(ref, 0, 255).astype(np.uint8)
    ref_satd = np.clip(ref + rng.integers(-3, 4, size=ref.shape), 0, 255).astype(np.uint8)
    bl = (pad + h - 1) * st + pad
    return np.ascontiguousarray(src), np.ascontiguousarray(ref_full), np.ascontiguousarray(ref_satd), bl, -st


def make_cases(rng, n, w=192, h=128, st=None, ystride=None):
    """n macro blocks at random 16-aligned positions with random histories."""
    cases = []
    mb_in = np.zeros(n, MB_IN)
    for i in range(n):
        mx = int(rng.integers(0, w // 16)) * 16
        my = int(rng.integers(0, h // 16)) * 16
        # four luma fragments of the MB in mb_maps order (bottom-left, bottom-right, top-left, top-right)
        offs = [(my + by) * ystride + mx + bx for by in (0, 8) for bx in (0, 8)]
        ncn = int(rng.integers(0, 5))
        rngv = 31 if i % 3 else 8
        nb_mvs = [mv_pack(int(rng.integers(-rngv, rngv + 1)), int(rng.integers(-rngv, rngv + 1))) for _ in range(ncn)]
        nb_err = [int(rng.integers(0, 3000)) for _ in range(ncn)]
        accum = mv_pack(int(rng.integers(-6, 7)), int(rng.integers(-6, 7))) if i % 4 == 0 else 0
        mv1 = mv_pack(int(rng.integers(-rngv, rngv + 1)), int(rng.integers(-rngv, rngv + 1)))
        mv2 = mv_pack(int(rng.integers(-rngv, rngv + 1)), int(rng.integers(-rngv, rngv + 1)))
        own_err = int(rng.integers(0, 2000))
        frame = int(rng.integers(0, 2))  # 0 GOLD, 1 PREV
        c, setb0, ncand = candidates(nb_mvs, accum, mv1, mv2)
        mb_in[i]["frag_off"] = offs
        for k, (x, y) in enumerate(c):
            mb_in[i]["cand"][k] = (x, y)
        mb_in[i]["setb0"] = setb0
        mb_in[i]["ncand"] = ncand
        mb_in[i]["t2_base"] = max([own_err] + nb_err[:3])
        mb_in[i]["is_prev"] = frame
        cases.append(dict(offs=offs, ncn=ncn, nb_mvs=nb_mvs, nb_err=nb_err, accum=accum, mv1=mv1, mv2=mv2,
                          own_err=own_err, frame=frame))
    return mb_in, cases


REF_IN = np.dtype([("frag_off", "<i4", (4,)), ("vec", "i1", (2,)), ("block_vec", "i1", (4, 2)), ("pad", "u1", (2,)),
                   ("satd", "<u4"), ("block_satd", "<u4", (4,))])
REF_OUT = np.dtype([("mv", "i1", (2,)), ("ref_mv", "i1", (4, 2)), ("pad", "u1", (2,)), ("satd", "<u4"),
                    ("block_satd", "<u4", (4,))])
assert REF_IN.itemsize == 48 and REF_OUT.itemsize == 32


def make_refine_cases(rng, n, w=192, h=128, ystride=None, vmax=15, entry="mixed"):
    """n macro blocks with random full-pel vectors (window +-vmax) and entry
    scores: 'mixed' draws them around typical SATD magnitudes so that some
    sites win and some do not; 0 makes every site lose, 'max' makes the best
    site always win."""
    mb = np.zeros(n, REF_IN)
    for i in range(n):
        mx = int(rng.integers(0, w // 16)) * 16
        my = int(rng.integers(0, h // 16)) * 16
        mb[i]["frag_off"] = [(my + by) * ystride + mx + bx for by in (0, 8) for bx in (0, 8)]
        mb[i]["vec"] = rng.integers(-vmax, vmax + 1, size=2)
        mb[i]["block_vec"] = rng.integers(-vmax, vmax + 1, size=(4, 2))
        if entry == "max":
            mb[i]["satd"] = 0xFFFFFFF
            mb[i]["block_satd"] = 0xFFFFFFF
        elif entry == 0:
            mb[i]["satd"] = 0
            mb[i]["block_satd"] = 0
        else:
            mb[i]["satd"] = int(rng.integers(2000, 30000))
            mb[i]["block_satd"] = rng.integers(500, 8000, size=4)
    return mb
