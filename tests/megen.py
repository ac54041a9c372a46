"""Inputs for the whole-frame motion-analysis tests: padded frame buffers in
the library's layout (include/theora_b200.h ocg_geometry) holding a textured
scene that moves a few pixels per frame, plus noisier "reconstructed" copies."""
import ctypes as C

import numpy as np

from theora_b200 import abi


def bind_ref_me(R):
    R.refh_me_open.restype = C.c_void_p
    R.refh_me_open.argtypes = [C.c_int, C.c_int, C.c_int]
    R.refh_me_close.argtypes = [C.c_void_p]
    R.refh_me_nmbs.argtypes = [C.c_void_p]
    R.refh_me_frame_size.argtypes = [C.c_void_p]
    R.refh_me_frame_size.restype = C.c_long
    R.refh_me_topology.argtypes = [C.c_void_p, C.c_void_p]
    R.refh_me_frame.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p]
    return R


def texture(rng, h, w):
    """Smooth-ish random texture with enough detail for SAD minima to be sharp."""
    t = rng.integers(0, 256, size=(h // 4 + 2, w // 4 + 2)).astype(np.float32)
    t = np.kron(t, np.ones((4, 4), np.float32))[:h, :w]
    t = (t + np.roll(t, 1, 0) + np.roll(t, 1, 1) + np.roll(t, (1, 1), (0, 1))) / 4
    return t


def scene_buffers(g, rng, nframes, motion=(3, 1), noise=4, recon_noise=3, regions=True):
    """Returns (orig[nframes], recon[nframes]) full ref-frame buffers (uint8, ref_frame_sz)."""
    p = g.planes[0]
    stride = -p.ystride
    rows = p.height + 2 * p.vpad
    margin = 64 + 4 * nframes * (abs(motion[0]) + abs(motion[1]) + 2)
    tex = texture(rng, rows + 2 * margin, stride + 2 * margin)
    tex2 = texture(rng, rows + 2 * margin, stride + 2 * margin)
    orig, recon = [], []
    for t in range(nframes):
        buf = rng.integers(0, 256, size=g.ref_frame_sz).astype(np.uint8)  # chroma / gaps: anything
        dx, dy = motion[0] * t, motion[1] * t
        a = tex[margin + dy:margin + dy + rows, margin + dx:margin + dx + stride]
        if regions:  # right half moves differently so 4MV / descents are exercised at the seam
            b = tex2[margin - 2 * dy:margin - 2 * dy + rows, margin + 2 * dx:margin + 2 * dx + stride]
            a = a.copy()
            a[:, stride // 2:] = b[:, stride // 2:]
        y = a + rng.integers(-noise, noise + 1, size=a.shape)
        y = np.clip(y, 0, 255).astype(np.uint8)
        buf[:rows * stride] = y.reshape(-1)
        orig.append(buf)
        r = buf.copy()
        rn = y.astype(np.int32) + rng.integers(-recon_noise, recon_noise + 1, size=y.shape)
        r[:rows * stride] = np.clip(rn, 0, 255).astype(np.uint8).reshape(-1)
        recon.append(r)
    return orig, recon


COMPARE_ALWAYS = ("analysis_mv", "error", "satd", "unref_mv", "unref_satd")


def assert_me_equal(got, want, valid, flags, where=""):
    v = valid.astype(bool)
    fields = list(COMPARE_ALWAYS)
    if not flags & abi.OCG_ME_FAST:
        fields += ["block_mv", "block_satd"]
        if flags & abi.OCG_ME_REFINE_4MV:
            fields += ["ref_mv", "ref_block_satd"]
    for f in fields:
        a, b = got[f][v], want[f][v]
        if not np.array_equal(a, b):
            bad = np.argwhere(a.reshape(len(a), -1) != b.reshape(len(b), -1))
            i = int(bad[0][0])
            raise AssertionError("%s field %s differs at valid MB #%d (of %d mismatching): got %s want %s" %
                                 (where, f, i, len(np.unique(bad[:, 0])), a[i].tolist(), b[i].tolist()))
