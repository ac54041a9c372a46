"""Shared test plumbing: ctypes views of the oracle, the compiled reference
(oracle/_ref, when present) and the product C-ABI library.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may touch
oracle/; the product package never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

u8p = C.POINTER(C.c_uint8)
i16p = C.POINTER(C.c_int16)
i32p = C.POINTER(C.c_int32)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)


def ptr(a, typ):
    return a.ctypes.data_as(typ)


# ---- C-ABI structs come from the product package ---------------------------
from theora_b200.abi import (DecFrame, ENC_FRAG_DTYPE, FrameWork, Geometry, INT32_MIN, PlaneGeom,  # noqa: E402,F401
                             REC_DTYPE, cls_of_last_zzi)


# ---- library loading ------------------------------------------------------
def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "theora_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return so


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        L = C.CDLL(build_oracle())
        L.oco_idct8x8.argtypes = [i16p, i16p, C.c_int]
        L.oco_mv_offsets.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int16]
        L.oco_state_frag_recon.argtypes = [u8p, u8p, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int16,
                                           i16p, C.c_int, C.c_uint16]
        L.oco_lflim.argtypes = [C.c_int, C.c_int]
        L.oco_loop_filter_init.argtypes = [C.POINTER(C.c_byte), C.c_int]
        for fn in (L.oco_loop_filter_plane_seq, L.oco_loop_filter_plane_cells):
            fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
        L.oco_borders_fill_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oco_geometry_init.argtypes = [C.POINTER(Geometry), C.c_int, C.c_int, C.c_int, C.c_int]
        L.oco_geometry_frag_buf_offs.argtypes = [C.POINTER(Geometry), i32p]
        L.oco_dec_frame.argtypes = [C.POINTER(Geometry), u8p, C.POINTER(DecFrame), C.c_int]
        L.oco_fdct8x8.argtypes = [i16p, i16p]
        L.oco_enquant_init.argtypes = [i16p, u16p]
        L.oco_quantize.argtypes = [i16p, i16p, u16p, i16p]
        L.oco_frag_sub.argtypes = [i16p, C.c_void_p, C.c_void_p, C.c_int]
        L.oco_frag_sub_128.argtypes = [i16p, C.c_void_p, C.c_int]
        for name in ("oco_frag_sad", "oco_frag_ssd"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            getattr(L, name).restype = C.c_uint
        L.oco_frag_sad_thresh.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
        L.oco_frag_sad_thresh.restype = C.c_uint
        L.oco_frag_sad2_thresh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
        L.oco_frag_sad2_thresh.restype = C.c_uint
        L.oco_frag_intra_sad.argtypes = [C.c_void_p, C.c_int]
        L.oco_frag_intra_sad.restype = C.c_uint
        L.oco_frag_satd.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_int]
        L.oco_frag_satd.restype = C.c_uint
        L.oco_frag_satd2.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oco_frag_satd2.restype = C.c_uint
        L.oco_frag_intra_satd.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_int]
        L.oco_frag_intra_satd.restype = C.c_uint
        L.oco_frag_border_ssd.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64]
        L.oco_frag_border_ssd.restype = C.c_uint
        L.oco_frag_copy2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oco_enc_metrics_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, u32p, i32p]
        L.oco_enc_fdct_quant_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, u16p, i16p,
                                               i16p, i16p, i32p]
        _oracle = L
    return _oracle


from th_harness_abi import bind_harness as _bind_harness  # noqa: E402


_ref = {}


def ref_available(kind="c"):
    return os.path.exists(os.path.join(REF_DIR, "libth_%s.so" % kind))


def ref(kind="c"):
    """The compiled, unmodified reference (+ harness). kind: 'c' or 'asm'."""
    if kind not in _ref:
        L = _bind_harness(C.CDLL(os.path.join(REF_DIR, "libth_%s.so" % kind)))
        if kind == "c":
            L.oc_idct8x8_c.argtypes = [i16p, i16p, C.c_int]
            L.oc_enc_fdct8x8_c.argtypes = [i16p, i16p]
            L.oc_enc_enquant_table_init_c.argtypes = [C.c_void_p, u16p]
            L.oc_enc_quantize_c.argtypes = [i16p, i16p, u16p, C.c_void_p]
            L.refh_mv_offsets.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
            L.refh_state_frag_recon.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int,
                                                C.c_int, C.c_int, i16p, C.c_int, C.c_int]
            L.refh_loop_filter_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
            L.refh_loop_filter_table.argtypes = [C.POINTER(C.c_byte), C.c_int]
            L.refh_borders_fill.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _ref[kind] = L
    return _ref[kind]


class Stream:
    """Packets of one Theora stream held by a harness library."""

    def __init__(self, lib, handle):
        self.lib, self.h = lib, handle

    @staticmethod
    def encode(lib, w, h, nframes, quality=48, kf=64, speed=1, noise_shift=30, seed=12345, f0=0, fmt=0):
        hnd = lib.refh_encode_synth_fmt(w, h, f0, nframes, quality, kf, speed, noise_shift, seed, fmt, None)
        assert hnd, "encoder failed"
        return Stream(lib, hnd)

    def to_bytes(self):
        n = self.lib.refh_stream_blob_size(self.h)
        buf = (C.c_uint8 * n)()
        assert self.lib.refh_stream_to_blob(self.h, buf, n) == n
        return bytes(buf)

    @staticmethod
    def from_bytes(lib, blob):
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        hnd = lib.refh_stream_from_blob(buf, len(blob))
        assert hnd, "bad stream blob"
        return Stream(lib, hnd)

    @property
    def nframes(self):
        return self.lib.refh_stream_npackets(self.h) - 3

    def packet_sizes(self):
        return [self.lib.refh_stream_packet_size(self.h, i) for i in range(self.lib.refh_stream_npackets(self.h))]

    def free(self):
        if self.h:
            self.lib.refh_stream_free(self.h)
            self.h = None


class Decoder:
    def __init__(self, lib, stream):
        self.lib = lib
        self.d = lib.refh_dec_open(stream.h)
        assert self.d, "decoder open failed"
        info = (C.c_int * 8)()
        lib.refh_dec_info(self.d, info)
        (self.fw, self.fh, self.pw, self.ph, self.px, self.py, self.fmt, self.nframes) = list(info)

    def next(self):
        return self.lib.refh_dec_next(self.d)

    def set_pplevel(self, level):
        return self.lib.refh_dec_set_pplevel(self.d, level)

    def hashes(self):
        h = (C.c_uint64 * 3)()
        self.lib.refh_dec_hash(self.d, h)
        return tuple(int(x) for x in h)

    def frame(self):
        cw = self.fw >> (0 if self.fmt & 1 else 1)
        ch = self.fh >> (0 if self.fmt & 2 else 1)
        out = np.empty(self.fw * self.fh + 2 * cw * ch, np.uint8)
        n = self.lib.refh_dec_copy_frame(self.d, out.ctypes.data)
        assert n == out.size
        return out

    def close(self):
        if self.d:
            self.lib.refh_dec_close(self.d)
            self.d = None


def fnv1a64(arr):
    """FNV-1a 64 over a uint8 array (vectorised in blocks is not possible; small inputs only)."""
    h = 14695981039346656037
    for b in np.asarray(arr, np.uint8).ravel().tolist():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def make_geometry(fw, fh, fmt=0, nrefs=3):
    g = Geometry()
    r = oracle().oco_geometry_init(C.byref(g), fw, fh, fmt, nrefs)
    assert r == 0, r
    return g


def planes_from_buffer(g, buf):
    """Top-down picture planes (frame_width x frame_height etc.) out of one padded buffer."""
    out = []
    for pli in range(3):
        p = g.planes[pli]
        stride = -p.ystride
        top_left = g.base_off + p.plane_off + (p.height - 1) * p.ystride
        rows = np.lib.stride_tricks.as_strided(buf[top_left:], shape=(p.height, p.width), strides=(stride, 1))
        out.append(np.ascontiguousarray(rows))
    return out
