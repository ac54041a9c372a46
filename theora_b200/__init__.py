"""theora_b200 -- B200 (sm_100a) back-end for libtheora's 8x8 fragment pipeline.

The product is `libtheora_b200.so` (hand-written CUDA kernels behind a C ABI,
include/theora_b200.h) plus the reference-side vtable binding in
theora_b200/backend/.  This Python package is thin plumbing over that ABI for
tests, the benchmark and the multi-stream driver; it contains no codec
arithmetic and no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import abi
from .abi import (DecFrame, FrameWork, Geometry, OcgError, REC_DTYPE, ENC_FRAG_DTYPE, check, lib)

__all__ = ["abi", "Geometry", "FrameWork", "Context", "Pack", "geometry", "frag_buf_offs", "run_batch",
           "OcgError", "REC_DTYPE", "ENC_FRAG_DTYPE"]


def geometry(frame_width, frame_height, pixel_fmt=0, nrefs=3):
    g = Geometry()
    check(lib().ocg_geometry_init(C.byref(g), frame_width, frame_height, pixel_fmt, nrefs), "ocg_geometry_init")
    return g


def frag_buf_offs(g):
    offs = np.empty(g.nfrags, np.int32)
    lib().ocg_geometry_frag_buf_offs(C.byref(g), offs.ctypes.data)
    return offs


class Context:
    """Device state of one decoder/encoder instance (ocg_ctx)."""

    def __init__(self, geom, device=0):
        self.geom = geom
        self.h = C.c_void_p()
        check(lib().ocg_ctx_create(C.byref(self.h), C.byref(geom), device), "ocg_ctx_create")

    def close(self):
        if self.h:
            lib().ocg_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return lib().ocg_ctx_stream(self.h)

    def frame_devptr(self, buf):
        return lib().ocg_ctx_frame_devptr(self.h, buf)

    def sync(self):
        check(lib().ocg_ctx_sync(self.h), "ocg_ctx_sync")

    def upload_frame(self, buf, host):
        assert host.dtype == np.uint8 and host.size == self.geom.ref_frame_sz
        check(lib().ocg_ctx_upload_frame(self.h, buf, host.ctypes.data), "ocg_ctx_upload_frame")

    def download_frame(self, buf, host=None):
        if host is None:
            host = np.empty(self.geom.ref_frame_sz, np.uint8)
        check(lib().ocg_ctx_download_frame(self.h, buf, host.ctypes.data), "ocg_ctx_download_frame")
        self.sync()
        return host

    def fill_frame(self, buf, value):
        check(lib().ocg_ctx_fill_frame(self.h, buf, value), "ocg_ctx_fill_frame")

    def submit(self, work, host_out=None):
        """ocg_dec_submit with host lists; asynchronous."""
        f = work.as_struct()
        check(lib().ocg_dec_submit(self.h, C.byref(f), host_out.ctypes.data if host_out is not None else None),
              "ocg_dec_submit")


class Pack:
    """Device-resident copy of a run of frames' work lists (ocg_pack)."""

    def __init__(self, works, nfrags, device=0):
        arr = (DecFrame * len(works))()
        self._keep = works
        for i, w in enumerate(works):
            arr[i] = w.as_struct()
        self.h = C.c_void_p()
        check(lib().ocg_pack_create(C.byref(self.h), arr, len(works), nfrags, device), "ocg_pack_create")
        self._keep = None
        self.nframes = len(works)

    def close(self):
        if self.h:
            lib().ocg_pack_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_batch(ctxs, packs, frame_idx, stream=None):
    n = len(ctxs)
    ca = (C.c_void_p * n)(*[c.h for c in ctxs])
    pa = (C.c_void_p * n)(*[p.h for p in packs])
    ia = (C.c_int32 * n)(*frame_idx)
    check(lib().ocg_dec_run_batch(ca, pa, ia, n, stream), "ocg_dec_run_batch")
