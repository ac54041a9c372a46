"""TEST / BENCH PLUMBING (not part of the product package).  Driver for the integrated library
(theora_b200/backend/libth_ocg.so): the reference's own th_decode_* host code with the B200 vtable
back-end plugged in, reached through the harness tools/th_harness.c (tools/libth_ocg_harness.so).

Used by the tests and bench.py to (a) decode real Theora packets end to end on
the GPU through the public API and (b) capture the per-frame block work lists
the back-end uploads, so they can be replayed from HBM (ocg_pack) or checked
against the oracle.  No codec arithmetic happens in this module.
"""
import ctypes as C
import os

import numpy as np

from theora_b200 import abi
from theora_b200.abi import DecFrame, FrameWork, REC_DTYPE, Staging

OCG_LIB = os.path.join(abi.PKG_DIR, "backend", "libth_ocg.so")
# the test/bench harness (tools/th_harness.c) built against the integrated library: NOT part of it
OCG_HARNESS = os.path.join(os.path.dirname(abi.PKG_DIR), "tools", "libth_ocg_harness.so")
BACKEND_GPU, BACKEND_RECORD = 0, 1

CAPTURE_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(DecFrame), C.POINTER(Staging))


class BackendStats(C.Structure):
    _fields_ = [("frames", C.c_long), ("coded_frags", C.c_long), ("uncoded_frags", C.c_long),
                ("coeff_rows", C.c_long), ("h2d_bytes", C.c_long), ("d2h_bytes", C.c_long),
                ("flush_seconds", C.c_double), ("wait_seconds", C.c_double)]


class EncBackendStats(C.Structure):
    """ocg_enc_backend_stats (theora_b200/backend/ocg_backend.h)."""
    _fields_ = [("frames", C.c_long), ("prepass_frames", C.c_long), ("coeff_rows", C.c_long),
                ("h2d_bytes", C.c_long), ("d2h_bytes", C.c_long), ("prepass_seconds", C.c_double),
                ("flush_seconds", C.c_double), ("me_frames", C.c_long), ("me_gold_refines", C.c_long),
                ("me_repairs", C.c_long), ("satd_lookups", C.c_long), ("satd_host", C.c_long),
                ("ssd_lookups", C.c_long), ("ssd_host", C.c_long), ("intra_satd_lookups", C.c_long),
                ("fdct_quant_lookups", C.c_long), ("fdct_quant_host", C.c_long),
                ("me_queue_seconds", C.c_double), ("me_sync_seconds", C.c_double),
                ("prev_wait_seconds", C.c_double), ("me_prep_seconds", C.c_double)]


ENC_AUTO, ENC_HOST = 0, 1
DC_DEVICE, DC_HOST, DC_DEVICE_AHEAD = 0, 1, 2
EXPAND_DEVICE, EXPAND_REFERENCE = 0, 1

_lib = None


def available():
    return os.path.exists(OCG_LIB) and os.path.exists(OCG_HARNESS) and os.path.exists(abi.LIB_PATH)


def lib():
    """libth_ocg.so with the harness (tools/th_harness.c) and back-end controls bound."""
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("%s not built (needs the reference sources at build time): "
                               "make -C theora_b200/backend" % OCG_LIB)
        abi.lib()  # make sure the product library is resolvable first
        # refh_*; th_* and the back-end controls resolve through its dependency on libth_ocg.so (local scope:
        # the compiled reference under oracle/_ref defines the same th_* names and must keep its own)
        L = C.CDLL(OCG_HARNESS)
        L.ocg_backend_set_mode.argtypes = [C.c_int]
        L.ocg_backend_set_device.argtypes = [C.c_int]
        L.ocg_backend_set_dc_mode.argtypes = [C.c_int]
        L.ocg_backend_set_expand_mode.argtypes = [C.c_int]
        L.ocg_backend_set_capture.argtypes = [CAPTURE_FN, C.c_void_p]
        L.ocg_backend_get_stats.argtypes = [C.POINTER(BackendStats), C.c_int]
        L.ocg_backend_set_enc_mode.argtypes = [C.c_int]
        L.ocg_backend_get_enc_stats.argtypes = [C.POINTER(EncBackendStats), C.c_int]
        _bind_harness(L)
        _lib = L
    return _lib


from th_harness_abi import bind_harness as _bind_harness  # noqa: E402


def _copy(ptr, nbytes, dtype):
    if nbytes == 0 or not ptr:
        return np.zeros(0, dtype)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


class Capture:
    """Collects every flushed frame's lists as FrameWork objects."""

    def __init__(self, nfrags):
        self.nfrags = nfrags
        self.frames = []
        self._cb = CAPTURE_FN(self._on_frame)

    def _on_frame(self, user, fptr, sptr):
        f, st = fptr.contents, sptr.contents
        recs = _copy(st.recs, self.nfrags * 16, REC_DTYPE)
        rows = _copy(st.coeff_rows, f.ncoeff_rows * 16, np.int16).reshape(-1, 8)
        dcq = [[f.dc_quant[i][j] for j in range(2)] for i in range(3)]
        self.frames.append(FrameWork([f.ref_idx[i] for i in range(3)], f.lf_limit, dcq, recs, rows,
                                     dc_residual=int(f.dc_residual)))

    def install(self):
        lib().ocg_backend_set_capture(self._cb, None)

    @staticmethod
    def uninstall():
        lib().ocg_backend_set_capture(C.cast(None, CAPTURE_FN), None)


def capture_stream_work(stream_blob, mode=BACKEND_RECORD, max_frames=None, dc_mode=DC_HOST, expand=None):
    """Decodes `stream_blob` (tools/th_harness.c serialisation) through the
    integrated library and returns (info, [FrameWork per decoded frame],
    [decoded frame bytes per frame or None in record mode]).  dc_mode: DC_DEVICE
    = the lists carry DC residuals (FrameWork.dc_residual), DC_HOST = final DCs.
    expand: EXPAND_REFERENCE = the host records per-fragment lists (what FrameWork holds);
    EXPAND_DEVICE (the product default on the GPU) = the device walks the token lists itself, there are no
    host lists and every FrameWork is None.  Default: lists in record mode, the product path on the GPU."""
    L = lib()
    if expand is None:
        expand = EXPAND_REFERENCE if mode == BACKEND_RECORD else EXPAND_DEVICE
    L.ocg_backend_set_mode(mode)
    L.ocg_backend_set_dc_mode(dc_mode)
    L.ocg_backend_set_expand_mode(expand)
    buf = (C.c_uint8 * len(stream_blob)).from_buffer_copy(stream_blob)
    sh = L.refh_stream_from_blob(buf, len(stream_blob))
    assert sh, "bad stream blob"
    d = L.refh_dec_open(sh)
    if not d:
        L.refh_stream_free(sh)
        L.ocg_backend_set_mode(BACKEND_GPU)
        raise RuntimeError("th_decode_alloc failed through the B200 back-end: %s" %
                           abi.lib().ocg_last_error().decode())
    info = (C.c_int * 8)()
    L.refh_dec_info(d, info)
    fw, fh, fmt = info[0], info[1], info[6]
    g = abi.Geometry()
    abi.check(abi.lib().ocg_geometry_init(C.byref(g), fw, fh, fmt, 3))
    cap = Capture(g.nfrags)
    if expand == EXPAND_REFERENCE:
        cap.install()
    outs = []
    try:
        n = 0
        while max_frames is None or n < max_frames:
            before = len(cap.frames)
            ret = L.refh_dec_next(d)
            if ret == 1000:
                break
            assert ret >= 0, "th_decode_packetin returned %d" % ret
            if len(cap.frames) == before:  # TH_DUPFRAME: nothing to flush / no host lists
                cap.frames.append(None)
            if mode == BACKEND_GPU:
                cw = fw >> (0 if fmt & 1 else 1)
                ch = fh >> (0 if fmt & 2 else 1)
                o = np.empty(fw * fh + 2 * cw * ch, np.uint8)
                L.refh_dec_copy_frame(d, o.ctypes.data)
                outs.append(o)
            else:
                outs.append(None)
            n += 1
    finally:
        Capture.uninstall()
        L.refh_dec_close(d)
        L.refh_stream_free(sh)
        L.ocg_backend_set_mode(BACKEND_GPU)
        L.ocg_backend_set_dc_mode(DC_HOST)
        L.ocg_backend_set_expand_mode(EXPAND_DEVICE)
    return g, cap.frames, outs
