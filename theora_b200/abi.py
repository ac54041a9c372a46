"""ctypes view of include/theora_b200.h (structs, dtypes, prototypes).

The shared library is the product; this module only describes its C ABI so
Python callers (tests, bench.py, the stream driver) can reach it.  There is no
Python or CPU implementation of any kernel here: if libtheora_b200.so is
missing, or no CUDA device is present, calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libtheora_b200.so")

OCG_FRAME_GOLD, OCG_FRAME_PREV, OCG_FRAME_SELF = 0, 1, 2
OCG_CLS_DC, OCG_CLS_3, OCG_CLS_10, OCG_CLS_FULL, OCG_NCLS = 0, 1, 2, 3, 4
OCG_MET_SAD, OCG_MET_SATD, OCG_MET_INTRA_SATD, OCG_MET_SSD, OCG_MET_INTRA_SAD = 0, 1, 2, 3, 4
OCG_MET_BORDER_SSD, OCG_MET_ACTIVITY, OCG_MET_SAD_THRESH = 5, 6, 7
INT32_MIN = -2 ** 31


class PlaneGeom(C.Structure):
    _fields_ = [("nhfrags", C.c_int32), ("nvfrags", C.c_int32), ("froffset", C.c_int32),
                ("nfrags", C.c_int32), ("ystride", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32), ("hpad", C.c_int32), ("vpad", C.c_int32),
                ("plane_off", C.c_int64)]


class Geometry(C.Structure):
    _fields_ = [("frame_width", C.c_int32), ("frame_height", C.c_int32), ("pixel_fmt", C.c_int32),
                ("nrefs", C.c_int32), ("nfrags", C.c_int32), ("reserved", C.c_int32),
                ("ref_frame_sz", C.c_int64), ("base_off", C.c_int64), ("planes", PlaneGeom * 3)]


class DecFrame(C.Structure):
    _fields_ = [("ref_idx", C.c_int32 * 3), ("lf_limit", C.c_int32), ("dc_quant", (C.c_uint16 * 2) * 3),
                ("ncoded", C.c_int32), ("intra_frame", C.c_int32),
                ("ncoeff_rows", C.c_int32), ("dc_residual", C.c_int32), ("recs", C.c_void_p), ("coeff_rows", C.c_void_p)]


class Staging(C.Structure):
    _fields_ = [("recs", C.c_void_p), ("coeff_rows", C.c_void_p)]


REC_DTYPE = np.dtype([("buf_off", "<i4"), ("mv", "<i2"), ("dc", "<i2"), ("coeff_row", "<u4"),
                      ("rowmask", "u1"), ("last_zzi", "u1"), ("refi", "u1"), ("pli_qti", "u1")])
ENC_FRAG_DTYPE = np.dtype([("src_off", "<i4"), ("ref_off0", "<i4"), ("ref_off1", "<i4"), ("aux", "<i4")])
assert REC_DTYPE.itemsize == 16 and ENC_FRAG_DTYPE.itemsize == 16
ME_TOPO_DTYPE = np.dtype([("frag_off", "<i4", (4,)), ("cn", "<i4", (4,)), ("ncn", "u1"), ("valid", "u1"),
                          ("pad", "u1", (6,))])
ME_MB_DTYPE = np.dtype([("analysis_mv", "<i2", (3, 2)), ("error", "<u2", (2,)), ("satd", "<u4", (2,)),
                        ("unref_mv", "<i2", (2,)), ("unref_satd", "<u4", (2,)), ("block_mv", "<i2", (4,)),
                        ("ref_mv", "<i2", (4,)), ("block_satd", "<u4", (4,)), ("ref_block_satd", "<u4", (4,)),
                        ("gold_ref_mv", "<i2"), ("pad0", "<u2"), ("gold_ref_satd", "<u4"), ("pad", "u1", (4,))])
assert ME_TOPO_DTYPE.itemsize == 40 and ME_MB_DTYPE.itemsize == 96
OCG_ME_REFINE_PREV, OCG_ME_REFINE_4MV, OCG_ME_NOSATD, OCG_ME_FAST, OCG_ME_DROPPED, OCG_ME_SPEC_GOLD = 1, 2, 4, 8, 16, 32


def cls_of_last_zzi(last_zzi):
    """state.c:967 / idct.c:327-329 class selection."""
    lz = np.asarray(last_zzi)
    return np.where(lz < 2, 0, np.where(lz <= 3, 1, np.where(lz <= 10, 2, 3))).astype(np.int32)


OCG_FRAG_UNCODED = 3


class FrameWork:
    """One frame of decoder block work held in numpy arrays (keeps them alive):
    `recs` has one record per fragment, in fragment-index order."""

    def __init__(self, ref_idx, lf_limit, dc_quant, recs, rows, dc_residual=0):
        self.dc_residual = int(dc_residual)
        self.ref_idx = tuple(int(x) for x in ref_idx)
        self.lf_limit = int(lf_limit)
        self.dc_quant = np.asarray(dc_quant, dtype=np.uint16).reshape(3, 2).copy()
        self.recs = np.ascontiguousarray(recs, dtype=REC_DTYPE)
        self.rows = np.ascontiguousarray(rows, dtype=np.int16).reshape(-1, 8)

    def as_struct(self):
        f = DecFrame()
        for i in range(3):
            f.ref_idx[i] = self.ref_idx[i]
            for j in range(2):
                f.dc_quant[i][j] = int(self.dc_quant[i, j])
        f.lf_limit = self.lf_limit
        f.ncoded = self.ncoded
        f.intra_frame = int(bool(np.all(self.recs["refi"] == OCG_FRAME_SELF)))
        f.ncoeff_rows = len(self.rows)
        f.dc_residual = self.dc_residual
        f.recs = self.recs.ctypes.data
        f.coeff_rows = self.rows.ctypes.data if len(self.rows) else None
        return f

    @property
    def coded_mask(self):
        return self.recs["refi"] != OCG_FRAG_UNCODED

    @property
    def ncoded(self):
        return int(self.coded_mask.sum())

    @property
    def nuncoded(self):
        return len(self.recs) - self.ncoded

    @property
    def ncls(self):
        """Coded fragments per sparsity class (state.c:967 / idct.c:327-329)."""
        cls = cls_of_last_zzi(self.recs["last_zzi"][self.coded_mask])
        return tuple(int((cls == k).sum()) for k in range(4))

    def list_bytes(self):
        return self.recs.nbytes + self.rows.nbytes

    def with_refs(self, ref_idx):
        return FrameWork(ref_idx, self.lf_limit, self.dc_quant, self.recs, self.rows, self.dc_residual)

    def to_dict(self, prefix):
        return {prefix + "ref_idx": np.array(self.ref_idx, np.int32), prefix + "lf": np.array([self.lf_limit], np.int32),
                prefix + "dcq": self.dc_quant, prefix + "recs": self.recs, prefix + "rows": self.rows}

    @staticmethod
    def from_dict(d, prefix):
        return FrameWork(d[prefix + "ref_idx"], d[prefix + "lf"][0], d[prefix + "dcq"], d[prefix + "recs"],
                         d[prefix + "rows"])


_PROTOS = {
    "ocg_version": (C.c_char_p, []),
    "ocg_last_error": (C.c_char_p, []),
    "ocg_device_count": (C.c_int, []),
    "ocg_geometry_init": (C.c_int, [C.POINTER(Geometry), C.c_int, C.c_int, C.c_int, C.c_int]),
    "ocg_geometry_frag_buf_offs": (None, [C.POINTER(Geometry), C.c_void_p]),
    "ocg_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Geometry), C.c_int]),
    "ocg_ctx_destroy": (None, [C.c_void_p]),
    "ocg_ctx_geometry": (C.POINTER(Geometry), [C.c_void_p]),
    "ocg_ctx_sync": (C.c_int, [C.c_void_p]),
    "ocg_dc_unpredict_supported": (C.c_int, [C.POINTER(Geometry)]),
    "ocg_dec_dc_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ocg_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "ocg_ctx_frame_devptr": (C.c_void_p, [C.c_void_p, C.c_int]),
    "ocg_ctx_upload_frame": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ocg_ctx_download_frame": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ocg_ctx_download_picture": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ocg_picture_bytes": (C.c_long, [C.POINTER(Geometry)]),
    "ocg_ctx_fill_frame": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ocg_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "ocg_host_unregister": (C.c_int, [C.c_void_p]),
    "ocg_dec_staging": (C.c_int, [C.c_void_p, C.POINTER(Staging)]),
    "ocg_dec_submit": (C.c_int, [C.c_void_p, C.POINTER(DecFrame), C.c_void_p]),
    "ocg_dec_flush": (C.c_int, [C.c_void_p, C.POINTER(DecFrame), C.c_void_p, C.c_int]),
    "ocg_dec_wait": (C.c_int, [C.c_void_p]),
    "ocg_dec_expand_setup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ocg_dec_flush_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "ocg_ctx_set_flush_graph": (None, [C.c_void_p, C.c_int]),
    "ocg_flush_profile_builds": (None, [C.POINTER(C.c_double), C.POINTER(C.c_long)]),
    "ocg_flush_profile": (None, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_long), C.c_int]),
    "ocg_pack_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(DecFrame), C.c_int, C.c_int, C.c_int]),
    "ocg_pack_destroy": (None, [C.c_void_p]),
    "ocg_pack_nframes": (C.c_int, [C.c_void_p]),
    "ocg_dec_run_batch": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int,
                                    C.c_void_p]),
    "ocg_set_stage_mask": (None, [C.c_int]),
    "ocg_launch_count": (C.c_long, []),
    "ocg_set_lf_tma": (None, [C.c_int]),
    "ocg_set_out_dma": (None, [C.c_int]),
    "ocg_set_blocking_sync": (None, [C.c_int]),
    "ocg_test_sleep_until": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int]),
    "ocg_me_nmbs": (C.c_int, [C.POINTER(Geometry)]),
    "ocg_me_topology": (C.c_int, [C.POINTER(Geometry), C.c_void_p]),
    "ocg_me_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "ocg_me_destroy": (None, [C.c_void_p]),
    "ocg_me_frame": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_void_p]),
    "ocg_me_frame_batch": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p]),
    "ocg_me_read": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ocg_me_write": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ocg_me_read_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ocg_me_write_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ocg_me_repair": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ocg_ctx_device": (C.c_int, [C.c_void_p]),
    "ocg_profile_enable": (None, [C.c_int]),
    "ocg_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_long)]),
    "ocg_enc_metrics_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "ocg_mcenc_search_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p]),
    "ocg_enc_fdct_quant_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ocg_mcenc_refine_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p]),
    "ocg_enc_inter_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int]),
    "ocg_enc_inter_destroy": (None, [C.c_void_p]),
    "ocg_enc_inter_border_slot": (C.c_int, [C.c_void_p, C.c_int]),
    "ocg_enc_inter_quant_tables": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "ocg_enc_inter_finish": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ocg_enc_inter_prepass": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ocg_pp_create": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ocg_pp_destroy": (None, [C.c_void_p]),
    "ocg_pp_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ocg_pp_download": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ocg_pp_download_variances": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ocg_enc_intra_reserve": (C.c_int, [C.c_void_p]),
    "ocg_enc_intra_prepass": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_void_p]),
}



class EncIntraTables(C.Structure):
    """ocg_enc_intra_tables (include/theora_b200.h): pinned host result tables."""
    _fields_ = [("satd", C.c_void_p), ("satd_dc", C.c_void_p), ("dct", C.c_void_p), ("qdct", C.c_void_p),
                ("nonzero", C.c_void_p)]


EXPORTED_SYMBOLS = tuple(sorted(_PROTOS))

_lib = None


def lib():
    """Loads libtheora_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C theora_b200/csrc` (this package has no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class OcgError(RuntimeError):
    pass


def check(code, what=""):
    if code < 0:
        raise OcgError("%s failed (%d): %s" % (what or "ocg call", code, lib().ocg_last_error().decode()))
    return code
