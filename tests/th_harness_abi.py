"""TEST / BENCH PLUMBING.  ctypes prototypes of the harness tools/th_harness.c (refh_*), which is compiled into the
reference builds under oracle/_ref and, for the integrated library, into tools/libth_ocg_harness.so.  Imports
nothing from the product package, so bench.py's reference arm can use it on its own."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def bind_harness(L):
    L.refh_encode_synth.restype = C.c_void_p
    L.refh_encode_synth.argtypes = [C.c_int] * 8 + [C.c_uint]
    L.refh_encode_synth_recon.restype = C.c_void_p
    L.refh_encode_synth_recon.argtypes = [C.c_int] * 8 + [C.c_uint, C.c_void_p]
    L.refh_encode_synth_fmt.restype = C.c_void_p
    L.refh_encode_synth_fmt.argtypes = [C.c_int] * 8 + [C.c_uint, C.c_int, C.c_void_p]
    L.refh_encode_time_mt.restype = C.c_double
    L.refh_encode_time_mt.argtypes = [C.c_int] * 7 + [C.c_uint, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_long)]
    L.refh_stream_free.argtypes = [C.c_void_p]
    L.refh_stream_npackets.argtypes = [C.c_void_p]
    L.refh_stream_packet_size.argtypes = [C.c_void_p, C.c_int]
    L.refh_stream_packet_size.restype = C.c_long
    L.refh_stream_blob_size.argtypes = [C.c_void_p]
    L.refh_stream_blob_size.restype = C.c_long
    L.refh_stream_to_blob.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    L.refh_stream_to_blob.restype = C.c_long
    L.refh_stream_from_blob.argtypes = [C.c_void_p, C.c_long]
    L.refh_stream_from_blob.restype = C.c_void_p
    L.refh_stream_append_data.argtypes = [C.c_void_p, C.c_void_p]
    L.refh_dec_open.restype = C.c_void_p
    L.refh_dec_open.argtypes = [C.c_void_p]
    L.refh_dec_close.argtypes = [C.c_void_p]
    L.refh_dec_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.refh_dec_next.argtypes = [C.c_void_p]
    L.refh_dec_rewind.argtypes = [C.c_void_p]
    L.refh_dec_hash.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.refh_dec_copy_frame.argtypes = [C.c_void_p, C.c_void_p]
    L.refh_dec_copy_frame.restype = C.c_long
    L.refh_dec_ctx.argtypes = [C.c_void_p]
    L.refh_dec_ctx.restype = C.c_void_p
    L.refh_dec_set_pplevel.argtypes = [C.c_void_p, C.c_int]
    L.refh_decode_time.restype = C.c_double
    L.refh_decode_time.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.refh_encode_time.restype = C.c_double
    L.refh_encode_time.argtypes = [C.c_int] * 7 + [C.c_uint, C.POINTER(C.c_long)]
    L.refh_synth_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    return L




def load_reference(kind=None):
    """The compiled, unmodified reference + harness: kind 'asm' (x86 SIMD build) or 'c'; None = the fastest present."""
    kinds = [kind] if kind else ["asm", "c"]
    for k in kinds:
        p = os.path.join(REF_DIR, "libth_%s.so" % k)
        if os.path.exists(p):
            return bind_harness(C.CDLL(p)), k
    raise RuntimeError("oracle/_ref is not built (make -C oracle ref)")
