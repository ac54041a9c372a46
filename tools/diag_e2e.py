import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theora_b200 import streams, workload as wl
L = streams.lib()
blob = wl.synth_stream(1920, 1080, 40, 32, 64)
buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
h = L.refh_stream_from_blob(buf, len(blob))
st = streams.BackendStats()
for mode in (streams.BACKEND_RECORD, streams.BACKEND_GPU):
    L.ocg_backend_set_mode(mode)
    for T in (1, 2, 4, 8, 16):
        L.ocg_backend_get_stats(C.byref(st), 1)
        secs = L.refh_decode_time(h, T, 2, None)
        L.ocg_backend_get_stats(C.byref(st), 1)
        print("mode", mode, "threads", T, "fps %.1f" % (T * 2 * 40 / secs), "per-thread ms/frame %.2f" % (secs / 80 * 1e3),
              "flush ms/frame %.3f" % (1e3 * st.flush_seconds / max(st.frames, 1)), flush=True)
