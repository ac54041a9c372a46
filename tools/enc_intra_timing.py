"""Intra-only 1080p encode through th_encode_*: device back-end vs the compiled
reference (x86 SIMD build), same frames, same thread counts.  Prints one JSON
line per configuration (diagnostic; bench.py holds the reported numbers)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import support as S  # noqa: E402
from theora_b200 import streams  # noqa: E402


def run(lib, name, threads, frames=9, q=32, speed=1):
    h, b = C.c_uint64(), C.c_long()
    secs = lib.refh_encode_time_mt(1920, 1080, frames, q, 1, speed, 30, 12345, threads, C.byref(h), C.byref(b))
    out = {"impl": name, "threads": threads, "frames_timed": (frames - 1) * threads, "secs": secs,
           "fps": (frames - 1) * threads / secs if secs > 0 else None, "hash": "%016x" % h.value,
           "bytes": b.value, "quality": q, "speed": speed}
    return out


def main():
    G = streams.lib()
    R = S.ref("asm")
    ncpu = os.cpu_count() or 1
    for q in (32, 48):
        for t in (1, min(8, ncpu), min(16, ncpu)):
            G.ocg_backend_get_enc_stats(None, 1)
            g = run(G, "b200", t, q=q)
            st = streams.EncBackendStats()
            G.ocg_backend_get_enc_stats(C.byref(st), 0)
            g["prepass_ms"] = 1e3 * st.prepass_seconds / max(1, st.prepass_frames)
            g["flush_ms"] = 1e3 * st.flush_seconds / max(1, st.frames)
            r = run(R, "reference_asm", t, q=q)
            g["same_bitstream"] = (g["hash"], g["bytes"]) == (r["hash"], r["bytes"])
            print(json.dumps(g))
            print(json.dumps(r))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
