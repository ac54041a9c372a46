"""BASELINE configs[1] end to end through th_decode_packetin / th_decode_ycbcr_out, measured in a plain
process (ctypes only: no torch, no second CUDA client in the process): T stream threads, packets in host
memory -> frames in host memory, ours and the reference interleaved.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def main():
    path, threads, ref_threads = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])  # ref_threads 0: no reference pass
    with_ref = ref_threads > 0
    dc_mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1   # streams.DC_HOST
    blocking = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    expand = int(sys.argv[6]) if len(sys.argv) > 6 else 0     # OCG_EXPAND_DEVICE
    pplevel = int(sys.argv[7]) if len(sys.argv) > 7 else 0    # TH_DECCTL_SET_PPLEVEL for every timed decoder
    import support as S
    import th_streams as streams
    blob = open(path, "rb").read()
    Lo = streams.lib()
    buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
    h = Lo.refh_stream_from_blob(buf, len(blob))
    nframes = Lo.refh_stream_npackets(h) - 3
    Lo.ocg_backend_set_mode(streams.BACKEND_GPU)
    Lo.ocg_backend_set_dc_mode(dc_mode)
    Lo.ocg_backend_set_expand_mode(expand)
    from theora_b200 import abi
    abi.lib().ocg_set_blocking_sync(blocking)
    if os.environ.get("OCG_OUT_DMA"):  # A/B of the hand-over path (diagnostic)
        abi.lib().ocg_set_out_dma(int(os.environ["OCG_OUT_DMA"]))
    Lo.refh_set_timed_pplevel(pplevel)
    Lo.refh_decode_time(h, min(threads, 2), 1, None)  # warm-up
    R = hr = None
    kind = None
    if with_ref:
        kind = "asm" if S.ref_available("asm") else "c"
        R = S.ref(kind)
        hr = R.refh_stream_from_blob(buf, len(blob))
        R.refh_set_timed_pplevel(pplevel)
    st = streams.BackendStats()
    ours, refs = [], []
    hsh, rh = C.c_uint64(0), C.c_uint64(0)
    for _ in range(3):
        Lo.ocg_backend_get_stats(C.byref(st), 1)
        secs = Lo.refh_decode_time(h, threads, 1, C.byref(hsh))
        s2 = streams.BackendStats()
        Lo.ocg_backend_get_stats(C.byref(s2), 1)
        assert secs > 0
        ours.append((secs, s2.h2d_bytes, s2.d2h_bytes, s2.flush_seconds / max(s2.frames, 1),
                     s2.wait_seconds / max(s2.frames, 1)))
        if R is not None:
            rs = R.refh_decode_time(hr, ref_threads, 1, C.byref(rh))
            assert rs > 0
            refs.append(rs)
    prep, launch, nfl = C.c_double(), C.c_double(), C.c_long()
    abi.lib().ocg_flush_profile(C.byref(prep), C.byref(launch), C.byref(nfl), 1)
    bs, nb = C.c_double(), C.c_long()
    abi.lib().ocg_flush_profile_builds(C.byref(bs), C.byref(nb))
    ours.sort()
    secs, h2d, d2h, flush, wait = ours[1]
    out = {"secs": secs, "frames": threads * nframes, "h2d_bytes": int(h2d), "d2h_bytes": int(d2h),
           "flush_ms_per_frame": 1e3 * flush, "wait_ms_per_frame": 1e3 * wait, "hash": int(hsh.value), "threads": threads,
           "flush_prepare_us": 1e6 * prep.value / max(nfl.value, 1), "graph_launch_us": 1e6 * launch.value / max(nfl.value, 1),
           "graph_builds": nb.value, "graph_build_ms_each": 1e3 * bs.value / max(nb.value, 1)}
    if refs:
        out["ref_secs"] = sorted(refs)[1]
        out["ref_hash"] = int(rh.value)
        out["ref_kind"] = kind
        out["ref_threads"] = ref_threads
    print(json.dumps(out))


if __name__ == "__main__":
    main()
