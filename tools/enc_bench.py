"""BASELINE configs[2] (kf=1: intra-only) and configs[3] (kf=64, speed level 1: inter frames with the
motion search) through th_encode_ycbcr_in / th_encode_packetout, measured in a plain process (ctypes only:
no torch, no second CUDA client in the process), ours and the reference interleaved.
Usage: enc_bench.py W H quality threads frames [kf [speed]].  Prints one JSON object; bench.py embeds it
as `encode_intra` / `encode_inter`."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def main():
    width, height, quality, threads, frames = (int(a) for a in sys.argv[1:6])
    kf = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    speed = int(sys.argv[7]) if len(sys.argv) > 7 else 1
    # encoder threads of OUR arm (independent streams; the reference arm always runs `threads`, one per core):
    # more streams than cores hide the wait for the device pre-pass, a waiting thread sleeps (blocking sync)
    ours_threads = int(sys.argv[8]) if len(sys.argv) > 8 else threads
    import support as S
    import th_streams as streams
    Lo = streams.lib()
    kind = "asm" if S.ref_available("asm") else "c"
    R = S.ref(kind)
    what = "intra-only encode (keyframe every frame)" if kf == 1 else "encode with inter frames (keyframe every %d, motion search)" % kf
    out = {"workload": "%dx%d 4:2:0 %s, q=%d, speed %d, %d timed frames per encoder stream, %d host cores"
           % (width, height, what, quality, speed, frames - 1, threads), "host_threads": threads, "unit": "frames/s"}

    def one(L, nthr=threads):
        h, b = C.c_uint64(), C.c_long()
        secs = L.refh_encode_time_mt(width, height, frames, quality, kf, speed, 30, 12345, nthr, C.byref(h), C.byref(b))
        assert secs > 0, "encode failed"
        return secs, h.value, b.value
    if ours_threads > threads:
        from theora_b200 import abi as _abi
        _abi.lib().ocg_set_blocking_sync(1)
    st = streams.EncBackendStats()
    Lo.ocg_backend_set_enc_mode(streams.ENC_AUTO)
    one(Lo, ours_threads)  # warm-up: contexts, pinned pools
    Lo.ocg_backend_get_enc_stats(None, 1)
    ours, refs = [], []
    for _ in range(3):  # interleaved, so that drifts of the host's speed hit both sides alike
        ours.append(one(Lo, ours_threads))
        refs.append(one(R))
    Lo.ocg_backend_get_enc_stats(C.byref(st), 0)
    secs, hsh, nbytes = sorted(ours)[1]
    rsecs, rhsh, rbytes = sorted(refs)[1]
    out["value"] = (frames - 1) * ours_threads / secs
    out["encoder_streams"] = ours_threads
    out["api"] = "th_encode_ycbcr_in + th_encode_packetout (reference host code, B200 back-end)"
    out["device_frames"] = int(st.frames)
    out["prepass_ms_per_frame"] = 1e3 * st.prepass_seconds / max(st.prepass_frames, 1)
    out["flush_ms_per_frame"] = 1e3 * st.flush_seconds / max(st.frames, 1)
    out["h2d_bytes_per_frame"] = int(st.h2d_bytes / max(st.prepass_frames, 1))
    out["d2h_bytes_per_frame"] = int(st.d2h_bytes / max(st.prepass_frames, 1))
    if kf > 1:
        calls = st.satd_lookups + st.satd_host
        out["prepass_breakdown_ms"] = {"queue": 1e3 * st.me_queue_seconds / max(st.me_frames, 1),
                                       "wait_for_results": 1e3 * st.me_sync_seconds / max(st.me_frames, 1),
                                       "wait_for_previous_flush": 1e3 * st.prev_wait_seconds / max(st.prepass_frames, 1)}
        out["motion_analysis"] = {"device_passes": int(st.me_frames), "gold_refinements": int(st.me_gold_refines),
                                  "gold_searches_redone": int(st.me_repairs)}
        out["block_metric_calls"] = {"satd_from_device_tables": int(st.satd_lookups), "satd_on_host": int(st.satd_host),
                                     "satd_table_hit_rate": st.satd_lookups / calls if calls else None,
                                     "skip_ssd_from_device_table": int(st.ssd_lookups),
                                     "coded_block_ssd_on_host": int(st.ssd_host),
                                     "intra_satd_from_device_table": int(st.intra_satd_lookups),
                                     "sub_fdct_quant_from_device_tables": int(st.fdct_quant_lookups),
                                     "sub_fdct_quant_on_host": int(st.fdct_quant_host)}
    from theora_b200 import abi
    prep, launch, nfl, bs, nb = C.c_double(), C.c_double(), C.c_long(), C.c_double(), C.c_long()
    abi.lib().ocg_flush_profile(C.byref(prep), C.byref(launch), C.byref(nfl), 0)
    abi.lib().ocg_flush_profile_builds(C.byref(bs), C.byref(nb))
    out["flush_profile"] = {"prepare_us": 1e6 * prep.value / max(nfl.value, 1), "graph_launch_us": 1e6 * launch.value / max(nfl.value, 1),
                            "flushes": nfl.value, "graph_builds": nb.value, "graph_build_ms_each": 1e3 * bs.value / max(nb.value, 1)}
    out["cpu_baseline"] = {"value": (frames - 1) * threads / rsecs, "cores": threads,
                           "kind": "reference" if kind == "asm" else "reference (C path)"}
    out["timing"] = "median of 3 passes each, ours and the reference interleaved, in a process of its own"
    out["packets_identical_to_reference"] = bool(all((o[1], o[2]) == (rhsh, rbytes) for o in ours))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
