"""BASELINE configs[4] on one rank: ONE 3840x2160 4:2:0 stream (seed 12345+rank) decoded and encoded through
th_decode_* / th_encode_* on this rank's GPU by one host thread each (a stream is serial); bench.py runs this
on every rank and aggregates ("8 independent streams, one per GPU").  With `ref_streams` > 0 the unmodified
reference also decodes/encodes that many streams on as many host cores.  Plain process, ctypes only.
Usage: config4_bench.py rank frames enc_frames ref_streams.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
W, H, Q = 3840, 2160, 32


def main():
    rank, frames, enc_frames, ref_streams = (int(a) for a in sys.argv[1:5])
    import th_harness_abi as HA
    import th_streams as streams
    import th_workload as wl
    Lo = streams.lib()
    seed = 12345 + rank
    # decode input: key frame every 8 so that the synthesis runs GOP-parallel (tooling, not timed)
    blob = wl.synth_stream(W, H, frames, Q, 8, seed=seed)
    buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
    h = Lo.refh_stream_from_blob(buf, len(blob))
    Lo.ocg_backend_set_mode(streams.BACKEND_GPU)
    Lo.ocg_backend_set_dc_mode(streams.DC_HOST)
    hsh = C.c_uint64(0)
    Lo.refh_decode_time(h, 1, 1, None)  # warm-up
    dec = sorted(Lo.refh_decode_time(h, 1, 1, C.byref(hsh)) for _ in range(3))[1]
    eh, eb = C.c_uint64(), C.c_long()
    Lo.ocg_backend_set_enc_mode(streams.ENC_AUTO)
    Lo.refh_encode_time_mt(W, H, 2, Q, 64, 1, 30, seed, 1, C.byref(eh), C.byref(eb))  # warm-up
    enc = sorted(Lo.refh_encode_time_mt(W, H, enc_frames, Q, 64, 1, 30, seed, 1, C.byref(eh), C.byref(eb)) for _ in range(2))[0]
    out = {"decode_secs": dec, "decode_frames": frames, "decode_hash": int(hsh.value),
           "encode_secs": enc, "encode_frames": enc_frames - 1, "encode_hash": int(eh.value)}
    if ref_streams > 0:
        R, kind = HA.load_reference()
        hr = R.refh_stream_from_blob(buf, len(blob))
        rh = C.c_uint64(0)
        rdec = sorted(R.refh_decode_time(hr, ref_streams, 1, C.byref(rh)) for _ in range(3))[1]
        reh, reb = C.c_uint64(), C.c_long()
        renc = R.refh_encode_time_mt(W, H, enc_frames, Q, 64, 1, 30, seed, ref_streams, C.byref(reh), C.byref(reb))
        out.update({"ref_kind": kind, "ref_streams": ref_streams, "ref_decode_secs": rdec, "ref_encode_secs": renc,
                    "ref_decode_hash": int(rh.value), "ref_encode_hash": int(reh.value)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
