#!/usr/bin/env python
"""bench.py -- 1080p 4:2:0 Theora decode on B200 (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            our arm
  python bench.py --impl reference --gpus N --steps K ...  reference CPU arm

Workload: a deterministic synthetic 1920x1080 4:2:0 stream, 300 frames, keyframe
every 64 (5 intra + 295 inter), quality 32 (loop filter on), speed level 1,
produced on the box by the reference encoder host code (cached in /tmp).

  value   frames/s of the block pipeline (fused recon+copy, loop filter,
          borders) with every stream's per-frame lists RESIDENT IN HBM: S
          independent streams per GPU, one batched launch set per frame index.
          A step = all 300 frames of all S streams.
  e2e     frames/s through the reference's public API th_decode_packetin
          (reference host entropy decode + vtable back-end): packets in host
          memory -> decoded frame in host memory, H2D of the lists and D2H of
          the frame inside the timed region, T host threads = T streams.
  roofline  dominant kernel's algorithmic bytes / CUDA-event time.
  cpu_baseline  the unmodified reference (x86 SIMD build, oracle/_ref) decoding
          the same packets on the host cores.
One process per GPU; ranks share nothing but the setup broadcast of the packets.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one CUDA stream per decoder/encoder thread: ask the driver for its maximum of hardware work queues before
# anything initialises CUDA (the library's constructor does the same for C callers; see ocg_api.cu)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


RANK = env_int("RANK", 0)
WORLD = env_int("WORLD_SIZE", 1)
LOCAL_RANK = env_int("LOCAL_RANK", 0)


def log(*a):
    print("[bench r%d]" % RANK, *a, file=sys.stderr, flush=True)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.p, self.lines = gpu, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def reference_lib():
    """The compiled, unmodified reference (oracle/_ref) + harness; imports nothing of the product."""
    tdir = os.path.join(ROOT, "tests")  # test infra
    if tdir not in sys.path:
        sys.path.insert(0, tdir)
    import th_harness_abi as HA
    return HA.load_reference()


def stream_handle(L, blob):
    buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
    h = L.refh_stream_from_blob(buf, len(blob))
    assert h, "bad stream blob"
    return h


def time_reference(blob, threads, passes):
    """Frames/s of the unmodified reference decoding `blob` on `threads` cores."""
    L, kind = reference_lib()
    h = stream_handle(L, blob)
    nframes = L.refh_stream_npackets(h) - 3
    hsh = C.c_uint64(0)
    secs = L.refh_decode_time(h, threads, passes, C.byref(hsh))
    L.refh_stream_free(h)
    assert secs > 0, "reference decode failed"
    return threads * passes * nframes / secs, secs, kind, int(hsh.value)


def bench_encode_api(Lo, threads, width, height, quality, frames=13, kf=1, speed=1, streams_per_core=1.5):
    """BASELINE configs[2]: intra-only encode through th_encode_ycbcr_in/packetout.
    `threads` independent encoders (unmodified reference host code) on the B200
    back-end -- device pre-pass look-ups + recorded reconstruction -- next to the
    reference x86 SIMD build on the same threads and the same frames; the first
    frame of every encoder is outside the timed region.  Packets must be
    byte-identical.  Runs in a process of its own (tools/enc_bench.py: ctypes
    only), the way a C program would use the library: inside this process torch's
    CUDA client and thread pools share the driver and the cores with the 16 encoder
    threads, which costs the device path ~25 %.  Our arm runs `streams_per_core` independent encoder
    streams per core (a stream that waits for its device pre-pass sleeps and another stream's analysis
    takes the core, as in the decode e2e pass); the reference arm runs one per core (more gain it nothing)."""
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = env.get("CUDA_VISIBLE_DEVICES", "").split(",")[LOCAL_RANK] if env.get(
        "CUDA_VISIBLE_DEVICES") else str(LOCAL_RANK)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "enc_bench.py"), str(width), str(height),
                        str(quality), str(threads), str(frames), str(kf), str(speed),
                        str(max(threads, int(round(threads * streams_per_core))))], capture_output=True, text=True,
                       env=env, timeout=900)
    if p.returncode != 0:
        raise RuntimeError("enc_bench failed: " + p.stderr[-400:])
    return json.loads(p.stdout.strip().splitlines()[-1])


def bench_encode_kernels(torch, dev, peak, nframes=40):
    """Throughput + algorithmic-bytes roofline of the encoder batch kernels on 1080p luma
    (BASELINE configs[2]/[3] block work; informational, the headline is decode)."""
    import theora_b200 as T
    from theora_b200 import abi
    tdir = os.path.join(ROOT, "tests")
    if tdir not in sys.path:
        sys.path.insert(0, tdir)
    import mcgen as M
    L = abi.lib()
    rng = np.random.default_rng(7)
    w, h = 1920, 1088
    src, rfull, rsatd, bl, ystride = M.make_scene(rng, w=w, h=h, pad=16, shift=(3, 1), noise=3)
    fsz = src.size
    dsrc = torch.from_numpy(np.tile(src.reshape(-1), nframes)).to(dev)
    dref = torch.from_numpy(np.tile(rfull.reshape(-1), nframes)).to(dev)
    dsat = torch.from_numpy(np.tile(rsatd.reshape(-1), nframes)).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    nfr = (w // 8) * (h // 8)
    fy, fx = np.divmod(np.arange(nfr), w // 8)
    off1 = (fy * 8 * ystride + fx * 8).astype(np.int64)
    offs = (off1[None, :] + (np.arange(nframes) * fsz)[:, None]).reshape(-1)
    n = offs.size
    fr = np.zeros(n, abi.ENC_FRAG_DTYPE)
    fr["src_off"] = offs
    fr["ref_off0"] = offs + 3 + 1 * ystride
    fr["ref_off1"] = abi.INT32_MIN
    fr["aux"] = 4
    dfr = torch.from_numpy(fr.view(np.int32).reshape(n, 4)).to(dev)
    deq = rng.integers(8, 300, size=(18, 64)).astype(np.uint16)
    enq = np.zeros((18, 128), np.int16)
    for t in range(18):  # enquant.c:184-192 (m,l) pairs
        d2 = deq[t].astype(np.int64) << 1
        lg = np.floor(np.log2(d2)).astype(np.int64)
        enq[t, 0::2] = ((1 + (1 << (16 + lg)) // d2) - 0x10000).astype(np.int16)
        enq[t, 1::2] = lg
    ddeq, denq = torch.from_numpy(deq.view(np.int16)).to(dev), torch.from_numpy(enq).to(dev)
    od = torch.empty((n, 64), dtype=torch.int16, device=dev)
    oq = torch.empty((n, 64), dtype=torch.int16, device=dev)
    onz = torch.empty(n, dtype=torch.int32, device=dev)
    ov = torch.empty(n, dtype=torch.int32, device=dev)
    odc = torch.empty(n, dtype=torch.int32, device=dev)
    nmb = (w // 16) * (h // 16)
    mb1 = np.zeros(nmb, M.MB_IN)
    i = 0
    for my in range(0, h, 16):
        for mx in range(0, w, 16):
            mb1[i]["frag_off"] = [(my + by) * ystride + mx + bx for by in (0, 8) for bx in (0, 8)]
            mb1[i]["cand"][3] = (0, 0)
            mb1[i]["setb0"], mb1[i]["ncand"], mb1[i]["t2_base"], mb1[i]["is_prev"] = 4, 5, 0, 1
            i += 1
    mbs = np.tile(mb1, nframes)
    mbs["frag_off"] += np.repeat(np.arange(nframes) * fsz, nmb)[:, None].astype(np.int32)
    dmb = torch.from_numpy(mbs.view(np.uint8).reshape(-1, 48)).to(dev)
    omb = torch.empty((len(mbs), 32), dtype=torch.uint8, device=dev)
    bs, br, bt = dsrc.data_ptr() + bl, dref.data_ptr() + bl, dsat.data_ptr() + bl

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {}
    ms = timed(lambda: abi.check(L.ocg_enc_fdct_quant_batch(bs, br, ystride, dfr.data_ptr(), n, ddeq.data_ptr(),
                                                            denq.data_ptr(), od.data_ptr(), oq.data_ptr(),
                                                            onz.data_ptr(), st)))
    out["fdct_quant_inter"] = {"ms": ms, "blocks_per_s": n / (ms * 1e-3), "alg_GBps": n * 384 / (ms * 1e-3) / 1e9,
                               "frac": n * 384 / (ms * 1e-3) / 1e9 / peak,
                               "luma_frames_per_s": nframes / (ms * 1e-3)}
    for name, metric, nbytes in (("sad", 0, 136), ("satd", 1, 136), ("ssd", 3, 136)):
        ms = timed(lambda: abi.check(L.ocg_enc_metrics_batch(metric, bs, br, ystride, dfr.data_ptr(), n,
                                                             ov.data_ptr(), odc.data_ptr(), st)))
        out[name] = {"ms": ms, "blocks_per_s": n / (ms * 1e-3), "alg_GBps": n * nbytes / (ms * 1e-3) / 1e9,
                     "frac": n * nbytes / (ms * 1e-3) / 1e9 / peak}
    ms = timed(lambda: abi.check(L.ocg_mcenc_search_batch(bs, br, bt, ystride, dmb.data_ptr(), omb.data_ptr(),
                                                          len(mbs), st)), reps=5)
    res = omb.cpu().numpy().view(M.MB_OUT).reshape(-1)
    out["mcenc_search"] = {"ms": ms, "macro_blocks_per_s": len(mbs) / (ms * 1e-3),
                           "frames_per_s_one_ref": nframes / (ms * 1e-3),
                           "found_planted_vector": float(np.mean((res["best_vec"][:, 0] == 3) &
                                                                 (np.abs(res["best_vec"][:, 1]) == 1)))}
    # half-pel refinement around the vectors the search just found (1MV + 4MV, SATD2)
    rin = np.zeros(len(mbs), M.REF_IN)
    rin["frag_off"] = mbs["frag_off"]
    rin["vec"] = res["best_vec"]
    rin["block_vec"] = res["block_vec"]
    rin["satd"] = res["satd"]
    rin["block_satd"] = res["block_satd"]
    drin = torch.from_numpy(rin.view(np.uint8).reshape(-1, 48)).to(dev)
    orf = torch.empty((len(mbs), 32), dtype=torch.uint8, device=dev)
    ms = timed(lambda: abi.check(L.ocg_mcenc_refine_batch(bs, bt, ystride, drin.data_ptr(), orf.data_ptr(),
                                                          len(mbs), 3, st)), reps=5)
    # 64 two-tap 8x8 SATD2 scores per macro block, 200 algorithmic bytes each
    out["mcenc_refine_1mv_4mv"] = {"ms": ms, "macro_blocks_per_s": len(mbs) / (ms * 1e-3),
                                   "satd2_blocks_per_s": 64 * len(mbs) / (ms * 1e-3),
                                   "alg_GBps": 64 * len(mbs) * 200 / (ms * 1e-3) / 1e9,
                                   "frames_per_s_one_ref": nframes / (ms * 1e-3)}
    out["config"] = ("1920x1088 luma, %d frames per launch (source + reference = %.0f MB, larger than the 126 MB L2), "
                     "inter residual vs (3,1)-displaced reference" % (nframes, 2 * nframes * fsz / 1e6))
    return out


def bench_me_frame(torch, dev, ncores, streams_n=64, nframes=4):
    """BASELINE configs[3] analysis front-end: oc_mcenc_search + refinements for EVERY macro block of
    1080p frames (both reference frames, half-pel refinement, 4MV), `streams_n` independent streams per
    launch set on the device, next to the unmodified reference running the same loop
    (oc_mcenc_search/refine1mv/refine4mv inside a th_encode_alloc context) on the host cores.
    Device inputs are resident in HBM; results are compared bit for bit on stream 0."""
    import theora_b200 as T
    from theora_b200 import abi
    tdir = os.path.join(ROOT, "tests")
    if tdir not in sys.path:
        sys.path.insert(0, tdir)
    import megen
    import support as S
    L = abi.lib()
    fw, fh = 1920, 1088
    g = S.make_geometry(fw, fh, 0, 6)
    rng = np.random.default_rng(11)
    orig, recon = megen.scene_buffers(g, rng, nframes + 1, motion=(3, 1))
    n = L.ocg_me_nmbs(C.byref(g))
    flags = abi.OCG_ME_REFINE_PREV | abi.OCG_ME_REFINE_4MV
    ctxs, mes = [], []
    for _ in range(streams_n):
        ctx = T.Context(g, dev.index or 0)
        me = C.c_void_p()
        abi.check(L.ocg_me_create(C.byref(me), ctx.h, None), "ocg_me_create")
        ctxs.append(ctx)
        mes.append(me)
    arr = (C.c_void_p * streams_n)(*[m.value for m in mes])
    bufs = (C.c_int * (5 * streams_n))(*([0, 1, 2, 3, 4] * streams_n))
    st = ctxs[0].stream
    stream = torch.cuda.ExternalStream(st, device=dev)
    zero = np.zeros(n, abi.ME_MB_DTYPE)

    def load(t):
        for ctx in ctxs:
            for role, buf in enumerate([orig[t], orig[t - 1], orig[0], recon[t - 1], recon[0]]):
                ctx.upload_frame(role, buf)
            ctx.sync()

    # timed: frame t=2 (history from frame 1 present), repeated from the same starting state
    load(1)
    abi.check(L.ocg_me_frame_batch(arr, bufs, streams_n, flags, st), "ocg_me_frame_batch")
    ctxs[0].sync()
    state1 = np.zeros(n, abi.ME_MB_DTYPE)
    abi.check(L.ocg_me_read(mes[0], state1.ctypes.data), "read")
    load(2)
    times = []
    for rep in range(5):
        for me in mes:
            abi.check(L.ocg_me_write(me, state1.ctypes.data), "write")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            abi.check(L.ocg_me_frame_batch(arr, bufs, streams_n, flags, st), "ocg_me_frame_batch")
            e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times[1:]))
    got = np.zeros(n, abi.ME_MB_DTYPE)
    abi.check(L.ocg_me_read(mes[0], got.ctypes.data), "read")
    # one stream alone: the latency of the wave-front
    lat = []
    one = (C.c_void_p * 1)(mes[0].value)
    for rep in range(4):
        abi.check(L.ocg_me_write(mes[0], state1.ctypes.data), "write")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            abi.check(L.ocg_me_frame_batch(one, bufs, 1, flags, st), "ocg_me_frame_batch")
            e1.record()
        e1.synchronize()
        lat.append(e0.elapsed_time(e1))
    for me in mes:
        L.ocg_me_destroy(me)
    for ctx in ctxs:
        ctx.close()
    out = {"workload": "1920x1088 luma, %d macro blocks/frame, oc_mcenc_search (PREV+GOLD) + refine1mv(PREV) + refine4mv, "
           "global motion (3,1) px/frame + noise" % int(n), "streams_per_launch": streams_n,
           "ms_per_launch": ms, "frames_per_s": streams_n / (ms * 1e-3),
           "macro_blocks_per_s": streams_n * 8160 / (ms * 1e-3), "single_stream_latency_ms": float(np.median(lat[1:]))}
    if S.ref_available("c"):
        kind = "asm" if S.ref_available("asm") else "c"
        R = megen.bind_ref_me(S.ref(kind))
        nthr = ncores
        handles = [R.refh_me_open(fw, fh, 0) for _ in range(nthr)]
        outs = [np.zeros(n, abi.ME_MB_DTYPE) for _ in range(nthr)]

        def frame(i, t):
            fr = [orig[t], orig[t - 1], orig[0], recon[t - 1], recon[0]]
            ptrs = (C.c_void_p * 5)(*[f.ctypes.data for f in fr])
            R.refh_me_frame(handles[i], ptrs, flags, None, outs[i].ctypes.data)

        def worker(i, secs):
            frame(i, 1)
            t0 = time.perf_counter()
            frame(i, 2)
            secs[i] = time.perf_counter() - t0
        secs = [0.0] * nthr
        th = [threading.Thread(target=worker, args=(i, secs)) for i in range(nthr)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for hnd in handles:
            R.refh_me_close(hnd)
        topo = np.zeros(n, abi.ME_TOPO_DTYPE)
        L.ocg_me_topology(C.byref(g), topo.ctypes.data)
        try:
            megen.assert_me_equal(got, outs[0], topo["valid"], flags, "bench")
            same = True
        except AssertionError:
            same = False
        out["cpu_baseline"] = {"frames_per_s": nthr / max(secs), "cores": nthr, "kind": ("reference" if kind == "asm" else "reference (C path)") +
                               " (each call also copies the five 3.3 MB frames into the context)", "ms_per_frame_per_core": 1e3 * float(np.mean(secs))}
        out["identical_to_reference"] = same
    return out



def stage_roofline(T, L, abi, ctxs, packs, nframes, S, stage_bytes, peak, peak_src, traffic_files=()):
    """Per-kernel device time (CUDA events on the launching stream; all S streams in one launch set on one
    CUDA stream, so the event pairs bracket one kernel group each) -> roofline object."""
    L.ocg_profile_enable(1)
    for f in range(nframes):
        T.run_batch(ctxs, packs, [f] * S, ctxs[0].stream)
    ms3, n3 = (C.c_double * 3)(), (C.c_long * 3)()
    abi.check(L.ocg_profile_collect(ms3, n3))
    L.ocg_profile_enable(0)
    stage_names = ["recon+copy", "loop_filter", "borders"]
    kern = {}
    for i in range(3):
        if n3[i]:
            avg_ms = ms3[i] / n3[i]
            kern[stage_names[i]] = {"avg_ms": avg_ms, "launches_per_step": int(n3[i]),
                                    "share": ms3[i] / max(sum(ms3), 1e-9),
                                    "alg_GBps": stage_bytes[i] / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else None}
    dom = max(range(3), key=lambda i: ms3[i])
    achieved = stage_bytes[dom] / (ms3[dom] / n3[dom] * 1e-3) / 1e9 if n3[dom] and stage_bytes[dom] else 0.0
    traffic = None
    for name in traffic_files:
        # measured DRAM bytes per launch from the committed ncu captures (same streams-per-launch only)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            if int(tj["streams_per_launch"]) == S:
                traffic = tj["dram_bytes_per_launch"].get(stage_names[dom])
                break
        except Exception:
            continue
    out = {"bound": "hbm", "kernel": stage_names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
           "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
           "alg_bytes_per_launch": stage_bytes[dom], "kernels": kern}
    if traffic:
        # bytes that actually crossed the HBM interface per launch / launch time / peak
        out["frac_dram"] = traffic / (ms3[dom] / n3[dom] * 1e-3) / 1e9 / peak
    return out


_REAL_STDOUT = None


def quiet_stdout():
    """Everything that is not the final JSON line goes to stderr, including text native
    libraries write straight to fd 1 (NCCL's version banner)."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=64, help="independent resident streams per GPU")
    ap.add_argument("--stream-groups", type=int, default=4, help="CUDA streams the resident decoder streams are spread over")
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--quality", type=int, default=32)
    ap.add_argument("--kf", type=int, default=64)
    ap.add_argument("--threads", type=int, default=0, help="host threads for e2e / CPU arms (0 = all cores)")
    ap.add_argument("--e2e-threads-per-core", type=int, default=3,
                    help="decoder stream threads per host core of the e2e pass (the flush is asynchronous: a thread "
                         "that waits for its frame yields the core to another stream's entropy decode)")
    ap.add_argument("--e2e-dc", default="host", choices=["device", "host"], help="where the e2e pass undoes the DC prediction")
    ap.add_argument("--e2e-variants", action="store_true", help="also time other e2e configurations (reported as alternatives)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-noisy", action="store_true", help="skip the dense-coefficient roofline workload")
    ap.add_argument("--no-config4", action="store_true", help="skip the 3840x2160 decode+encode section (BASELINE configs[4])")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-postproc", action="store_true", help="skip the post-processing (pp level 6) e2e section")
    ap.add_argument("--no-encode-kernels", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    ncores = args.threads or host_cores()
    if args.impl == "ours" and WORLD > 1 and not args.threads:
        ncores = max(1, ncores // WORLD)  # the ranks share the host's cores
    workload = "1080p 4:2:0 decode, %d synthetic frames (kf=%d, q=%d)" % (args.frames, args.kf, args.quality)
    if (args.width, args.height) != (1920, 1080):
        workload = "%dx%d 4:2:0 decode, %d synthetic frames (kf=%d, q=%d)" % (args.width, args.height, args.frames,
                                                                               args.kf, args.quality)

    tdir = os.path.join(ROOT, "tests")
    if tdir not in sys.path:
        sys.path.insert(0, tdir)
    import th_workload as wl

    # ------------------------------------------------------------------ reference arm
    # (loads oracle/_ref only: neither the product library nor the integrated build)
    if args.impl == "reference":
        if RANK != 0:
            return 0
        blob = wl.synth_stream(args.width, args.height, args.frames, args.quality, args.kf, lib=reference_lib()[0])
        vals = []
        for i in range(args.warmup + args.steps):
            fps, secs, kind, _ = time_reference(blob, ncores, 1)
            if i >= args.warmup:
                vals.append((fps, secs))
            if i == 0 and secs > 60:  # keep the arm bounded
                break
        if not vals:
            vals = [(fps, secs)]
        fps = float(np.mean([v[0] for v in vals]))
        ms = float(np.mean([v[1] for v in vals])) * 1e3
        line = {"impl": "reference", "metric": "1080p decode frames/sec", "value": fps, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int16",
                "data": "synthetic", "config": {"workload": workload, "streams": ncores, "frames": args.frames},
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ncores,
                                 "kind": "reference" if kind == "asm" else "reference (C path)",
                                 "sample": "%d streams x %d frames, th_decode_packetin, x86 SIMD build" % (ncores, args.frames)},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    import theora_b200 as T
    from theora_b200 import abi, sharding
    import th_streams as streams
    if os.environ.get("OCG_LF_VARIANT"):  # A/B of the loop-filter kernels (diagnostic; 0 = default)
        T.lib().ocg_set_lf_tma(int(os.environ["OCG_LF_VARIANT"]))

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(LOCAL_RANK)
    dev = torch.device("cuda", LOCAL_RANK)
    if WORLD > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if WORLD > 1:
            dist.barrier()

    # setup: rank 0 synthesises the stream, one NCCL broadcast hands the packets to every rank
    t0 = time.time()
    if RANK == 0:
        blob = wl.synth_stream(args.width, args.height, args.frames, args.quality, args.kf)
        log("stream ready: %.1f MB in %.1fs" % (len(blob) / 1e6, time.time() - t0))
    else:
        blob = b""
    blob = sharding.broadcast_bytes(blob, 0, dev)

    # capture the per-frame lists by decoding once through the public API on this GPU
    Lo = streams.lib()
    Lo.ocg_backend_set_device(LOCAL_RANK)
    t0 = time.time()
    # resident packs are replayed, so they hold final DC values (DC_HOST, also the e2e default)
    g, works, outs = streams.capture_stream_work(blob, streams.BACKEND_GPU, dc_mode=streams.DC_HOST,
                                                 expand=streams.EXPAND_REFERENCE)
    works = [w for w in works if w is not None]
    nframes = len(works)
    outs = None
    log("captured %d frames of lists in %.1fs" % (nframes, time.time() - t0))
    alg = [wl.algorithmic_bytes(w) for w in works]
    recon_bytes_frame = float(np.mean([a[0] for a in alg]))
    lf_bytes_frame = float(np.mean([a[1] for a in alg]))
    list_bytes_frame = float(np.mean([w.list_bytes() for w in works]))

    # resident state: S contexts + S private copies of the lists
    S = args.streams
    t0 = time.time()
    ctxs = [T.Context(g, LOCAL_RANK) for _ in range(S)]
    packs = [T.Pack(works, g.nfrags, LOCAL_RANK) for _ in range(S)]
    torch.cuda.synchronize()
    log("resident: %d streams, %.2f GB of lists, %.1fs" % (S, S * list_bytes_frame * nframes / 1e9, time.time() - t0))
    stream = torch.cuda.ExternalStream(ctxs[0].stream, device=dev)
    L = abi.lib()

    # `groups` CUDA streams, S/groups decoder streams each: one group's ALU-bound loop filter overlaps
    # another group's memory-bound reconstruction
    G = max(1, min(args.stream_groups, S))
    bounds = [S * k // G for k in range(G + 1)]
    gstreams = [ctxs[bounds[k]].stream for k in range(G)]
    fork, joins = torch.cuda.Event(), [torch.cuda.Event() for _ in range(G)]

    def step():
        for f in range(nframes):
            for k in range(G):
                T.run_batch(ctxs[bounds[k]:bounds[k + 1]], packs[bounds[k]:bounds[k + 1]],
                            [f] * (bounds[k + 1] - bounds[k]), gstreams[k])

    def fork_groups():  # the other groups' streams start after everything queued on group 0's stream so far
        if G > 1:
            with torch.cuda.stream(stream):
                fork.record()
            for k in range(1, G):
                torch.cuda.ExternalStream(gstreams[k], device=dev).wait_event(fork)

    def join_groups():  # group 0's stream (where the timing events live) waits for the other groups
        for k in range(1, G):
            es = torch.cuda.ExternalStream(gstreams[k], device=dev)
            with torch.cuda.stream(es):
                joins[k].record()
            stream.wait_event(joins[k])

    sampler = ClockSampler(LOCAL_RANK)  # samples through warm-up + timed region (both under the same load)
    sampler.start()
    for _ in range(args.warmup):
        step()
    ctxs[0].sync()
    torch.cuda.synchronize()
    barrier()
    launches0 = L.ocg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    fork_groups()
    for _ in range(args.steps):
        step()
    join_groups()
    with torch.cuda.stream(stream):
        e1.record()
    e1.synchronize()
    torch.cuda.synchronize()
    launches = L.ocg_launch_count() - launches0
    clocks = sampler.stop()
    barrier()
    ms_total = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    value = WORLD * S * nframes * args.steps / (ms_total * 1e-3)

    torch.cuda.synchronize()
    peak, peak_src = load_peaks()
    roofline = stage_roofline(T, L, abi, ctxs, packs, nframes, S, [recon_bytes_frame * S, lf_bytes_frame * S, 0.0], peak, peak_src,
                              ("r2_traffic.json", "r1b_traffic.json", "r1_final_traffic.json"))
    for p in packs:
        p.close()
    packs = None
    # second workload for the kernel roofline only: a dense-coefficient stream (SURVEY 8(d) "high-noise"
    # variant), where the transform pass (pass B) carries the reconstruction stage instead of the copy pass
    roofline_noisy = None
    if not args.no_noisy:
        try:
            nblob = sharding.broadcast_bytes(
                wl.synth_stream(args.width, args.height, 32, 48, 64, noise_shift=26, lib=reference_lib()[0]) if RANK == 0 else b"", 0, dev)
            _, nworks, _ = streams.capture_stream_work(nblob, streams.BACKEND_GPU, dc_mode=streams.DC_HOST,
                                                       expand=streams.EXPAND_REFERENCE)
            nworks = [w for w in nworks if w is not None]
            nalg = [wl.algorithmic_bytes(w) for w in nworks]
            npacks = [T.Pack(nworks, g.nfrags, LOCAL_RANK) for _ in range(S)]
            for f in range(len(nworks)):  # warm-up
                T.run_batch(ctxs, npacks, [f] * S, ctxs[0].stream)
            torch.cuda.synchronize()
            roofline_noisy = stage_roofline(T, L, abi, ctxs, npacks, len(nworks), S,
                                            [float(np.mean([a[0] for a in nalg])) * S, float(np.mean([a[1] for a in nalg])) * S, 0.0],
                                            peak, peak_src, ("r2_traffic_noisy.json",))
            roofline_noisy["workload"] = ("%dx%d, %d frames, q=48 (loop filter off), noise_shift=26: %.0f KB of packets per frame, "
                                          "%.0f %% of the coded fragments need a transform" % (
                                              args.width, args.height, len(nworks), len(nblob) / len(nworks) / 1e3,
                                              100.0 * float(np.mean([1.0 - (w.ncls[0] / max(w.ncoded, 1)) for w in nworks]))))
            for p in npacks:
                p.close()
        except Exception as e:  # informational: never take the headline down
            roofline_noisy = {"error": repr(e)}
    for c in ctxs:
        c.close()

    # e2e through the public API: T host threads, packets in RAM -> frames in RAM.  Measured by
    # tools/dec_e2e_bench.py in a process of its own per rank (ctypes only, the way a C program uses the
    # library; inside this process torch's CUDA client and thread pools share the driver and the cores with
    # the stream threads), ours and -- on a single GPU -- the reference interleaved pass by pass.
    e2e = None
    cpu = None
    postproc = None
    if not args.no_e2e:
        blob_path = os.path.join("/tmp", "theora_b200_bench_rank%d.ogs" % RANK)
        with open(blob_path, "wb") as f:
            f.write(blob)
        env = dict(os.environ)
        vis = env.get("CUDA_VISIBLE_DEVICES", "")
        env["CUDA_VISIBLE_DEVICES"] = vis.split(",")[LOCAL_RANK] if vis else str(LOCAL_RANK)
        with_ref = int(RANK == 0 and WORLD == 1 and not args.no_cpu)

        npass = [0]

        def e2e_pass(nthreads, blocking, dc_mode, ref_threads, pplevel=0):
            # ranks start together and finished ranks SLEEP until the last one is done (a NCCL barrier would
            # spin a host core per waiting rank and slow the ranks that are still measuring)
            npass[0] += 1
            barrier()
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dec_e2e_bench.py"), blob_path, str(nthreads),
                                str(int(ref_threads)), str(dc_mode), str(int(blocking)), "0", str(int(pplevel))],
                               capture_output=True, text=True,
                               env=env, timeout=1200)
            if p.returncode != 0:
                raise RuntimeError("dec_e2e_bench failed: " + p.stderr[-400:])
            d = json.loads(p.stdout.strip().splitlines()[-1])
            sharding.quiet_barrier("e2e_pass_%d" % npass[0])
            secs = sharding.max_over_ranks(d["secs"], dev)
            r = {"value": WORLD * nthreads * nframes / secs, "unit": "frames/s",
                 "h2d_bytes_per_step": int(d["h2d_bytes"]), "d2h_bytes_per_step": int(d["d2h_bytes"]),
                 "host_threads": nthreads, "host_cores": ncores,
                 "wait": "yield (a waiting stream thread gives its core away)" if blocking else "spin",
                 "dc_unprediction": "device (wave-front kernel inside the flush)" if dc_mode == streams.DC_DEVICE else "host (table-driven routine in the hook)",
                 "token_expansion": "device (ocg_dec_flush_tokens: fragment words, vectors and token lists read in place)",
                 "flush": "one CUDA graph of kernels per frame, asynchronous until th_decode_ycbcr_out",
                 "api": "th_decode_packetin + th_decode_ycbcr_out (reference host code, B200 back-end)",
                 "flush_ms_per_frame": d["flush_ms_per_frame"], "wait_ms_per_frame": d["wait_ms_per_frame"],
                 "d2h_GBps": WORLD * d["d2h_bytes"] / secs / 1e9, "final_frame_hash": d["hash"],
                 "timing": "median of 3 passes, in a process of its own per rank"}
            return r, d
        # THE e2e configuration (fixed, not picked from measurements): stream threads per core and wait policy
        tpc = max(1, args.e2e_threads_per_core)
        dcm = streams.DC_DEVICE if args.e2e_dc == "device" else streams.DC_HOST
        e2e, raw = e2e_pass(ncores * tpc, tpc > 1, dcm, ncores if with_ref else 0)
        alts = []
        if args.e2e_variants:
            for nthr, blk, dm in ((ncores, False, streams.DC_HOST), (ncores, False, streams.DC_DEVICE),
                                  (2 * ncores, True, dcm), (4 * ncores, True, dcm)):
                a = e2e_pass(nthr, blk, dm, 0)[0]
                alts.append({k: a[k] for k in ("value", "host_threads", "wait", "dc_unprediction", "flush_ms_per_frame",
                                               "wait_ms_per_frame", "final_frame_hash")})
        # the same streams with the out-of-loop de-blocking + de-ringing filters on (TH_DECCTL_SET_PPLEVEL 6:
        # luma and chroma), device filters vs the reference's host filters
        if with_ref and not args.no_postproc:
            try:
                ppr, ppraw = e2e_pass(ncores * tpc, tpc > 1, dcm, ncores, pplevel=6)
                postproc = {"workload": workload + ", post-processing level 6 (de-blocking + de-ringing, luma and chroma)",
                            "value": ppr["value"], "unit": "frames/s", "host_threads": ppr["host_threads"],
                            "api": "th_decode_ctl(TH_DECCTL_SET_PPLEVEL) + th_decode_packetin + th_decode_ycbcr_out",
                            "cpu_baseline": {"value": ncores * nframes / ppraw["ref_secs"], "cores": ncores, "kind": "reference"},
                            "identical_to_reference": bool(ppraw["hash"] == ppraw["ref_hash"])}
            except Exception as e:
                postproc = {"error": repr(e)}
        e2e["alternatives"] = alts
        e2e["all_variants_same_output"] = all(a["final_frame_hash"] == e2e["final_frame_hash"] for a in alts)
        if "ref_secs" in raw:
            rfps = ncores * nframes / raw["ref_secs"]
            cpu = {"value": rfps, "unit": "frames/s", "cores": ncores,
                   "kind": "reference" if raw["ref_kind"] == "asm" else "reference (C path)",
                   "sample": "%d streams x %d frames via th_decode_packetin, %.1fs; median of 3 passes interleaved with ours"
                   % (ncores, nframes, raw["ref_secs"]), "final_frame_hash": raw["ref_hash"]}
            e2e["parity_with_cpu_baseline"] = bool(e2e["final_frame_hash"] == raw["ref_hash"])
        try:
            os.remove(blob_path)
        except OSError:
            pass
    if cpu is None and RANK == 0 and WORLD == 1 and not args.no_cpu:
        fps, secs, kind, ref_hash = sorted(time_reference(blob, ncores, 1) for _ in range(3))[1]
        cpu = {"value": fps, "unit": "frames/s", "cores": ncores,
               "kind": "reference" if kind == "asm" else "reference (C path)",
               "sample": "%d streams x %d frames via th_decode_packetin, %.1fs; median of 3" % (ncores, nframes, secs),
               "final_frame_hash": ref_hash}

    enc = None
    if RANK == 0 and not args.no_encode_kernels:
        try:
            enc = bench_encode_kernels(torch, dev, peak)
        except Exception as e:  # informational section: never take the headline down with it
            enc = {"error": repr(e)}

    me_frame = None
    if RANK == 0 and not args.no_encode_kernels:
        try:
            me_frame = bench_me_frame(torch, dev, ncores)
        except Exception as e:
            me_frame = {"error": repr(e)}

    enc_intra = enc_inter = None
    if RANK == 0 and WORLD == 1 and not args.no_e2e and not args.no_cpu:
        try:
            enc_intra = bench_encode_api(Lo, ncores, args.width, args.height, args.quality, streams_per_core=1.0)
        except Exception as e:
            enc_intra = {"error": repr(e)}
        try:
            # BASELINE configs[3]: key frame + inter frames with the motion search, speed level 1
            enc_inter = bench_encode_api(Lo, ncores, args.width, args.height, args.quality, frames=17, kf=64, speed=1, streams_per_core=1.5)
        except Exception as e:
            enc_inter = {"error": repr(e)}

    # BASELINE configs[4]: one 3840x2160 stream per GPU, decode and encode through the public API, every rank
    config4 = None
    if not args.no_e2e and not args.no_config4:
        try:
            env = dict(os.environ)
            vis = env.get("CUDA_VISIBLE_DEVICES", "")
            env["CUDA_VISIBLE_DEVICES"] = vis.split(",")[LOCAL_RANK] if vis else str(LOCAL_RANK)
            barrier()
            c4_frames, c4_enc = 24, 4
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "config4_bench.py"), str(RANK), str(c4_frames),
                                str(c4_enc), str(WORLD if RANK == 0 and not args.no_cpu else 0)], capture_output=True, text=True,
                               env=env, timeout=1200)
            if p.returncode != 0:
                raise RuntimeError("config4_bench failed: " + p.stderr[-400:])
            d = json.loads(p.stdout.strip().splitlines()[-1])
            sharding.quiet_barrier("config4")
            dsecs = sharding.max_over_ranks(d["decode_secs"], dev)
            esecs = sharding.max_over_ranks(d["encode_secs"], dev)
            config4 = {"workload": "%d independent 3840x2160 4:2:0 streams, one per GPU: %d frames decoded and %d frames "
                       "encoded (key frame + inter frames, speed 1) per stream through th_decode_* / th_encode_*, one host "
                       "thread per stream" % (WORLD, c4_frames, c4_enc - 1),
                       "decode_frames_per_s": WORLD * c4_frames / dsecs, "encode_frames_per_s": WORLD * (c4_enc - 1) / esecs,
                       "unit": "frames/s", "n_gpus": WORLD}
            if "ref_decode_secs" in d:
                config4["cpu_baseline"] = {"decode_frames_per_s": WORLD * c4_frames / d["ref_decode_secs"],
                                           "encode_frames_per_s": WORLD * (c4_enc - 1) / d["ref_encode_secs"],
                                           "cores": WORLD, "kind": "reference" if d["ref_kind"] == "asm" else "reference (C path)",
                                           "sample": "the same %d streams on %d host cores" % (WORLD, WORLD)}
                config4["identical_to_reference"] = bool(d["decode_hash"] == d["ref_decode_hash"] and
                                                         d["encode_hash"] == d["ref_encode_hash"])
        except Exception as e:
            config4 = {"error": repr(e)}

    if RANK == 0:
        line = {"metric": "1080p decode frames/sec", "value": value, "unit": "frames/s", "n_gpus": WORLD,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int16",
                "data": "synthetic",
                "config": {"workload": workload, "streams_per_gpu": S, "frames_per_stream": nframes,
                           "frame_units_per_step": S * nframes, "l2": "working set of a launch (%d streams x 3 x %.1f MB "
                           "frames + lists) exceeds the 126 MB L2" % (S, g.ref_frame_sz / 1e6),
                           "parallelism": "independent streams, %d per GPU on %d CUDA stream(s)" % (S, G)},
                "roofline": roofline, "roofline_dense_coefficients": roofline_noisy, "cpu_baseline": cpu, "e2e": e2e,
                "postprocess": postproc,
                "config4_2160p": config4, "encode_kernels": enc,
                "encode_intra": enc_intra, "encode_inter": enc_inter, "motion_analysis": me_frame,
                "gpu_launches": int(launches),
                "clocks": clocks}
        emit(line)
    if WORLD > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
