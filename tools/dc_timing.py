"""Device time of a single-frame submit with and without the DC wave-front
kernel (diagnostic): real 1080p lists, one stream, CUDA events around
ocg_dec_submit (no frame download)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import theora_b200 as T  # noqa: E402
import th_streams as streams  # noqa: E402
import th_workload as wl  # noqa: E402


def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
    blob = wl.synth_stream(w, h, 24, 32, 64)
    for name, mode in (("host-DC lists", streams.DC_HOST), ("device-DC lists", streams.DC_DEVICE)):
        g, works, _ = streams.capture_stream_work(blob, streams.BACKEND_GPU, dc_mode=mode, expand=streams.EXPAND_REFERENCE)
        ctx = T.Context(g, 0)
        stream = torch.cuda.ExternalStream(ctx.stream)
        per = []
        for fi, wk in enumerate(works):
            if wk is None:
                continue
            if wk.dc_residual:
                wk.dc_residual = 1  # stand-alone submit: the wave-front kernel runs inside the flush
            ts = []
            for rep in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record()
                    ctx.submit(wk, None)
                    e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            per.append((fi, wk.ncoded, float(np.median(ts[1:]))))
        ctx.close()
        print(name, "median ms/frame %.4f" % np.median([p[2] for p in per]),
              "keyframe %.4f" % per[0][2], "frames", [(p[0], p[1], round(p[2], 3)) for p in per[:6]])


if __name__ == "__main__":
    main()
