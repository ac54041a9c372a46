"""Times the resident decode stages (CUDA events per kernel group, all streams in one launch set) for a list
of environment settings of the tuning switches, in one process: python tools/kernel_tune.py OCG_LF2_CFG 0 1 2 ...
Optional: --noisy (dense-coefficient stream), --frames N, --streams S."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    args = sys.argv[1:]
    noisy = "--noisy" in args
    if noisy:
        args.remove("--noisy")
    nfr, S = 24, 64
    if "--frames" in args:
        i = args.index("--frames"); nfr = int(args[i + 1]); del args[i:i + 2]
    if "--streams" in args:
        i = args.index("--streams"); S = int(args[i + 1]); del args[i:i + 2]
    var, vals = (args[0], args[1:]) if args else ("OCG_NONE", ["0"])
    import torch
    import theora_b200 as T
    from theora_b200 import abi
    import th_streams as streams
    import th_workload as wl
    L = abi.lib()
    if noisy:
        blob = wl.synth_stream(1920, 1080, nfr, 48, 64, noise_shift=26)
    else:
        blob = wl.synth_stream(1920, 1080, nfr, 32, 64)
    g, works, _ = streams.capture_stream_work(blob, streams.BACKEND_GPU, dc_mode=streams.DC_HOST, expand=streams.EXPAND_REFERENCE)
    works = [w for w in works if w is not None]
    ctxs = [T.Context(g, 0) for _ in range(S)]
    packs = [T.Pack(works, g.nfrags, 0) for _ in range(S)]
    torch.cuda.synchronize()
    for v in vals:
        os.environ[var] = v
        res = []
        for rep in range(3):
            if rep:
                L.ocg_profile_enable(1)
            for f in range(len(works)):
                T.run_batch(ctxs, packs, [f] * S, ctxs[0].stream)
            ctxs[0].sync()
            if rep:
                ms3, n3 = (C.c_double * 3)(), (C.c_long * 3)()
                abi.check(L.ocg_profile_collect(ms3, n3))
                L.ocg_profile_enable(0)
                res.append([ms3[i] / max(n3[i], 1) * 1e3 for i in range(3)])
        r = np.min(np.array(res), axis=0)
        print(json.dumps({var: v, "recon_us": round(float(r[0]), 2), "lf_us": round(float(r[1]), 2), "border_us": round(float(r[2]), 2)}), flush=True)


if __name__ == "__main__":
    main()
