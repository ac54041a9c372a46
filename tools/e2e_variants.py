"""Diagnostic: e2e decode (tools/dec_e2e_bench.py) over a grid of host-thread counts / wait policies /
DC modes on this box.  Usage: e2e_variants.py [WxH] [frames] [quality] ; prints one JSON line per variant."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    w, h = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080").split("x"))
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    q = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    import th_workload as wl
    blob = wl.synth_stream(w, h, frames, q, 64)
    path = "/tmp/e2e_variants.ogs"
    open(path, "wb").write(blob)
    cores = len(os.sched_getaffinity(0))
    # (threads, wait policy 0 spin / 1 yield / 2 sleep, dc 1 host / 0 device, reference pass)
    grid = [(1, 0, 1, 1), (cores, 0, 1, 1), (2 * cores, 2, 1, 0), (3 * cores, 2, 1, 0), (4 * cores, 2, 1, 0),
            (3 * cores, 1, 1, 0), (2 * cores, 2, 0, 0), (3 * cores, 2, 0, 0), (4 * cores, 2, 0, 0), (6 * cores, 2, 0, 0)]
    for threads, blocking, dc_mode, ref in grid:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dec_e2e_bench.py"), path, str(threads), str(threads if ref else 0),
                            str(dc_mode), str(blocking)], capture_output=True, text=True, timeout=150)
        if p.returncode != 0:
            print(json.dumps({"threads": threads, "error": p.stderr[-300:]}), flush=True)
            continue
        d = json.loads(p.stdout.strip().splitlines()[-1])
        d.update({"blocking": blocking, "dc_mode": "host" if dc_mode else "device", "cores": cores,
                  "fps": d["frames"] / d["secs"], "d2h_GBps": d["d2h_bytes"] / d["secs"] / 1e9})
        if "ref_secs" in d:
            d["ref_fps"] = threads * frames / d["ref_secs"]
        print(json.dumps(d), flush=True)


if __name__ == "__main__":
    main()
