#!/usr/bin/env python
"""Diagnostic: host<->device copy bandwidth of the box (pinned memory), per GPU and aggregate over
GPUs, plus host memcpy bandwidth.  Usage: pcie_probe.py [ngpus]"""
import json
import multiprocessing as mp
import os
import sys
import time


def worker(dev, q, go, nbytes, nstreams, secs):
    import torch
    torch.cuda.set_device(dev)
    res = {}
    for kind in ("d2h", "h2d"):
        hs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(nstreams)]
        ds = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(nstreams)]
        sts = [torch.cuda.Stream() for _ in range(nstreams)]
        torch.cuda.synchronize()
        go.wait()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < secs:
            for i in range(nstreams):
                with torch.cuda.stream(sts[i]):
                    for _ in range(8):
                        if kind == "d2h":
                            hs[i].copy_(ds[i], non_blocking=True)
                        else:
                            ds[i].copy_(hs[i], non_blocking=True)
                        n += 1
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[kind] = n * nbytes / dt / 1e9
    q.put((dev, res))


def memcpy_worker(q, go, nbytes, secs):
    import numpy as np
    a = np.ones(nbytes, np.uint8)
    b = np.empty(nbytes, np.uint8)
    go.wait()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < secs:
        np.copyto(b, a)
        n += 1
    q.put(n * nbytes / (time.perf_counter() - t0) / 1e9)


def main():
    ng = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    out = {"cores": len(os.sched_getaffinity(0))}
    ctx = mp.get_context("spawn")
    for nbytes, nstreams in ((3279360, 1), (3279360, 4), (3279360, 16), (64 << 20, 2)):
        for g in sorted({1, ng}):
            q, go = ctx.Queue(), ctx.Event()
            ps = [ctx.Process(target=worker, args=(d, q, go, nbytes, nstreams, 1.5)) for d in range(g)]
            for p in ps:
                p.start()
            time.sleep(20 if nbytes == 3279360 and nstreams == 1 else 8)
            go.set()
            r = [q.get() for _ in ps]
            for p in ps:
                p.join()
            out["copy_%dB_x%dstreams_%dgpu" % (nbytes, nstreams, g)] = {
                "d2h_GBps_total": sum(x[1]["d2h"] for x in r), "h2d_GBps_total": sum(x[1]["h2d"] for x in r)}
            print(json.dumps(out), flush=True)
    for nthr in (1, 4, out["cores"]):
        q, go = ctx.Queue(), ctx.Event()
        ps = [ctx.Process(target=memcpy_worker, args=(q, go, 3279360 * 8, 1.5)) for _ in range(nthr)]
        for p in ps:
            p.start()
        time.sleep(3)
        go.set()
        r = [q.get() for _ in ps]
        for p in ps:
            p.join()
        out["host_memcpy_GBps_%dthreads" % nthr] = sum(r)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
