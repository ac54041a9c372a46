"""PCIe D2H ceiling of the box: pinned-memory copies by the copy engines, frame-sized chunks on several streams."""
import time, torch
dev = torch.device("cuda", 0)
for chunk_mb, nstreams in ((3.11, 1), (3.11, 4), (3.11, 16), (32, 4)):
    n = int(chunk_mb * 1e6)
    src = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(nstreams)]
    dst = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(nstreams)]
    st = [torch.cuda.Stream() for _ in range(nstreams)]
    reps = max(8, int(2e9 / n / nstreams))
    for w in range(2):
        torch.cuda.synchronize()
        t0 = time.time()
        for r in range(reps):
            for k in range(nstreams):
                with torch.cuda.stream(st[k]):
                    dst[k].copy_(src[k], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
    print("D2H %.2f MB chunks x %d streams: %.1f GB/s" % (chunk_mb, nstreams, reps * nstreams * n / dt / 1e9), flush=True)
# both directions at once (the e2e path also uploads ~0.3 MB per frame)
n = int(3.11e6)
src = torch.empty(n, dtype=torch.uint8, device=dev); dst = torch.empty(n, dtype=torch.uint8).pin_memory()
up_s = torch.empty(n, dtype=torch.uint8).pin_memory(); up_d = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.time()
for r in range(400):
    with torch.cuda.stream(s1): dst.copy_(src, non_blocking=True)
    with torch.cuda.stream(s2): up_d.copy_(up_s, non_blocking=True)
torch.cuda.synchronize(); dt = time.time() - t0
print("D2H with concurrent H2D: %.1f GB/s each way" % (400 * n / dt / 1e9))
