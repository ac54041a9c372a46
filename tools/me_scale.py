import sys, os, json
sys.path.insert(0, os.getcwd())
import bench, torch
dev = torch.device("cuda", 0)
for s in (32, 64, 128, 256):
    o = bench.bench_me_frame(torch, dev, 1, streams_n=s)
    print(s, round(o["ms_per_launch"], 3), round(o["frames_per_s"]), o.get("identical_to_reference"))
