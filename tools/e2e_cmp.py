import sys, os, json, subprocess
sys.path.insert(0, os.getcwd())
from theora_b200 import workload as wl
blob = wl.synth_stream(1920, 1080, 300, 32, 64)
open("/tmp/s.ogs", "wb").write(blob)
for i in range(3):
    for dc in (1, 0):
        p = subprocess.run([sys.executable, "tools/dec_e2e_bench.py", "/tmp/s.ogs", "16", "1", str(dc), "0", "0"], capture_output=True, text=True)
        d = json.loads(p.stdout.strip().splitlines()[-1])
        ex = dc
        print("dc_mode=%d (1 host, 0 device-ahead): ours %.0f fps, ref %.0f fps, ratio %.2f, flush %.3f ms, same output %s" % (ex, d["frames"] / d["secs"], d["frames"] / d["ref_secs"], d["ref_secs"] / d["secs"], d["flush_ms_per_frame"], d["hash"] == d["ref_hash"]))
