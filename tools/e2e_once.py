"""One interleaved e2e sample (ours vs reference) in a plain process: diagnostic."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theora_b200 import workload as wl  # noqa: E402

blob = wl.synth_stream(1920, 1080, 300, 32, 64)
open("/tmp/s.ogs", "wb").write(blob)
threads = str(len(os.sched_getaffinity(0)))
p = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "dec_e2e_bench.py"), "/tmp/s.ogs",
                    threads, "1"], capture_output=True, text=True)
d = json.loads(p.stdout.strip().splitlines()[-1])
print("ours %.0f fps, reference %.0f fps, ratio %.2f, flush %.3f ms, same output %s" % (
    d["frames"] / d["secs"], d["frames"] / d["ref_secs"], d["ref_secs"] / d["secs"], d["flush_ms_per_frame"],
    d["hash"] == d["ref_hash"]))
