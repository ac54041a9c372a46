python - <<PY
import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import th_workload as wl
blob = wl.synth_stream(1920, 1080, 60, 32, 64)
open("/tmp/blob60.ogs", "wb").write(blob)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/r2_e2e_launches.csv python tools/dec_e2e_bench.py /tmp/blob60.ogs 1 0 1 0 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2_e2e_launches.csv
