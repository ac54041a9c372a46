python - <<PY
import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import th_workload as wl
open("/tmp/blob.ogs", "wb").write(wl.synth_stream(1920, 1080, 300, 32, 64))
PY
for dma in 0 1 1 0; do
  OCG_OUT_DMA=$dma python tools/dec_e2e_bench.py /tmp/blob.ogs 48 0 1 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dma=$dma', round(d['frames']/d['secs']), 'fps', round(d['d2h_bytes']/d['secs']/1e9,1), 'GB/s', d['hash'])"
done
