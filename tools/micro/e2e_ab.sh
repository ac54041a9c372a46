python - <<PY
import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import th_workload as wl
blob = wl.synth_stream(1920, 1080, 300, 32, 64)
open("/tmp/blob.ogs", "wb").write(blob)
PY
for dma in 0 1 0 1; do
  OCG_OUT_DMA=$dma python tools/dec_e2e_bench.py /tmp/blob.ogs 48 0 1 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dma=$dma', round(d['frames']/d['secs']), 'fps', round(d['d2h_bytes']/d['secs']/1e9,1), 'GB/s flush_ms', round(d['flush_ms_per_frame'],3), 'wait_ms', round(d['wait_ms_per_frame'],2), d['hash'])"
done
OCG_OUT_DMA=1 python tools/dec_e2e_bench.py /tmp/blob.ogs 32 0 1 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dma=1 32thr', round(d['frames']/d['secs']), 'fps')"
OCG_OUT_DMA=1 python tools/dec_e2e_bench.py /tmp/blob.ogs 64 0 1 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dma=1 64thr', round(d['frames']/d['secs']), 'fps')"
