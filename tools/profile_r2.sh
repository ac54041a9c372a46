set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 640 --csv --log-file gpurun_out/r2_launches.csv python bench.py --frames 40 --steps 1 --no-e2e --no-cpu --no-encode-kernels --no-config4 --no-noisy > gpurun_out/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ocg_recon|ocg_lf2|ocg_border" -s 200 -c 4 -o gpurun_out/r2_decode -f python bench.py --frames 40 --steps 1 --warmup 1 --stream-groups 1 --no-e2e --no-cpu --no-encode-kernels --no-config4 --no-noisy > gpurun_out/r2_decode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ocg_recon" -s 24 -c 2 -o gpurun_out/r2_noisy -f python tools/kernel_tune.py --noisy --frames 8 > gpurun_out/r2_noisy.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ocg_enc_metrics_kernel|ocg_enc_fdct_quant|ocg_me_wavefront" -s 6 -c 10 -o gpurun_out/r2_encode -f python bench.py --frames 8 --streams 8 --steps 1 --warmup 1 --no-e2e --no-cpu --no-config4 --no-noisy > gpurun_out/r2_encode.log 2>&1
ls -la gpurun_out/*.ncu-rep
