mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -k "qattn or unit_vectors" 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-recon 2>&1 | tail -1 | cut -c1-420
EDADM_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --no-graph --no-recon --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_r01b.csv
