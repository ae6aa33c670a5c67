timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q 2>&1 | tail -4
NOLIB=1 timeout 300 python scratch/bench_gemm.py 2>&1 | grep TOP
timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-recon 2>&1 | tail -1 | cut -c1-330
