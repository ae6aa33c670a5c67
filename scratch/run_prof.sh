timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -15
NOLIB=1 ONLY=imagenet timeout 300 python scratch/bench_gemm.py 2>&1 | grep TOP
timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-recon 2>&1 | tail -1 | cut -c1-330
timeout 900 python bench.py --workload imagenet --steps 5 --no-cpu-baseline --no-recon 2>&1 | tail -1 | cut -c1-330
