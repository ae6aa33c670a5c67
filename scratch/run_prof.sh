timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "qgemm" 2>&1 | tail -2
ONLY=imagenet timeout 300 python scratch/bench_gemm.py 2>&1 | cut -c1-110
timeout 900 python bench.py --workload imagenet --steps 5 --no-cpu-baseline --no-recon 2>&1 | tail -1 | cut -c1-330
