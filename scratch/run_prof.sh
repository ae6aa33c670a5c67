timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "qgemm" 2>&1 | tail -3
timeout 300 python scratch/bench_gemm.py
