timeout 300 python scratch/bench_actq.py
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -2
