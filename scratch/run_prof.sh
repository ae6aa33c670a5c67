timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -8
timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-recon 2>&1 | tail -1 | cut -c1-330
timeout 900 python bench.py --workload cifar --steps 10 --no-cpu-baseline --no-recon 2>&1 | tail -1 | cut -c1-330
