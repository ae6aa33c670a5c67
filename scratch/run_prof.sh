timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "layernorm or transformer or token" 2>&1 | tail -3
timeout 900 python bench.py --workload imagenet --steps 5 --no-cpu-baseline --no-recon 2>&1 | tail -1 | cut -c1-330
