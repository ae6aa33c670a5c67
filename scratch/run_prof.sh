timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k recon 2>&1 | grep -E "Error|error|passed|failed" | head -8
timeout 900 python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -12 | cut -c1-400
