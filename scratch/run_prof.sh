timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-recon 2>&1 | tail -1 | cut -c1-330
