timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python bench.py --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('church', d['recon'])"
timeout 900 python bench.py --workload imagenet --steps 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('imagenet', d['recon'])"
