#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests/test_gpu_parity_full.py -q -m gpu -s --tb=short 2>&1 > gpurun_out/r02/pytest_parity_full3.txt
grep -E "fuse_norm=|injected|passed|failed|Error|worst rel|seed, flipped|assert" gpurun_out/r02/pytest_parity_full3.txt | cut -c1-500
python bench.py --workload imagenet --no-recon --no-cpu-baseline --steps 10 > gpurun_out/r02/bench_imagenet_v3.json 2> gpurun_out/r02/bench_imagenet_v3.err; python -c "
import json; d=json.loads(open('gpurun_out/r02/bench_imagenet_v3.json').read().strip().splitlines()[-1]); print('imagenet', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['roofline']['launches_per_step'])"
python bench.py --workload church --no-recon --no-cpu-baseline --steps 10 > gpurun_out/r02/bench_church_v3.json 2> gpurun_out/r02/bench_church_v3.err; python -c "
import json; d=json.loads(open('gpurun_out/r02/bench_church_v3.json').read().strip().splitlines()[-1]); print('church', d['ms_per_step'], d['value'], d['roofline']['frac'])"
