#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "qattn or epilogue_fusions or transformer_block" --tb=short 2>&1 | tail -6
python bench.py --workload imagenet --no-recon --no-cpu-baseline --no-secondary --steps 10 > gpurun_out/r02/bench_imagenet_v4.json 2> gpurun_out/r02/bench_imagenet_v4.err; python -c "
import json; d=json.loads(open('gpurun_out/r02/bench_imagenet_v4.json').read().strip().splitlines()[-1]); print('imagenet', d['ms_per_step'], d['value'], d['roofline']['frac'])"
python bench.py --workload church --no-recon --no-cpu-baseline --steps 10 > gpurun_out/r02/bench_church_v4.json 2> gpurun_out/r02/bench_church_v4.err; python -c "
import json; d=json.loads(open('gpurun_out/r02/bench_church_v4.json').read().strip().splitlines()[-1]); print('church', d['ms_per_step'], d['value'], d['roofline']['frac'])"
