#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --tb=short -k "codes_epilogue" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-recon > gpurun_out/bench_p13.json 2> gpurun_out/bench_p13.err; tail -2 gpurun_out/bench_p13.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_p13.json').read().strip().splitlines() if l.startswith('{')][-1])
print('imagenet', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])
s=d['secondary']; print('church', s['ms_per_step'], s['value'], s['e2e']['value'], s['roofline']['frac'])
PY
