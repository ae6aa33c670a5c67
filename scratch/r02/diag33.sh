#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -m gpu --tb=short 2>&1 | tail -4
for mb in 100000 16 32 48 64; do
EDADM_GN_CHUNK_MB=$mb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-recon --no-secondary > gpurun_out/bench_chunk.json 2> gpurun_out/bench_chunk.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_chunk.json').read().strip().splitlines() if l.startswith('{')][-1])
print('chunk MB $mb', 'imagenet', round(d['ms_per_step'],2), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
PY
done
