#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_api.py -x -q -m gpu --tb=short 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --recon-iters 40 > gpurun_out/bench_wg.json 2> gpurun_out/bench_wg.err; tail -2 gpurun_out/bench_wg.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_wg.json').read().strip().splitlines() if l.startswith('{')][-1])
for name, r in (('imagenet', d['recon']), ('church', d['secondary']['recon'])):
    for k in r:
        if isinstance(r[k],dict): print(name, k, round(r[k]['geomean_iters_per_s'],1), {u:round(v['iters_per_s'],1) for u,v in r[k]['units'].items()})
PY
