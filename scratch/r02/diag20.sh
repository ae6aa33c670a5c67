#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --tb=short -k "conv_bf16x3 or linear_bf16x3" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_api.py -x -q -m gpu --tb=short 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --recon-iters 40 > gpurun_out/bench_conv4.json 2> gpurun_out/bench_conv4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_conv4.json').read().strip().splitlines()[-1])
r=d['recon']
for k,v in r.items():
    if isinstance(v,dict):
        print(k, {u:(round(x['iters_per_s'],1), x['fp_taps_memoised']) for u,x in v['units'].items()}, round(v['geomean_iters_per_s'],1))
PY
