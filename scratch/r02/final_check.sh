#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 2400 python -m pytest tests -x -q -m gpu --tb=short 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/bench_final_b.json 2> gpurun_out/bench_final_b.err; tail -2 gpurun_out/bench_final_b.err
python - <<'PY'
import json
def last(f): return json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
d=last('gpurun_out/bench_final_b.json')
print('imagenet', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('frac_of_burst_peak'), d['gpu_launches'], d['clocks'], 'wall', d.get('wall_s'))
r=d['recon']
for k in r:
    if isinstance(r[k],dict): print(k, round(r[k]['geomean_iters_per_s'],1), {u:round(v['iters_per_s'],1) for u,v in r[k]['units'].items()})
s=d['secondary']; print('church', s['ms_per_step'], s['value'], s['e2e']['value'], s['roofline']['frac'])
PY
