#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for wl in cifar bedroom; do
timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --recon-iters 20 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -2 gpurun_out/bench_$wl.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_$wl.json').read().strip().splitlines() if l.startswith('{')][-1])
print('$wl', round(d['ms_per_step'],2), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), d['config'].get('on_int8_tcgen05_path'), d['config'].get('quant_modules'))
r=d['recon']
for k in r:
    if isinstance(r[k],dict): print('  ', k, round(r[k]['geomean_iters_per_s'],1), {u:round(v['iters_per_s'],1) for u,v in r[k]['units'].items()})
PY
done
