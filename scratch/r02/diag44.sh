#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1500 python bench.py > gpurun_out/bench_final_a.json 2> gpurun_out/bench_final_a.err; tail -2 gpurun_out/bench_final_a.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -2 gpurun_out/bench_final_ref.err
timeout 600 python bench.py --workload sd --no-secondary --no-cpu-baseline --no-recon > gpurun_out/bench_final_sd.json 2> gpurun_out/bench_final_sd.err; tail -2 gpurun_out/bench_final_sd.err
python - <<'PY'
import json
def last(f): return json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
d=last('gpurun_out/bench_final_a.json')
print('imagenet', d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d['gpu_launches'], d['clocks'], 'wall', d.get('wall_s'))
print('calib', d['config'].get('calibration_s'))
r=d['recon']
for k in r:
    if isinstance(r[k],dict): print(k, round(r[k]['geomean_iters_per_s'],1), {u:round(v['iters_per_s'],1) for u,v in r[k]['units'].items()})
s=d['secondary']; print('church', s['ms_per_step'], s['value'], s['e2e']['value'], s['roofline']['frac'])
rs=s['recon']
for k in rs:
    if isinstance(rs[k],dict): print(' church', k, round(rs[k]['geomean_iters_per_s'],1), {u:round(v['iters_per_s'],1) for u,v in rs[k]['units'].items()})
print('cpu', d['cpu_baseline'])
print('ref', last('gpurun_out/bench_final_ref.json'))
x=last('gpurun_out/bench_final_sd.json'); print('sd', x['ms_per_step'], x['value'], x['config'].get('on_int8_tcgen05_path'), x['config'].get('quant_modules'))
PY
