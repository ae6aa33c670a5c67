#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
timeout 1200 python bench.py --steps 10 --recon-iters 30 > gpurun_out/r02/bench_default.json 2> gpurun_out/r02/bench_default.err; tail -3 gpurun_out/r02/bench_default.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_default.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'wall', d['wall_s'])
print('calib', d['config']['calibration_s'])
for mode in ('weak','strong'):
    r=d['recon'][mode]; print(mode, 'geomean', r['geomean_iters_per_s'], {k:(round(v['iters_per_s'],1), v['cuda_graph']) for k,v in r['units'].items()})
print('cpu', d['cpu_baseline'])
s=d['secondary']; print('church', s['ms_per_step'], s['value'], s['roofline']['frac'], {k:(round(v['iters_per_s'],1), v['cuda_graph']) for k,v in s['recon']['weak']['units'].items()})
PY
timeout 600 python bench.py --workload sd --steps 5 --no-recon --no-cpu-baseline > gpurun_out/r02/bench_sd.json 2> gpurun_out/r02/bench_sd.err; tail -3 gpurun_out/r02/bench_sd.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_sd.json').read().strip().splitlines()[-1])
print('sd ms', d['ms_per_step'], 'value', d['value'], 'frac', d['roofline']['frac'], d['config']['on_int8_tcgen05_path'], d['config']['quant_modules'], d['config']['calibration_s'])
PY
