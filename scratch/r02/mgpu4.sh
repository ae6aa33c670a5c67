#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
N=4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --recon-iters 40 --no-cpu-baseline > gpurun_out/r02/bench_g$N.json 2> gpurun_out/r02/bench_g$N.err
tail -3 gpurun_out/r02/bench_g$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02/bench_g$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'wall', d['wall_s'])
for mode in ('weak','strong'):
    r=d['recon'][mode]; print(mode, 'geomean it/s', round(r['geomean_iters_per_s'],1), 'samples/s', round(r['geomean_samples_per_s'],1), {k:(round(v['iters_per_s'],1), round(v['ms_per_iter'],3), v['allreduce_bytes_per_iter']) for k,v in r['units'].items()})
s=d['secondary']; print('church', s['ms_per_step'], s['value'], {k:(round(v['iters_per_s'],1), round(v['ms_per_iter'],3)) for k,v in s['recon']['weak']['units'].items()})
PY
