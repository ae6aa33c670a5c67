timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 2>&1 | tail -1 > gpurun_out/bench_church.json; cut -c1-200 gpurun_out/bench_church.json
python -c "
import json; d=json.load(open('gpurun_out/bench_church.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step']); print('roofline',d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['share_of_eager_step']); print('recon',d['recon']); print('cpu',d['cpu_baseline'])"
