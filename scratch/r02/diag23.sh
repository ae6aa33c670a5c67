#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 ./scratch/r02/rcpcheck
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --tb=short 2>&1 | tail -5
python - <<'PY'
import sys, torch
sys.path[:0]=['.', 'eda-dm_b200']
from edadm import ops
dev=torch.device('cuda:0')
x=torch.randn(128,192,64,64,device=dev); d=torch.tensor([0.03],device=dev); z=torch.tensor([128.],device=dev); aq=ops.ActQuant(d,z,256)
a_,s_=ops.gn_fold(x,torch.randn(192,device=dev),torch.randn(192,device=dev),32,1e-5)
for mode in (1, 2, 17, 0):
    for _ in range(3): ops.norm_act_quant_nhwc(x,a_,s_,mode,aq,1)
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(10): ops.norm_act_quant_nhwc(x,a_,s_,mode,aq,1)
    e1.record(); torch.cuda.synchronize(); us=e0.elapsed_time(e1)*100
    print('producer silu mode', mode, '%.1f us' % us, '%.0f GB/s' % (x.numel()*5/us/1e3))
for _ in range(3): ops.act_quant_nhwc(x, aq, 1)
torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
for _ in range(10): ops.act_quant_nhwc(x, aq, 1)
e1.record(); torch.cuda.synchronize(); us=e0.elapsed_time(e1)*100
print('plain producer', '%.1f us' % us, '%.0f GB/s' % (x.numel()*5/us/1e3))
PY
