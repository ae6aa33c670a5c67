#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --tb=short -k "conv3x3_small or head" 2>&1 | tail -3
python - <<'PY'
import sys, torch
sys.path[:0]=['.', 'eda-dm_b200']
from edadm import ops
dev=torch.device('cuda:0')
for shape in ((128,192,64,64),(100,128,32,32)):
    x=torch.randn(*shape,device=dev); w=torch.randn(3,shape[1],3,3,device=dev)*0.05; b=torch.randn(3,device=dev)
    a_,s_=ops.gn_fold(x,torch.randn(shape[1],device=dev),torch.randn(shape[1],device=dev),32,1e-5)
    for aff in (None,(a_,s_,1),(a_,s_,17)):
        f=lambda: ops.conv3x3_small_n(x,w,b,affine=aff)
        for _ in range(3): f()
        torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize(); us=e0.elapsed_time(e1)*100
        print(shape, 'affine', None if aff is None else aff[2], '%.1f us' % us, '%.0f GB/s' % (x.numel()*4/us/1e3))
PY
