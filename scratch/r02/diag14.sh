#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_api.py -x -q -m gpu --tb=short -k "bf16x3 or recon or traces or checkpointed or walk or attn_block" 2>&1 | tail -6
python - <<'PY'
import sys, torch
sys.path[:0]=['.', 'eda-dm_b200']
import bench
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
from qdiff.quant_layer import backend
dev=torch.device('cuda:0')
kind,batch,shape,ctx,_=bench.WORKLOADS['imagenet']
fp = bench.build_fp_unet(kind).to(dev)
qnn = QuantModel(fp, bench.WQ, bench.AQ, sm_abit=8).to(dev).eval()
qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization(); bench.set_split(qnn.model, kind)
cali=[c.to(dev) for c in bench.synth_inputs(shape,ctx,64,seed=1234)]
set_weight_quantize_params(qnn, cali); set_act_quantize_params(qnn, cali, batch_size=32, all_attention=True)
for flag in (True, False):
    backend.calib_gemm_bf16x3 = flag
    r = bench.bench_recon(qnn, kind, shape, ctx, dev, 1, 30, ["weak"])
    print('calib_gemm_bf16x3', flag, {k:(round(v['iters_per_s'],1), round(v['ms_per_iter'],2)) for k,v in r['weak']['units'].items()})
PY
