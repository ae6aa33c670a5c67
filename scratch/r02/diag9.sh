#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -m gpu -k "search or checkpointed or traces" --tb=short 2>&1 | tail -8
python - <<'PY'
import sys, time, torch
sys.path[:0]=['.', 'eda-dm_b200']
import bench
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
from qdiff.quant_layer import backend
dev=torch.device('cuda:0')
for use in (True, False):
    backend.search_kernel = use
    fp = bench.build_fp_unet('imagenet').to(dev)
    qnn = QuantModel(fp, bench.WQ, bench.AQ, sm_abit=8).to(dev).eval()
    qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization(); bench.set_split(qnn.model, 'imagenet')
    cali=[c.to(dev) for c in bench.synth_inputs((3,64,64),(1,512),64,seed=1234)]
    torch.cuda.synchronize(); t0=time.perf_counter(); set_weight_quantize_params(qnn, cali); torch.cuda.synchronize(); tw=time.perf_counter()-t0
    t0=time.perf_counter(); set_act_quantize_params(qnn, cali, batch_size=32, all_attention=True); torch.cuda.synchronize(); ta=time.perf_counter()-t0
    print('search_kernel', use, 'set_weight %.2f s set_act %.2f s' % (tw, ta))
    del qnn, fp
PY
