#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_api.py -x -q -m gpu --tb=short 2>&1 | tail -4; export EDADM_LOG=1
python - <<'PY'
# whole-model reconstruction wall-clock with and without prefix reuse (church LDM-8, 64 calibration samples, 2 iterations per unit)
import sys, time, torch, random
sys.path[:0]=['.', 'eda-dm_b200']
import bench, logging; logging.basicConfig(level=logging.INFO)
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params, recon_block_Qmodel, Change_LDM_model_attnblock
from qdiff.quant_layer import backend
dev=torch.device('cuda:0')
for reuse in (True, False):
    backend.cache_prefix_reuse = reuse
    fp = bench.build_fp_unet('church').to(dev)
    qnn = QuantModel(fp, bench.WQ, bench.AQ, sm_abit=8).to(dev).eval()
    qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization(); bench.set_split(qnn.model, 'church')
    cali=[c.to(dev) for c in bench.synth_inputs((4,32,32),None,64,seed=1234)]
    set_weight_quantize_params(qnn, cali); set_act_quantize_params(qnn, cali, batch_size=32, all_attention=True)
    kw=dict(cali_data=cali, iters=2, batch_size=32, weight=0.01, asym=True, b_range=(20,2), warmup=0.2, act_quant=True, opt_mode='mse', lr_a=4e-4, lr_w=1e-2, p=2.0, input_prob=0.5, keep_gpu=True, recon_w=True, recon_a=True, add_loss=0.8)
    random.seed(1); torch.manual_seed(1)
    torch.cuda.synchronize(); t0=time.perf_counter()
    recon_block_Qmodel(None, qnn, cali, kw).recon()
    torch.cuda.synchronize(); print('church whole-model walk (2 iters/unit), prefix reuse', reuse, '%.1f s' % (time.perf_counter()-t0), 'block_count', qnn.block_count)
    del qnn, fp
PY
