"""Scale-init drivers for the class-conditional LDM (interface of the reference's
qdiff_control/set_quantize_params_Conditional.py:11-139).  `module` is the reference's LatentDiffusion wrapper; batches go
through its DDIMSampler_control in the single-step `quant_unet=True` mode (needs the reference's `ldm` package)."""
import logging

import torch

from qdiff.set_quantize_params import _act_quantizers, _weight_quantizers

logger = logging.getLogger(__name__)


def _sample(module, args, batch_size, cali_slice):
    from ldm.models.diffusion.ddim_control import DDIMSampler_control  # reference L0, reused as-is
    uc = None
    if args.scale != 1.0:
        uc = module.get_learned_conditioning({module.cond_stage_key: torch.tensor(batch_size * [1000]).to(module.device)})
    c = module.get_learned_conditioning({module.cond_stage_key: args.data[:batch_size].to(module.device)})
    sampler = DDIMSampler_control(module)
    sampler.sample(S=args.custom_steps, conditioning=c, batch_size=batch_size, shape=[3, 64, 64], verbose=False,
                   unconditional_guidance_scale=args.scale, unconditional_conditioning=uc, eta=args.ddim_eta,
                   quant_unet=True, cali_data=cali_slice)


def set_act_quantize_params_Conditional(module, cali_data, args, batch_size: int = 32):
    logger.info("set_act_quantize_params")
    unet = module.model.diffusion_model
    unet.set_quant_state(True, True)
    for q in _act_quantizers(unet, ldm_matmuls=True, transformer=True):
        q.set_inited(False)
    batch_size = min(batch_size, cali_data[0].size(0))
    with torch.no_grad():
        for i in range(int(cali_data[0].size(0) / batch_size)):
            _sample(module, args, batch_size, [_[i * batch_size:(i + 1) * batch_size].cuda() for _ in cali_data])
    for q in _act_quantizers(module, ldm_matmuls=True, transformer=True):
        q.set_inited(True)


def set_weight_quantize_params_Conditional(model, cali_data, args):
    logger.info("set_weight_quantize_params")
    unet = model.model.diffusion_model
    unet.set_quant_state(True, False)
    for q in _weight_quantizers(unet, with_split_twin=False):
        q.set_inited(False)
    batch_size = 2
    with torch.no_grad():
        _sample(model, args, batch_size, [_[:batch_size].cuda() for _ in cali_data])
    for q in _weight_quantizers(model, with_split_twin=True):
        q.set_inited(True)
