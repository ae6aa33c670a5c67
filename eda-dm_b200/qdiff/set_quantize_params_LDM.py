"""LDM variants of the scale-init drivers (interface of qdiff/set_quantize_params_LDM.py:11-103).  The model
argument is the reference's LatentDiffusion wrapper; calibration batches go through its DDIMSampler in the
single-step `quant_unet=True` mode, so this module needs the reference's `ldm` package on sys.path."""
import logging

import torch

from .set_quantize_params import _act_quantizers, _weight_quantizers

logger = logging.getLogger(__name__)


def _sampler(module):
    from ldm.models.diffusion.ddim import DDIMSampler  # reference L0 (samplers are out of scope, reused as-is)
    return DDIMSampler(module)


def set_act_quantize_params_LDM(module, cali_data, args, batch_size: int = 32):
    logger.info("set_act_quantize_params")
    unet = module.model.diffusion_model
    unet.set_quant_state(True, True)
    for q in _act_quantizers(unet, ldm_matmuls=True):
        q.set_inited(False)
    batch_size = min(batch_size, cali_data[0].size(0))
    shape = [unet.in_channels, unet.image_size, unet.image_size]
    ddim = _sampler(module)
    with torch.no_grad():
        for i in range(int(cali_data[0].size(0) / batch_size)):
            ddim.sample(args.custom_steps, batch_size=batch_size, shape=shape, eta=args.eta, verbose=False, quant_unet=True,
                        cali_data=[_[i * batch_size:(i + 1) * batch_size].cuda() for _ in cali_data])
    for q in _act_quantizers(module, ldm_matmuls=True):
        q.set_inited(True)


def set_weight_quantize_params_LDM(model, cali_data, args):
    logger.info("set_weight_quantize_params")
    unet = model.model.diffusion_model
    unet.set_quant_state(True, False)
    for q in _weight_quantizers(unet, with_split_twin=False):
        q.set_inited(False)
    batch_size = 8
    shape = [unet.in_channels, unet.image_size, unet.image_size]
    ddim = _sampler(model)
    with torch.no_grad():
        ddim.sample(args.custom_steps, batch_size=batch_size, shape=shape, eta=args.eta, verbose=False, quant_unet=True,
                    cali_data=[_[:batch_size].cuda() for _ in cali_data])
    for q in _weight_quantizers(model, with_split_twin=True):
        q.set_inited(True)
