"""block_reconstruction, CFG variant (signature of the reference's qdiff_control/block_recon.py:13-18): the q/k/v/softmax
step sizes of QuantBasicTransformerBlock become trainable (:82-107) and the cache is built with `batch_size` (:139)."""
import torch

from qdiff.quant_model import QuantModel
from qdiff.quant_block import BaseQuantBlock
from qdiff._recon_engine import reconstruct, LossFunction, LinearTempDecay  # noqa: F401
from .data_utils import save_inp_oup_data


def block_reconstruction(model: QuantModel, block: BaseQuantBlock, cali_data: torch.Tensor,
                         batch_size: int = 32, iters: int = 20000, weight: float = 0.01, opt_mode: str = 'mse',
                         asym: bool = False, b_range: tuple = (20, 2),
                         warmup: float = 0.0, act_quant: bool = False, lr_a: float = 4e-5, lr_w=1e-2, p: float = 2.0,
                         input_prob: float = 1.0, keep_gpu: bool = True,
                         recon_w: bool = False, recon_a: bool = False, add_loss: float = 0.0, **extra):
    return reconstruct(model, block, cali_data, batch_size=batch_size, iters=iters, weight=weight, opt_mode=opt_mode,
                       asym=asym, b_range=b_range, warmup=warmup, act_quant=act_quant, lr_a=lr_a, lr_w=lr_w, p=p,
                       input_prob=input_prob, keep_gpu=keep_gpu, recon_w=recon_w, recon_a=recon_a, add_loss=add_loss,
                       cache_builder=save_inp_oup_data, cache_batch_size=batch_size, transformer=True, **extra)
