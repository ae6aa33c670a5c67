"""CFG cache builder (interface of the reference's qdiff_control/data_utils.py:7-76): every calibration slice
(x, t, index, cond, uncond) is fed as [x;x], [t;t], [uncond;cond], so the cache holds 2N rows."""
import torch

from qdiff.data_utils import (save_inp_oup_data as _save, GetLayerInpOut, DataSaverHook,  # noqa: F401
                              StopForwardException)


def cfg_batch(batch):
    """reference qdiff_control/data_utils.py:28-31"""
    return [torch.cat([batch[0]] * 2), torch.cat([batch[1]] * 2), torch.cat([batch[4], batch[3]])]


def save_inp_oup_data(model, layer, cali_data, asym: bool = False, act_quant: bool = False, batch_size: int = 32,
                      input_prob: bool = False, keep_gpu: bool = True):
    return _save(model, layer, cali_data, asym, act_quant, batch_size=batch_size, input_prob=input_prob, keep_gpu=keep_gpu,
                 batch_transform=cfg_batch)
