"""qdiff -- B200-native drop-in for EDA-DM's `qdiff` package (same import surface as the reference's
qdiff/__init__.py:1-7)."""
from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule
from .quant_model import QuantModel
from .set_quantize_params import set_weight_quantize_params, set_act_quantize_params
from .recon_block_Qmodel import recon_block_Qmodel, Change_LDM_model_attnblock
from .recon_layer_Qmodel import recon_layer_Qmodel
from .set_quantize_params_LDM import set_weight_quantize_params_LDM, set_act_quantize_params_LDM
