"""qdiff_control -- classifier-free-guidance variants of the reconstruction drivers (ImageNet, Stable Diffusion); same
import surface as the reference's qdiff_control/__init__.py:1-4."""
from .set_quantize_params_Stable import set_weight_quantize_params_Stable, set_act_quantize_params_Stable
from .coco_prompt import get_prompts, center_resize_image
from .recon_block_Qmodel import recon_block_Qmodel
from .set_quantize_params_Conditional import set_weight_quantize_params_Conditional, set_act_quantize_params_Conditional
