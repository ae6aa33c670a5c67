import sys, os, torch
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
from qdiff.quant_layer import backend, QuantModule
from unet_zoo.ldm_unet import UNetModel
cuda=torch.device('cuda:0')
def build(split):
    torch.manual_seed(29)
    model = UNetModel(image_size=16, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=2, attention_resolutions=(), channel_mult=(1, 2, 3), num_heads=1).to(cuda).eval()
    for p in model.parameters():
        if p.dim() > 1 and float(p.detach().abs().max()) == 0: torch.nn.init.normal_(p, std=0.02)
    model.split_shortcut = split
    wq = dict(n_bits=4, symmetric=True, channel_wise=True, scale_method='mse')
    aq = dict(n_bits=8, symmetric=True, channel_wise=False, scale_method='mse', leaf_param=True, prob=1.0)
    qnn = QuantModel(model, wq, aq, sm_abit=8).to(cuda).eval()
    qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization()
    return qnn
g = torch.Generator().manual_seed(5)
x = torch.randn(8, 4, 16, 16, generator=g).to(cuda); t = torch.randint(0, 1000, (8,), generator=g).to(cuda)
for split in (False, True):
    qnn = build(split)
    with torch.no_grad():
        set_weight_quantize_params(qnn, (x, t)); set_act_quantize_params(qnn, (x, t)); qnn.set_quant_state(True, True)
        for knob in (None, 'fused_attention', 'fuse_norm', 'fuse_epilogue', 'lazy_cat', 'integer_path'):
            if knob: setattr(backend, knob, False)
            ys = [qnn(x, t).clone() for _ in range(4)]
            if knob: setattr(backend, knob, True)
            print('split', split, 'knob off:', knob, 'run-to-run', [(ys[0]-y).abs().max().item() for y in ys[1:]])
        mb = qnn.model.middle_block[1]
        print(type(mb).__name__, [(n, type(m).__name__, getattr(m, 'inited', None)) for n, m in mb.named_modules() if 'quantizer' in n][:12])
