import sys, os, torch
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
from qdiff.quant_layer import backend, QuantModule
from qdiff.quant_block import QuantResBlock
from edadm import ops
from unet_zoo.ldm_unet import UNetModel
cuda=torch.device('cuda:0')
torch.manual_seed(29)
model = UNetModel(image_size=16, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=2, attention_resolutions=(), channel_mult=(1, 2, 3), num_heads=1).to(cuda).eval()
for p in model.parameters():
    if p.dim() > 1 and float(p.detach().abs().max()) == 0: torch.nn.init.normal_(p, std=0.02)
model.split_shortcut = True
wq = dict(n_bits=4, symmetric=True, channel_wise=True, scale_method='mse')
aq = dict(n_bits=8, symmetric=True, channel_wise=False, scale_method='mse', leaf_param=True, prob=1.0)
qnn = QuantModel(model, wq, aq, sm_abit=8).to(cuda).eval()
qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization()
g = torch.Generator().manual_seed(5)
x = torch.randn(8, 4, 16, 16, generator=g).to(cuda); t = torch.randint(0, 1000, (8,), generator=g).to(cuda)
with torch.no_grad():
    set_weight_quantize_params(qnn, (x, t)); set_act_quantize_params(qnn, (x, t)); qnn.set_quant_state(True, True)
    taps = {}
    def tap(mod, q, pad): taps.setdefault(id(mod), []).append(q.clone())
    names = {id(m): n for n, m in qnn.named_modules()}
    backend.code_tap = tap
    ya = qnn(x, t); taps.clear()
    yb = qnn(x, t); taps.clear()
    print('lazy run-to-run', (ya-yb).abs().max().item())
    y1 = qnn(x, t); t1 = dict(taps); taps.clear()
    backend.lazy_cat = False
    yc = qnn(x, t); taps.clear()
    y0 = qnn(x, t); t0 = dict(taps)
    print('eager-cat run-to-run', (yc-y0).abs().max().item())
    backend.code_tap = None
print('out diff', (y1-y0).abs().max().item(), (y1-y0).norm().item()/y0.norm().item())
for k in t0:
    for a, b in zip(t0[k], t1.get(k, [])):
        if a.shape != b.shape: print(names[k], 'shape', a.shape, b.shape); continue
        nd = (a != b).sum().item()
        if nd:
            idx = (a != b).nonzero()
            print(names[k], 'codes differ', nd, 'of', a.numel(), 'first', idx[0].tolist(), 'chan range', idx[:, -1].min().item(), idx[:, -1].max().item(), 'pix h', idx[:,1].min().item(), idx[:,1].max().item())
            break
