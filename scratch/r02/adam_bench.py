"""edadm_fused_adam at the size of an ImageNet transformer-block unit (17.5 M AdaRound alphas + 20 step sizes): CUDA-event timing
of the one-launch step against torch.optim.Adam (capturable, foreach) on the same tensors.  32 B per element algorithmic
(read g, p, m, v; write p, m, v, cleared g).  Run under ncu with `-k regex:fused_adam -c 1` for the DRAM bytes."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "eda-dm_b200")]
import torch
from qdiff.dist import GradBucket
from qdiff._fused_adam import FusedAdam

dev = torch.device("cuda:0")
shapes = [(384, 384)] * 8 + [(3072, 384), (384, 1536)] + [(384, 384, 3, 3)] * 11 + [()] * 20
params = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
n = sum(p.numel() for p in params)
bucket = GradBucket(params)
lrs = torch.tensor([1e-2, 4e-4], device=dev)
adam = FusedAdam(bucket, len(shapes) - 20, lrs)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    ts = []
    for _ in range(reps + 3):
        bucket.flat.normal_()
        flush.zero_()                                  # L2 flush: 256 MB > 126 MB
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[3:])
    return ts[len(ts) // 2]


us_fused = timed(adam.step, reps=int(os.environ.get('ADAM_REPS', 20)))
if os.environ.get('ADAM_ONLY'):
    print(us_fused); sys.exit(0)
opt_w = torch.optim.Adam(params[:-20], lr=lrs[0], capturable=True)
opt_a = torch.optim.Adam(params[-20:], lr=lrs[1], capturable=True)
us_torch = timed(lambda: (opt_w.step(), opt_a.step()))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
print(json.dumps({"elements": n, "fused_adam_us": us_fused, "fused_adam_GBps": 32.0 * n / us_fused * 1e-3,
                  "torch_optim_adam_x2_us": us_torch, "segments": adam.n_segments, "l2": "flushed between launches",
                  "peaks": peaks}))
