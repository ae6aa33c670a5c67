"""The two Adam optimisers of a reconstruction unit as one kernel launch (edadm_fused_adam, csrc/optim.cu).

The reference builds `torch.optim.Adam(w_para, lr=lr_w)` and `torch.optim.Adam(a_para, lr=lr_a)` per unit
(qdiff/block_recon.py:113-117) and steps both every iteration (:199-206).  Here the gradients of both parameter
groups already live in ONE flat bucket (`qdiff.dist.GradBucket`: the buffer the data-parallel all-reduce works on), so
the update is a single streaming pass over {gradient, parameter, exp_avg, exp_avg_sq}: same arithmetic as
torch's `_single_tensor_adam` (betas 0.9 / 0.999, eps 1e-8, no weight decay, no amsgrad -- the defaults the
reference uses), learning rates and the step count read from device memory so a captured CUDA graph sees the
cosine schedule.  The consumed gradients are cleared by the same pass.
"""
import numpy as np
import torch

from edadm import ops

SEGMENT = 8192      # elements per thread block (csrc/optim.cu: 256 threads x 8 float4)


def segment_table(params, n_group0: int, segment: int = SEGMENT):
    """int64 [n_segments, 3] host table in the layout of csrc/optim.cu AdamSegment: {param address of the segment's first
    element, its offset in the flat buffers, count | group << 32}.  `params` in GradBucket order; the first `n_group0`
    tensors belong to group 0 (lr[0]), the rest to group 1."""
    rows, off = [], 0
    for i, p in enumerate(params):
        assert p.dtype == torch.float32 and p.is_contiguous(), "Adam parameters are contiguous fp32 tensors"
        n, base, grp = p.numel(), p.data_ptr(), (0 if i < n_group0 else 1)
        for s in range(0, n, segment):
            cnt = min(segment, n - s)
            rows.append((base + 4 * s, off + s, cnt | (grp << 32)))
        off += n
    return np.asarray(rows, dtype=np.int64).reshape(-1, 3), off


class FusedAdam:
    """Adam over the parameters of a `GradBucket` (group 0: the first `n_group0` tensors at lr[0]; group 1: the rest at
    lr[1]).  `lr` is a 2-element fp32 device tensor the caller updates (or a captured graph reads); `.step()` is graph-safe."""

    def __init__(self, bucket, n_group0: int, lr: torch.Tensor, betas=(0.9, 0.999), eps: float = 1e-8, zero_grad: bool = True):
        flat = bucket.flat
        assert flat.is_cuda and lr.is_cuda and lr.dtype == torch.float32 and lr.numel() == 2
        table, total = segment_table(bucket.params, n_group0)
        assert total == flat.numel()
        self.bucket, self.params, self.lr = bucket, list(bucket.params), lr
        self.segments = torch.from_numpy(table).to(flat.device)
        self.n_segments = int(table.shape[0])
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=flat.device)
        self.betas, self.eps, self.zero_grad = (float(betas[0]), float(betas[1])), float(eps), bool(zero_grad)

    def step(self):
        self.step_count.add_(1)
        ops.fused_adam(self.segments, self.n_segments, self.bucket.flat, self.exp_avg, self.exp_avg_sq, self.lr,
                       self.step_count, self.betas[0], self.betas[1], self.eps, self.zero_grad)

    def finish(self):
        """the kernel writes through raw pointers: tell autograd's version counters (packed-weight and calibration-cache keys
        are built from them) that every parameter changed"""
        if self.params:
            torch.autograd.graph.increment_version(self.params)
