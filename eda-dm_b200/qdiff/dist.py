"""Data-parallel plumbing of the reconstruction loop: one process per GPU, calibration
rows sharded across ranks, ONE all-reduce per reconstruction step over a flat fp32 bucket that holds every
AdaRound alpha gradient and every activation step-size gradient of the unit.  torch.distributed (NCCL over
NVLink on the box, gloo in the CPU tests) carries the collective; with no process group everything is a no-op.
"""
import torch
import torch.distributed as dist


def is_active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size() -> int:
    return dist.get_world_size() if is_active() else 1


def rank() -> int:
    return dist.get_rank() if is_active() else 0


def barrier():
    if is_active():
        dist.barrier()


def shard_rows(n: int, rank_: int = None, world: int = None):
    """Contiguous row range [lo, hi) of rank `rank_`: rank r owns rows [r*n/W, (r+1)*n/W)."""
    r = rank() if rank_ is None else rank_
    w = world_size() if world is None else world
    return (r * n) // w, ((r + 1) * n) // w


def shard_calibration(cali_data):
    """Each rank keeps its own contiguous slice of every calibration tensor (TDAC already shuffles timesteps,
    scripts/calibration.py:105-106 of the reference, so slices are i.i.d. over timesteps)."""
    if not is_active():
        return cali_data
    lo, hi = shard_rows(cali_data[0].size(0))
    return [t[lo:hi] for t in cali_data]


class GradBucket:
    """Flat fp32 gradient bucket: `.grad` of every parameter is a view into one contiguous buffer, so the
    per-step exchange is a single all-reduce with no packing copies."""

    def __init__(self, params):
        self.params = [p for p in params]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device('cpu')
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        if is_active():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world_size())

    def nbytes(self):
        return self.flat.numel() * 4
