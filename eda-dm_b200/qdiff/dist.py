"""Data-parallel plumbing of the reconstruction loop: one process per GPU, calibration
rows sharded across ranks, ONE all-reduce per reconstruction step over a flat fp32 bucket that holds every
AdaRound alpha gradient and every activation step-size gradient of the unit.  torch.distributed (NCCL over
NVLink on the box, gloo in the CPU tests) carries the collective; with no process group everything is a no-op.
"""
import contextlib

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

        self._works, self._done, self._hooks = [], set(), []

    def zero(self, memset: bool = True):
        """start of an iteration; memset=False when the previous optimiser pass already cleared the buffer (FusedAdam)"""
        if memset:
            self.flat.zero_()
        self._works, self._done = [], set()

    def overlap_backward(self, min_numel: int = 1 << 16, stream=None):
        """Start the exchange of every large gradient (an AdaRound alpha) the moment autograd has finished accumulating it,
        instead of one all-reduce after the whole backward: the reduction of the last layers' gradients then runs on NCCL's
        stream underneath the backward of the earlier layers.  Every rank builds the same autograd graph, so the collectives are
        issued in the same order everywhere.  Small gradients (step sizes) and whatever was not reduced early go out in one final
        call from `all_reduce_mean`.  Works inside CUDA-graph capture (the asynchronous collectives become a parallel branch)."""
        if not is_active() or self._hooks:
            return
        for i, p in enumerate(self.params):
            if p.numel() < min_numel:
                continue

            def hook(param, i=i):
                if i not in self._done and param.grad is not None and param.grad.data_ptr() == self._view_ptr[i]:
                    self._done.add(i)
                    # issued against the loop's compute stream (hooks run on an autograd worker thread whose current stream
                    # need not be the one the accumulation was enqueued on): NCCL's stream waits for everything queued so far
                    ctx = torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()
                    with ctx:
                        self._works.append(dist.all_reduce(param.grad, op=dist.ReduceOp.SUM, async_op=True))
            self._hooks.append(p.register_post_accumulate_grad_hook(hook))
        self._view_ptr = [p.grad.data_ptr() for p in self.params]

    def release(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def all_reduce_mean(self):
        if not is_active():
            return
        if not self._done:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        else:
            # contiguous runs of parameters that were not reduced during the backward
            off, run = 0, None
            for i, p in enumerate(self.params + [None]):
                pending = p is not None and i not in self._done
                if pending and run is None:
                    run = off
                if not pending and run is not None:
                    self._works.append(dist.all_reduce(self.flat[run:off], op=dist.ReduceOp.SUM, async_op=True))
                    run = None
                if p is not None:
                    off += p.numel()
            for w in self._works:
                w.wait()            # the compute stream waits for NCCL's stream (no host synchronisation)
            self._works, self._done = [], set()
        self.flat.div_(world_size())

    def nbytes(self):
        return self.flat.numel() * 4
