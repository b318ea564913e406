"""Data-parallel plumbing of the hot path.

The path shards over images only: inference needs no collective; training has ONE exchange step,
the average of the trainable gradients (LoRA A/B, projection, decoder: 6.6-32 M elements) plus the
criterion's 1-float `num_boxes` sum (reference losses.py:228-229).  The model also works unchanged
under stock `DistributedDataParallel(find_unused_parameters=True)` (reference train.py:677); this
module is the lean alternative: one flat fp32 buffer that the gradients are accumulated into in
place and ONE all-reduce over NCCL / NVLink per step instead of 25 MB buckets with per-bucket hooks.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [start, end) of n_items for `rank`; the first n % world ranks get one more."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class FlatGradSync:
    """Keeps `p.grad` of every trainable parameter as a view into one flat fp32 buffer.

    usage per step:   sync.zero();  loss.backward();  sync.all_reduce();  optimizer.step()
    (call `optimizer.zero_grad(set_to_none=False)` or `sync.zero()`, never set_to_none=True, or the
    views are dropped; `attach()` re-installs them).
    """

    def __init__(self, params, process_group=None):
        # parameters flagged `_dod_unused` are never an input of the train node (decoder.reference_points,
        # reference detr_decoder.py:44-45): like under torch autograd their .grad stays None
        self.params = [p for p in params if p.requires_grad and not getattr(p, "_dod_unused", False)]
        self.group = process_group
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.attach()

    def attach(self):
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, average=True):
        """Sum (and average, like DDP) the flat gradient over the process group."""
        if not (dist.is_available() and dist.is_initialized()):
            return self.flat
        for p in self.params:                      # a set_to_none=True zero_grad dropped the views
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
                    p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * 4:
                raise RuntimeError("FlatGradSync: p.grad is no longer a view of the flat buffer; call attach()")
        if average and dist.get_backend(self.group) == "nccl":
            # NCCL averages inside the collective (no separate pass over the buffer afterwards)
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            return self.flat
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if average:
            self.flat.div_(dist.get_world_size(self.group))
        return self.flat


def all_reduce_num_boxes(num_boxes: torch.Tensor, group=None):
    """reference losses.py:228-230: SUM over ranks, NOT divided by the world size, clamped to >= 1."""
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(num_boxes, op=dist.ReduceOp.SUM, group=group)
    return torch.clamp(num_boxes, min=1)
