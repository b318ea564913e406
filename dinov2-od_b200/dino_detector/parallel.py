"""Data-parallel plumbing of the hot path.

The path shards over images only: inference needs no collective; training has ONE exchange step,
the average of the trainable gradients (LoRA A/B, projection, decoder: 6.6-32 M elements) plus the
criterion's 1-float `num_boxes` sum (reference losses.py:228-229).  The model also works unchanged
under stock `DistributedDataParallel(find_unused_parameters=True)` (reference train.py:677); this
module is the lean alternative: one flat fp32 buffer that the gradients are accumulated into in
place and ONE all-reduce over NCCL / NVLink per step instead of 25 MB buckets with per-bucket hooks.
"""
from __future__ import annotations

import os
import weakref

import torch
import torch.distributed as dist

# FlatGradSync instances by the identity of the parameters they own: the hand-written backward (_train.py) looks its
# sink up through a decoder parameter and hands it the decoder / projection gradients BEFORE it runs the backward of
# the two LoRA encoder blocks, so that their all-reduce travels on a side stream under that backward.
_SINKS = {}


def sink_for(param):
    ref = _SINKS.get(id(param))
    return ref() if ref is not None else None


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
        self._offsets = {}
        off = 0
        for i, p in enumerate(self.params):
            self._offsets[id(p)] = (i, off)
            off += p.numel()
            _SINKS[id(p)] = weakref.ref(self)
        self.early = True          # per-instance switch of the early all-reduce (early_enabled)
        self._side = None          # side stream of the early all-reduce
        self._early = None         # (event, first flat element reduced early) of the current step

    def early_enabled(self):
        """DOD_EARLY_ALLREDUCE=1: the tail of the flat buffer (projection + decoder gradients, complete before the LoRA
        blocks' backward) is all-reduced on a side stream as soon as it is written (SURVEY 8e asks for the overlap).
        NCCL only.  Off by default -- measured on 2 B200s (L/14 r=8, 32 images per GPU): 50.55 ms per step with it,
        49.74 ms without (49.58 / 49.01 ms from one CUDA graph): the 26.6 MB collective costs ~0.3 ms where it stands,
        while its kernel running under the backward takes SMs from it and the 57 in-place accumulations move onto the
        critical path (profiles/r02_summary.md)."""
        return (self.early and os.environ.get("DOD_EARLY_ALLREDUCE", "0") == "1" and dist.is_available()
                and dist.is_initialized() and dist.get_world_size(self.group) > 1
                and dist.get_backend(self.group) == "nccl" and self.flat.is_cuda)

    def early_tail(self, grads_by_id):
        """Called from inside the backward: accumulate the given gradients (id(param) -> tensor) into the flat buffer
        and, if those parameters are exactly a suffix of the buffer, start its all-reduce (AVG) on the side stream.
        Returns the set of parameter ids whose gradient is now in place (the autograd node returns None for them)."""
        idx = sorted(self._offsets[pid][0] for pid in grads_by_id if pid in self._offsets)
        if not idx or idx != list(range(idx[0], len(self.params))):
            return set()
        if self._early is not None:
            # a second backward before all_reduce() (gradient accumulation): the tail holds the AVERAGE of the first
            # micro-step already; adding this micro-step's local gradients and averaging again is exact, because an
            # average is the same on every rank.  The additions must follow the collective that is still in flight.
            torch.cuda.current_stream(self.flat.device).wait_event(self._early[0])
            self._early = None
        done = set()
        for pid, g in grads_by_id.items():
            ent = self._offsets.get(pid)
            if ent is None:
                continue
            p = self.params[ent[0]]
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * ent[1]:
                return set() if not done else done      # views dropped: leave the rest to autograd + all_reduce()
            p.grad.add_(g.reshape(p.shape))
            done.add(pid)
        start = self._offsets[id(self.params[idx[0]])][1]
        cur = torch.cuda.current_stream(self.flat.device)
        if self._side is None:
            self._side = torch.cuda.Stream(self.flat.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.flat[start:], op=dist.ReduceOp.AVG, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._early = (ev, start)
        return done

    def attach(self):
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        if self._early is not None:                  # a backward without a following all_reduce(): join, then drop
            torch.cuda.current_stream(self.flat.device).wait_event(self._early[0])
            self._early = None
        self.flat.zero_()

    def all_reduce(self, average=True):
        """Sum (and average, like DDP) the flat gradient over the process group."""
        if not (dist.is_available() and dist.is_initialized()):
            return self.flat
        for p in self.params:                      # a set_to_none=True zero_grad dropped the views
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
                    p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * 4:
                raise RuntimeError("FlatGradSync: p.grad is no longer a view of the flat buffer; call attach()")
        if self._early is not None:
            # the tail went out from inside the backward (early_tail): reduce the head, then join the side stream
            ev, start = self._early
            self._early = None
            if not average:
                raise RuntimeError("FlatGradSync: the early all-reduce averages; call all_reduce(average=True)")
            if start > 0:
                dist.all_reduce(self.flat[:start], op=dist.ReduceOp.AVG, group=self.group)
            torch.cuda.current_stream(self.flat.device).wait_event(ev)
            return self.flat
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
