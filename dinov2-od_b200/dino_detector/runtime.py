"""CUDA-graph replay of the inference forward and of the whole training step (launch-bound small batches).

The detector forward is ~200 kernel launches; at batch 2 / 224x224 (BASELINE config 1) the GPU work is
well under the host time needed to issue them.  `GraphedDetector` captures one forward for a fixed
input shape into a CUDA graph (all libdod launches go to torch's current stream, all buffers come from
the graph's private pool) and replays it with a single launch.  Weights must not change between
capture and replay (the packed copies are baked into the graph); call `recapture()` after an update.
"""
from __future__ import annotations

import torch


class GraphedDetector:
    def __init__(self, model, example_input, warmup=2):
        if torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters()) and model.training:
            raise ValueError("GraphedDetector captures the inference forward: call model.eval() first")
        self.model = model
        self.static_in = example_input.detach().clone().float().contiguous()
        self.graph = None
        self.static_out = None
        self._capture(warmup)

    def _capture(self, warmup):
        side = torch.cuda.Stream(self.static_in.device)
        side.wait_stream(torch.cuda.current_stream(self.static_in.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(1, warmup)):            # builds weight packs, sets kernel attributes
                self.model(self.static_in)
        torch.cuda.current_stream(self.static_in.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = self.model(self.static_in)

    def recapture(self):
        self._capture(1)

    @torch.no_grad()
    def __call__(self, pixel_values):
        if tuple(pixel_values.shape) != tuple(self.static_in.shape):
            raise ValueError(f"graph was captured for input shape {tuple(self.static_in.shape)}, "
                             f"got {tuple(pixel_values.shape)}")
        self.static_in.copy_(pixel_values, non_blocking=True)
        self.graph.replay()
        return self.static_out          # static buffers: overwritten by the next call, clone to keep


class GraphedTrainStep:
    """One whole training step -- forward, GPU matcher, fused criterion, hand-written backward, flat
    gradient all-reduce, global-norm clip + Adam -- captured into ONE CUDA graph and replayed.

    A train step is ~1300 kernel launches issued from Python; below ~16 images per GPU the host cannot
    issue them as fast as the GPU retires them (profiles/r01_summary.md: L/14, batch 8: 20 ms of kernels in
    a 31 ms step).  Everything on the step is device-resident already (the assignment never visits the
    host), so the step replays from a graph once three things live on the device instead of in launch
    arguments: the dropout seed and the Adam step number (int64 counters advanced inside the graph) and
    the targets (CSR buffers of fixed capacity that `__call__` refills before each replay).

        step = GraphedTrainStep(model, criterion, FusedAdam(model.parameters(), ...), images, max_targets=100)
        losses = step(images, targets)     # dict of device scalars (weighted, like SetCriterion.forward)

    The returned loss tensors are the graph's static outputs: the next call overwrites them, so read
    (or clone) them before stepping again.
    Shapes are fixed at capture (batch, H, W, max_targets per image).  The solver status is not checked
    on the host (criterion.strict is ignored): NaN / infeasible cost matrices leave those images unmatched.
    Under torch.distributed the NCCL all-reduces (num_boxes, flat gradient) are part of the graph.
    """

    def __init__(self, model, criterion, optimizer, example_images, max_targets=100, warmup=3):
        from . import ops
        from .optim import FusedAdam
        if not isinstance(optimizer, FusedAdam):
            raise TypeError("GraphedTrainStep needs optim.FusedAdam (flat device-resident state)")
        self.model, self.criterion, self.opt = model, criterion, optimizer
        dev = example_images.device
        b = example_images.shape[0]
        self.batch, self.max_t = b, int(max_targets)
        self.images = example_images.detach().clone().contiguous()
        cap = b * self.max_t
        self.labels = torch.zeros(cap, dtype=torch.int64, device=dev)
        self.boxes = torch.full((cap, 4), 0.5, dtype=torch.float32, device=dev)
        self.offsets = torch.zeros(b + 1, dtype=torch.int32, device=dev)
        self.num_boxes = torch.zeros(1, dtype=torch.float32, device=dev)
        # pinned staging for the per-step target upload (one async copy per buffer)
        self._h_labels = torch.zeros(cap, dtype=torch.int64).pin_memory()
        self._h_boxes = torch.zeros((cap, 4), dtype=torch.float32).pin_memory()
        self._h_offsets = torch.zeros(b + 1, dtype=torch.int32).pin_memory()
        self._h_num = torch.zeros(1, dtype=torch.float32).pin_memory()
        # recorded after the four H2D copies of a step; the staging buffers are not rewritten before it has
        # completed (the host runs ahead of the GPU: without this, step N's copies -- still queued behind the
        # replay of step N-1 -- would read the targets of step N+1)
        self._staged = torch.cuda.Event()
        self._staged_pending = False
        # [0] dropout seed epoch, [1] Adam step number; continue from the optimizer's host count
        self.counters = torch.tensor([1, optimizer.step_count], dtype=torch.int64, device=dev)
        self._ops = ops
        self.losses = None
        self.graph = None
        self._capture(warmup)

    def _step_body(self):
        ops = self._ops
        ops.counter_add(self.counters, 1)
        self.opt.zero_grad()
        out = self.model(self.images)
        packed = (self.labels, self.boxes, self.offsets, None, self.max_t)
        losses = self.criterion.forward_packed(out, packed, self.num_boxes)
        total = losses["loss_ce"] + losses["loss_bbox"] + losses["loss_giou"]
        total.backward()
        self.opt.step()
        return {k: v.detach() for k, v in losses.items()}

    def _capture(self, warmup):
        ops = self._ops
        dev = self.images.device
        strict, self.criterion.strict = self.criterion.strict, False
        ops.set_device_seed(self.counters)
        self.opt.device_step = self.counters[1:2]
        # the warm-up steps really run (on the example images, no targets): snapshot the optimizer's
        # flat state and the counters so that capturing does not train the model
        opt = self.opt
        snap = [t.clone() for t in (opt.flat_param, opt.exp_avg, opt.exp_avg_sq, self.counters)]
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):        # kernel attributes, allocator pools, NCCL warm-up
                    self._step_body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.losses = self._step_body()
            for dst, src in zip((opt.flat_param, opt.exp_avg, opt.exp_avg_sq, self.counters), snap):
                dst.copy_(src)
        finally:
            ops.set_device_seed(None)
            self.opt.device_step = None
            self.criterion.strict = strict
        self.opt.step_count = int(self.counters[1].item())

    def load_targets(self, targets):
        """Refill the static CSR target buffers from a list of target dicts (dataset.py:102-111)."""
        if len(targets) != self.batch:
            raise ValueError(f"graph was captured for {self.batch} images, got {len(targets)} target dicts")
        if self._staged_pending:
            self._staged.synchronize()
            self._staged_pending = False
        off = 0
        self._h_offsets[0] = 0
        for i, t in enumerate(targets):
            n = int(t["labels"].shape[0])
            if n > self.max_t:
                raise ValueError(f"image {i} has {n} targets, the graph was captured for at most {self.max_t}")
            if n:
                self._h_labels[off:off + n].copy_(t["labels"].reshape(-1))
                self._h_boxes[off:off + n].copy_(t["boxes"].reshape(-1, 4))
            off += n
            self._h_offsets[i + 1] = off
        self._h_num[0] = float(off)
        self.labels.copy_(self._h_labels, non_blocking=True)
        self.boxes.copy_(self._h_boxes, non_blocking=True)
        self.offsets.copy_(self._h_offsets, non_blocking=True)
        self.num_boxes.copy_(self._h_num, non_blocking=True)
        self._staged.record(torch.cuda.current_stream(self.labels.device))
        self._staged_pending = True

    def __call__(self, images, targets):
        if tuple(images.shape) != tuple(self.images.shape):
            raise ValueError(f"graph was captured for images of shape {tuple(self.images.shape)}, got {tuple(images.shape)}")
        self.images.copy_(images, non_blocking=True)
        self.load_targets(targets)
        self.graph.replay()
        self.opt.step_count += 1
        return self.losses
