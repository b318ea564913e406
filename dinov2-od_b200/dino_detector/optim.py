"""Fused clip + Adam on flat buffers (reference train.py:1000-1004, 1104-1110).

`FusedAdam` re-homes the trainable parameters into one flat fp32 buffer (each `p.data` becomes a view),
keeps their gradients in the `FlatGradSync` buffer that is also the NCCL all-reduce payload, and runs
`clip_grad_norm_(params, max_norm)` + `Adam.step()` as two kernel launches for the whole model with no
host sync.  Semantics are torch.optim.Adam's (L2 weight decay added to the gradient, bias correction,
eps outside the square root) and torch.nn.utils.clip_grad_norm_'s (norm over the trainable gradients,
coef = min(1, max_norm / (norm + 1e-6))).
"""
from __future__ import annotations

import torch

from . import _engine, ops
from .parallel import FlatGradSync


class FusedAdam:
    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_grad_norm=1.0,
                 process_group=None):
        params = list(params)
        # torch.optim.Adam's parameter numbering (state_dict layout): every requires_grad parameter, in order
        self.all_params = [p for p in params if p.requires_grad]
        self.sync = FlatGradSync(params, process_group)
        ps = self.sync.params
        dev = ps[0].device
        self.flat_param = torch.empty(self.sync.numel, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in ps:
                n = p.numel()
                view = self.flat_param[off:off + n].view_as(p)
                view.copy_(p.data)
                p.data = view                         # parameters now live in the flat buffer
                off += n
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.norm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        # optional int64 device tensor holding the step number (runtime.GraphedTrainStep advances it
        # inside the captured graph); None = the host-side step_count is passed by value
        self.device_step = None

    def zero_grad(self):
        self.sync.zero()

    def step(self, all_reduce=True):
        """(all-reduce the flat gradient,) clip by global norm and apply Adam."""
        if all_reduce:
            self.sync.all_reduce(average=True)
        self.step_count += 1
        self.norm_sq.zero_()
        if self.max_grad_norm and self.max_grad_norm > 0:
            ops.sumsq(self.sync.flat, self.norm_sq)
        ops.adam_step(self.flat_param, self.sync.flat, self.exp_avg, self.exp_avg_sq, self.step_count,
                      lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                      weight_decay=self.weight_decay, grad_sumsq=self.norm_sq,
                      max_grad_norm=float(self.max_grad_norm or 0.0), step_ptr=self.device_step)
        # the update went through the flat buffer, not torch's per-tensor version counters
        _engine.bump_weight_epoch()

    # ---- checkpointing in torch.optim.Adam's layout (reference train.py:1011-1016, 1281-1287) ----
    def _spans(self):
        """(index in all_params, offset, numel) of every parameter held in the flat buffers."""
        index = {id(p): i for i, p in enumerate(self.all_params)}
        off, out = 0, []
        for p in self.sync.params:
            out.append((index[id(p)], off, p.numel(), p.shape))
            off += p.numel()
        return out

    def state_dict(self):
        """Same structure as `torch.optim.Adam(filter(requires_grad, model.parameters())).state_dict()`:
        per-parameter `step` / `exp_avg` / `exp_avg_sq` (grad-less parameters have no state, as in torch)."""
        if self.device_step is not None:
            self.step_count = int(self.device_step.item())
        state = {}
        if self.step_count > 0:
            for i, off, n, shape in self._spans():
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[off:off + n].view(shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view(shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": False, "params": list(range(len(self.all_params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.all_params):
            raise ValueError("loaded state dict contains a parameter group that doesn't match the size of "
                             "optimizer's group")               # torch.optim's message (train.py:1017 catches it)
        g = groups[0]
        self.lr, self.betas, self.eps = g["lr"], tuple(g["betas"]), g["eps"]
        self.weight_decay = g["weight_decay"]
        steps = set()
        with torch.no_grad():
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            for i, off, n, shape in self._spans():
                st = sd["state"].get(i, sd["state"].get(str(i)))
                if st is None:
                    continue
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"FusedAdam keeps one step count for all parameters, the checkpoint has {sorted(steps)}")
        self.step_count = steps.pop() if steps else 0
        if self.device_step is not None:
            self.device_step.fill_(self.step_count)

    def grad_norm(self):
        """Global gradient norm of the last step (device tensor; reading it syncs)."""
        return self.norm_sq.sqrt()
