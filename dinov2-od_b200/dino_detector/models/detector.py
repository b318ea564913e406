"""DINOv2ObjectDetector (reference models/detector.py:8-69): same constructor arguments and
defaults, same forward(images) -> {pred_logits, pred_boxes} contract, same state_dict keys."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import config
from .detr_decoder import DETRDecoder
from .dinov2_backbone import DINOv2Backbone


class DINOv2ObjectDetector(nn.Module):
    def __init__(self,
                 num_classes=config.num_classes,
                 dino_model_name=config.dino_model_name,
                 lora_r=config.lora_r,
                 lora_alpha=config.lora_alpha,
                 hidden_dim=config.hidden_dim,
                 num_queries=config.num_queries,
                 nheads=config.nheads,
                 num_decoder_layers=config.num_decoder_layers,
                 dim_feedforward=config.dim_feedforward,
                 dropout=config.dropout,
                 n_points=config.n_points,
                 use_deformable=config.use_deformable):
        super().__init__()
        if hidden_dim is None:
            # reference detector.py:25-35
            if 'small' in dino_model_name:
                hidden_dim = 384
            elif 'base' in dino_model_name:
                hidden_dim = 768
            elif 'large' in dino_model_name:
                hidden_dim = 1024
            elif 'giant' in dino_model_name:
                hidden_dim = 1536
            else:
                hidden_dim = 768
        self.backbone = DINOv2Backbone(model_name=dino_model_name, lora_r=lora_r,
                                       lora_alpha=lora_alpha, target_dim=hidden_dim)
        self.decoder = DETRDecoder(num_queries=num_queries, hidden_dim=hidden_dim, nheads=nheads,
                                   num_decoder_layers=num_decoder_layers, num_classes=num_classes,
                                   dim_feedforward=dim_feedforward, dropout=dropout,
                                   n_points=n_points, use_deformable=use_deformable)
        self._precision = None

    @property
    def precision(self):
        return self._precision

    @precision.setter
    def precision(self, value):
        """'bf16' | 'fp32' | None (DOD_PRECISION env / config.precision)."""
        self._precision = value
        self.backbone.precision = value
        self.decoder.precision = value

    def forward(self, pixel_values):
        """pixel_values [B, 3, H, W] fp32 in [0, 1] -> {"pred_logits": [B, Q, C], "pred_boxes": [B, Q, 4]}."""
        # the hand-written training forward (activation stash + autograd node) is for model.train() under grad
        # mode, as in the reference's loop (train.py:1042-1101); model.eval()(x) runs the inference sequence
        # even outside torch.no_grad() -- its outputs then carry no graph (the reference's would)
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .._train import detector_forward_train
            return detector_forward_train(self, pixel_values)
        mem, b, n = self.backbone.forward_rows(pixel_values)
        return self.decoder.forward_rows(mem, b, n)
