"""DETR-style decoder + heads (reference models/detr_decoder.py:7-82) on libdod kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _engine
from ..utils import MLP
from .deformable_attention import DeformableDecoderLayer, DeformableTransformerDecoder


class DETRDecoder(nn.Module):
    def __init__(self, num_queries, hidden_dim, nheads, num_decoder_layers, num_classes,
                 dim_feedforward=2048, dropout=0.1, n_points=4, use_deformable=True):
        super().__init__()
        self.num_queries = num_queries
        self.use_deformable = use_deformable
        self.hidden_dim, self.nheads, self.n_points = hidden_dim, nheads, n_points
        self.dropout_p = dropout
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        if use_deformable:
            print("Using Deformable Attention in Decoder")
            layer = DeformableDecoderLayer(d_model=hidden_dim, n_heads=nheads,
                                           dim_feedforward=dim_feedforward, dropout=dropout,
                                           n_points=n_points)
            self.decoder = DeformableTransformerDecoder(layer, num_layers=num_decoder_layers)
        else:
            layer = nn.TransformerDecoderLayer(d_model=hidden_dim, nhead=nheads,
                                               dim_feedforward=dim_feedforward, dropout=dropout)
            self.decoder = nn.TransformerDecoder(layer, num_layers=num_decoder_layers)
        self.class_embed = nn.Linear(hidden_dim, num_classes)
        self.bbox_embed = MLP(hidden_dim, hidden_dim // 2, 4, num_layers=2)
        if use_deformable:
            self.reference_points = nn.Linear(hidden_dim, 2)   # unused, like reference :44-45
            for p in self.reference_points.parameters():
                # never receives a gradient (the reference needs find_unused_parameters=True for it): the flat
                # gradient buffer / FusedAdam leave it out, as torch.optim.Adam skips grad-less parameters
                p._dod_unused = True
        self.precision = None
        self._pack = None
        self._pack_key = None

    def _get_pack(self):
        mode = _engine.resolve_precision(self.precision)
        key = (mode, _engine.params_version(self))
        if self._pack is None or self._pack_key != key:
            self._pack = _engine.DecoderPack(self, mode)
            self._pack_key = key
        return self._pack

    def forward_rows(self, memory_rows, b, n):
        logits, boxes = _engine.decoder_forward(self._get_pack(), memory_rows, b, n)
        return {"pred_logits": logits, "pred_boxes": boxes}

    def forward(self, src):
        """src: [batch, seq_len, hidden_dim] -> {"pred_logits", "pred_boxes"} (reference :47-83)."""
        b, n, d = src.shape
        pack = self._get_pack()
        adt = torch.bfloat16 if pack.mode == "bf16" else torch.float32
        rows = src.detach().reshape(b * n, d)
        if rows.dtype != adt:
            rows = rows.to(adt)
        return self.forward_rows(rows.contiguous(), b, n)
