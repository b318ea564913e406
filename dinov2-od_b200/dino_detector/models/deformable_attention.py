"""Parameter containers of the reference's "deformable" decoder
(reference models/deformable_attention.py:8-308).  The arithmetic lives in
`_engine.decoder_forward` (dod_deform_sample replaces the 4-deep python loop at :147-170).
"""
from __future__ import annotations

import torch.nn as nn
from torch.nn.init import constant_, xavier_uniform_


class DeformableAttention(nn.Module):
    def __init__(self, d_model=256, n_heads=8, n_points=4):
        super().__init__()
        self.d_model, self.n_heads, self.n_points = d_model, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # reference deformable_attention.py:38-51
        constant_(self.sampling_offsets.weight.data, 0.)
        constant_(self.sampling_offsets.bias.data, 0.)
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)


class DeformableDecoderLayer(nn.Module):
    def __init__(self, d_model=256, n_heads=8, dim_feedforward=2048, dropout=0.1, n_points=4):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.cross_attn = DeformableAttention(d_model, n_heads, n_points)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.activation = nn.ReLU()
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)
        self.reference_points_proj = nn.Linear(d_model, 2)


class DeformableTransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers):
        super().__init__()
        # reference deformable_attention.py:284: the SAME layer object n times -> shared weights,
        # aliased state_dict keys layers.0.* ... layers.{n-1}.*
        self.layers = nn.ModuleList([decoder_layer for _ in range(num_layers)])
