"""DINOv2 backbone with LoRA on the last two blocks (reference models/dinov2_backbone.py:7-66).

The reference wraps HF `Dinov2Model` (transformers modeling_dinov2.py:38-485); here the same
parameter tree (identical state_dict keys and shapes) is held by plain containers and the
forward pass is the libdod kernel sequence in `_engine.backbone_forward`.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import _engine
from ..utils import add_lora_to_module

# HF configs of facebook/dinov2-{small,base,large,giant} (image_size 518, patch 14, head dim 64)
_VARIANTS = {
    "small": dict(hidden_size=384, num_hidden_layers=12, num_attention_heads=6, use_swiglu_ffn=False),
    "base": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, use_swiglu_ffn=False),
    "large": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, use_swiglu_ffn=False),
    "giant": dict(hidden_size=1536, num_hidden_layers=40, num_attention_heads=24, use_swiglu_ffn=True),
}


# test hook: override the encoder depth of the next constructed backbone (reduced-depth
# giant parity case); None in normal use.
_LAYER_OVERRIDE = None


def _variant(model_variant: str) -> str:
    # reference dinov2_backbone.py:17-27: substring match, default base
    for v in ("small", "base", "large", "giant"):
        if v in model_variant:
            return v
    return "base"


class _Bag(nn.Module):
    """Attribute container (no forward): only gives parameters their HF key names."""


def _trunc_normal_(t, std=0.02):
    nn.init.trunc_normal_(t, mean=0.0, std=std)


class Dinov2Weights(nn.Module):
    """Parameter tree of HF Dinov2Model (keys `embeddings.*`, `encoder.layer.N.*`, `layernorm.*`),
    initialised like HF `_init_weights` (modeling_dinov2.py:406-422)."""

    def __init__(self, hidden_size, num_hidden_layers, num_attention_heads, use_swiglu_ffn,
                 image_size=518, patch_size=14, mlp_ratio=4, layerscale_value=1.0):
        super().__init__()
        d = hidden_size
        self.hidden_size, self.num_heads, self.use_swiglu = d, num_attention_heads, use_swiglu_ffn
        self.patch_size = patch_size
        emb = _Bag()
        emb.cls_token = nn.Parameter(torch.empty(1, 1, d))
        emb.mask_token = nn.Parameter(torch.zeros(1, d))
        emb.position_embeddings = nn.Parameter(torch.empty(1, (image_size // patch_size) ** 2 + 1, d))
        emb.patch_embeddings = _Bag()
        emb.patch_embeddings.projection = nn.Conv2d(3, d, kernel_size=patch_size, stride=patch_size)
        self.embeddings = emb
        self.encoder = _Bag()
        layers = []
        for _ in range(num_hidden_layers):
            lyr = _Bag()
            lyr.norm1 = nn.LayerNorm(d, eps=1e-6)
            lyr.attention = _Bag()
            lyr.attention.attention = _Bag()
            lyr.attention.attention.query = nn.Linear(d, d)
            lyr.attention.attention.key = nn.Linear(d, d)
            lyr.attention.attention.value = nn.Linear(d, d)
            lyr.attention.output = _Bag()
            lyr.attention.output.dense = nn.Linear(d, d)
            lyr.layer_scale1 = _Bag()
            lyr.layer_scale1.lambda1 = nn.Parameter(layerscale_value * torch.ones(d))
            lyr.norm2 = nn.LayerNorm(d, eps=1e-6)
            lyr.mlp = _Bag()
            if use_swiglu_ffn:
                hid = (int(d * mlp_ratio * 2 / 3) + 7) // 8 * 8      # modeling_dinov2.py:331-338
                lyr.mlp.weights_in = nn.Linear(d, 2 * hid)
                lyr.mlp.weights_out = nn.Linear(hid, d)
            else:
                lyr.mlp.fc1 = nn.Linear(d, d * mlp_ratio)
                lyr.mlp.fc2 = nn.Linear(d * mlp_ratio, d)
            lyr.layer_scale2 = _Bag()
            lyr.layer_scale2.lambda1 = nn.Parameter(layerscale_value * torch.ones(d))
            layers.append(lyr)
        self.encoder.layer = nn.ModuleList(layers)
        self.layernorm = nn.LayerNorm(d, eps=1e-6)
        self._init_weights()

    @torch.no_grad()
    def _init_weights(self):
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                _trunc_normal_(m.weight)
                if m.bias is not None:
                    m.bias.zero_()
            elif isinstance(m, nn.LayerNorm):
                m.weight.fill_(1.0)
                m.bias.zero_()
        _trunc_normal_(self.embeddings.position_embeddings)
        _trunc_normal_(self.embeddings.cls_token)

    @classmethod
    def from_pretrained(cls, model_name):
        """Build the named architecture.  There is no network on the target boxes, so weights are
        random-init unless DOD_PRETRAINED_DIR points at a directory holding `<variant>.pt` state
        dicts in HF key layout (pure I/O, loaded with load_state_dict)."""
        v = _variant(model_name.split("/")[-1])
        kw = dict(_VARIANTS[v])
        if _LAYER_OVERRIDE is not None:
            kw["num_hidden_layers"] = _LAYER_OVERRIDE
        model = cls(**kw)
        root = os.environ.get("DOD_PRETRAINED_DIR")
        if root:
            path = os.path.join(root, f"dinov2-{v}.pt")
            if os.path.exists(path):
                model.load_state_dict(torch.load(path, map_location="cpu"), strict=True)
        return model


class DINOv2Backbone(nn.Module):
    def __init__(self, model_name="facebook/dinov2-base", lora_r=4, lora_alpha=1.0, target_dim=None):
        super().__init__()
        self.dino = Dinov2Weights.from_pretrained(model_name)
        self.model_variant = model_name.split("/")[-1]
        self.hidden_dim = _VARIANTS[_variant(self.model_variant)]["hidden_size"]
        print(f"DINOv2 backbone variant: {self.model_variant}, hidden dimension: {self.hidden_dim}")
        self.target_dim = target_dim
        if target_dim is not None and target_dim != self.hidden_dim:
            print(f"Creating projection layer from dimension {self.hidden_dim} to {target_dim}")
            self.projection = nn.Linear(self.hidden_dim, target_dim)
        else:
            self.projection = None
        # freeze everything in the backbone (reference dinov2_backbone.py:40-41) ...
        for param in self.dino.parameters():
            param.requires_grad = False
        # ... then LoRA on every nn.Linear of the last two blocks (reference :45-51)
        num_layers = len(self.dino.encoder.layer)
        for i in range(num_layers - min(2, num_layers), num_layers):
            print(f"Applying LoRA to encoder layer {i}")
            add_lora_to_module(self.dino.encoder.layer[i], r=lora_r, alpha=lora_alpha)
        self.precision = None          # None -> DOD_PRECISION env / config.precision
        self._pack = None
        self._pack_key = None

    def _get_pack(self):
        mode = _engine.resolve_precision(self.precision)
        key = (mode, _engine.params_version(self))
        if self._pack is None or self._pack_key != key:
            self._pack = _engine.BackbonePack(self, mode)
            self._pack_key = key
        return self._pack

    def _get_frozen_pack(self, n_layers):
        """Embeddings + the first n_layers (frozen) blocks for the training forward.  Keyed on THOSE parameters
        only: the LoRA / projection updates of every optimizer step do not trigger a re-pack of 100+ MB of
        frozen weights (the whole-backbone key of _get_pack did, 140 cast launches per training step)."""
        mode = _engine.resolve_precision(self.precision)
        mods = [self.dino.embeddings] + list(self.dino.encoder.layer)[:n_layers]
        key = (mode, n_layers) + tuple((p.data_ptr(), p._version) for m in mods for p in m.parameters())
        if getattr(self, "_fpack", None) is None or self._fpack_key != key:
            self._fpack = _engine.BackbonePack(self, mode, n_layers=n_layers)
            self._fpack_key = key
        return self._fpack

    def ln_fold_max_mean_ratio(self, reset=False):
        """Largest |row mean| / row std any folded LayerNorm of this backbone has seen so far (reading it syncs).
        The folded form rounds the un-normalised residual stream to bf16, so its noise relative to the normalised
        signal grows like sqrt(1 + ratio^2): well below 1 it equals the standalone kernel's; if a checkpoint
        drives it above ~3, run with DOD_LN_FOLD=0."""
        best = 0.0
        for pk in (self._pack, getattr(self, "_fpack", None)):
            if pk is not None and getattr(pk, "ln_monitor", None) is not None:
                best = max(best, float(pk.ln_monitor.item()))
                if reset:
                    pk.ln_monitor.zero_()
        return best

    def forward_rows(self, pixel_values):
        """-> (memory [B*N, out_dim] in the activation dtype, B, N)."""
        return _engine.backbone_forward(self._get_pack(), pixel_values)

    def forward(self, pixel_values):
        """reference dinov2_backbone.py:58-67: last_hidden_state (+ projection), [B, N, dim]."""
        mem, b, n = self.forward_rows(pixel_values)
        return mem.view(b, n, -1)
