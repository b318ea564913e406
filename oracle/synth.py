"""Deterministic synthetic weights / inputs for parity tests -- TEST INFRASTRUCTURE ONLY.

`make_state_dict` builds a reference-layout state_dict (exact key names and shapes
of `DINOv2ObjectDetector(...).state_dict()`) from a seed without
importing the reference, so the same weights can be rebuilt on the GPU box where
/root/reference does not exist.  oracle/make_golden.py loads these dicts into the
real reference with strict=True, which pins the key/shape list itself.

All tensors that the reference zero- or identity-initialises (lora_B,
sampling_offsets, attention_weights, LayerScale, LN affine, biases) are
randomised so that no term of the forward pass is hidden by an init value.
"""
from __future__ import annotations

import torch

from detector_oracle import VARIANTS, variant_of

PATCH = 14
POS_TOKENS = 1 + 37 * 37  # HF checkpoints: image_size 518


class _Gen:
    def __init__(self, seed):
        self.g = torch.Generator(device="cpu")
        self.g.manual_seed(seed)

    def randn(self, *shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=self.g) * std + mean

    def uniform(self, *shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=self.g) * (hi - lo) + lo


def _lin(sd, g, prefix, out_f, in_f, bias=True, wstd=None, bstd=0.05):
    sd[prefix + ".weight"] = g.randn(out_f, in_f, std=wstd if wstd is not None else in_f ** -0.5)
    if bias:
        sd[prefix + ".bias"] = g.randn(out_f, std=bstd)


def _ln(sd, g, prefix, d):
    sd[prefix + ".weight"] = g.randn(d, std=0.1, mean=1.0)
    sd[prefix + ".bias"] = g.randn(d, std=0.05)


def _maybe_lora(sd, g, prefix, out_f, in_f, r):
    """nn.Linear, or the LoraLinear key layout (X.linear.*, X.lora_A, X.lora_B) when r > 0."""
    if r:
        _lin(sd, g, prefix + ".linear", out_f, in_f)
        sd[prefix + ".lora_A.weight"] = g.randn(r, in_f, std=in_f ** -0.5)
        sd[prefix + ".lora_B.weight"] = g.randn(out_f, r, std=0.3 * r ** -0.5)
    else:
        _lin(sd, g, prefix, out_f, in_f)


def make_state_dict(*, dino_model_name="facebook/dinov2-base", num_classes=91, lora_r=2, hidden_dim=768,
                    num_queries=50, nheads=8, num_decoder_layers=3, dim_feedforward=1024, n_points=2,
                    use_deformable=True, backbone_layers=None, seed=0):
    """Reference-layout fp32 CPU state_dict for the given constructor arguments.
    `backbone_layers` overrides the encoder depth (used for reduced-depth giant cases)."""
    cfg = VARIANTS[variant_of(dino_model_name)]
    d, heads, swiglu = cfg["dim"], cfg["heads"], cfg["swiglu"]
    n_layers = backbone_layers or cfg["layers"]
    if hidden_dim is None:
        hidden_dim = d
    g = _Gen(seed)
    sd = {}
    e = "backbone.dino.embeddings."
    sd[e + "cls_token"] = g.randn(1, 1, d, std=0.5)
    sd[e + "mask_token"] = torch.zeros(1, d)
    sd[e + "position_embeddings"] = g.randn(1, POS_TOKENS, d, std=0.3)
    sd[e + "patch_embeddings.projection.weight"] = g.randn(d, 3, PATCH, PATCH, std=(3 * PATCH * PATCH) ** -0.5 * 3)
    sd[e + "patch_embeddings.projection.bias"] = g.randn(d, std=0.05)
    for i in range(n_layers):
        p = f"backbone.dino.encoder.layer.{i}."
        r = lora_r if i >= n_layers - min(2, n_layers) else 0
        _ln(sd, g, p + "norm1", d)
        for nm in ("query", "key", "value"):
            _maybe_lora(sd, g, p + "attention.attention." + nm, d, d, r)
        _maybe_lora(sd, g, p + "attention.output.dense", d, d, r)
        sd[p + "layer_scale1.lambda1"] = g.uniform(d, lo=0.2, hi=1.0)
        _ln(sd, g, p + "norm2", d)
        if swiglu:
            hid = (int(d * 4 * 2 / 3) + 7) // 8 * 8      # HF Dinov2SwiGLUFFN: 4096 for d=1536
            _maybe_lora(sd, g, p + "mlp.weights_in", 2 * hid, d, r)
            _maybe_lora(sd, g, p + "mlp.weights_out", d, hid, r)
        else:
            _maybe_lora(sd, g, p + "mlp.fc1", 4 * d, d, r)
            _maybe_lora(sd, g, p + "mlp.fc2", d, 4 * d, r)
        sd[p + "layer_scale2.lambda1"] = g.uniform(d, lo=0.2, hi=1.0)
    _ln(sd, g, "backbone.dino.layernorm", d)
    if hidden_dim != d:
        _lin(sd, g, "backbone.projection", hidden_dim, d)
    h = hidden_dim
    sd["decoder.query_embed.weight"] = g.randn(num_queries, h)
    for i in range(num_decoder_layers):
        p = f"decoder.decoder.layers.{i}."
        sd[p + "self_attn.in_proj_weight"] = g.randn(3 * h, h, std=h ** -0.5)
        sd[p + "self_attn.in_proj_bias"] = g.randn(3 * h, std=0.05)
        _lin(sd, g, p + "self_attn.out_proj", h, h)
        if use_deformable:
            # sampling positions are kept weakly dependent on the query (small weights, O(1)
            # biases): bilinear sampling of an unsmooth random feature map amplifies position
            # noise by (grid_w - 1) * |v1 - v0| per layer, which makes fp32-vs-fp64 runs of the
            # REFERENCE ITSELF disagree at 1e-2 when these weights are O(1/sqrt(h)).
            _lin(sd, g, p + "cross_attn.sampling_offsets", nheads * n_points * 2, h, wstd=0.02 * h ** -0.5, bstd=0.2)
            _lin(sd, g, p + "cross_attn.attention_weights", nheads * n_points, h, bstd=0.5)
            _lin(sd, g, p + "cross_attn.value_proj", h, h)
            _lin(sd, g, p + "cross_attn.output_proj", h, h)
            _lin(sd, g, p + "reference_points_proj", 2, h, wstd=0.02 * h ** -0.5, bstd=1.0)
        else:
            sd[p + "multihead_attn.in_proj_weight"] = g.randn(3 * h, h, std=h ** -0.5)
            sd[p + "multihead_attn.in_proj_bias"] = g.randn(3 * h, std=0.05)
            _lin(sd, g, p + "multihead_attn.out_proj", h, h)
        _lin(sd, g, p + "linear1", dim_feedforward, h)
        _lin(sd, g, p + "linear2", h, dim_feedforward)
        for nm in ("norm1", "norm2", "norm3"):
            _ln(sd, g, p + nm, h)
    _lin(sd, g, "decoder.class_embed", num_classes, h)
    _lin(sd, g, "decoder.bbox_embed.mlp.0", h // 2, h)
    _lin(sd, g, "decoder.bbox_embed.mlp.2", 4, h // 2)
    if use_deformable:
        _lin(sd, g, "decoder.reference_points", 2, h)
    return sd


def make_images(batch, height, width, seed=0):
    """fp32 RGB in [0, 1) like ToTensor() (ref: train.py:584-587)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.rand((batch, 3, height, width), generator=g)


def make_targets(batch, *, max_gt=50, num_classes=91, seed=0, min_gt=0):
    """COCO-style targets (ref: dataset.py:102-111): labels int64 [n], boxes cxcywh fp32 [n,4],
    non-degenerate so that GIoU has no NaN."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out = []
    for _ in range(batch):
        n = int(torch.randint(min_gt, max_gt + 1, (1,), generator=g))
        cxcy = torch.rand((n, 2), generator=g) * 0.6 + 0.2
        wh = torch.rand((n, 2), generator=g) * 0.3 + 0.02
        out.append({"labels": torch.randint(0, num_classes, (n,), generator=g),
                    "boxes": torch.cat([cxcy, wh], dim=1)})
    return out


def make_predictions(batch, num_queries, num_classes=91, seed=0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return {"pred_logits": torch.randn((batch, num_queries, num_classes), generator=g),
            "pred_boxes": torch.rand((batch, num_queries, 4), generator=g) * 0.5 + 0.25}


# Named parity cases: constructor kwargs + input shape.  Shared by make_golden.py and tests/.
CASES = {
    # BASELINE.json configs[0]: lightweight S/14 + decoder, 100 queries, fp32, B=2, 224x224
    "c1_small_deform": dict(ctor=dict(dino_model_name="facebook/dinov2-small", hidden_dim=256, num_queries=100,
                                      num_decoder_layers=2, dim_feedforward=512, lora_r=1, nheads=4),
                            batch=2, hw=(224, 224)),
    "c1_small_std": dict(ctor=dict(dino_model_name="facebook/dinov2-small", hidden_dim=256, num_queries=100,
                                   num_decoder_layers=2, dim_feedforward=512, lora_r=1, nheads=4,
                                   use_deformable=False),
                         batch=2, hw=(224, 224)),
    # default constructor (B/14, deformable) at the 518x518 size the metric is quoted on
    "base_518": dict(ctor=dict(), batch=1, hw=(518, 518)),
    # non-square, non-native size: bicubic position resize + (4, 79) grid factorisation
    "small_nonsquare": dict(ctor=dict(dino_model_name="facebook/dinov2-small", hidden_dim=None, num_queries=25,
                                      num_decoder_layers=2, dim_feedforward=768, lora_r=4, nheads=6, n_points=4),
                            batch=2, hw=(210, 294)),
    # L/14 r=8 with projection 1024 -> 768 and the standard decoder (config 4's model)
    "large_proj_std": dict(ctor=dict(dino_model_name="facebook/dinov2-large", lora_r=8, num_queries=100,
                                     use_deformable=False),
                           batch=1, hw=(224, 224)),
    # g/14 SwiGLU blocks, depth reduced to 3 to keep the CPU oracle in seconds
    "giant3_swiglu": dict(ctor=dict(dino_model_name="facebook/dinov2-giant", hidden_dim=None, lora_r=2,
                                    num_queries=50, nheads=8),
                          batch=1, hw=(224, 224), backbone_layers=3),
    # full-depth g/14 (40 SwiGLU blocks, 24 heads), default constructor otherwise (projection 1536 -> 768):
    # the model tools / bench.py time as config 5
    "giant40_full": dict(ctor=dict(dino_model_name="facebook/dinov2-giant"), batch=1, hw=(224, 224)),
    # config 4's model exactly as stated: L/14, LoRA r=8, default (deformable) decoder, 518x518
    "large_r8_deform_518": dict(ctor=dict(dino_model_name="facebook/dinov2-large", lora_r=8), batch=1, hw=(518, 518)),
}

CTOR_DEFAULTS = dict(num_classes=91, dino_model_name="facebook/dinov2-base", lora_r=2, lora_alpha=1.0,
                     hidden_dim=768, num_queries=50, nheads=8, num_decoder_layers=3, dim_feedforward=1024,
                     dropout=0.1, n_points=2, use_deformable=True)


def case_ctor(name):
    kw = dict(CTOR_DEFAULTS)
    kw.update(CASES[name]["ctor"])
    return kw


def case_state_dict(name, seed=0):
    kw = case_ctor(name)
    return make_state_dict(dino_model_name=kw["dino_model_name"], num_classes=kw["num_classes"],
                           lora_r=kw["lora_r"], hidden_dim=kw["hidden_dim"], num_queries=kw["num_queries"],
                           nheads=kw["nheads"], num_decoder_layers=kw["num_decoder_layers"],
                           dim_feedforward=kw["dim_feedforward"], n_points=kw["n_points"],
                           use_deformable=kw["use_deformable"],
                           backbone_layers=CASES[name].get("backbone_layers"), seed=seed)
