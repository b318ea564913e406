"""CPU: the drop-in boundary -- C ABI exports, constructor/state_dict contract, error behaviour."""
import ctypes
import inspect
import os
import re

import pytest
import torch

from helpers import ROOT, build_product_model, synth


def test_library_loads_and_exports_every_declared_symbol():
    from dino_detector import _dod
    with open(os.path.join(ROOT, "include", "dod.h")) as fh:
        declared = re.findall(r"DOD_API\s+[\w\s\*]+?\b(dod_\w+)\s*\(", fh.read())
    assert len(declared) >= 19
    lib = ctypes.CDLL(_dod.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libdod.so does not export {name}"
    assert _dod.lib().dod_version() >= 100
    assert set(_dod.FUNCTIONS) == set(declared)


def test_no_cpu_fallback():
    from dino_detector import ops
    from dino_detector._dod import DodError
    with pytest.raises(DodError):
        ops.layernorm(torch.zeros(4, 8), torch.ones(8), torch.zeros(8), 1e-6)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dinov2-od_b200", "dino_detector")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|detector_oracle|matcher_oracle|synth)\b",
                                     src, flags=re.M), f


def test_constructor_signature_matches_reference_defaults():
    from dino_detector.models import DINOv2ObjectDetector
    sig = inspect.signature(DINOv2ObjectDetector.__init__)
    got = {k: v.default for k, v in sig.parameters.items() if k != "self"}
    assert got == synth.CTOR_DEFAULTS          # reference detector.py:9-21 + config.py:21-36
    assert list(got) == list(synth.CTOR_DEFAULTS)


@pytest.mark.parametrize("case", ["c1_small_deform", "c1_small_std", "giant3_swiglu"])
def test_state_dict_keys_shapes_and_trainable_set(case):
    """Keys/shapes equal the reference's (the synth dict is loaded strict=True into the real
    reference by oracle/make_golden.py); trainable set = LoRA + projection + decoder."""
    model, sd, kw = build_product_model(case)
    ours = model.state_dict()
    assert set(ours) == set(sd)
    for k in sd:
        assert tuple(ours[k].shape) == tuple(sd[k].shape), k
    for name, p in model.named_parameters():
        want = ("lora_A" in name or "lora_B" in name or name.startswith("backbone.projection")
                or name.startswith("decoder."))
        assert p.requires_grad == want, name
    if kw["use_deformable"]:
        # reference deformable_attention.py:284: one layer object, aliased keys
        layers = model.decoder.decoder.layers
        assert all(l is layers[0] for l in layers)
        assert len({id(p) for p in model.parameters()}) == len(list(model.parameters()))


def test_deformable_load_last_key_wins():
    model, sd, _ = build_product_model("c1_small_deform")
    last = sd["decoder.decoder.layers.1.linear1.weight"]
    assert torch.equal(model.state_dict()["decoder.decoder.layers.0.linear1.weight"], last)


def test_hidden_dim_none_infers_backbone_width():
    model, _, _ = build_product_model("small_nonsquare")
    assert model.backbone.projection is None and model.decoder.hidden_dim == 384


def test_forward_without_gpu_fails_loudly():
    from dino_detector._dod import DodError
    model, _, _ = build_product_model("c1_small_std")
    with torch.no_grad(), pytest.raises(DodError):
        model(torch.rand(1, 3, 224, 224))


def test_layernorm_fold_and_lora_merge_identities():
    """The algebra behind the default bf16 inference packs (_engine.PackedLinear ln=..., _lin_parts merge=True),
    checked in float64 on the CPU:
      LN(h) W^T + b == rstd * (h W''^T) + (b + W beta)   with W'' = gamma (.) W minus its row means,
      (W + alpha B A) x == W x + alpha B (A x)            (reference utils.py:68-70)."""
    import torch
    g = torch.Generator().manual_seed(0)
    m, d, n, r = 37, 96, 40, 4
    h = torch.randn(m, d, generator=g, dtype=torch.float64) * 3 + 1.5
    w = torch.randn(n, d, generator=g, dtype=torch.float64)
    b = torch.randn(n, generator=g, dtype=torch.float64)
    gamma = 1 + 0.3 * torch.randn(d, generator=g, dtype=torch.float64)
    beta = 0.2 * torch.randn(d, generator=g, dtype=torch.float64)
    eps = 1e-6
    ref = torch.nn.functional.layer_norm(h, (d,), gamma, beta, eps) @ w.t() + b
    w2 = w * gamma[None, :]
    w2 = w2 - w2.mean(dim=1, keepdim=True)
    s1, s2 = h.sum(1), (h * h).sum(1)                     # what the producer GEMM's epilogue accumulates
    mean = s1 / d
    rstd = torch.rsqrt((s2 / d - mean * mean).clamp_min(0) + eps)
    out = rstd[:, None] * (h @ w2.t()) + (b + w @ beta)
    assert torch.allclose(out, ref, rtol=1e-9, atol=1e-9)

    a_l = torch.randn(r, d, generator=g, dtype=torch.float64)
    b_l = torch.randn(n, r, generator=g, dtype=torch.float64)
    alpha = 0.7
    x = torch.randn(m, d, generator=g, dtype=torch.float64)
    assert torch.allclose(x @ (w + alpha * b_l @ a_l).t() + b, x @ w.t() + b + alpha * (x @ a_l.t()) @ b_l.t(),
                          rtol=1e-10, atol=1e-10)
