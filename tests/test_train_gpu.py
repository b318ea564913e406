"""GPU parity of the training path: gradients of every trainable parameter from the libdod
backward vs torch autograd through the CPU oracle (fp32) on identical weights / images and the
same upstream gradient.  dropout = 0 (RNG streams cannot be matched)."""
import pytest
import torch

from helpers import build_product_model, detector_oracle, manifest, synth

pytestmark = pytest.mark.gpu


def _loss(out):
    """A smooth detection-like objective (coherent gradients, unlike a random projection whose
    per-token contributions cancel and turn bf16 rounding into O(1) relative noise)."""
    return torch.nn.functional.softplus(out["pred_logits"]).sum() + ((out["pred_boxes"] - 0.3) ** 2).sum()


def _oracle_grads(sd, x, kw):
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    with torch.enable_grad():
        mem = detector_oracle.backbone(sd, x, detector_oracle.variant_of(kw["dino_model_name"]), kw["lora_alpha"])
        out = detector_oracle.decoder(sd, mem, kw["nheads"], kw["n_points"], kw["use_deformable"])
        _loss(out).backward()
    return sd, out


# Tolerances.  Gradients are compared per tensor with (a) the cosine against the fp32 oracle gradient
# and (b) max |diff| / max |ref|.  bf16 operands put ~1e-2 relative noise on every activation, so:
#  * standard decoder: every tensor must reach cos >= 0.99; (b) <= 0.15 except for sums over tokens that
#    cancel (e.g. a rank-1 LoRA A gradient sum_m dt[m] * ctx[m, :] when attention rows are nearly
#    identical keeps its direction -- cos 0.9999 -- but not its magnitude); at most two such tensors.
#  * deformable decoder: the sampling-position gradient (v10 - v00) * wy0 + ... switches to another
#    pair of tokens whenever bf16 noise moves a sample across an integer grid boundary, so it is
#    discontinuous in the activations and only direction-level agreement (cos >= 0.9) is meaningful;
#    the sampling backward itself is checked exactly in fp32 in test_train_kernels_gpu.py.
# (min cosine, max relative error, tensors allowed above it, median relative error over tensors)
CASE_TOL = {"c1_small_std": (0.99, 0.15, 2, 0.1), "large_proj_std": (0.99, 0.15, 2, 0.1),
            # giant3: three applications of the shared layer on a (1, 257) grid, the most position-sensitive case
            "c1_small_deform": (0.90, 1.0, 0, 0.2), "giant3_swiglu": (0.80, 1.0, 0, 0.3),
            # config 4's model as stated: L/14, LoRA r=8, deformable decoder, 518x518 (a (10, 137) grid)
            "large_r8_deform_518": (0.90, 1.0, 0, 0.2)}


@pytest.mark.parametrize("case", ["c1_small_std", "c1_small_deform", "giant3_swiglu", "large_proj_std",
                                  "large_r8_deform_518"])
def test_gradients_match_oracle_autograd(case):
    man = manifest()[case]
    model, sd, kw = build_product_model(case, device="cuda", dropout=0.0)
    model.train()
    x = synth.make_images(man["batch"], *man["hw"], seed=man["image_seed"])
    out = model(x.cuda())
    assert out["pred_logits"].requires_grad and out["pred_boxes"].requires_grad
    _loss(out).backward()
    torch.cuda.synchronize()
    ref_sd, ref_out = _oracle_grads(sd, x, kw)
    assert (out["pred_logits"].detach().cpu() - ref_out["pred_logits"].detach()).abs().max() < 0.05 * ref_out["pred_logits"].abs().max()
    n_dec = kw["num_decoder_layers"]
    min_cos, max_err, n_exempt, median_tol = CASE_TOL[case]
    rows = []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            assert p.grad is None, name
            continue
        key = name
        if kw["use_deformable"] and name.startswith("decoder.decoder.layers.0."):
            key = name.replace("layers.0.", f"layers.{n_dec - 1}.")      # shared layer: last key wins
        ref = ref_sd[key].grad
        if name.startswith("decoder.reference_points."):
            assert p.grad is None and ref is None                          # unused in the forward
            continue
        assert p.grad is not None, name
        assert ref is not None, key
        got = p.grad.detach().float().cpu()
        err = ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-9)).item()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
        rows.append((cos, err, name))
    assert len(rows) > 20
    all_got = torch.cat([p.grad.detach().float().cpu().flatten() for n, p in model.named_parameters()
                         if p.grad is not None])
    # the two projections that produce sampling POSITIONS get the discontinuous part of the gradient
    # directly: a one-ulp change of a bf16 activation (e.g. a different but equally exact summation order
    # in the decoder self-attention) moves a sample across a grid line and shifts their cosine by ~0.1
    def floor(name):
        pos = "reference_points_proj" in name or "sampling_offsets" in name
        return min(min_cos, 0.6) if pos and kw["use_deformable"] else min_cos
    low_cos = [r for r in rows if r[0] < floor(r[2])]
    assert not low_cos, low_cos
    big_err = [r for r in rows if r[1] > max_err]
    assert len(big_err) <= n_exempt, big_err
    errs = sorted(r[1] for r in rows)
    assert errs[len(errs) // 2] < median_tol, errs                         # median over tensors
    assert torch.isfinite(all_got).all()
    # achieved numbers on record: per-tensor extremes and the WHOLE flat gradient against the oracle's
    from helpers import record_margin
    all_ref = torch.cat([ref_sd[(n.replace("layers.0.", f"layers.{n_dec - 1}.") if kw["use_deformable"] and
                                 n.startswith("decoder.decoder.layers.0.") else n)].grad.flatten()
                         for n, p in model.named_parameters() if p.grad is not None])
    flat_cos = torch.nn.functional.cosine_similarity(all_got, all_ref, dim=0).item()
    flat_rel = ((all_got - all_ref).norm() / all_ref.norm()).item()
    record_margin(case, "train grads: 1 - min cosine over tensors (bf16 vs fp32 oracle autograd)",
                  1.0 - min(r[0] for r in rows), 1.0 - min(min_cos, 0.6) if kw["use_deformable"] else 1.0 - min_cos,
                  worst_tensor=min(rows)[2])
    record_margin(case, "train grads: median over tensors of max|diff|/max|ref|", errs[len(errs) // 2], median_tol)
    record_margin(case, "train grads: whole flat gradient |g - g_ref| / |g_ref|", flat_rel, 1.0, flat_cosine=flat_cos)


def test_optimizer_step_changes_outputs_and_frozen_weights_stay():
    """A reference-style training step (train.py:1000-1004, 1075-1110) runs unchanged."""
    case = "c1_small_deform"
    model, sd, kw = build_product_model(case, device="cuda", dropout=0.1)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-4)
    x = synth.make_images(2, 224, 224, seed=3).cuda()
    frozen_before = model.backbone.dino.encoder.layer[0].attention.attention.query.weight.clone()
    losses = []
    for _ in range(3):
        opt.zero_grad()
        out = model(x)
        loss = out["pred_logits"].square().mean() + (out["pred_boxes"] - 0.5).square().mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    assert torch.equal(frozen_before, model.backbone.dino.encoder.layer[0].attention.attention.query.weight)
    assert model.decoder.reference_points.weight.grad is None
