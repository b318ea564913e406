"""GPU parity of the training path: gradients of every trainable parameter from the libdod
backward vs torch autograd through the CPU oracle (fp32) on identical weights / images and the
same upstream gradient.  dropout = 0 (RNG streams cannot be matched)."""
import pytest
import torch

from helpers import build_product_model, detector_oracle, manifest, synth

pytestmark = pytest.mark.gpu


def _oracle_grads(sd, x, kw, g_logits, g_boxes):
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    with torch.enable_grad():
        mem = detector_oracle.backbone(sd, x, detector_oracle.variant_of(kw["dino_model_name"]), kw["lora_alpha"])
        out = detector_oracle.decoder(sd, mem, kw["nheads"], kw["n_points"], kw["use_deformable"])
        loss = (out["pred_logits"] * g_logits).sum() + (out["pred_boxes"] * g_boxes).sum()
        loss.backward()
    return sd, out


@pytest.mark.parametrize("case", ["c1_small_deform", "c1_small_std", "giant3_swiglu"])
def test_gradients_match_oracle_autograd(case):
    man = manifest()[case]
    model, sd, kw = build_product_model(case, device="cuda", dropout=0.0)
    model.train()
    x = synth.make_images(man["batch"], *man["hw"], seed=man["image_seed"])
    out = model(x.cuda())
    assert out["pred_logits"].requires_grad and out["pred_boxes"].requires_grad
    g = torch.Generator().manual_seed(11)
    g_logits = torch.randn(out["pred_logits"].shape, generator=g)
    g_boxes = torch.randn(out["pred_boxes"].shape, generator=g)
    loss = (out["pred_logits"] * g_logits.cuda()).sum() + (out["pred_boxes"] * g_boxes.cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    ref_sd, ref_out = _oracle_grads(sd, x, kw, g_logits, g_boxes)
    assert (out["pred_logits"].detach().cpu() - ref_out["pred_logits"].detach()).abs().max() < 0.05 * ref_out["pred_logits"].abs().max()
    n_dec = kw["num_decoder_layers"]
    worst = []
    checked = 0
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
        denom = ref.abs().max().clamp_min(1e-6)
        err = ((got - ref).abs().max() / denom).item()
        worst.append((err, name))
        checked += 1
    worst.sort(reverse=True)
    print("worst relative gradient errors:", worst[:8])
    assert checked > 20
    bad = [(e, n) for e, n in worst if e > 6e-2]
    assert not bad, bad


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
