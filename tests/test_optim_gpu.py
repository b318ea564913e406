"""GPU: fused clip + Adam vs torch.nn.utils.clip_grad_norm_ + torch.optim.Adam on the same gradients."""
import pytest
import torch

from helpers import build_product_model, synth

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam_with_clipping():
    from dino_detector.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(64, 32), (32,), (7, 5, 3), (1000,)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt = FusedAdam(ours, lr=1e-2, weight_decay=1e-4, max_grad_norm=1.0)
    ropt = torch.optim.Adam(ref, lr=1e-2, weight_decay=1e-4)
    for step in range(5):
        grads = [torch.randn(s, device="cuda") * (3.0 if step % 2 else 0.01) for s in shapes]   # clipped / not
        opt.zero_grad()
        for p, g in zip(ours, grads):
            p.grad.add_(g)
        for p, g in zip(ref, grads):
            p.grad = g.clone()
        norm = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        ropt.step()
        opt.step(all_reduce=False)
        assert torch.allclose(opt.grad_norm(), norm.reshape(1), rtol=1e-5)
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, atol=1e-6, rtol=1e-5)


def test_fused_adam_trains_the_detector():
    from dino_detector.losses import SetCriterion
    from dino_detector.matching import HungarianMatcher
    from dino_detector.optim import FusedAdam
    model, sd, kw = build_product_model("c1_small_std", device="cuda", dropout=0.0)
    model.train()
    crit = SetCriterion(HungarianMatcher(), 91, {"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0})
    crit.strict = False                                   # no host sync anywhere in the step
    opt = FusedAdam(model.parameters(), lr=2e-4, weight_decay=1e-4, max_grad_norm=1.0)
    x = synth.make_images(2, 224, 224, seed=3).cuda()
    targets = [{k: v.cuda() for k, v in t.items()} for t in synth.make_targets(2, max_gt=8, seed=12, min_gt=2)]
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = sum(crit(model(x), targets).values())
        loss.backward()
        opt.step(all_reduce=False)
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    # eval after training sees the updated weights (pack invalidation through the weight epoch)
    model.eval()
    with torch.no_grad():
        a = model(x)["pred_logits"]
    model.train()
    b = model(x)["pred_logits"].detach()
    assert (a - b).abs().max() < 0.05 * b.abs().max()
