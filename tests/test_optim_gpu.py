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


def test_fused_adam_state_dict_is_torch_adams_and_skips_gradless_parameters():
    """optimizer_state_dict save / resume of the reference (train.py:1011-1016, 1281-1287): FusedAdam's
    state_dict loads into torch.optim.Adam over the same parameter list and back, and a parameter that never
    receives a gradient (decoder.reference_points) is left alone exactly like torch.optim.Adam leaves it."""
    from dino_detector.optim import FusedAdam
    torch.manual_seed(1)
    shapes = [(16, 8), (8,), (5, 3), (3,)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ours[2]._dod_unused = True                                 # like decoder.reference_points.weight
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt = FusedAdam(ours, lr=1e-2, weight_decay=1e-2, max_grad_norm=0.0)
    ropt = torch.optim.Adam(ref, lr=1e-2, weight_decay=1e-2)
    assert ours[2].grad is None

    def one_step(o, params, fused):
        grads = [torch.full(s, 0.1, device="cuda") for s in shapes]
        if fused:
            o.zero_grad()
            for i, (p, g) in enumerate(zip(params, grads)):
                if i != 2:
                    p.grad.add_(g)
            o.step(all_reduce=False)
        else:
            for i, (p, g) in enumerate(zip(params, grads)):
                p.grad = None if i == 2 else g
            o.step()

    for _ in range(3):
        one_step(opt, ours, True)
        one_step(ropt, ref, False)
    assert torch.equal(ours[2], ref[2]), "a grad-less parameter must not be decayed"
    sd, rsd = opt.state_dict(), ropt.state_dict()
    assert sorted(sd["state"]) == sorted(rsd["state"]) == [0, 1, 3]
    assert sd["param_groups"][0]["params"] == rsd["param_groups"][0]["params"]
    for i in sd["state"]:
        assert float(sd["state"][i]["step"]) == float(rsd["state"][i]["step"]) == 3.0
        assert torch.allclose(sd["state"][i]["exp_avg"], rsd["state"][i]["exp_avg"], atol=1e-7)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], rsd["state"][i]["exp_avg_sq"], atol=1e-9)
    # ours -> torch -> continue, torch -> ours -> continue: same trajectory
    ropt2 = torch.optim.Adam(ref, lr=1e-2, weight_decay=1e-2)
    ropt2.load_state_dict(sd)
    opt2 = FusedAdam(ours, lr=1.0, weight_decay=0.0, max_grad_norm=0.0)
    opt2.load_state_dict(rsd)
    assert opt2.step_count == 3 and opt2.lr == 1e-2 and opt2.weight_decay == 1e-2
    one_step(opt2, ours, True)
    one_step(ropt2, ref, False)
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-5)
    with pytest.raises(ValueError):
        FusedAdam(ours[:2], lr=1e-2).load_state_dict(rsd)
