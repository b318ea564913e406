"""GPU: the building blocks of the hand-written backward (dino_detector/_train.py) against torch
autograd on the same bf16-rounded operands."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _g(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _randn(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).cuda()


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-9)).item()


@pytest.mark.parametrize("m,n,k", [(200, 91, 256), (200, 4, 128), (514, 1024, 384), (200, 50, 256)])
def test_tlinear_backward(m, n, k):
    from dino_detector import _train, ops
    g = _g(m + n)
    lin = torch.nn.Linear(k, n).cuda()
    with torch.no_grad():
        lin.weight.copy_(_randn((n, k), g, k ** -0.5))
        lin.bias.copy_(_randn((n,), g))
    x = _randn((m, k), g).bfloat16()
    t = _train.TLinear(lin.weight, lin.bias, "t")
    y = t.fwd(x, out_dtype=torch.float32)
    wq = lin.weight.detach().bfloat16().float().requires_grad_(True)
    bq = lin.bias.detach().clone().requires_grad_(True)
    xq = x.float().requires_grad_(True)
    yref = xq @ wq.t() + bq
    assert _rel(y[:, :n], yref) < 1e-4
    dy = torch.zeros((m, t.n_pad), device="cuda")
    dy[:, :n] = _randn((m, n), g)
    dy16 = dy.bfloat16()
    yref.backward(dy16[:, :n].float())
    grads = _train.Grads()
    dx = t.bwd(dy16, x, grads)
    out = {}
    t.collect(grads, out)
    assert _rel(dx[:, :k], xq.grad) < 1e-2
    assert _rel(out[id(lin.weight)], wq.grad) < 1e-3
    assert _rel(out[id(lin.bias)], bq.grad) < 1e-3


@pytest.mark.parametrize("r", [1, 8])
def test_llinear_backward(r):
    from dino_detector import _train, ops
    from dino_detector.utils import LoraLinear
    g = _g(r)
    m, k, n = 514, 384, 384
    mods = []
    for i in range(3):
        lin = torch.nn.Linear(k, n)
        ll = LoraLinear(lin, r=r, alpha=0.7).cuda()
        with torch.no_grad():
            ll.linear.weight.copy_(_randn((n, k), g, k ** -0.5))
            ll.lora_A.weight.copy_(_randn((r, k), g, k ** -0.5))
            ll.lora_B.weight.copy_(_randn((n, r), g, 0.3))
        mods.append(ll)
    x = _randn((m, k), g).bfloat16()
    L = _train.LLinear(mods, {}, "k")
    y, t = L.fwd(x, out_dtype=torch.float32)
    xq = x.float().requires_grad_(True)
    params = []
    outs = []
    for ll in mods:
        w = ll.linear.weight.detach().bfloat16().float()
        a = ll.lora_A.weight.detach().bfloat16().float().requires_grad_(True)
        bb = (0.7 * ll.lora_B.weight.detach()).bfloat16().float().requires_grad_(True)   # alpha folded like the pack
        params.append((a, bb))
        outs.append(xq @ w.t() + ll.linear.bias.detach() + (xq @ a.t()).bfloat16().float() @ bb.t())
    yref = torch.cat(outs, dim=1)
    assert _rel(y, yref) < 2e-3
    dy = _randn((m, 3 * n), g).bfloat16()
    yref.backward(dy.float())
    grads = _train.Grads()
    dx = L.bwd(dy, x, t, grads)
    out = {}
    L.collect(grads, out)
    assert _rel(dx, xq.grad) < 2e-2
    for ll, (a, bb) in zip(mods, params):
        assert _rel(out[id(ll.lora_A.weight)], a.grad) < 2e-2
        # d/dB_param = alpha * d/d(alpha B)
        assert _rel(out[id(ll.lora_B.weight)], 0.7 * bb.grad) < 2e-2


@pytest.mark.parametrize("b,lq,lk,h,dh", [(2, 100, 100, 4, 64), (2, 257, 257, 6, 64), (2, 50, 1370, 8, 96)])
def test_attention_backward(b, lq, lk, h, dh):
    from dino_detector import _train
    g = _g(lq + lk)
    d = h * dh
    q = _randn((b, lq, d), g).bfloat16()
    k = _randn((b, lk, d), g).bfloat16()
    v = _randn((b, lk, d), g).bfloat16()
    do = _randn((b, lq, d), g).bfloat16()
    scale = 1 / math.sqrt(dh)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    qh = qf.view(b, lq, h, dh).transpose(1, 2)
    kh = kf.view(b, lk, h, dh).transpose(1, 2)
    vh = vf.view(b, lk, h, dh).transpose(1, 2)
    o = (torch.softmax(qh @ kh.transpose(-1, -2) * scale, -1) @ vh).transpose(1, 2).reshape(b, lq, d)
    o.backward(do.float())
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    _train.attention_bwd(q, k, v, do, dq, dk, dv, h, dh, scale)
    assert _rel(dq, qf.grad) < 3e-2
    assert _rel(dk, kf.grad) < 3e-2
    assert _rel(dv, vf.grad) < 3e-2


def test_attention_dropout_forward_backward_consistent_with_mask():
    """Attention-probability dropout: forward and backward use the same counter-hash mask."""
    from dino_detector import _train, ops
    g = _g(3)
    b, lq, lk, h, dh, p_drop, seed = 2, 50, 50, 1, 64, 0.3, 12345
    q, k, v, do = (_randn((b, n, h * dh), g).bfloat16() for n in (lq, lk, lk, lq))
    scale = 1 / math.sqrt(dh)
    lkp = (lk + 7) // 8 * 8
    ctx = _train.attention_fwd_dropout(q, k, v, h, dh, scale, p_drop, seed)
    # recover the mask the kernels drew (depends only on seed, row, column)
    ones = torch.zeros((b * lq, lkp), device="cuda")
    pd = ops.softmax_rows(ones, lk, 1.0, ldp=lkp, drop_p=p_drop, seed=seed)   # rows = (b, h=0, q)
    mask = (pd[:, :lk] > 0).float().view(b, lq, lk)
    assert abs(mask.mean().item() - (1 - p_drop)) < 0.03
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    P = torch.softmax(qf @ kf.transpose(1, 2) * scale, -1)
    o = (P * mask / (1 - p_drop)) @ vf
    assert _rel(ctx, o) < 2e-2
    o.backward(do.float())
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    _train.attention_bwd(q, k, v, do, dq, dk, dv, h, dh, scale, drop_p=p_drop, seed=seed)
    assert _rel(dq, qf.grad) < 3e-2 and _rel(dk, kf.grad) < 3e-2 and _rel(dv, vf.grad) < 3e-2
