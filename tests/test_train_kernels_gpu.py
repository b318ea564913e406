"""Kernel-level parity of the backward / training helpers (csrc/train.cu) against torch autograd
or a plain torch restatement on the same seeded inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from dino_detector import ops as _ops
    return _ops


def _g(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _randn(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).cuda()


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-9)).item()


def test_gemm_batched_strided_heads(ops):
    """S[b] = Q_h[b] K_h[b]^T on head slices of a fused qkv buffer (training attention)."""
    g = _g(0)
    b, s, h = 3, 257, 4
    d = h * 64
    qkv = _randn((b * s, 3 * d), g).bfloat16()
    q3 = qkv.view(b, s, 3 * d)
    sp = (s + 7) // 8 * 8
    for head in range(h):
        out = torch.empty((b, s, sp), dtype=torch.float32, device="cuda")[:, :, :s]
        ops.gemm_batched(q3[:, :, head * 64:(head + 1) * 64], q3[:, :, d + head * 64:d + (head + 1) * 64], out)
        ref = q3[:, :, head * 64:(head + 1) * 64].float() @ q3[:, :, d + head * 64:d + (head + 1) * 64].float().transpose(1, 2)
        assert _rel(out, ref) < 2e-5


def test_gemm_batched_two_levels(ops):
    """All (image, head) problems in one launch through head-strided 4-D views."""
    g = _g(3)
    b, s, h = 2, 300, 3
    d = h * 64
    qkv = _randn((b * s, 3 * d), g).bfloat16()
    q3 = qkv.view(b, s, 3 * d)
    q4 = q3[:, :, :d].as_strided((b, h, s, 64), (s * 3 * d, 64, 3 * d, 1), q3[:, :, :d].storage_offset())
    k4 = q3[:, :, d:2 * d].as_strided((b, h, s, 64), (s * 3 * d, 64, 3 * d, 1), q3[:, :, d:2 * d].storage_offset())
    sp = (s + 7) // 8 * 8
    out = torch.empty((b, h, s, sp), dtype=torch.float32, device="cuda")
    ops.gemm_batched(q4, k4, out[..., :s])
    ref = q4.float() @ k4.float().transpose(-1, -2)
    assert _rel(out[..., :s], ref) < 2e-5
    t = ops.transpose(k4)
    assert t.shape == (b, h, 64, s) and torch.equal(t, k4.transpose(-1, -2))
    # strided bf16 output into head slices of a [B, L, D] buffer
    dst = torch.zeros((b, s, d), dtype=torch.bfloat16, device="cuda")
    d4 = dst.as_strided((b, h, s, 64), (s * d, 64, d, 1), 0)
    p_full = torch.zeros((b, h, s, sp), dtype=torch.bfloat16, device="cuda")     # row stride % 8 == 0
    p_full[..., :s] = torch.softmax(ref, -1).bfloat16()
    p = p_full[..., :s]
    ops.gemm_batched(p, ops.transpose(k4), d4)
    assert _rel(d4, p.float() @ k4.float()) < 1e-2


@pytest.mark.parametrize("m,n,k", [(200, 64, 333), (136, 120, 64), (1370, 64, 1370), (640, 768, 1000),
                                   (1024, 1024, 4384), (520, 264, 72)])
@pytest.mark.parametrize("a_trans,w_trans", [(True, False), (False, True), (True, True)])
def test_gemm_transposed_views(ops, m, n, k, a_trans, w_trans):
    """a_trans / w_trans read the stored transposes as MN-major tcgen05 operands (dW = dY^T X,
    dX = dY W without a transposed copy); both the 1-CTA and the CTA-pair kernel."""
    g = _g(m + n + k)
    pad = lambda v: (v + 7) // 8 * 8
    a = _randn((k, pad(m)) if a_trans else (m, pad(k)), g).bfloat16()
    w = _randn((k, pad(n)) if w_trans else (n, pad(k)), g).bfloat16()
    a_v = a[:, :m] if a_trans else a[:, :k]
    w_v = w[:, :n] if w_trans else w[:, :k]
    out = torch.empty((m, pad(n)), dtype=torch.float32, device="cuda")[:, :n]
    ops.gemm(a_v, w_v, None, out=out, a_trans=a_trans, w_trans=w_trans)
    af = a_v.float().t() if a_trans else a_v.float()
    wf = w_v.float() if w_trans else w_v.float().t()
    assert _rel(out, af @ wf) < 2e-5
    # accumulate into an fp32 buffer through the residual path (weight-gradient use)
    if n % 8 == 0:
        acc = _randn((m, n), g)
        want = acc + af @ wf
        ops.gemm(a_v, w_v, None, residual=acc, out=acc, a_trans=a_trans, w_trans=w_trans)
        assert _rel(acc, want) < 2e-5


def test_gemm_batched_transposed_views(ops):
    """dK = dS^T Q and dQ = dS K over (image, head) batches without transposed copies."""
    g = _g(11)
    b, h, lq, lk = 2, 3, 300, 257
    d = h * 64
    lkp = (lk + 7) // 8 * 8
    ds_full = torch.zeros((b, h, lq, lkp), dtype=torch.bfloat16, device="cuda")
    ds_full[..., :lk] = _randn((b, h, lq, lk), g).bfloat16()
    ds = ds_full[..., :lk]
    q3 = _randn((b, lq, d), g).bfloat16()
    k3 = _randn((b, lk, d), g).bfloat16()
    heads = lambda x3, l: x3.as_strided((b, h, l, 64), (l * d, 64, d, 1), 0)
    q4, k4 = heads(q3, lq), heads(k3, lk)
    dk3 = torch.zeros((b, lk, d), dtype=torch.bfloat16, device="cuda")
    ops.gemm_batched(ds, q4, heads(dk3, lk), a_trans=True, w_trans=True)      # dK = dS^T Q
    assert _rel(heads(dk3, lk), ds.float().transpose(-1, -2) @ q4.float()) < 1e-2
    dq3 = torch.zeros((b, lq, d), dtype=torch.bfloat16, device="cuda")
    ops.gemm_batched(ds, k4, heads(dq3, lq), w_trans=True)                    # dQ = dS K
    assert _rel(heads(dq3, lq), ds.float() @ k4.float()) < 1e-2


@pytest.mark.parametrize("b,s,h", [(2, 257, 3), (1, 1370, 2), (3, 128, 1), (2, 100, 2)])
def test_fmha_bwd_matches_autograd(ops, b, s, h):
    """Fused attention backward (recomputed probabilities, TMEM-resident dK/dV, TMA reduce-add dQ)
    against torch autograd of softmax(QK^T/8)V on the same bf16 inputs."""
    g = _g(b * 1000 + s)
    d = h * 64
    ld = 3 * d + 8                                     # padded row: strides are not the tight ones
    qkv = torch.zeros((b * s, ld), dtype=torch.bfloat16, device="cuda")
    qkv[:, :3 * d] = _randn((b * s, 3 * d), g, 0.8).bfloat16()
    dctx = _randn((b * s, d), g).bfloat16()
    lse = torch.empty((b, h, s), dtype=torch.float32, device="cuda")
    ctx = ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125, lse=lse)
    x = qkv[:, :3 * d].float().view(b, s, 3, h, 64).permute(2, 0, 3, 1, 4).detach().requires_grad_(True)
    sc = (x[0] @ x[1].transpose(-1, -2)) * 0.125
    ref = torch.softmax(sc, -1) @ x[2]                 # [b, h, s, 64]
    want_lse = torch.logsumexp(sc, -1) * 1.4426950408889634
    assert (lse - want_lse).abs().max().item() < 2e-3
    ref.backward(dctx.float().view(b, s, h, 64).permute(0, 2, 1, 3))
    dqkv = torch.full((b * s, ld), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.fmha_bwd(qkv, ctx, dctx, lse, dqkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125)
    got = dqkv[:, :3 * d].float().view(b, s, 3, h, 64).permute(2, 0, 3, 1, 4)
    for i, name in enumerate(("dq", "dk", "dv")):
        err = _rel(got[i], x.grad[i])
        assert err < 2e-2, (name, err)
    assert (dqkv[:, 3 * d:] == 7.0).all()              # padding columns untouched


def test_transpose(ops):
    g = _g(1)
    x = _randn((3, 257, 100), g).bfloat16()
    t = ops.transpose(x)
    assert t.shape == (3, 100, 257) and t.stride(1) % 8 == 0
    assert torch.equal(t, x.transpose(1, 2))
    y = _randn((1000, 72), g).bfloat16()
    assert torch.equal(ops.transpose(y), y.t())


@pytest.mark.parametrize("c", [777, 776, 1024])        # 777: unaligned rows -> thread-per-column fallback kernel
@pytest.mark.parametrize("r", [1, 2, 8, 12, 24, 40])
def test_lowrank_wgrad_and_colsum(ops, r, c):
    g = _g(r + c)
    m = 3001
    big = _randn((m, c), g).bfloat16()
    small = torch.zeros((m, 64), dtype=torch.bfloat16, device="cuda")
    small[:, :r] = _randn((m, r), g).bfloat16()
    out = torch.zeros((c, r), device="cuda")
    ops.lowrank_wgrad(big, small, r, out, transposed=False, alpha=0.5)
    ref = 0.5 * big.float().t() @ small[:, :r].float()
    assert _rel(out, ref) < 1e-4
    out_t = torch.zeros((r, c), device="cuda")
    ops.lowrank_wgrad(big, small, r, out_t, transposed=True)
    assert _rel(out_t, 2 * ref.t()) < 1e-4
    cs = torch.zeros(c, device="cuda")
    ops.colsum(big, cs)
    assert _rel(cs, big.float().sum(0)) < 1e-4


@pytest.mark.parametrize("splits,mc,blocks,cb,r", [(4, 1370, 1, 1024, 8), (3, 257, 3, 384, 8), (2, 1370, 3, 1024, 8),
                                                   (5, 100, 1, 4096, 8), (2, 640, 1, 768, 24), (1, 300, 2, 256, 16)])
def test_lowrank_wgrad_tc(ops, splits, mc, blocks, cb, r):
    """LoRA dA / dB as one batched M-reduction tcgen05 GEMM over per-image row slices and diagonal blocks
    (fused q / k / v), against fp32 torch on the same bf16 operands; accumulates into `out`."""
    g = _g(splits * 1000 + mc + blocks)
    m = splits * mc
    big = _randn((m, blocks * cb), g).bfloat16()
    small = torch.zeros((m, 64), dtype=torch.bfloat16, device="cuda")
    small[:, :blocks * r] = _randn((m, blocks * r), g).bfloat16()
    ref = torch.stack([big[:, i * cb:(i + 1) * cb].float().t() @ small[:, i * r:(i + 1) * r].float()
                       for i in range(blocks)])                                   # [blocks, cb, r]
    out = torch.ones((blocks, cb, r), device="cuda")
    ops.lowrank_wgrad_tc(big, small, r, out, transposed=False, splits=splits, blocks=blocks)
    assert _rel(out - 1.0, ref) < 1e-4
    out_t = torch.zeros((blocks, r, cb), device="cuda")
    ops.lowrank_wgrad_tc(big, small, r, out_t, transposed=True, splits=splits, blocks=blocks)
    assert _rel(out_t, ref.transpose(1, 2)) < 1e-4


@pytest.mark.parametrize("d", [256, 768, 1024])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_layernorm_bwd(ops, d, dt):
    g = _g(d)
    rows = 333
    x = (_randn((rows, d), g) * 2 + 0.5).requires_grad_(True)
    gamma = _randn((d,), g).requires_grad_(True)
    beta = _randn((d,), g).requires_grad_(True)
    dy = _randn((rows, d), g).to(dt)
    dres = _randn((rows, d), g)
    y = torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-5)
    y.backward(dy.float())
    dgamma, dbeta = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    dx = ops.layernorm_bwd(dy, x.detach(), gamma.detach(), 1e-5, dres=dres, dgamma=dgamma, dbeta=dbeta)
    assert _rel(dx - dres, x.grad) < 1e-4
    assert _rel(dgamma, gamma.grad) < 1e-4 and _rel(dbeta, beta.grad) < 1e-4


def test_eltwise_modes(ops):
    g = _g(5)
    rows, cols = 257, 384
    a, b = _randn((rows, cols), g), _randn((rows, cols), g)
    vec = _randn((cols,), g)
    assert torch.equal(ops.eltwise(ops.ELT_CAST, a, out_dtype=torch.bfloat16), a.bfloat16())
    assert torch.allclose(ops.eltwise(ops.ELT_SCALE_COLS, a, vec=vec, out_dtype=torch.float32), a * vec)
    assert torch.allclose(ops.eltwise(ops.ELT_ADD, a, b, out_dtype=torch.float32), a + b)
    z = a.clone().requires_grad_(True)
    torch.nn.functional.gelu(z).backward(b)
    assert torch.allclose(ops.eltwise(ops.ELT_GELU_FWD, a, out_dtype=torch.float32), torch.nn.functional.gelu(a), atol=1e-6)
    assert torch.allclose(ops.eltwise(ops.ELT_GELU_BWD, b, a, out_dtype=torch.float32), z.grad, atol=1e-5)
    # bf16 outputs use the one-MUFU tanh / ex2 forms: within bf16 rounding of the exact result
    wide_a = a * 3.0
    zz = wide_a.clone().requires_grad_(True)
    torch.nn.functional.gelu(zz).backward(b)
    out16 = ops.eltwise(ops.ELT_GELU_FWD, wide_a, out_dtype=torch.bfloat16).float()
    ref16 = torch.nn.functional.gelu(wide_a)
    assert ((out16 - ref16).abs() <= 2.0 ** -8 * ref16.abs() + 3e-4).all()
    g16 = ops.eltwise(ops.ELT_GELU_BWD, b, wide_a, out_dtype=torch.bfloat16).float()
    assert ((g16 - zz.grad).abs() <= 2.0 ** -8 * zz.grad.abs() + 6e-4 * b.abs() + 1e-6).all()
    relu = torch.relu(a)
    assert torch.equal(ops.eltwise(ops.ELT_RELU_BWD, b, relu, out_dtype=torch.float32), b * (relu > 0))
    sg = torch.sigmoid(a)
    assert torch.allclose(ops.eltwise(ops.ELT_SIGMOID_BWD, b, sg, out_dtype=torch.float32), b * sg * (1 - sg), atol=1e-6)
    # ragged width (scalar path: 91 columns, rows not 16-byte aligned) and strided views (vector path)
    odd = _randn((33, 91), g)
    assert torch.equal(ops.eltwise(ops.ELT_CAST, odd, out_dtype=torch.bfloat16), odd.bfloat16())
    assert torch.allclose(ops.eltwise(ops.ELT_ADD, odd, odd, out_dtype=torch.float32), 2 * odd)
    wide = _randn((rows, 3 * cols), g)
    dst = torch.zeros((rows, 2 * cols), dtype=torch.bfloat16, device="cuda")
    ops.eltwise(ops.ELT_CAST, wide[:, cols:2 * cols], out=dst[:, cols:])
    assert torch.equal(dst[:, cols:], wide[:, cols:2 * cols].bfloat16()) and (dst[:, :cols] == 0).all()
    # SwiGLU
    zz = _randn((rows, 2 * cols), g).requires_grad_(True)
    x1, x2 = zz.chunk(2, dim=-1)
    y = torch.nn.functional.silu(x1) * x2
    y.backward(b)
    assert torch.allclose(ops.eltwise(ops.ELT_SWIGLU_FWD, zz.detach(), cols=cols, out_dtype=torch.float32), y.detach(), atol=1e-6)
    assert torch.allclose(ops.eltwise(ops.ELT_SWIGLU_BWD, b, zz.detach(), cols=cols, out_dtype=torch.float32), zz.grad, atol=1e-5)
    # dropout: deterministic in (seed, index), keeps ~1-p, scaled by 1/(1-p)
    d1 = ops.eltwise(ops.ELT_DROPOUT, a, p0=0.25, seed=7, out_dtype=torch.float32)
    d2 = ops.eltwise(ops.ELT_DROPOUT, a, p0=0.25, seed=7, out_dtype=torch.float32)
    assert torch.equal(d1, d2)
    kept = d1 != 0
    assert abs(kept.float().mean().item() - 0.75) < 0.01
    assert torch.allclose(d1[kept], a[kept] / 0.75)


def test_softmax_rows_fwd_bwd(ops):
    g = _g(6)
    rows, n = 515, 1370
    s = _randn((rows, n), g, 3.0)
    p = ops.softmax_rows(s, n, 0.125)
    assert p.shape == (rows, 1376) and (p[:, n:] == 0).all()
    ref = torch.softmax(0.125 * s, dim=-1)
    assert (p[:, :n].float() - ref).abs().max().item() < 2e-3
    dp = _randn((rows, 1376), g)
    ds = ops.softmax_bwd_rows(p, dp, n, 0.125)
    pf = p[:, :n].float()
    want = 0.125 * pf * (dp[:, :n] - (pf * dp[:, :n]).sum(-1, keepdim=True))
    assert _rel(ds[:, :n], want) < 1e-2 and (ds[:, n:] == 0).all()


@pytest.mark.parametrize("b,q,h,p,dh,gh,gw", [(2, 50, 8, 2, 96, 10, 137), (3, 33, 4, 4, 64, 1, 257), (32, 50, 8, 2, 96, 10, 137)])
@pytest.mark.parametrize("do_dt", [torch.bfloat16, torch.float32])
def test_deform_sample_bwd_vectorised(ops, b, q, h, p, dh, gh, gw, do_dt, monkeypatch):
    """The 8-channels-per-thread backward (bf16 values: vector reductions into d value, atomics-free sums for the
    position / weight gradients) against the thread-per-channel kernel and against torch autograd on the same
    bf16-rounded values; its position gradients are the same bits on every run."""
    from test_kernels_gpu import _deform_ref
    g = _g(b * 7 + q + dh)
    hp, d = h * p, h * dh
    v16 = _randn((b * gh * gw, d), g).bfloat16()
    cols = (3 * hp + 2 + 7) // 8 * 8
    raw = torch.zeros((b * q, cols), device="cuda")
    raw[:, :3 * hp + 2] = _randn((b * q, 3 * hp + 2), g) * 0.5
    dout = _randn((b * q, d), g).to(do_dt)

    def run(vec):
        monkeypatch.setenv("DOD_DEFORM_VEC", vec)
        dvalue = torch.zeros((b * gh * gw, d), device="cuda")
        dq = torch.zeros((b * q, cols), device="cuda")
        ops.deform_sample_bwd(v16, raw[:, 3 * hp:3 * hp + 2], raw[:, :2 * hp], raw[:, 2 * hp:3 * hp], dout, dvalue, dq,
                              b, q, h, p, dh, gh, gw, ref_is_logit=True)
        return dvalue, dq

    dv_old, dq_old = run("0")
    dv_new, dq_new = run("1")
    assert _rel(dv_new, dv_old) < 1e-5
    assert _rel(dq_new[:, :3 * hp + 2], dq_old[:, :3 * hp + 2]) < 1e-4
    dv_again, dq_again = run("1")
    assert torch.equal(dq_new, dq_again)                       # no atomics on this path
    value = v16.float().requires_grad_(True)
    rawg = raw.clone().requires_grad_(True)
    out = _deform_ref(value, rawg[:, 3 * hp:3 * hp + 2].sigmoid(), rawg[:, :2 * hp], rawg[:, 2 * hp:3 * hp], b, q, h, p, dh, gh, gw)
    out.backward(dout.float())
    assert _rel(dv_new, value.grad) < 1e-4
    assert _rel(dq_new[:, :3 * hp + 2], rawg.grad[:, :3 * hp + 2]) < 1e-3


@pytest.mark.parametrize("gh,gw", [(10, 137), (1, 257)])
def test_deform_sample_bwd_matches_autograd(ops, gh, gw):
    from test_kernels_gpu import _deform_ref
    g = _g(gh)
    b, q, h, p, dh = 2, 20, 4, 2, 64
    hp = h * p
    value = _randn((b * gh * gw, h * dh), g).requires_grad_(True)
    raw = torch.zeros((b * q, 32), device="cuda")
    raw[:, :3 * hp + 2] = _randn((b * q, 3 * hp + 2), g) * 0.5
    raw.requires_grad_(True)
    out = _deform_ref(value, raw[:, 3 * hp:3 * hp + 2].sigmoid(), raw[:, :2 * hp], raw[:, 2 * hp:3 * hp], b, q, h, p, dh, gh, gw)
    dout = _randn(out.shape, g)
    out.backward(dout)
    dvalue = torch.zeros_like(value)
    dq = torch.zeros((b * q, 32), device="cuda")
    rd = raw.detach()
    ops.deform_sample_bwd(value.detach(), rd[:, 3 * hp:3 * hp + 2], rd[:, :2 * hp], rd[:, 2 * hp:3 * hp], dout, dvalue, dq,
                          b, q, h, p, dh, gh, gw, ref_is_logit=True)
    assert _rel(dvalue, value.grad) < 1e-4
    assert _rel(dq[:, :3 * hp + 2], raw.grad[:, :3 * hp + 2]) < 1e-3
