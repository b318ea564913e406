"""Kernel-level parity: every libdod op against a plain PyTorch fp32 restatement
of the same arithmetic on the same seeded inputs (B200 only)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from dino_detector import ops as _ops
    return _ops


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _randn(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).cuda()


# --------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("m,n,k", [
    (128, 64, 64), (128, 128, 128), (256, 256, 256), (300, 768, 768), (1000, 2304, 768),
    (87, 96, 392), (2740, 3072, 768), (513, 56, 768), (4096, 768, 3072), (129, 8, 64),
    (5000, 1024, 592), (4700, 520, 200), (20000, 520, 136), (16600, 256, 72),
])
def test_gemm_plain(ops, m, n, k):
    g = _gen(m * 7 + n * 3 + k)
    a = _randn((m, k), g).bfloat16()
    w = _randn((n, k), g, 1 / math.sqrt(k)).bfloat16()
    bias = _randn((n,), g)
    out = ops.gemm(a, w, bias, out_dtype=torch.float32)
    ref = a.float() @ w.float().t() + bias
    assert _rel(out, ref) < 2e-5
    out_b = ops.gemm(a, w, bias)
    assert out_b.dtype == torch.bfloat16
    assert _rel(out_b, ref) < 1e-2


@pytest.mark.parametrize("n,k,res", [(2304, 768, False), (768, 3072, True)])
def test_gemm_bench_shapes_full_size(ops, n, k, res):
    """The bench workload's own GEMM shapes (64 images x 1370 tokens = 87 680 rows, 41 tiles per CTA pair):
    every output element against an fp32 product of the same bf16 operands."""
    m = 64 * 1370
    g = _gen(n + k)
    a = _randn((m, k), g).bfloat16()
    w = _randn((n, k), g, 1 / math.sqrt(k)).bfloat16()
    bias = _randn((n,), g)
    ref = a.float() @ w.float().t() + bias
    if res:
        scale, r = _randn((n,), g), _randn((m, n), g)
        out = ops.gemm(a, w, bias, scale=scale, residual=r, out_dtype=torch.float32)
        assert _rel(out, r + scale * ref) < 2e-5
    else:
        assert _rel(ops.gemm(a, w, bias, out_dtype=torch.float32), ref) < 2e-5
        assert _rel(ops.gemm(a, w, bias), ref) < 1e-2


@pytest.mark.parametrize("m", [300, 777, 2049, 4700, 17000])
@pytest.mark.parametrize("act", ["gelu", "relu"])
def test_gemm_act_scale_residual(ops, act, m):
    n, k = 384, 320
    g = _gen(11)
    a = _randn((m, k), g).bfloat16()
    w = _randn((n, k), g, 1 / math.sqrt(k)).bfloat16()
    bias, scale = _randn((n,), g), _randn((n,), g)
    res = _randn((m, n), g)
    code = ops.ACT_GELU_ERF if act == "gelu" else ops.ACT_RELU
    out = ops.gemm(a, w, bias, act=code, scale=scale, residual=res, out_dtype=torch.float32)
    z = a.float() @ w.float().t() + bias
    z = torch.nn.functional.gelu(z) if act == "gelu" else torch.relu(z)
    ref = res + scale * z
    assert _rel(out, ref) < 2e-5


def test_gemm_second_k_segment(ops):
    """LoRA as a second K segment: x.W^T + (x.A^T).(alpha B)^T."""
    m, n, k, r = 900, 768, 768, 8
    g = _gen(5)
    x = _randn((m, k), g).bfloat16()
    w = _randn((n, k), g, 1 / math.sqrt(k)).bfloat16()
    t = torch.zeros((m, 64), dtype=torch.bfloat16, device="cuda")
    t[:, :r] = _randn((m, r), g).bfloat16()
    bw = torch.zeros((n, 64), dtype=torch.bfloat16, device="cuda")
    bw[:, :r] = _randn((n, r), g).bfloat16()
    out = ops.gemm(x, w, None, a2=t, w2=bw, out_dtype=torch.float32)
    ref = x.float() @ w.float().t() + t.float() @ bw.float().t()
    assert _rel(out, ref) < 2e-5


@pytest.mark.parametrize("m", [500, 1111])     # 1-CTA kernel / CTA-pair kernel
def test_gemm_swiglu(ops, m):
    k, hidden = 256, 512
    g = _gen(6)
    x = _randn((m, k), g).bfloat16()
    w_in = _randn((2 * hidden, k), g, 1 / math.sqrt(k))
    b_in = _randn((2 * hidden,), g)
    # interleave gate / linear halves in blocks of 128 rows (see dod.h)
    idx = []
    for blk in range(hidden // 128):
        idx += list(range(blk * 128, blk * 128 + 128))
        idx += list(range(hidden + blk * 128, hidden + blk * 128 + 128))
    idx = torch.tensor(idx, device="cuda")
    out = ops.gemm(x, w_in[idx].bfloat16().contiguous(), b_in[idx].contiguous(), act=ops.ACT_SWIGLU)
    z = x.float() @ w_in.bfloat16().float().t() + b_in
    ref = torch.nn.functional.silu(z[:, :hidden]) * z[:, hidden:]
    assert out.shape == (m, hidden)
    assert _rel(out, ref) < 1e-2


@pytest.mark.parametrize("batch,rows,n,k", [(3, 600, 384, 592), (5, 1369, 768, 592), (2, 513, 1024, 72)])
def test_gemm_per_image_shared_weight_and_residual(ops, batch, rows, n, k):
    """One launch, one GEMM per image, W and the fp32 residual shared (the patch embedding): image i's rows land
    `tokens` rows apart behind a row the GEMM must not touch (the CLS row)."""
    g = _gen(batch * rows + n)
    a = _randn((batch * rows, k), g).bfloat16()
    w = _randn((n, k), g, 1 / math.sqrt(k)).bfloat16()
    bias = _randn((n,), g)
    res = _randn((rows + 1, n), g)
    tokens = rows + 1
    out = torch.full((batch * tokens, n), 7.0, device="cuda")
    ops.gemm_per_image(a, w, bias, res[1:], out[1:], batch, tokens * n)
    ref = (a.float() @ w.float().t() + bias).view(batch, rows, n) + res[1:]
    got = out.view(batch, tokens, n)
    assert torch.equal(got[:, 0], torch.full((batch, n), 7.0, device="cuda"))   # row 0 of every image untouched
    assert _rel(got[:, 1:], ref) < 2e-5


def test_gemm_patch_rows(ops):
    """Patch-embedding row map: GEMM row m -> token row m + m/P + 1, residual row 1 + m%P."""
    b, p, k, d = 3, 16, 592, 384
    g = _gen(8)
    a = _randn((b * p, k), g).bfloat16()
    w = _randn((d, k), g, 1 / math.sqrt(k)).bfloat16()
    bias = _randn((d,), g)
    pos = _randn((p + 1, d), g)
    tokens = torch.full((b * (p + 1), d), -7.0, device="cuda")
    ops.gemm(a, w, bias, residual=pos, out=tokens, patch_rows=p)
    ref = (a.float() @ w.float().t() + bias).view(b, p, d) + pos[1:]
    got = tokens.view(b, p + 1, d)
    assert _rel(got[:, 1:], ref) < 2e-5
    assert (got[:, 0] == -7.0).all()


def test_gemm_fp32_mode_split3(ops):
    """3-term bf16 split on both operands reproduces the fp32 product."""
    m, n, k = 515, 384, 384
    g = _gen(9)
    a, w = _randn((m, k), g), _randn((n, k), g, 1 / math.sqrt(k))
    a6 = ops.split3_bf16(a, k, w_side=False)
    w6 = ops.split3_bf16(w, k, w_side=True)
    out = ops.gemm(a6, w6, None, out_dtype=torch.float32)
    ref = (a.double() @ w.double().t()).float()
    # TMEM accumulation is not a correctly rounded fp32 sum (alignment truncation inside the
    # tensor core), so the split product lands at ~1e-5 of max rather than 2^-22
    assert _rel(out, ref) < 2e-5


# ----------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("m,d,n2,act", [(514, 384, 1152, "none"), (2740, 768, 3072, "gelu"), (1000, 1024, 3072, "none"),
                                       (700, 1536, 4096, "swiglu")])
def test_gemm_folded_layernorm(ops, m, d, n2, act):
    """LayerNorm folded across two GEMMs (include/dod.h): a residual GEMM that also emits bf16(h) and
    per-row partial sums, then a projection of h with gamma folded into zero-sum weight rows and the rstd
    scaling in its epilogue -- against  act(LayerNorm(h) @ W^T + b)  in fp32 torch."""
    g = _gen(m + d + n2)
    k1 = 256
    a = _randn((m, k1), g).bfloat16()
    w1 = _randn((d, k1), g, 1 / math.sqrt(k1)).bfloat16()
    b1 = _randn((d,), g)
    ls = _randn((d,), g)
    x = _randn((m, d), g) + 0.3                      # residual stream with a non-zero row mean
    h_ref = x + ls * (a.float() @ w1.float().t() + b1)
    h16 = torch.empty((m, d), dtype=torch.bfloat16, device="cuda")
    slots = 2 * ((d + 255) // 256)
    stats = torch.full((slots, m, 2), float("nan"), device="cuda")
    h = ops.gemm(a, w1, b1, scale=ls, residual=x, out_dtype=torch.float32, ln_out=(h16, stats))
    assert _rel(h, h_ref) < 2e-5
    assert torch.equal(h16, h.bfloat16())
    assert torch.allclose(stats[:, :, 0].sum(0), h.sum(1), rtol=1e-4, atol=1e-3)
    assert torch.allclose(stats[:, :, 1].sum(0), (h * h).sum(1), rtol=1e-4, atol=1e-3)

    gamma = 1.0 + 0.2 * _randn((d,), g)
    beta = 0.1 * _randn((d,), g)
    w2 = _randn((n2, d), g, 1 / math.sqrt(d))
    b2 = _randn((n2,), g)
    eps = 1e-6
    ln = torch.nn.functional.layer_norm(h, (d,), gamma, beta, eps)
    ref = ln @ w2.t() + b2
    w2f = w2 * gamma[None, :]
    w2f = (w2f - w2f.mean(dim=1, keepdim=True)).bfloat16()      # zero-sum rows: the product drops the row mean
    b2f = b2 + w2 @ beta
    dod_act = ops.ACT_NONE
    if act == "gelu":
        ref = torch.nn.functional.gelu(ref)
        dod_act = ops.ACT_GELU_ERF
    elif act == "swiglu":
        half = n2 // 2
        ref = torch.nn.functional.silu(ref[:, :half]) * ref[:, half:]
        idx = torch.arange(n2, device="cuda").view(2, half // 128, 128).permute(1, 0, 2).reshape(-1)
        w2f, b2f = w2f[idx].contiguous(), b2f[idx].contiguous()
        dod_act = ops.ACT_SWIGLU
    out = ops.gemm(h16, w2f, b2f, act=dod_act, row_scale=ops.ln_rstd(stats, d, eps))
    # same bf16-operand tolerance as the unfolded path (bf16 activations / weights / output)
    assert _rel(out, ref) < 2e-2
    unfolded = ops.gemm(ln.bfloat16(), w2.bfloat16() if act != "swiglu" else w2.bfloat16()[idx].contiguous(),
                        b2 if act != "swiglu" else b2[idx].contiguous(), act=dod_act)
    err_f = (out.float() - ref).abs().mean().item()
    err_u = (unfolded.float() - ref).abs().mean().item()
    assert err_f < 2.0 * err_u + 1e-4, (err_f, err_u)


@pytest.mark.parametrize("mean_ratio,outliers", [(0.0, False), (1.0, False), (10.0, False), (0.0, True), (1.0, True)])
def test_folded_layernorm_stress(ops, mean_ratio, outliers):
    """Folded vs standalone LayerNorm on residual streams unlike the friendly test statistics: rows whose mean is
    `mean_ratio` standard deviations away from zero, and four channels at 100x the magnitude of the rest (the outlier
    channels of real ViT residual streams).  Both bf16 paths are compared with the exact fp64 result; the achieved
    errors go on record (profiles/r02_parity_margins.json).  Bar: the folded form may not be worse than 1.5x the
    standalone form + the growth sqrt(1 + ratio^2) that rounding the un-normalised stream implies -- and the
    monitor (dod_ln_rstd max_mean_ratio) must report the ratio so that a deployment can see it."""
    from helpers import record_margin
    m, d, n2 = 1370, 768, 2304
    g = _gen(int(mean_ratio * 10) + 7 * outliers)
    h = _randn((m, d), g)
    if outliers:
        h[:, [5, 111, 400, 767]] *= 100.0
    std = h.std(dim=1, keepdim=True)
    h = h - h.mean(dim=1, keepdim=True) + mean_ratio * std          # row mean = mean_ratio * row std
    gamma = 1.0 + 0.2 * _randn((d,), g)
    beta = 0.1 * _randn((d,), g)
    w2 = _randn((n2, d), g, 1 / math.sqrt(d))
    b2 = _randn((n2,), g)
    eps = 1e-6
    exact = (torch.nn.functional.layer_norm(h.double(), (d,), gamma.double(), beta.double(), eps) @ w2.double().t()
             + b2.double())
    # producer: h = x + 1 * (a @ w1^T + 0) with a = 0  ->  the GEMM's epilogue just forwards the residual stream
    a0 = torch.zeros((m, 64), dtype=torch.bfloat16, device="cuda")
    w0 = torch.zeros((d, 64), dtype=torch.bfloat16, device="cuda")
    h16 = torch.empty((m, d), dtype=torch.bfloat16, device="cuda")
    stats = torch.empty((2 * ((d + 255) // 256), m, 2), device="cuda")
    hh = ops.gemm(a0, w0, torch.zeros(d, device="cuda"), scale=torch.ones(d, device="cuda"), residual=h,
                  out_dtype=torch.float32, ln_out=(h16, stats))
    assert torch.equal(hh, h)
    w2f = w2 * gamma[None, :]
    w2f = (w2f - w2f.mean(dim=1, keepdim=True)).bfloat16()
    monitor = torch.zeros(1, device="cuda")
    folded = ops.gemm(h16, w2f, b2 + w2 @ beta, row_scale=ops.ln_rstd(stats, d, eps, max_mean_ratio=monitor))
    standalone = ops.gemm(ops.layernorm(h, gamma, beta, eps), w2.bfloat16(), b2)
    scale = exact.abs().max().item()
    err_f = (folded.double() - exact).abs().max().item() / scale
    err_s = (standalone.double() - exact).abs().max().item() / scale
    rms_f = ((folded.double() - exact).pow(2).mean().sqrt() / exact.pow(2).mean().sqrt()).item()
    rms_s = ((standalone.double() - exact).pow(2).mean().sqrt() / exact.pow(2).mean().sqrt()).item()
    tag = f"mean/std={mean_ratio:g}{' +4 outlier channels x100' if outliers else ''}"
    record_margin("ln_fold_stress", f"folded LN max-rel, {tag}", err_f, 2e-2, rms_rel=rms_f)
    record_margin("ln_fold_stress", f"standalone LN max-rel, {tag}", err_s, 2e-2, rms_rel=rms_s)
    want = (h.mean(1).abs() / h.var(1, unbiased=False).add(eps).sqrt()).max().item()
    assert abs(monitor.item() - want) <= 1e-3 * max(1.0, want), (monitor.item(), want)
    growth = math.sqrt(1.0 + mean_ratio ** 2)
    assert rms_f <= 1.5 * growth * rms_s + 1e-5, (rms_f, rms_s)
    if mean_ratio <= 1.0:
        assert err_f < 2e-2 and err_s < 2e-2


@pytest.mark.parametrize("rel_scale", [0.3, 1e-2, 1e-3, 1e-4])
def test_lora_merge_vs_two_segment_stress(ops, rel_scale):
    """bf16 inference packs merge LoRA into the base weight (W + alpha B A rounded once); the two-segment GEMM keeps
    A and B apart like the reference's op order.  With the update at `rel_scale` of |W| both forms are compared with
    the exact fp64 result: the TOTAL output error must be the same (one bf16 rounding of the weights either way);
    the error measured relative to the LoRA contribution alone is recorded too -- it grows as 2^-9 / rel_scale in the
    merged form, which is why training always uses the two-segment form (A and B must feel their gradients)."""
    from dino_detector import _engine
    from helpers import record_margin
    g = _gen(int(-math.log10(rel_scale) * 10))
    m, k, n, r, alpha = 1370, 768, 768, 8, 1.0
    x = _randn((m, k), g).bfloat16()
    w = _randn((n, k), g, 1 / math.sqrt(k))
    bias = _randn((n,), g, 0.05)
    a_l = _randn((r, k), g, 1 / math.sqrt(k))
    # |alpha B A| ~ rel_scale * |W| elementwise (rms)
    b_l = _randn((n, r), g) * (rel_scale / math.sqrt(r))
    dw = alpha * (b_l @ a_l)
    exact = x.double() @ (w.double() + dw.double()).t() + bias.double()
    lora_part = x.double() @ dw.double().t()
    f32 = torch.float32                                 # fp32 outputs: the subject is the weight rounding
    merged = _engine.PackedLinear(w + dw, bias, "bf16")(x, out_dtype=f32).double()
    two_seg = _engine.PackedLinear(w, bias, "bf16", lora=(a_l, alpha * b_l))(x, out_dtype=f32).double()
    base_only = _engine.PackedLinear(w, bias, "bf16")(x, out_dtype=f32).double()
    tot = exact.abs().max().item()
    e_m, e_t = (merged - exact).abs().max().item() / tot, (two_seg - exact).abs().max().item() / tot
    lp = lora_part.pow(2).mean().sqrt().item()
    l_m = ((merged - base_only) - lora_part).pow(2).mean().sqrt().item() / lp
    l_t = ((two_seg - base_only) - lora_part).pow(2).mean().sqrt().item() / lp
    record_margin("lora_merge_stress", f"merged W+aBA total max-rel, |dW|/|W|={rel_scale:g}", e_m, 2e-2, lora_part_rms_rel=l_m)
    record_margin("lora_merge_stress", f"two-segment total max-rel, |dW|/|W|={rel_scale:g}", e_t, 2e-2, lora_part_rms_rel=l_t)
    assert e_m < 2e-2 and e_t < 2e-2
    assert e_m < 1.5 * e_t + 1e-4, (e_m, e_t)          # same total error: one rounding of the weights either way
    assert l_t < 2e-2                                   # the two-segment form resolves the update itself


@pytest.mark.parametrize("d", [256, 384, 768, 1024, 1536])
@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16])
def test_layernorm(ops, d, xdt):
    g = _gen(d)
    x = (_randn((1001, d), g) * 3 + 1).to(xdt)
    gamma, beta = _randn((d,), g), _randn((d,), g)
    y32, y16 = ops.layernorm(x, gamma, beta, 1e-6, out_dtype=torch.float32, also_other=True)
    ref = torch.nn.functional.layer_norm(x.float(), (d,), gamma, beta, 1e-6)
    assert _rel(y32, ref) < 1e-5
    assert _rel(y16, ref) < 1e-2


# ------------------------------------------------------------------- patch embedding
def test_patchify_and_pos_resize(ops):
    g = _gen(3)
    b, h, w, d = 2, 42, 56, 64
    px = torch.rand((b, 3, h, w), generator=g).cuda()
    cls, pos = _randn((d,), g), _randn((1 + 12, d), g)
    tokens = torch.zeros((b * 13, d), device="cuda")
    patches = ops.patchify14(px, 592, cls=cls, pos=pos, tokens=tokens)
    ref = torch.nn.functional.unfold(px, 14, stride=14).transpose(1, 2).reshape(b * 12, 588)
    assert torch.equal(patches[:, :588], ref.bfloat16())
    assert (patches[:, 588:] == 0).all()
    assert torch.allclose(tokens.view(b, 13, d)[:, 0], (cls + pos[0]).expand(b, d))
    # bicubic resize of a 37x37 grid to 16x16 (HF interpolate_pos_encoding, 224px input)
    pe = _randn((1 + 37 * 37, d), g)
    out = ops.pos_resize_bicubic(pe, 37, 16, 16)
    grid = pe[1:].view(1, 37, 37, d).permute(0, 3, 1, 2)
    ref = torch.nn.functional.interpolate(grid, size=(16, 16), mode="bicubic", align_corners=False)
    ref = ref.permute(0, 2, 3, 1).reshape(256, d)
    assert torch.allclose(out[1:], ref, atol=2e-5, rtol=1e-5)
    assert torch.equal(out[0], pe[0])


# ----------------------------------------------------------------------- attention
@pytest.mark.parametrize("kernel", ["128-key tiles, 2 CTAs/SM", "64-key tiles, 4 CTAs/SM"])
@pytest.mark.parametrize("b,s,h", [(1, 128, 1), (2, 257, 6), (1, 1370, 12), (3, 100, 2), (2, 384, 3), (5, 65, 2),
                                   (2, 1, 1)])
def test_fmha(ops, b, s, h, kernel, monkeypatch):
    """Both forward kernels (default: 128-key tiles; DOD_FMHA64=1: 64-key tiles, four CTAs per SM) on every
    shape, log-sum-exp output included."""
    monkeypatch.setenv("DOD_FMHA64", "1" if kernel.startswith("64") else "0")
    g = _gen(s + h)
    d = h * 64
    qkv = _randn((b * s, 3 * d), g).bfloat16()
    out = ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125)
    q, k, v = (qkv.float().view(b, s, 3, h, 64).permute(2, 0, 3, 1, 4))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    ref = ref.permute(0, 2, 1, 3).reshape(b * s, d)
    assert _rel(out, ref) < 2e-2
    assert (out.float() - ref).abs().max().item() < 2e-2
    lse = torch.empty((b, h, s), dtype=torch.float32, device="cuda")
    out2 = ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125, lse=lse)
    assert torch.equal(out, out2)
    want = torch.logsumexp(q @ k.transpose(-1, -2) * 0.125, dim=-1) * 1.4426950408889634       # log2 domain
    assert (lse - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item())


def test_fmha_bench_shape_full_size(ops):
    """The bench workload's attention (64 images x 12 heads x 1370 tokens, 8448 CTAs): every (image, head)
    against torch SDPA on the same bf16 inputs, and a few slices against an fp32 softmax."""
    b, s, h = 64, 1370, 12
    d = h * 64
    g = _gen(5)
    qkv = _randn((b * s, 3 * d), g, 0.7).bfloat16()
    out = ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125).view(b, s, h, 64)
    q, k, v = qkv.view(b, s, 3, h, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3)     # bf16 library path
    assert (out.float() - ref.float()).abs().max().item() < 2e-2
    for bi, hi in [(0, 0), (17, 5), (63, 11)]:
        qf, kf, vf = (t[bi, hi].float() for t in (q, k, v))
        exact = torch.softmax(qf @ kf.t() * 0.125, dim=-1) @ vf
        assert (out[bi, :, hi].float() - exact).abs().max().item() < 1e-2


@pytest.mark.parametrize("b,lq,lk,h,dh", [(3, 100, 257, 4, 64), (2, 100, 1370, 8, 96), (1, 25, 1370, 8, 128),
                                          (2, 50, 300, 6, 192), (64, 100, 1370, 8, 96)])
def test_cross_attention_on_tensor_cores(b, lq, lk, h, dh):
    """Standard-decoder cross-attention (few queries x long memory) as batched tcgen05 GEMMs + row softmax
    (_engine.cross_attention_tc) vs fp32 torch on the same bf16 inputs; K and V are column slices of ONE fused
    projection buffer like in the decoder (stride 2 * hidden * layers)."""
    from dino_detector import _engine
    g = _gen(lq * 3 + lk + dh)
    d = h * dh
    q = _randn((b * lq, d), g).bfloat16()
    kv = _randn((b * lk, 3 * 2 * d), g).bfloat16()               # three layers' K | V side by side
    k, v = kv[:, 2 * d:3 * d], kv[:, 3 * d:4 * d]                 # layer 1
    out = _engine.cross_attention_tc(q, k, v, b, lq, lk, h, dh, 1 / math.sqrt(dh))
    for bi in sorted({0, b // 2, b - 1}):
        qf = q[bi * lq:(bi + 1) * lq].float().view(lq, h, dh).transpose(0, 1)
        kf = k[bi * lk:(bi + 1) * lk].float().view(lk, h, dh).transpose(0, 1)
        vf = v[bi * lk:(bi + 1) * lk].float().view(lk, h, dh).transpose(0, 1)
        ref = (torch.softmax(qf @ kf.transpose(1, 2) / math.sqrt(dh), dim=-1) @ vf).transpose(0, 1).reshape(lq, d)
        got = out[bi * lq:(bi + 1) * lq].float()
        assert (got - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item()), (bi,)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("b,lq,lk,h,dh", [(2, 50, 50, 8, 96), (3, 100, 257, 4, 64), (1, 25, 1370, 8, 96),
                                          (2, 17, 33, 4, 192), (1, 100, 100, 8, 96), (2, 128, 128, 4, 64)])
def test_mha_small(ops, dt, b, lq, lk, h, dh):
    g = _gen(lq + lk)
    d = h * dh
    q, k, v = (_randn((b * lq, d), g).to(dt), _randn((b * lk, d), g).to(dt), _randn((b * lk, d), g).to(dt))
    out = ops.mha_small(q, k, v, b, lq, lk, h, dh, 1 / math.sqrt(dh))
    qf = q.float().view(b, lq, h, dh).transpose(1, 2)
    kf = k.float().view(b, lk, h, dh).transpose(1, 2)
    vf = v.float().view(b, lk, h, dh).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(b * lq, d)
    assert _rel(out, ref) < (1e-2 if dt == torch.bfloat16 else 1e-5)


def _deform_ref(value, ref_pts, offs, logits, b, q, h, p, dh, gh, gw):
    """Vectorised restatement of reference deformable_attention.py:100-178."""
    offs = offs.view(b, q, h, p, 2)
    wts = logits.view(b, q, h, p).softmax(-1)
    loc = (ref_pts.view(b, q, 1, 1, 2) + offs).clamp(0, 1)
    sx, sy = loc[..., 0] * (gw - 1), loc[..., 1] * (gh - 1)
    x0, y0 = torch.floor(sx).long(), torch.floor(sy).long()
    x1, y1 = (x0 + 1).clamp(0, gw - 1), (y0 + 1).clamp(0, gh - 1)
    x0, y0 = x0.clamp(0, gw - 1), y0.clamp(0, gh - 1)
    wx1, wy1 = sx - x0.float(), sy - y0.float()
    wx0, wy0 = 1 - wx1, 1 - wy1
    vh = value.float().view(b, gh * gw, h, dh)
    bi = torch.arange(b, device=value.device).view(b, 1, 1, 1)
    hi = torch.arange(h, device=value.device).view(1, 1, h, 1)
    def gat(yy, xx):
        return vh[bi, yy * gw + xx, hi]
    samp = (gat(y0, x0) * (wx0 * wy0)[..., None] + gat(y1, x0) * (wx0 * wy1)[..., None]
            + gat(y0, x1) * (wx1 * wy0)[..., None] + gat(y1, x1) * (wx1 * wy1)[..., None])
    return (samp * wts[..., None]).sum(3).reshape(b * q, h * dh)


@pytest.mark.parametrize("b,l,h,dh", [(64, 50, 8, 96), (3, 100, 4, 64), (2, 128, 8, 32), (5, 7, 2, 192)])
def test_query_self_attention_bf16_smem_form_is_bit_identical(ops, b, l, h, dh):
    """mha_tiny_kernel keeps K / V as bf16 in shared memory when the rows are 16-byte aligned (four CTAs per SM);
    rows that are not (ld % 8 != 0) take the fp32-staged form: same arithmetic, same bits, and both match fp32 torch."""
    g = _gen(b + l + dh)
    d = h * dh
    vals = _randn((b * l, 3 * d), g).bfloat16()
    outs = []
    for pad in (8, 4):                                  # ld = 3d + 8: vector form; 3d + 4: scalar form
        buf = torch.zeros((b * l, 3 * d + pad), dtype=torch.bfloat16, device="cuda")
        buf[:, :3 * d] = vals
        outs.append(ops.mha_small(buf[:, :d], buf[:, d:2 * d], buf[:, 2 * d:3 * d], b, l, l, h, dh, 1 / math.sqrt(dh)))
    assert torch.equal(outs[0], outs[1])
    q, k, v = (vals.float().view(b, l, 3, h, dh).permute(2, 0, 3, 1, 4))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(b * l, d)
    assert _rel(outs[0], ref) < 1e-2


@pytest.mark.parametrize("gh,gw", [(10, 137), (1, 257), (16, 16)])
def test_deform_sample(ops, gh, gw):
    g = _gen(gh)
    b, q, h, p, dh = 2, 50, 8, 2, 96
    value = _randn((b * gh * gw, h * dh), g)
    raw = _randn((b * q, 56), g)  # [offsets 32 | logits 16 | ref logits 2 | pad]
    out = ops.deform_sample(value, raw[:, 48:50], raw[:, :32], raw[:, 32:48], b, q, h, p, dh, gh, gw,
                            ref_is_logit=True, out_dtype=torch.float32)
    ref = _deform_ref(value, raw[:, 48:50].sigmoid(), raw[:, :32], raw[:, 32:48], b, q, h, p, dh, gh, gw)
    assert torch.allclose(out, ref, atol=1e-5, rtol=1e-5)
    out16 = ops.deform_sample(value.bfloat16(), raw[:, 48:50], raw[:, :32], raw[:, 32:48], b, q, h, p,
                              dh, gh, gw)
    assert _rel(out16, ref) < 2e-2


@pytest.mark.parametrize("b,q,h,p,dh,gh,gw", [(2, 50, 8, 2, 96, 10, 137), (3, 100, 4, 4, 64, 1, 257), (64, 50, 8, 2, 96, 10, 137),
                                             (1, 7, 8, 2, 32, 16, 16)])
def test_deform_sample_vectorised_is_bit_identical(ops, b, q, h, p, dh, gh, gw, monkeypatch):
    """The 8-channels-per-thread kernel (bf16 values) against the thread-per-channel kernel: same bits, both
    output types, value rows taken as a column slice of a wider buffer (ldv > d_model) like in the decoder."""
    g = _gen(b + q + dh)
    d = h * dh
    wide = _randn((b * gh * gw, d + 64), g).bfloat16()
    value = wide[:, 32:32 + d]
    raw = _randn((b * q, 3 * h * p + 8), g)
    args = (value, raw[:, 3 * h * p:3 * h * p + 2], raw[:, :2 * h * p], raw[:, 2 * h * p:3 * h * p], b, q, h, p, dh, gh, gw)
    for odt in (torch.bfloat16, torch.float32):
        monkeypatch.setenv("DOD_DEFORM_VEC", "0")
        want = ops.deform_sample(*args, ref_is_logit=True, out_dtype=odt)
        monkeypatch.setenv("DOD_DEFORM_VEC", "1")
        got = ops.deform_sample(*args, ref_is_logit=True, out_dtype=odt)
        assert torch.equal(got, want)


def test_row_utils(ops):
    g = _gen(1)
    x = _randn((77, 96), g)
    assert torch.equal(ops.rowcopy(x, 91), x[:, :91])
    assert torch.allclose(ops.rowcopy(x, 4, sigmoid=True), x[:, :4].sigmoid(), atol=1e-6)
    src = _randn((50, 64), g)
    f, h = ops.broadcast_rows(src, 3)
    assert torch.equal(f.view(3, 50, 64), src.expand(3, 50, 64))
    assert torch.equal(h, f.bfloat16())
    c = ops.cast_pad_bf16(x, 128, scale=2.0, dst_rows=80)
    assert torch.equal(c[:77, :96], (2 * x).bfloat16()) and (c[:, 96:] == 0).all() and (c[77:] == 0).all()


# ------------------------------------------------------------------------- matcher
def _targets(bs, g, max_n=50, classes=91):
    labels, boxes, offs = [], [], [0]
    for _ in range(bs):
        n = int(torch.randint(0, max_n + 1, (1,), generator=g))
        labels.append(torch.randint(0, classes, (n,), generator=g))
        cxcy = torch.rand((n, 2), generator=g) * 0.6 + 0.2
        wh = torch.rand((n, 2), generator=g) * 0.3 + 0.02
        boxes.append(torch.cat([cxcy, wh], 1))
        offs.append(offs[-1] + n)
    return labels, boxes, offs


@pytest.mark.parametrize("image0", [True, False])
def test_match_cost_and_lsap_vs_scipy(ops, image0):
    from scipy.optimize import linear_sum_assignment
    g = _gen(0)
    bs, q, c = 64, 100, 91
    logits = torch.randn((bs, q, c), generator=g).cuda()
    pboxes = (torch.rand((bs, q, 4), generator=g) * 0.5 + 0.25).cuda()
    labels, boxes, offs = _targets(bs, g)
    max_t = max(len(l) for l in labels)
    cost = ops.match_cost(logits, pboxes, torch.cat(labels).cuda(), torch.cat(boxes).cuda(),
                          torch.tensor(offs, dtype=torch.int32).cuda(), max_t, w_class=1.0, w_bbox=5.0,
                          w_giou=2.0, alpha=0.25, gamma=2.0, use_image0_rows=image0)
    oq, ot, status = ops.lsap(cost, torch.tensor(offs, dtype=torch.int32).cuda(), max_t)
    cost_h, oq, ot, status = cost.cpu(), oq.cpu(), ot.cpu(), status.cpu()
    assert (status == 0).all()
    for b in range(bs):
        n = len(labels[b])
        # (i) cost vs the reference formula (matching.py:80-98) in torch fp32 on the CPU
        src = 0 if image0 else b
        p = logits[src].cpu().sigmoid()
        neg = 0.75 * (p ** 2.0) * (-(1 - p + 1e-8).log())
        pos = 0.25 * ((1 - p) ** 2.0) * (-(p + 1e-8).log())
        cc = pos[:, labels[b]] - neg[:, labels[b]]
        pb = pboxes[src].cpu()
        cb = torch.cdist(pb, boxes[b], p=1) if n else torch.zeros((q, 0))
        def xyxy(x):
            return torch.stack([x[:, 0] - 0.5 * x[:, 2], x[:, 1] - 0.5 * x[:, 3],
                                x[:, 0] + 0.5 * x[:, 2], x[:, 1] + 0.5 * x[:, 3]], -1)
        b1, b2 = xyxy(pb), xyxy(boxes[b])
        a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
        a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
        wh = (torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])).clamp(min=0)
        inter = wh[..., 0] * wh[..., 1]
        union = a1[:, None] + a2 - inter
        whe = (torch.max(b1[:, None, 2:], b2[:, 2:]) - torch.min(b1[:, None, :2], b2[:, :2])).clamp(min=0)
        ae = whe[..., 0] * whe[..., 1]
        giou = inter / union - (ae - union) / ae
        ref = 1.0 * cc + 5.0 * cb + 2.0 * (-giou)
        assert torch.allclose(cost_h[b, :, :n], ref, atol=2e-5, rtol=1e-5)
        # (ii) assignment bit-identical to scipy on OUR fp32 cost matrix
        ri, ci = linear_sum_assignment(cost_h[b, :, :n].numpy())
        k = min(q, n)
        assert np.array_equal(oq[b, :k].numpy(), ri) and np.array_equal(ot[b, :k].numpy(), ci)


def test_lsap_ties_and_shapes(ops):
    """Adversarial assignment problems: constant / integer-tie / duplicate-column costs,
    n = 0, n = Q and n > Q (no transpose)."""
    from scipy.optimize import linear_sum_assignment
    g = _gen(4)
    q = 20
    cases = []
    for n in [0, 1, 5, 19, 20, 21, 37]:
        cases.append(torch.zeros((q, n)))
        cases.append(torch.randint(0, 3, (q, n), generator=g).float())
        cases.append(torch.rand((q, n), generator=g))
        if n >= 2:
            dup = torch.rand((q, n), generator=g)
            dup[:, 1] = dup[:, 0]
            dup[3] = dup[2]
            cases.append(dup)
    max_t = max(c.shape[1] for c in cases)
    bs = len(cases)
    cost = torch.zeros((bs, q, max_t))
    offs = [0]
    for i, c in enumerate(cases):
        cost[i, :, :c.shape[1]] = c
        offs.append(offs[-1] + c.shape[1])
    oq, ot, status = ops.lsap(cost.cuda(), torch.tensor(offs, dtype=torch.int32).cuda(), max_t)
    oq, ot, status = oq.cpu().numpy(), ot.cpu().numpy(), status.cpu().numpy()
    assert (status == 0).all()
    for i, c in enumerate(cases):
        ri, ci = linear_sum_assignment(c.numpy())
        k = min(q, c.shape[1])
        assert np.array_equal(oq[i, :k], ri), (i, c.shape)
        assert np.array_equal(ot[i, :k], ci), (i, c.shape)
    # NaN cost -> status 1 (scipy raises ValueError)
    bad = torch.rand((1, q, 4))
    bad[0, 3, 2] = float("nan")
    _, _, st = ops.lsap(bad.cuda(), torch.tensor([0, 4], dtype=torch.int32).cuda(), 4)
    assert st.cpu().item() == 1


def test_add_layernorm(ops):
    """y = LayerNorm(x + r) (decoder post-norm helper of the C ABI)."""
    from dino_detector import _dod
    g = _gen(77)
    rows, d = 301, 256
    x, r = _randn((rows, d), g), _randn((rows, d), g).bfloat16()
    gamma, beta = _randn((d,), g), _randn((d,), g)
    y = torch.empty((rows, d), device="cuda")
    y16 = torch.empty((rows, d), device="cuda", dtype=torch.bfloat16)
    _dod.call("dod_add_layernorm", torch.cuda.current_stream().cuda_stream, x=x, r=r, r_dtype=0, gamma=gamma,
              beta=beta, y=y, y_bf16=y16, rows=rows, d=d, eps=1e-5)
    ref = torch.nn.functional.layer_norm(x + r.float(), (d,), gamma, beta, 1e-5)
    assert _rel(y, ref) < 1e-5 and _rel(y16, ref) < 1e-2
