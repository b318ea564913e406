"""Tensor-level wrappers over the libdod C ABI.

PyTorch is used for device memory and streams only: every function here
validates its arguments, allocates the output with torch.empty on the caller's
device and launches the hand-written sm_100a kernel on torch's current stream.
Nothing in this file computes with torch ops and nothing falls back to them.
"""
from __future__ import annotations

import torch

from . import _dod
from ._dod import (ACT_GELU_ERF, ACT_NONE, ACT_RELU, ACT_SWIGLU, DOD_BF16, DOD_F32, DodError)

_DT = {torch.bfloat16: DOD_BF16, torch.float32: DOD_F32}

# Optional per-launch timing (bench.py): when a list is installed, the tensor-core ops bracket
# their launch with CUDA events on the launching stream and append (kind, flops, start, end).
_profile = None


def profile_begin():
    global _profile
    _profile = []


def profile_end():
    """-> {kind: (launches, total_flops, total_ms)}; call after torch.cuda.synchronize()."""
    global _profile
    rec, _profile = _profile or [], None
    out = {}
    for kind, flops, e0, e1 in rec:
        n, f, t = out.get(kind, (0, 0.0, 0.0))
        out[kind] = (n + 1, f + flops, t + e0.elapsed_time(e1))
    return out


class _Timed:
    def __init__(self, kind, flops):
        self.kind, self.flops = kind, flops

    def __enter__(self):
        if _profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _profile is not None and exc[0] is None:
            self.e1.record()
            _profile.append((self.kind, self.flops, self.e0, self.e1))


def _stream(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise DodError("libdod ops need CUDA tensors (no CPU fallback exists)")
    cur = torch.cuda.current_device()
    idx = t.device.index if t.device.index is not None else cur
    if idx != cur:
        # the C ABI launches in the calling thread's current CUDA context: a tensor of another device would run
        # the kernel on the wrong GPU, unsynchronised with its stream (one process per GPU is the model here)
        raise DodError(f"tensor lives on cuda:{idx} but the current CUDA device is cuda:{cur}: call "
                       f"torch.cuda.set_device({idx}) (or use `with torch.cuda.device({idx}):`) before libdod ops")
    _dod.check_device(idx)
    return torch.cuda.current_stream(t.device).cuda_stream


def _rowmajor(t: torch.Tensor, name: str):
    if t.dim() != 2 or t.stride(1) != 1:
        raise DodError(f"{name}: expected a 2-D tensor with unit inner stride, got {tuple(t.shape)} "
                       f"strides {t.stride()}")
    return t.stride(0)


def gemm(a, w, bias=None, *, act=ACT_NONE, scale=None, residual=None, out=None,
         out_dtype=torch.bfloat16, a2=None, w2=None, patch_rows=0, out_rows=None, a_trans=False, w_trans=False,
         ln_out=None, row_scale=None):
    """out = residual + scale * act(a @ w.T (+ a2 @ w2.T) + bias); a, w (a2, w2) bf16.
    a_trans: `a` is the stored transpose [K, M] (out = a.T @ ...); w_trans: `w` is stored [K, N]
    (out = ... @ w) -- the kernel reads them as MN-major operands, no transposed copy is made.
    Folded LayerNorm (include/dod.h): ln_out=(h_bf16 [M, N], row_stats [2*ceil(N/256), M, 2]) makes a residual
    GEMM also emit the bf16 copy of its output and per-row partial sums; row_scale (f32 [M], = ln_rstd(row_stats))
    makes the epilogue multiply the accumulator of row m by row_scale[m] before bias / activation (the weights
    then carry gamma and have zero-sum rows, the bias carries W beta)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    lda, ldw = _rowmajor(a, "a"), _rowmajor(w, "w")
    k, m = (a.shape if a_trans else a.shape[::-1])
    k_w, n = (w.shape if w_trans else w.shape[::-1])
    assert k_w == k, (a.shape, w.shape, a_trans, w_trans)
    n_out = n // 2 if act == ACT_SWIGLU else n
    if out is None:
        out = torch.empty((out_rows if out_rows is not None else m, n_out), dtype=out_dtype,
                          device=a.device)
    ldo = _rowmajor(out, "out")
    kw = dict(a=a, w=w, m=m, n=n, k=k, lda=lda, ldw=ldw, bias=bias, act=act, scale=scale,
              residual=residual, ldr=_rowmajor(residual, "residual") if residual is not None else 0,
              out=out, ldo=ldo, out_dtype=_DT[out.dtype], patch_rows=patch_rows,
              a_trans=int(a_trans), w_trans=int(w_trans))
    if a2 is not None:
        assert a2.dtype == torch.bfloat16 and w2.dtype == torch.bfloat16
        assert a2.shape[0] == m and w2.shape[0] == n and a2.shape[1] == w2.shape[1]
        kw.update(a2=a2, w2=w2, k2=a2.shape[1], lda2=_rowmajor(a2, "a2"), ldw2=_rowmajor(w2, "w2"))
    if ln_out is not None:
        h16, stats = ln_out
        assert h16.dtype == torch.bfloat16 and stats.dtype == torch.float32 and h16.shape[0] >= m
        assert stats.is_contiguous() and tuple(stats.shape) == (2 * ((n + 255) // 256), m, 2)
        kw.update(out_bf16=h16, ldo_bf16=_rowmajor(h16, "ln_out"), row_stats_out=stats)
    if row_scale is not None:
        assert row_scale.dtype == torch.float32 and row_scale.is_contiguous() and row_scale.numel() >= m
        assert bias is not None
        kw.update(row_scale=row_scale)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= n
    if scale is not None:
        assert scale.dtype == torch.float32 and scale.numel() >= n
    if residual is not None:
        assert residual.dtype == torch.float32
    k_tot = k + (a2.shape[1] if a2 is not None else 0)
    with _Timed("gemm", 2.0 * m * n * k_tot):
        _dod.call("dod_gemm_bf16", _stream(a), **kw)
    return out


def ln_rstd(row_stats, dim, eps, out=None, max_mean_ratio=None):
    """rstd [M] from the partial sums [slots, M, 2] a residual GEMM wrote through ln_out (folded LayerNorm).
    max_mean_ratio: optional f32 [1] device tensor that keeps the running maximum of |mean| * rstd over all rows."""
    slots, m, _ = row_stats.shape
    assert row_stats.dtype == torch.float32 and row_stats.is_contiguous()
    if out is None:
        out = torch.empty(m, dtype=torch.float32, device=row_stats.device)
    if max_mean_ratio is not None:
        assert max_mean_ratio.dtype == torch.float32 and max_mean_ratio.numel() == 1
    _dod.call("dod_ln_rstd", _stream(row_stats), row_stats=row_stats, rstd=out, slots=slots, rows=m, dim=int(dim),
              eps=float(eps), max_mean_ratio=max_mean_ratio)
    return out


def layernorm(x, gamma, beta, eps, *, out_dtype=torch.bfloat16, also_other=False, out=None):
    """LayerNorm over the last dim of a 2-D tensor (f32 or bf16 in).  Returns y, or
    (y, y_other_dtype) when also_other."""
    ldx = _rowmajor(x, "x")
    rows, d = x.shape
    y = out if out is not None else torch.empty((rows, d), dtype=out_dtype, device=x.device)
    y2 = None
    if also_other:
        other = torch.float32 if y.dtype == torch.bfloat16 else torch.bfloat16
        y2 = torch.empty((rows, d), dtype=other, device=x.device)
    nbytes = rows * d * (x.element_size() + y.element_size() + (y2.element_size() if y2 is not None else 0))
    with _Timed("layernorm", float(nbytes)):
        _dod.call("dod_layernorm", _stream(x), x=x, x_dtype=_DT[x.dtype], gamma=gamma, beta=beta, y=y,
                  y_dtype=_DT[y.dtype], y2=y2, rows=rows, d=d, ldx=ldx, ldy=_rowmajor(y, "y"), eps=eps)
    return (y, y2) if also_other else y


def patchify14(pixels, kpad, *, cls=None, pos=None, tokens=None):
    """im2col of 14x14 patches -> bf16 [B*P, kpad]; optionally writes the CLS rows of tokens.
    pixels: f32 [B, 3, H, W] in [0, 1], or uint8 [B, H, W, 3] (0..255, divided by 255 in the kernel)."""
    assert pixels.is_contiguous() and pixels.dim() == 4
    if pixels.dtype == torch.uint8:
        b, h, w, c = pixels.shape
        fmt = 1
    else:
        assert pixels.dtype == torch.float32
        b, c, h, w = pixels.shape
        fmt = 0
    if c != 3:
        # same error as HF Dinov2PatchEmbeddings.forward (modeling_dinov2.py:143-147)
        raise ValueError(
            "Make sure that the channel dimension of the pixel values match with the one set in the "
            f"configuration. Expected 3 but got {c}.")
    p = (h // 14) * (w // 14)
    patches = torch.empty((b * p, kpad), dtype=torch.bfloat16, device=pixels.device)
    _dod.call("dod_patchify14", _stream(pixels), pixels=pixels, patches=patches, batch=b, height=h,
              width=w, kpad=kpad, cls=cls, pos=pos, tokens=tokens,
              d=tokens.shape[-1] if tokens is not None else 0, pixel_format=fmt)
    return patches


def pos_resize_bicubic(pos, g0, gh, gw):
    """pos: f32 [1 + g0*g0, D] -> f32 [1 + gh*gw, D]."""
    assert pos.dtype == torch.float32 and pos.is_contiguous()
    d = pos.shape[1]
    out = torch.empty((1 + gh * gw, d), dtype=torch.float32, device=pos.device)
    _dod.call("dod_pos_resize_bicubic", _stream(pos), src=pos, dst=out, g0=g0, gh=gh, gw=gw, d=d)
    return out


def fmha(qkv, batch, seq, heads, *, q_off, k_off, v_off, scale, out=None, lse=None):
    """Self-attention over a fused projection buffer qkv [B*S, ld] (bf16), head dim 64.
    lse: optional f32 [batch, heads, seq] that receives the per-row log-sum-exp (for fmha_bwd)."""
    assert qkv.dtype == torch.bfloat16
    ld = _rowmajor(qkv, "qkv")
    assert qkv.shape[0] == batch * seq
    if out is None:
        out = torch.empty((batch * seq, heads * 64), dtype=torch.bfloat16, device=qkv.device)
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == batch * heads * seq
    with _Timed("fmha", 4.0 * batch * heads * seq * seq * 64):
        _dod.call("dod_fmha_fwd", _stream(qkv), qkv=qkv, ctx=out, batch=batch, seq=seq, heads=heads,
                  ld=ld, ldo=_rowmajor(out, "out"), q_off=q_off, k_off=k_off, v_off=v_off, scale=scale, lse=lse)
    return out


def fmha_bwd(qkv, ctx, dctx, lse, dqkv, batch, seq, heads, *, q_off, k_off, v_off, scale):
    """Backward of fmha(): writes bf16 dQ / dK / dV into the head slices of dqkv [B*S, ld] at the same
    column offsets as q / k / v in qkv.  ctx is the forward output, lse the forward's log-sum-exp."""
    assert qkv.dtype == ctx.dtype == dctx.dtype == dqkv.dtype == torch.bfloat16
    assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == batch * heads * seq
    hd = heads * 64
    dsum = torch.empty_like(lse)
    dq_acc = torch.zeros((batch * seq, hd), dtype=torch.float32, device=qkv.device)
    with _Timed("fmha_bwd", 10.0 * batch * heads * seq * seq * 64):
        _dod.call("dod_fmha_bwd", _stream(qkv), qkv=qkv, ctx=ctx, dctx=dctx, lse=lse, dsum=dsum, dq_acc=dq_acc,
                  dqkv=dqkv, batch=batch, seq=seq, heads=heads, ld=_rowmajor(qkv, "qkv"), ldo=_rowmajor(ctx, "ctx"),
                  lddo=_rowmajor(dctx, "dctx"), ld_dq=hd, ld_dqkv=_rowmajor(dqkv, "dqkv"), q_off=q_off, k_off=k_off,
                  v_off=v_off, scale=scale)
    eltwise(ELT_CAST, dq_acc, out=dqkv[:, q_off:q_off + hd])
    return dqkv


def mha_small(q, k, v, batch, lq, lk, heads, head_dim, scale, out=None):
    """Generic-head-dim attention with fp32 math; q/k/v are 2-D row views (bf16 or f32)."""
    assert q.dtype == k.dtype == v.dtype
    if out is None:
        out = torch.empty((batch * lq, heads * head_dim), dtype=q.dtype, device=q.device)
    _dod.call("dod_mha_small", _stream(q), q=q, k=k, v=v, out=out, batch=batch, lq=lq, lk=lk,
              heads=heads, head_dim=head_dim, ldq=_rowmajor(q, "q"), ldk=_rowmajor(k, "k"),
              ldv=_rowmajor(v, "v"), ldo=_rowmajor(out, "out"), scale=scale, dtype=_DT[q.dtype])
    return out


def deform_sample(value, ref, offs, logits, batch, queries, heads, points, head_dim, grid_h, grid_w,
                  *, ref_is_logit=True, out_dtype=torch.bfloat16):
    out = torch.empty((batch * queries, heads * head_dim), dtype=out_dtype, device=value.device)
    _dod.call("dod_deform_sample", _stream(value), value=value, ref=ref, offs=offs, logits=logits,
              out=out, batch=batch, queries=queries, heads=heads, points=points, head_dim=head_dim,
              grid_h=grid_h, grid_w=grid_w, ldv=_rowmajor(value, "value"), ldref=_rowmajor(ref, "ref"),
              ldoffs=_rowmajor(offs, "offs"), ldlog=_rowmajor(logits, "logits"),
              ldo=_rowmajor(out, "out"), value_dtype=_DT[value.dtype], out_dtype=_DT[out.dtype],
              ref_is_logit=int(ref_is_logit))
    return out


def rowcopy(x, n, *, sigmoid=False):
    """out[r, :n] = act(x[r, :n]) as a contiguous f32 tensor."""
    assert x.dtype == torch.float32
    rows = x.shape[0]
    out = torch.empty((rows, n), dtype=torch.float32, device=x.device)
    _dod.call("dod_rowcopy", _stream(x), **{"in": x, "out": out, "rows": rows, "n": n,
                                          "ld_in": _rowmajor(x, "x"), "ld_out": n,
                                          "act": int(sigmoid)})
    return out


def broadcast_rows(src, batch, *, want_f32=True, want_bf16=True):
    assert src.dtype == torch.float32 and src.is_contiguous()
    rows, d = src.shape
    of = torch.empty((batch * rows, d), dtype=torch.float32, device=src.device) if want_f32 else None
    ob = torch.empty((batch * rows, d), dtype=torch.bfloat16, device=src.device) if want_bf16 else None
    _dod.call("dod_broadcast_rows", _stream(src), src=src, out=of, out_bf16=ob, batch=batch,
              rows=rows, d=d)
    return of, ob


def cast_pad_bf16(src, dst_cols=None, *, scale=1.0, dst_rows=None):
    """f32 [rows, cols] -> bf16 [dst_rows >= rows, dst_cols >= cols], zero padded."""
    assert src.dtype == torch.float32
    rows, cols = src.shape
    dst_cols = dst_cols or cols
    dst_rows = dst_rows or rows
    if dst_rows > rows:
        dst = torch.zeros((dst_rows, dst_cols), dtype=torch.bfloat16, device=src.device)
    else:
        dst = torch.empty((dst_rows, dst_cols), dtype=torch.bfloat16, device=src.device)
    _dod.call("dod_cast_pad_bf16", _stream(src), src=src, dst=dst, rows=rows, cols=cols,
              ld_src=_rowmajor(src, "src"), ld_dst=dst_cols, dst_cols=dst_cols, scale=scale)
    return dst


def split3_bf16(src, kseg, *, w_side, dst_rows=None):
    """fp32 mode operand: f32 [rows, cols] -> bf16 [rows, 6*kseg] (see dod.h)."""
    assert src.dtype == torch.float32
    rows, cols = src.shape
    dst_rows = dst_rows or rows
    alloc = torch.zeros if dst_rows > rows else torch.empty
    dst = alloc((dst_rows, 6 * kseg), dtype=torch.bfloat16, device=src.device)
    _dod.call("dod_split3_bf16", _stream(src), src=src, dst=dst, rows=rows, cols=cols,
              ld_src=_rowmajor(src, "src"), ld_dst=6 * kseg, kseg=kseg, w_side=int(w_side))
    return dst


def match_cost(logits, boxes, tgt_labels, tgt_boxes, tgt_offsets, max_t, *, w_class, w_bbox, w_giou,
               alpha, gamma, use_image0_rows):
    b, q, c = logits.shape
    assert logits.dtype == torch.float32 and logits.is_contiguous()
    assert boxes.dtype == torch.float32 and boxes.is_contiguous()
    cost = torch.empty((b, q, max(max_t, 1)), dtype=torch.float32, device=logits.device)
    if max_t > 0:
        assert tgt_labels.dtype == torch.int64 and tgt_boxes.dtype == torch.float32
        assert tgt_offsets.dtype == torch.int32
        _dod.call("dod_match_cost", _stream(logits), logits=logits, boxes=boxes,
                  tgt_labels=tgt_labels, tgt_boxes=tgt_boxes, tgt_offsets=tgt_offsets, cost=cost,
                  batch=b, queries=q, classes=c, max_t=max_t, w_class=w_class, w_bbox=w_bbox,
                  w_giou=w_giou, alpha=alpha, gamma=gamma, use_image0_rows=int(use_image0_rows))
    return cost


def lsap(cost, tgt_offsets, max_t):
    """-> (out_q [B, K], out_t [B, K], status [B]) int32 on the device, K = min(Q, max_t)."""
    b, q, ld = cost.shape
    assert cost.dtype == torch.float32 and cost.is_contiguous() and ld >= max_t
    k = max(min(q, max_t), 1)
    # -1 = "no pair": images the solver rejects (status 1) leave their rows untouched
    out_q = torch.full((b, k), -1, dtype=torch.int32, device=cost.device)
    out_t = torch.full((b, k), -1, dtype=torch.int32, device=cost.device)
    status = torch.empty((b,), dtype=torch.int32, device=cost.device)
    _dod.call("dod_lsap_jv", _stream(cost), cost=cost, tgt_offsets=tgt_offsets, out_q=out_q,
              out_t=out_t, status=status, batch=b, queries=q, max_t=ld if max_t > 0 else 0, max_k=k)
    return out_q, out_t, status


# ---------------------------------------------------------------- backward / training helpers
ELT_CAST, ELT_SCALE_COLS, ELT_ADD, ELT_GELU_FWD, ELT_GELU_BWD, ELT_RELU_BWD, ELT_SIGMOID_BWD, \
    ELT_SWIGLU_FWD, ELT_SWIGLU_BWD, ELT_DROPOUT, ELT_AXPBY = range(11)


# Device-resident dropout seed counter (runtime.GraphedTrainStep): when set, the dropout / attention
# dropout kernels add (*counter << 44) to their launch-time seed, so a replayed CUDA graph draws a new
# mask every step although its launch arguments are frozen.
_seed_ptr = None


def set_device_seed(counter):
    """counter: int64 CUDA tensor [>= 1] (element 0 is used) or None to return to host-side seeds."""
    global _seed_ptr
    if counter is not None:
        assert counter.dtype == torch.int64 and counter.is_cuda
    _seed_ptr = counter


def device_seed():
    return _seed_ptr


def counter_add(counters, delta=1):
    """counters[:] += delta on the device (int64, <= 32 elements)."""
    assert counters.dtype == torch.int64 and counters.is_contiguous()
    _dod.call("dod_counter_add", _stream(counters), counters=counters, n=counters.numel(), delta=delta)


def gemm_per_image(a, w, bias, residual, out, batch, out_batch_stride):
    """batch GEMMs  out_i = residual + a_i @ w.T + bias  that share w and the fp32 residual: a [batch * rows, K]
    (image i = rows i * rows ...), out = 2-D f32 view starting at image 0's first output row, image i's rows start
    out_batch_stride elements later.  rows >= 512 (CTA-pair kernel); rows past `rows` of an image are never written."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and out.dtype == torch.float32
    assert residual.dtype == torch.float32 and a.shape[0] % batch == 0
    rows, k = a.shape[0] // batch, a.shape[1]
    n = w.shape[0]
    lda = _rowmajor(a, "a")
    with _Timed("gemm", 2.0 * batch * rows * n * k):
        _dod.call("dod_gemm_bf16", _stream(a), a=a, w=w, m=rows, n=n, k=k, lda=lda, ldw=_rowmajor(w, "w"), bias=bias,
                  residual=residual, ldr=_rowmajor(residual, "residual"), out=out, ldo=_rowmajor(out, "out"),
                  out_dtype=_DT[out.dtype], batch=batch, batch_stride_a=rows * lda, batch_stride_w=0,
                  batch_stride_out=out_batch_stride)
    return out


def gemm_batched(a, w, out, *, bias=None, act=ACT_NONE, a_trans=False, w_trans=False):
    """out[..] = act(a[..] @ w[..].T + bias) over one or two leading batch dims:
    a [B, M, K] / [B, H, M, K], w [B, N, K] / [B, H, N, K], out [B, M, N] / [B, H, M, N]; strided views
    with unit inner stride and row / batch strides that are multiples of 8 elements (e.g. head slices
    of a fused qkv buffer viewed as [B, H, L, dh]).  a_trans / w_trans: the operand is given as its
    stored transpose ([.., K, M] / [.., K, N]) and read MN-major by the kernel."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.dim() == w.dim() == out.dim() and a.dim() in (3, 4)
    assert a.stride(-1) == 1 and w.stride(-1) == 1 and out.stride(-1) == 1
    if a.dim() == 3:
        a, w, out = a.unsqueeze(1), w.unsqueeze(1), out.unsqueeze(1)
    nb, nh = a.shape[:2]
    k, m = (a.shape[2:] if a_trans else a.shape[:1:-1])
    k_w, n = (w.shape[2:] if w_trans else w.shape[:1:-1])
    assert k_w == k and tuple(w.shape[:2]) == (nb, nh) and tuple(out.shape) == (nb, nh, m, n), \
        (a.shape, w.shape, out.shape, a_trans, w_trans)
    with _Timed("gemm", 2.0 * nb * nh * m * n * k):
        _dod.call("dod_gemm_bf16", _stream(a), a=a, w=w, m=m, n=n, k=k, lda=a.stride(2), ldw=w.stride(2),
                  bias=bias, act=act, out=out, ldo=out.stride(2), out_dtype=_DT[out.dtype], batch=nb,
                  batch_stride_a=a.stride(0) if nb > 1 else 0, batch_stride_w=w.stride(0) if nb > 1 else 0,
                  batch_stride_out=out.stride(0) if nb > 1 else 0, batch_inner=nh,
                  inner_stride_a=a.stride(1) if nh > 1 else 0, inner_stride_w=w.stride(1) if nh > 1 else 0,
                  inner_stride_out=out.stride(1) if nh > 1 else 0, a_trans=int(a_trans), w_trans=int(w_trans))
    return out


def transpose(x, out=None):
    """bf16 [.., R, C] -> [.., C, R] for 2-D, 3-D [B, R, C] or 4-D [B, H, R, C] strided views with unit
    inner stride; the result is contiguous over the batch dims with its row stride padded to a multiple
    of 8 so it can feed the GEMM."""
    assert x.dtype == torch.bfloat16 and x.stride(-1) == 1
    dims = x.dim()
    x4 = x if dims == 4 else (x.unsqueeze(0) if dims == 3 else x.unsqueeze(0).unsqueeze(0))
    if dims == 3:
        x4 = x.unsqueeze(1)
    nb, nh, r, c = x4.shape
    if out is None:
        out = torch.empty((nb, nh, c, (r + 7) // 8 * 8), dtype=torch.bfloat16, device=x.device)[:, :, :, :r]
    o4 = out
    while o4.dim() < 4:
        o4 = o4.unsqueeze(0) if o4.dim() == 2 else o4.unsqueeze(1)
    _dod.call("dod_transpose_bf16", _stream(x), **{"in": x4, "out": o4, "rows": r, "cols": c,
                                                 "ld_in": x4.stride(2), "ld_out": o4.stride(2), "batch": nb,
                                                 "batch_stride_in": x4.stride(0),
                                                 "batch_stride_out": o4.stride(1),
                                                 "batch_inner": nh, "inner_stride_in": x4.stride(1)})
    if dims == 4:
        return o4
    return o4[:, 0] if dims == 3 else o4[0, 0]


def lowrank_wgrad(big, small, r, out, *, transposed, alpha=1.0):
    """out[c, j] += alpha * sum_m big[m, c] * small[m, j], j < r; transposed: out[j, c]."""
    assert big.dtype == torch.bfloat16 and small.dtype == torch.bfloat16 and out.dtype == torch.float32
    m, cols = big.shape
    assert small.shape[0] == m and small.shape[1] >= r
    _dod.call("dod_lowrank_wgrad", _stream(big), big=big, small=small, out=out, m=m, cols=cols, r=r,
              ld_big=_rowmajor(big, "big"), ld_small=_rowmajor(small, "small"), ldo=_rowmajor(out, "out"),
              transposed=int(transposed), alpha=alpha)
    return out


def lowrank_wgrad_tc(big, small, r, out, *, transposed, splits, blocks=1, scratch=None):
    """The LoRA weight gradients on the tensor cores: per diagonal block i < blocks (fused q / k / v projections)

        out[i, c, j] += sum_m big[m, i * cb + c] * small[m, i * r + j]        (transposed: out[i, j, c])

    as ONE batched M-reduction GEMM: the rows are cut into `splits` equal slices (one per image), every
    (slice, block) is a batch entry whose operands are read as stored (MN-major tcgen05 tiles, a_trans / w_trans),
    the fp32 partial products [splits, blocks, ...] meet in one column-sum pass.  `big` is read from HBM exactly
    once (dod_lowrank_wgrad re-read it for every eight LoRA columns and computed the off-diagonal blocks too).
    big bf16 [M, blocks * cb], small bf16 [M, >= blocks * r] (unit inner stride), r % 8 == 0, M % splits == 0,
    out f32 contiguous [blocks, cb, r] (or [blocks, r, cb])."""
    assert big.dtype == torch.bfloat16 and small.dtype == torch.bfloat16 and out.dtype == torch.float32
    m, cols = big.shape
    assert cols % blocks == 0 and m % splits == 0 and r % 8 == 0 and small.shape[1] >= blocks * r
    cb, mc = cols // blocks, m // splits
    ldb, lds = _rowmajor(big, "big"), _rowmajor(small, "small")
    b4 = big.as_strided((splits, blocks, mc, cb), (mc * ldb, cb, ldb, 1), big.storage_offset())
    s4 = small.as_strided((splits, blocks, mc, r), (mc * lds, r, lds, 1), small.storage_offset())
    shape = (splits, blocks, r, cb) if transposed else (splits, blocks, cb, r)
    assert out.is_contiguous() and out.numel() == blocks * cb * r
    if scratch is None or scratch.numel() < splits * blocks * cb * r:
        scratch = torch.empty(splits * blocks * cb * r, dtype=torch.float32, device=big.device)
    part = scratch[:splits * blocks * cb * r].view(shape)
    if transposed:
        gemm_batched(s4, b4, part, a_trans=True, w_trans=True)
    else:
        gemm_batched(b4, s4, part, a_trans=True, w_trans=True)
    colsum(part.view(splits, blocks * cb * r), out.view(-1))
    return out


def colsum(x, out):
    """out[c] += sum_m x[m, c] (f32 accumulate)."""
    m, cols = x.shape
    assert out.dtype == torch.float32 and out.numel() >= cols
    _dod.call("dod_colsum", _stream(x), x=x, x_dtype=_DT[x.dtype], out=out, m=m, cols=cols, ld=_rowmajor(x, "x"))
    return out


def layernorm_bwd(dy, x, gamma, eps, *, dres=None, dgamma=None, dbeta=None):
    """dx (f32) of y = LN(x) given dy; dres (f32) is added; dgamma/dbeta accumulated if given."""
    rows, d = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and dy.is_contiguous() and dy.shape == x.shape
    dx = torch.empty_like(x)
    if dres is not None:
        assert dres.dtype == torch.float32 and dres.is_contiguous()
    _dod.call("dod_layernorm_bwd", _stream(x), dy=dy, dy_dtype=_DT[dy.dtype], x=x, gamma=gamma, dres=dres,
              dx=dx, dgamma=dgamma, dbeta=dbeta, rows=rows, d=d, eps=eps)
    return dx


def eltwise(mode, a, b=None, *, vec=None, out_dtype=None, out=None, cols=None, p0=0.0, seed=0,
            out2_dtype=None):
    """Elementwise helper (see dod_eltwise_mode); a, b 2-D with unit inner stride."""
    rows = a.shape[0]
    cols = cols if cols is not None else a.shape[1]
    out_cols = 2 * cols if mode == ELT_SWIGLU_BWD else cols
    if out is None:
        out = torch.empty((rows, out_cols), dtype=out_dtype or a.dtype, device=a.device)
    out2 = None
    if out2_dtype is not None:
        out2 = torch.empty((rows, out_cols), dtype=out2_dtype, device=a.device)
    _dod.call("dod_eltwise", _stream(a), mode=mode, a=a, a_dtype=_DT[a.dtype], b=b,
              b_dtype=_DT[b.dtype] if b is not None else 0, vec=vec, out=out, out_dtype=_DT[out.dtype],
              out2=out2, out2_dtype=_DT[out2.dtype] if out2 is not None else 0, rows=rows, cols=cols,
              ld_a=_rowmajor(a, "a"), ld_b=_rowmajor(b, "b") if b is not None else 0,
              ld_out=_rowmajor(out, "out"), p0=p0, seed=seed,
              seed_ptr=_seed_ptr if mode == ELT_DROPOUT else None)
    return (out, out2) if out2 is not None else out


def softmax_rows(s, n, scale, *, ldp=None, drop_p=0.0, seed=0):
    """P = softmax(scale * S[:, :n]) -> bf16 [rows, ldp] (zero padded)."""
    rows = s.shape[0]
    ldp = ldp or (n + 7) // 8 * 8
    p = torch.empty((rows, ldp), dtype=torch.bfloat16, device=s.device)
    _dod.call("dod_softmax_rows", _stream(s), s=s, s_dtype=_DT[s.dtype], p=p, rows=rows, n=n,
              lds=_rowmajor(s, "s"), ldp=ldp, scale=scale, drop_p=drop_p, seed=seed,
              seed_ptr=_seed_ptr if drop_p > 0 else None)
    return p


def softmax_bwd_rows(p, dp, n, scale, *, drop_p=0.0, seed=0):
    rows = p.shape[0]
    ds = torch.empty_like(p)
    _dod.call("dod_softmax_bwd_rows", _stream(p), p=p, dp=dp, dp_dtype=_DT[dp.dtype], ds=ds, rows=rows, n=n,
              ldp=_rowmajor(p, "p"), lddp=_rowmajor(dp, "dp"), ldds=_rowmajor(ds, "ds"), scale=scale,
              drop_p=drop_p, seed=seed, seed_ptr=_seed_ptr if drop_p > 0 else None)
    return ds


def deform_sample_bwd(value, ref, offs, logits, dout, dvalue, dqproj, batch, queries, heads, points, head_dim,
                      grid_h, grid_w, *, ref_is_logit=True):
    assert dvalue.dtype == torch.float32 and dqproj.dtype == torch.float32
    _dod.call("dod_deform_sample_bwd", _stream(value), value=value, value_dtype=_DT[value.dtype], ref=ref,
              offs=offs, logits=logits, dout=dout, dout_dtype=_DT[dout.dtype], dvalue=dvalue, dqproj=dqproj,
              batch=batch, queries=queries, heads=heads, points=points, head_dim=head_dim, grid_h=grid_h,
              grid_w=grid_w, ldv=_rowmajor(value, "value"), ldref=_rowmajor(ref, "ref"),
              ldoffs=_rowmajor(offs, "offs"), ldlog=_rowmajor(logits, "logits"), lddo=_rowmajor(dout, "dout"),
              lddv=_rowmajor(dvalue, "dvalue"), lddq=_rowmajor(dqproj, "dqproj"),
              ref_is_logit=int(ref_is_logit))


def criterion(logits, boxes, tgt_labels, tgt_boxes, tgt_offsets, out_q, out_t, num_boxes, *, alpha, gamma,
              w_ce, w_bbox, w_giou):
    """Fused SetCriterion forward+backward -> (losses [3], dlogits, dboxes_l1, dboxes_giou)."""
    b, q, c = logits.shape
    dev = logits.device
    assert logits.dtype == torch.float32 and logits.is_contiguous() and boxes.is_contiguous()
    losses = torch.zeros(3, dtype=torch.float32, device=dev)
    dlogits = torch.empty_like(logits)
    dboxes = torch.zeros_like(boxes)
    dboxes_giou = torch.zeros_like(boxes)
    tclass = torch.empty((b, q), dtype=torch.int32, device=dev)
    _dod.call("dod_criterion", _stream(logits), logits=logits, boxes=boxes, tgt_labels=tgt_labels,
              tgt_boxes=tgt_boxes, tgt_offsets=tgt_offsets, out_q=out_q, out_t=out_t, num_boxes=num_boxes,
              tclass=tclass, losses=losses, dlogits=dlogits, dboxes=dboxes, dboxes_giou=dboxes_giou, batch=b,
              queries=q, classes=c, max_k=out_q.shape[1], focal_alpha=alpha, focal_gamma=gamma, w_ce=w_ce,
              w_bbox=w_bbox, w_giou=w_giou)
    return losses, dlogits, dboxes, dboxes_giou


def sumsq(x, out):
    """out[0] += sum(x^2) for a flat f32 tensor."""
    assert x.dtype == torch.float32 and x.is_contiguous() and out.dtype == torch.float32
    _dod.call("dod_sumsq", _stream(x), x=x, n=x.numel(), out=out)
    return out


def adam_step(param, grad, exp_avg, exp_avg_sq, step, *, lr, beta1, beta2, eps, weight_decay,
              grad_sumsq=None, max_grad_norm=0.0, step_ptr=None):
    """step_ptr: optional int64 device counter used instead of `step` (graph replay)."""
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == param.numel()
    _dod.call("dod_adam_step", _stream(param), param=param, grad=grad, exp_avg=exp_avg, exp_avg_sq=exp_avg_sq,
              n=param.numel(), step=step, grad_sumsq=grad_sumsq, max_grad_norm=max_grad_norm, lr=lr,
              beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay, step_ptr=step_ptr)


def postprocess(logits, boxes, threshold=0.05):
    """-> (scores [B, cap], boxes_xywh [B, cap, 4], classes [B, cap], counts [B]) on the device."""
    b, q, c = logits.shape
    logits = logits.float().contiguous()
    boxes = boxes.float().contiguous()
    cap = q * (c - 1)
    dev = logits.device
    out_score = torch.empty((b, cap), dtype=torch.float32, device=dev)
    out_box = torch.empty((b, cap, 4), dtype=torch.float32, device=dev)
    out_class = torch.empty((b, cap), dtype=torch.int32, device=dev)
    counts = torch.empty((b,), dtype=torch.int32, device=dev)
    _dod.call("dod_postprocess", _stream(logits), logits=logits, boxes=boxes, out_score=out_score,
              out_box=out_box, out_class=out_class, counts=counts, batch=b, queries=q, classes=c, capacity=cap,
              threshold=threshold)
    return out_score, out_box, out_class, counts
