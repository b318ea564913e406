"""Training path: forward with saved activations + hand-written backward, wired into autograd.

`DINOv2ObjectDetector.forward` routes here when gradients are enabled and parameters require
them (reference train.py:1079-1101: `outputs = model(images)` ... `loss.backward()`).  The returned
`pred_logits` / `pred_boxes` carry a single autograd node (`_DetectorFn`) whose backward produces the
gradients of exactly the reference's trainable set: LoRA A/B of the last two encoder
blocks, `backbone.projection`, and everything under `decoder.*` (`decoder.reference_points` is
unused in the forward and gets no gradient, like the reference -- DDP find_unused_parameters=True).

What runs where
  * frozen blocks 0..L-3: the inference kernel sequence, nothing saved;
  * blocks L-2, L-1: same kernels (incl. the fused tcgen05 attention) with activations kept; the
    attention backward recomputes P = softmax(QK^T) per head with the batched GEMM and runs
    dV = P^T dO, dP = dO V^T, dS, dQ = dS K, dK = dS^T Q as batched tcgen05 GEMMs;
  * every dense contraction of the backward pass is dod_gemm_bf16: dgrad with transposed weight
    copies, full wgrad (decoder, projection) as an M-reduction GEMM over transposed operands that
    accumulates into the fp32 gradient, LoRA wgrad with dod_lowrank_wgrad (HBM-bound, r <= 64);
  * gradients w.r.t. activations travel as fp32 on the residual stream and bf16 into GEMMs.

bf16 mode only (fp32 mode is an inference/parity mode).  Train-mode dropout (the four residual-branch
dropouts of each decoder layer and nn.MultiheadAttention's attention-probability dropout) uses a
counter-hash mask that the backward regenerates; the RNG stream differs from torch's by design.
"""
from __future__ import annotations

import itertools
import math

import torch

from . import _engine, ops
from .ops import ACT_NONE, ACT_RELU

BF16, F32 = torch.bfloat16, torch.float32
_seed_counter = itertools.count(1)


def _pad8(n):
    return (n + 7) // 8 * 8


def _to_bf16(x):
    return x if x.dtype == BF16 else ops.eltwise(ops.ELT_CAST, x, out_dtype=BF16)


class Grads:
    """fp32 gradient accumulators keyed by parameter identity."""

    def __init__(self):
        self.bufs = {}

    def buf(self, key, shape, device):
        if key not in self.bufs:
            self.bufs[key] = torch.zeros(shape, dtype=F32, device=device)
        return self.bufs[key]


# ---------------------------------------------------------------------------
# linear layers with gradients
# ---------------------------------------------------------------------------
class TLinear:
    """Fully trainable nn.Linear (decoder / projection): y = act(x W^T + b)."""

    def __init__(self, weight, bias, name):
        self.weight, self.bias_p, self.name = weight, bias, name
        n, k = weight.shape
        self.n, self.k, self.n_pad = n, k, _pad8(n)
        self.w = ops.cast_pad_bf16(weight.detach().float().contiguous(), _pad8(k), dst_rows=self.n_pad)
        self.bias = None
        if bias is not None:
            self.bias = torch.zeros(self.n_pad, dtype=F32, device=weight.device)
            self.bias[:n].copy_(bias.detach())

    def fwd(self, x, *, act=ACT_NONE, residual=None, out_dtype=BF16):
        return ops.gemm(x, self.w, self.bias, act=act, residual=residual, out_dtype=out_dtype)

    def bwd(self, dy, x, grads, *, need_dx=True):
        """dy bf16 [M, n_pad] (zero in padded columns), x bf16 [M, k] -> dx bf16 [M, k_pad]."""
        dev = dy.device
        gw = grads.buf(("w", id(self.weight)), (self.n_pad, self.w.shape[1]), dev)
        # dW += dy^T x: both operands are read as stored ([M, n] / [M, k], the reduction dim M outermost)
        ops.gemm(dy, x, None, residual=gw, out=gw, a_trans=True, w_trans=True)
        if self.bias_p is not None:
            ops.colsum(dy, grads.buf(("b", id(self.bias_p)), (self.n_pad,), dev))
        if not need_dx:
            return None
        return ops.gemm(dy, self.w, None, w_trans=True)                                 # dx = dy W

    def collect(self, grads, out):
        gw = grads.bufs.get(("w", id(self.weight)))
        if gw is not None:
            out[id(self.weight)] = gw[:self.n, :self.k]
        if self.bias_p is not None:
            gb = grads.bufs.get(("b", id(self.bias_p)))
            if gb is not None:
                out[id(self.bias_p)] = gb[:self.n]


class LLinear:
    """One or more LoraLinear containers fused along N (frozen W, trainable A/B, utils.py:46-70)."""

    def __init__(self, mods, frozen_cache, key):
        self.mods = mods
        parts = [_engine._lin_parts(m) for m in mods]
        self.n = sum(p[0].shape[0] for p in parts)
        self.k = parts[0][0].shape[1]
        dev = parts[0][0].device
        fz = frozen_cache.get(key)
        if fz is None:
            w = torch.cat([p[0] for p in parts], dim=0)
            wb = ops.cast_pad_bf16(w.contiguous())
            bias = torch.cat([p[1] for p in parts]).contiguous() if parts[0][1] is not None else None
            fz = frozen_cache[key] = (wb, ops.transpose(wb), bias)
        self.w, self.wT, self.bias = fz
        self.r_parts = [p[2].shape[0] for p in parts]
        self.r = sum(self.r_parts)
        assert self.r <= 64, "LoRA rank (summed over fused projections) must be <= 64"
        a_cat = torch.zeros((64, self.k), dtype=F32, device=dev)
        b_full = torch.zeros((self.n, 64), dtype=F32, device=dev)
        r0 = n0 = 0
        for p in parts:
            n, r = p[0].shape[0], p[2].shape[0]
            a_cat[r0:r0 + r].copy_(p[2])
            b_full[n0:n0 + n, r0:r0 + r].copy_(p[3])          # alpha already folded in
            r0 += r
            n0 += n
        self.a = ops.cast_pad_bf16(a_cat)                     # [64, k]
        self.aT = ops.transpose(self.a)                       # [k, 64]
        self.b = ops.cast_pad_bf16(b_full)                    # [n, 64]
        self.bT = ops.transpose(self.b)                       # [64, n]

    def fwd(self, x, **kw):
        t = ops.gemm(x, self.a, None)                         # x A^T  [M, 64]
        return ops.gemm(x, self.w, self.bias, a2=t, w2=self.b, **kw), t

    def bwd(self, dy, x, t, grads, *, need_dx=True, splits=None):
        """splits: number of equal row slices (images) of the M-reduction.  With it the two LoRA weight gradients
        run as batched tcgen05 GEMMs (ops.lowrank_wgrad_tc: each operand read once, only the diagonal blocks of
        fused q / k / v); without it, or for ranks / part sizes the GEMM form does not cover, dod_lowrank_wgrad."""
        dev = dy.device
        dt = ops.gemm(dy, self.bT, None)                      # alpha * dy B   [M, 64]
        da = grads.buf(("la", id(self)), (self.r, self.k), dev)
        nb, rp = len(self.r_parts), self.r_parts[0]
        tc = (splits is not None and dy.shape[0] % splits == 0 and self.r % 8 == 0 and rp % 8 == 0 and
              self.n % nb == 0 and all(r == rp for r in self.r_parts) and
              all(m.out_features == self.n // nb for m in self.mods) and self.k % 8 == 0 and (self.n // nb) % 8 == 0)
        if tc:
            dbc = grads.buf(("lbc", id(self)), (nb, self.n // nb, rp), dev)
            ops.lowrank_wgrad_tc(dy, t, rp, dbc, transposed=False, splits=splits, blocks=nb)   # dy^T (x A^T), per part
            ops.lowrank_wgrad_tc(x, dt, self.r, da, transposed=True, splits=splits)            # (alpha dy B)^T x
        else:
            db = grads.buf(("lb", id(self)), (self.n, self.r), dev)
            ops.lowrank_wgrad(dy, t, self.r, db, transposed=False)
            ops.lowrank_wgrad(x, dt, self.r, da, transposed=True)
        if not need_dx:
            return None
        return ops.gemm(dy, self.wT, None, a2=dt, w2=self.aT)                  # dy W + dt A

    def collect(self, grads, out):
        db, da = grads.bufs.get(("lb", id(self))), grads.bufs.get(("la", id(self)))
        dbc = grads.bufs.get(("lbc", id(self)))          # per-part [parts, n, r] (tensor-core form)
        if db is None and dbc is None:
            return
        r0 = n0 = 0
        for i, (m, r) in enumerate(zip(self.mods, self.r_parts)):
            n = m.out_features
            # d/dB of alpha * B (A x) = alpha * dy^T t ; b_full carries alpha -> dt does, db does not
            g_b = dbc[i] if dbc is not None else db[n0:n0 + n, r0:r0 + r]
            if dbc is not None and db is not None:
                g_b = g_b + db[n0:n0 + n, r0:r0 + r]
            out[id(m.lora_B.weight)] = g_b * float(m.alpha)
            out[id(m.lora_A.weight)] = da[r0:r0 + r]
            r0 += r
            n0 += n


# ---------------------------------------------------------------------------
# attention backward (materialised probabilities, batched tcgen05 GEMMs)
# ---------------------------------------------------------------------------
def _heads(x3, heads, dh):
    """[B, L, H*dh] (strided, unit inner stride) -> [B, H, L, dh] view, no copy."""
    b, l, _ = x3.shape
    return x3.as_strided((b, heads, l, dh), (x3.stride(0), dh, x3.stride(1), 1), x3.storage_offset())


def attention_fwd_dropout(q3, k3, v3, heads, dh, scale, drop_p, seed):
    """Materialised attention forward with dropout on the probabilities (nn.MultiheadAttention in
    train mode): ctx = dropout(softmax(Q K^T * scale)) V, all heads in one batched GEMM each."""
    b, lq, d = q3.shape
    lk = k3.shape[1]
    lkp = _pad8(lk)
    q4, k4, v4 = _heads(q3, heads, dh), _heads(k3, heads, dh), _heads(v3, heads, dh)
    ctx = torch.empty((b, lq, d), dtype=BF16, device=q3.device)
    s_full = torch.empty((b, heads, lq, lkp), dtype=F32, device=q3.device)
    ops.gemm_batched(q4, k4, s_full[..., :lk])
    pd = ops.softmax_rows(s_full.view(-1, lkp), lk, scale, ldp=lkp, drop_p=drop_p, seed=seed)
    ops.gemm_batched(pd.view(b, heads, lq, lkp)[..., :lk], v4, _heads(ctx, heads, dh), w_trans=True)
    return ctx


def attention_bwd(q3, k3, v3, dctx3, dq3, dk3, dv3, heads, dh, scale, drop_p=0.0, seed=0):
    """q3 [B, Lq, H*dh], k3/v3 [B, Lk, H*dh], dctx3 [B, Lq, H*dh] (bf16 3-D views, unit inner
    stride); writes bf16 gradients into the dq3/dk3/dv3 views.  All (image, head) problems run as
    ONE batched tcgen05 GEMM per product through 4-D TMA maps over the head-strided views.
    drop_p/seed: the attention dropout mask of attention_fwd_dropout is regenerated from the hash."""
    b, lq, _ = q3.shape
    lk = k3.shape[1]
    lkp = _pad8(lk)
    dev = q3.device
    q4, k4, v4, do4 = (_heads(t, heads, dh) for t in (q3, k3, v3, dctx3))
    s_full = torch.empty((b, heads, lq, lkp), dtype=F32, device=dev)
    s2d = s_full.view(-1, lkp)
    ops.gemm_batched(q4, k4, s_full[..., :lk])                                       # S = Q K^T
    p = ops.softmax_rows(s2d, lk, scale, ldp=lkp)                                    # P  [B*H*Lq, lkp]
    pd = p if drop_p <= 0 else ops.softmax_rows(s2d, lk, scale, ldp=lkp, drop_p=drop_p, seed=seed)
    ops.gemm_batched(do4, v4, s_full[..., :lk])                                      # dP = dO V^T
    ds = ops.softmax_bwd_rows(p, s2d, lk, scale, drop_p=drop_p, seed=seed)           # dS
    ds4 = ds.view(b, heads, lq, lkp)[..., :lk]
    pd4 = pd.view(b, heads, lq, lkp)[..., :lk]
    # the transposed products read dS / P / K / Q / dO as stored (MN-major tcgen05 operands)
    ops.gemm_batched(ds4, k4, _heads(dq3, heads, dh), w_trans=True)                  # dQ = dS K
    ops.gemm_batched(ds4, q4, _heads(dk3, heads, dh), a_trans=True, w_trans=True)    # dK = dS^T Q
    ops.gemm_batched(pd4, do4, _heads(dv3, heads, dh), a_trans=True, w_trans=True)   # dV = dropout(P)^T dO


# ---------------------------------------------------------------------------
# encoder blocks with LoRA
# ---------------------------------------------------------------------------
class EncTrainLayer:
    def __init__(self, lyr, swiglu, frozen_cache, idx):
        att = lyr.attention
        self.swiglu = swiglu
        self.n1 = (_engine.f32c(lyr.norm1.weight), _engine.f32c(lyr.norm1.bias))
        self.n2 = (_engine.f32c(lyr.norm2.weight), _engine.f32c(lyr.norm2.bias))
        self.ls1, self.ls2 = _engine.f32c(lyr.layer_scale1.lambda1), _engine.f32c(lyr.layer_scale2.lambda1)
        self.qkv = LLinear([att.attention.query, att.attention.key, att.attention.value], frozen_cache, (idx, "qkv"))
        self.proj = LLinear([att.output.dense], frozen_cache, (idx, "proj"))
        if swiglu:
            self.fc1 = LLinear([lyr.mlp.weights_in], frozen_cache, (idx, "w_in"))
            self.fc2 = LLinear([lyr.mlp.weights_out], frozen_cache, (idx, "w_out"))
        else:
            self.fc1 = LLinear([lyr.mlp.fc1], frozen_cache, (idx, "fc1"))
            self.fc2 = LLinear([lyr.mlp.fc2], frozen_cache, (idx, "fc2"))

    def fwd(self, x, b, n, heads):
        d = x.shape[1]
        sv = {"x": x}
        sv["h1"] = ops.layernorm(x, *self.n1, 1e-6)
        sv["qkv"], sv["t1"] = self.qkv.fwd(sv["h1"])
        sv["lse"] = torch.empty((b, heads, n), dtype=F32, device=x.device)
        sv["ctx"] = ops.fmha(sv["qkv"], b, n, heads, q_off=0, k_off=d, v_off=2 * d, scale=0.125, lse=sv["lse"])
        sv["x_mid"], sv["t2"] = self.proj.fwd(sv["ctx"], scale=self.ls1, residual=x, out_dtype=F32)
        sv["h2"] = ops.layernorm(sv["x_mid"], *self.n2, 1e-6)
        sv["z"], sv["t3"] = self.fc1.fwd(sv["h2"])
        if self.swiglu:
            sv["a"] = ops.eltwise(ops.ELT_SWIGLU_FWD, sv["z"], cols=sv["z"].shape[1] // 2)
        else:
            sv["a"] = ops.eltwise(ops.ELT_GELU_FWD, sv["z"])
        x_out, sv["t4"] = self.fc2.fwd(sv["a"], scale=self.ls2, residual=sv["x_mid"], out_dtype=F32)
        return x_out, sv

    def bwd(self, dx_out, sv, b, n, heads, grads, need_dx_in):
        d = dx_out.shape[1]
        dy = ops.eltwise(ops.ELT_SCALE_COLS, dx_out, vec=self.ls2, out_dtype=BF16)
        da = self.fc2.bwd(dy, sv["a"], sv["t4"], grads, splits=b)
        if self.swiglu:
            dz = ops.eltwise(ops.ELT_SWIGLU_BWD, da, sv["z"], cols=da.shape[1])
        else:
            dz = ops.eltwise(ops.ELT_GELU_BWD, da, sv["z"])
        dh2 = self.fc1.bwd(dz, sv["h2"], sv["t3"], grads, splits=b)
        dx_mid = ops.layernorm_bwd(dh2, sv["x_mid"], self.n2[0], 1e-6, dres=dx_out)
        dy = ops.eltwise(ops.ELT_SCALE_COLS, dx_mid, vec=self.ls1, out_dtype=BF16)
        dctx = self.proj.bwd(dy, sv["ctx"], sv["t2"], grads, splits=b)
        dqkv = torch.empty_like(sv["qkv"])
        # fused flash-style backward: probabilities recomputed from q, k and the saved log-sum-exp
        ops.fmha_bwd(sv["qkv"], sv["ctx"], dctx, sv["lse"], dqkv, b, n, heads, q_off=0, k_off=d, v_off=2 * d,
                     scale=0.125)
        dh1 = self.qkv.bwd(dqkv, sv["h1"], sv["t1"], grads, need_dx=need_dx_in, splits=b)
        if not need_dx_in:
            return None
        return ops.layernorm_bwd(dh1, sv["x"], self.n1[0], 1e-6, dres=dx_mid)

    def collect(self, grads, out):
        for l in (self.qkv, self.proj, self.fc1, self.fc2):
            l.collect(grads, out)


# ---------------------------------------------------------------------------
# decoder
# ---------------------------------------------------------------------------
class _LN:
    def __init__(self, mod):
        self.mod = mod
        self.g, self.b = _engine.f32c(mod.weight), _engine.f32c(mod.bias)

    def fwd(self, x):
        y32, y16 = ops.layernorm(x, self.g, self.b, 1e-5, out_dtype=F32, also_other=True)
        return y32, y16

    def bwd(self, dy, x, grads):
        dev = x.device
        d = x.shape[1]
        return ops.layernorm_bwd(dy, x, self.g, 1e-5, dgamma=grads.buf(("g", id(self.mod.weight)), (d,), dev),
                                 dbeta=grads.buf(("bb", id(self.mod.bias)), (d,), dev))

    def collect(self, grads, out):
        g = grads.bufs.get(("g", id(self.mod.weight)))
        if g is not None:
            out[id(self.mod.weight)] = g
            out[id(self.mod.bias)] = grads.bufs[("bb", id(self.mod.bias))]


class _FusedLinear(TLinear):
    """Several nn.Linear containers concatenated along N (in_proj halves, query projections)."""

    def __init__(self, pieces, name):
        # pieces: list of (weight tensor view, bias tensor view, owner param weight, owner param bias, row slice in owner)
        self.pieces = pieces
        w = torch.cat([p[0].detach().float() for p in pieces], dim=0).contiguous()
        bias = torch.cat([p[1].detach().float() for p in pieces]).contiguous()
        n, k = w.shape
        self.weight, self.bias_p, self.name = w, bias, name       # identity keys for the grad buffers
        self.n, self.k, self.n_pad = n, k, _pad8(n)
        self.w = ops.cast_pad_bf16(w, _pad8(k), dst_rows=self.n_pad)
        self.bias = torch.zeros(self.n_pad, dtype=F32, device=w.device)
        self.bias[:n].copy_(bias)

    def collect(self, grads, out):
        gw, gb = grads.bufs.get(("w", id(self.weight))), grads.bufs.get(("b", id(self.bias_p)))
        if gw is None:
            return
        r0 = 0
        for wv, bv, wparam, bparam, rows in self.pieces:
            n = wv.shape[0]
            gwp = out.setdefault(id(wparam), torch.zeros(wparam.shape, dtype=F32, device=gw.device))
            gbp = out.setdefault(id(bparam), torch.zeros(bparam.shape, dtype=F32, device=gw.device))
            gwp[rows].copy_(gw[r0:r0 + n, :self.k])
            gbp[rows].copy_(gb[r0:r0 + n])
            r0 += n


def _piece(lin):
    return (lin.weight, lin.bias, lin.weight, lin.bias, slice(0, lin.weight.shape[0]))


class DecTrainLayer:
    def __init__(self, lyr, hd, deformable):
        self.deformable = deformable
        sa = lyr.self_attn
        self.sa_in = _FusedLinear([(sa.in_proj_weight, sa.in_proj_bias, sa.in_proj_weight, sa.in_proj_bias,
                                    slice(0, 3 * hd))], "sa_in")
        self.sa_out = TLinear(sa.out_proj.weight, sa.out_proj.bias, "sa_out")
        self.n1, self.n2, self.n3 = _LN(lyr.norm1), _LN(lyr.norm2), _LN(lyr.norm3)
        self.l1 = TLinear(lyr.linear1.weight, lyr.linear1.bias, "l1")
        self.l2 = TLinear(lyr.linear2.weight, lyr.linear2.bias, "l2")
        if deformable:
            ca = lyr.cross_attn
            self.qproj = _FusedLinear([_piece(ca.sampling_offsets), _piece(ca.attention_weights),
                                       _piece(lyr.reference_points_proj)], "qproj")
            self.value = TLinear(ca.value_proj.weight, ca.value_proj.bias, "value")
            self.out = TLinear(ca.output_proj.weight, ca.output_proj.bias, "out")
        else:
            ca = lyr.multihead_attn
            w, bb = ca.in_proj_weight, ca.in_proj_bias
            self.ca_q = _FusedLinear([(w[:hd], bb[:hd], w, bb, slice(0, hd))], "ca_q")
            self.ca_kv = _FusedLinear([(w[hd:], bb[hd:], w, bb, slice(hd, 3 * hd))], "ca_kv")
            self.ca_out = TLinear(ca.out_proj.weight, ca.out_proj.bias, "ca_out")

    def modules(self):
        mods = [self.sa_in, self.sa_out, self.n1, self.n2, self.n3, self.l1, self.l2]
        mods += [self.qproj, self.value, self.out] if self.deformable else [self.ca_q, self.ca_kv, self.ca_out]
        return mods


class TrainState:
    pass


def _dropout_add(x_sub, residual32, p, seed):
    """residual + dropout(x_sub) -> f32 (x_sub f32)."""
    if p <= 0.0:
        return ops.eltwise(ops.ELT_ADD, residual32, x_sub, out_dtype=F32)
    return ops.eltwise(ops.ELT_ADD, residual32, ops.eltwise(ops.ELT_DROPOUT, x_sub, p0=p, seed=seed, out_dtype=F32),
                       out_dtype=F32)


def _dropout_bwd(dy, p, seed, out_dtype):
    if p <= 0.0:
        return dy if dy.dtype == out_dtype else ops.eltwise(ops.ELT_CAST, dy, out_dtype=out_dtype)
    return ops.eltwise(ops.ELT_DROPOUT, dy, p0=p, seed=seed, out_dtype=out_dtype)


def train_forward(model, pixel_values):
    bk, dec = model.backbone, model.decoder
    if _engine.resolve_precision(model.precision) != "bf16":
        raise NotImplementedError("the libdod training path runs in bf16 mode (fp32 mode is inference-only)")
    st = TrainState()
    dino = bk.dino
    n_layers = len(dino.encoder.layer)
    n_train = min(2, n_layers)
    heads = dino.num_heads
    # ---- embeddings + frozen blocks (inference kernels, nothing saved; packed once, not per step) ----
    x, b, n = _engine.backbone_forward(bk._get_frozen_pack(n_layers - n_train), pixel_values, final_norm=False)
    st.b, st.n, st.heads = b, n, heads
    # ---- LoRA blocks ----
    cache = bk.__dict__.setdefault("_train_frozen_cache", {})
    fver = tuple((p.data_ptr(), p._version) for p in dino.parameters() if not p.requires_grad)
    if cache.get("_ver") != fver:
        cache.clear()
        cache["_ver"] = fver
    st.enc = []
    for i in range(n_layers - n_train, n_layers):
        lt = EncTrainLayer(dino.encoder.layer[i], dino.use_swiglu, cache, i)
        x, sv = lt.fwd(x, b, n, heads)
        st.enc.append((lt, sv))
    # ---- final LayerNorm (+ projection) ----
    st.x_final = x
    st.fln = (_engine.f32c(dino.layernorm.weight), _engine.f32c(dino.layernorm.bias))
    mem = ops.layernorm(x, *st.fln, 1e-6)
    st.lnf_out = mem
    st.proj = None
    if bk.projection is not None:
        st.proj = TLinear(bk.projection.weight, bk.projection.bias, "projection")
        mem = st.proj.fwd(mem)
    st.memory = mem
    # ---- decoder ----
    q, hd, nh = dec.num_queries, dec.hidden_dim, dec.nheads
    dh = hd // nh
    scale = 1.0 / math.sqrt(dh)
    p_drop = float(dec.dropout_p) if dec.training else 0.0
    # with a device seed counter (graph replay) the per-step part of the seed lives on the device
    st.p_drop, st.seed = p_drop, (0 if ops.device_seed() is not None else next(_seed_counter) << 44)
    st.q, st.hd, st.nh, st.dh, st.scale = q, hd, nh, dh, scale
    st.deformable = dec.use_deformable
    layers = list(dec.decoder.layers)
    uniq = {}
    st.dec_layers = []
    for lyr in layers:                          # deformable: one shared module -> one DecTrainLayer
        if id(lyr) not in uniq:
            uniq[id(lyr)] = DecTrainLayer(lyr, hd, dec.use_deformable)
        st.dec_layers.append(uniq[id(lyr)])
    st.query = _engine.f32c(dec.query_embed.weight)
    tgt32, tgt = ops.broadcast_rows(st.query, b)
    st.value = None
    if dec.use_deformable:
        gh, gw = _engine.grid_shape(n)
        if gh * gw != n:
            raise ValueError(f"Cannot reshape input of size {n} into a square feature map")
        st.grid = (gh, gw)
    st.saved = []
    for li, L in enumerate(st.dec_layers):
        sv = {"tgt": tgt, "tgt32": tgt32}
        seed = st.seed + (li << 36)
        sv["qkv"] = L.sa_in.fwd(tgt)
        if p_drop > 0:
            qkv3 = sv["qkv"].view(b, q, -1)
            ctx = attention_fwd_dropout(qkv3[:, :, :hd], qkv3[:, :, hd:2 * hd], qkv3[:, :, 2 * hd:3 * hd], nh, dh,
                                        scale, p_drop, seed + 5).view(b * q, hd)
        else:
            ctx = ops.mha_small(sv["qkv"][:, :hd], sv["qkv"][:, hd:2 * hd], sv["qkv"][:, 2 * hd:], b, q, q, nh, dh, scale)
        sv["ctx_sa"] = ctx
        sv["x1"] = _dropout_add(L.sa_out.fwd(ctx, out_dtype=F32), tgt32, p_drop, seed + 1)
        t1_32, t1 = L.n1.fwd(sv["x1"])
        sv["t1"], sv["t1_32"] = t1, t1_32
        if st.deformable:
            if st.value is None:
                st.value = L.value.fwd(mem)
            hp = nh * dec.n_points
            sv["qp"] = L.qproj.fwd(t1, out_dtype=F32)
            qp = sv["qp"]
            sv["samp"] = ops.deform_sample(st.value, qp[:, 3 * hp:3 * hp + 2], qp[:, :2 * hp], qp[:, 2 * hp:3 * hp],
                                           b, q, nh, dec.n_points, dh, *st.grid, ref_is_logit=True)
            sub = L.out.fwd(sv["samp"], out_dtype=F32)
        else:
            sv["cq"] = L.ca_q.fwd(t1)
            sv["kv"] = L.ca_kv.fwd(mem)
            if p_drop > 0:
                kv3 = sv["kv"].view(b, n, 2 * hd)
                sv["ctx_ca"] = attention_fwd_dropout(sv["cq"].view(b, q, hd), kv3[:, :, :hd], kv3[:, :, hd:], nh, dh,
                                                     scale, p_drop, seed + 6).view(b * q, hd)
            else:
                sv["ctx_ca"] = ops.mha_small(sv["cq"], sv["kv"][:, :hd], sv["kv"][:, hd:], b, q, n, nh, dh, scale)
            sub = L.ca_out.fwd(sv["ctx_ca"], out_dtype=F32)
        sv["x2"] = _dropout_add(sub, t1_32, p_drop, seed + 2)
        t2_32, t2 = L.n2.fwd(sv["x2"])
        sv["t2"] = t2
        a = L.l1.fwd(t2, act=ACT_RELU)
        sv["a"] = a
        a_d = a if p_drop <= 0 else ops.eltwise(ops.ELT_DROPOUT, a, p0=p_drop, seed=seed + 3, out_dtype=BF16)
        sv["a_d"] = a_d
        sv["x3"] = _dropout_add(L.l2.fwd(a_d, out_dtype=F32), t2_32, p_drop, seed + 4)
        tgt32, tgt = L.n3.fwd(sv["x3"])
        st.saved.append(sv)
    st.hs = tgt
    st.cls = TLinear(dec.class_embed.weight, dec.class_embed.bias, "cls")
    st.box0 = TLinear(dec.bbox_embed.mlp[0].weight, dec.bbox_embed.mlp[0].bias, "box0")
    st.box1 = TLinear(dec.bbox_embed.mlp[2].weight, dec.bbox_embed.mlp[2].bias, "box1")
    nc = dec.class_embed.out_features
    logits = st.cls.fwd(tgt, out_dtype=F32)
    st.box_hid = st.box0.fwd(tgt, act=ACT_RELU)
    box_raw = st.box1.fwd(st.box_hid, out_dtype=F32)
    st.nc = nc
    st.logits = ops.rowcopy(logits, nc).view(b, q, nc)
    st.boxes = ops.rowcopy(box_raw, 4, sigmoid=True).view(b, q, 4)
    return st


def train_backward(model, st, dlogits, dboxes):
    """-> {id(param): fp32 gradient tensor}."""
    bk, dec = model.backbone, model.decoder
    b, n, q, hd, nh, dh = st.b, st.n, st.q, st.hd, st.nh, st.dh
    r = b * q
    dev = st.hs.device
    grads = Grads()
    p_drop = st.p_drop
    # ---- heads ----
    dl = torch.zeros((r, st.cls.n_pad), dtype=F32, device=dev)
    dl[:, :st.nc].copy_(dlogits.reshape(r, st.nc))
    dl16 = _to_bf16(dl)
    dt = st.cls.bwd(dl16, st.hs, grads)                                            # bf16 [r, hd]
    db = torch.zeros((r, st.box1.n_pad), dtype=F32, device=dev)
    db[:, :4].copy_(dboxes.reshape(r, 4))
    sg = torch.zeros((r, st.box1.n_pad), dtype=F32, device=dev)
    sg[:, :4].copy_(st.boxes.reshape(r, 4))
    draw = ops.eltwise(ops.ELT_SIGMOID_BWD, db, sg, out_dtype=BF16)
    dhid = st.box1.bwd(draw, st.box_hid, grads)
    dhid = ops.eltwise(ops.ELT_RELU_BWD, dhid, st.box_hid, out_dtype=BF16)
    dt2 = st.box0.bwd(dhid, st.hs, grads)
    dtgt = ops.eltwise(ops.ELT_ADD, dt, dt2, out_dtype=F32)                        # f32 [r, hd]
    # ---- decoder layers, last to first ----
    dvalue = None
    dmem32 = None
    if st.deformable:
        dvalue = torch.zeros((b * n, hd), dtype=F32, device=dev)
    else:
        dmem32 = torch.zeros((b * n, hd), dtype=F32, device=dev)
    for li in reversed(range(len(st.dec_layers))):
        L, sv = st.dec_layers[li], st.saved[li]
        seed = st.seed + (li << 36)
        dx3 = L.n3.bwd(dtgt, sv["x3"], grads)                                      # f32
        dy = _dropout_bwd(dx3, p_drop, seed + 4, BF16)
        da = L.l2.bwd(dy, sv["a_d"], grads)
        if p_drop > 0:
            da = ops.eltwise(ops.ELT_DROPOUT, da, p0=p_drop, seed=seed + 3, out_dtype=BF16)
        dz = ops.eltwise(ops.ELT_RELU_BWD, da, sv["a"], out_dtype=BF16)
        dt2_ffn = L.l1.bwd(dz, sv["t2"], grads)
        dt2_tot = ops.eltwise(ops.ELT_ADD, dx3, dt2_ffn, out_dtype=F32)
        dx2 = L.n2.bwd(dt2_tot, sv["x2"], grads)
        dy = _dropout_bwd(dx2, p_drop, seed + 2, BF16)
        if st.deformable:
            hp = nh * dec.n_points
            dsamp = L.out.bwd(dy, sv["samp"], grads)
            dqp = torch.zeros((r, L.qproj.n_pad), dtype=F32, device=dev)
            qp = sv["qp"]
            ops.deform_sample_bwd(st.value, qp[:, 3 * hp:3 * hp + 2], qp[:, :2 * hp], qp[:, 2 * hp:3 * hp], dsamp,
                                  dvalue, dqp, b, q, nh, dec.n_points, dh, *st.grid, ref_is_logit=True)
            dt1_ca = L.qproj.bwd(_to_bf16(dqp), sv["t1"], grads)
        else:
            dctx = L.ca_out.bwd(dy, sv["ctx_ca"], grads)
            dcq = torch.empty_like(sv["cq"])
            dkv = torch.empty_like(sv["kv"])
            kv3 = sv["kv"].view(b, n, 2 * hd)
            dkv3 = dkv.view(b, n, 2 * hd)
            attention_bwd(sv["cq"].view(b, q, hd), kv3[:, :, :hd], kv3[:, :, hd:], dctx.view(b, q, hd),
                          dcq.view(b, q, hd), dkv3[:, :, :hd], dkv3[:, :, hd:], nh, dh, st.scale,
                          drop_p=p_drop, seed=seed + 6)
            dt1_ca = L.ca_q.bwd(dcq, sv["t1"], grads)
            dm = L.ca_kv.bwd(dkv, st.memory, grads)
            dmem32 = ops.eltwise(ops.ELT_ADD, dmem32, dm, out_dtype=F32)
        dt1_tot = ops.eltwise(ops.ELT_ADD, dx2, dt1_ca, out_dtype=F32)
        dx1 = L.n1.bwd(dt1_tot, sv["x1"], grads)
        dy = _dropout_bwd(dx1, p_drop, seed + 1, BF16)
        dctx = L.sa_out.bwd(dy, sv["ctx_sa"], grads)
        dqkv = torch.empty_like(sv["qkv"])
        qkv3, dq3 = sv["qkv"].view(b, q, -1), dqkv.view(b, q, -1)
        attention_bwd(qkv3[:, :, :hd], qkv3[:, :, hd:2 * hd], qkv3[:, :, 2 * hd:3 * hd], dctx.view(b, q, hd),
                      dq3[:, :, :hd], dq3[:, :, hd:2 * hd], dq3[:, :, 2 * hd:3 * hd], nh, dh, st.scale,
                      drop_p=p_drop, seed=seed + 5)
        if dqkv.shape[1] > 3 * hd:
            dqkv[:, 3 * hd:].zero_()
        dt0 = L.sa_in.bwd(dqkv, sv["tgt"], grads)
        dtgt = ops.eltwise(ops.ELT_ADD, dx1, dt0, out_dtype=F32)
    # ---- query embedding: tgt0[b, q, :] = query[q, :]  ->  sum over images ----
    gq = torch.zeros((q * hd,), dtype=F32, device=dev)
    ops.colsum(dtgt.view(b, q * hd), gq)
    out = {id(dec.query_embed.weight): gq.view(q, hd)}
    # ---- value projection / memory gradient ----
    if st.deformable:
        L = st.dec_layers[0]
        dmem = L.value.bwd(_to_bf16(dvalue), st.memory, grads)                     # bf16 [b*n, hd]
    else:
        dmem = _to_bf16(dmem32)
    # ---- projection + final LayerNorm ----
    if st.proj is not None:
        dmem = st.proj.bwd(dmem, st.lnf_out, grads)
    dx = ops.layernorm_bwd(dmem, st.x_final, st.fln[0], 1e-6)
    # ---- projection / decoder / head gradients are complete: collect them now ----
    if st.proj is not None:
        st.proj.collect(grads, out)
    seen = set()
    for L in st.dec_layers:
        if id(L) in seen:
            continue
        seen.add(id(L))
        for m in L.modules():
            m.collect(grads, out)
    for m in (st.cls, st.box0, st.box1):
        m.collect(grads, out)
    # parallel.FlatGradSync (one flat fp32 gradient buffer): this part of the buffer is all-reduced on a
    # side stream UNDER the backward of the LoRA blocks below; the node then returns None for these parameters
    from . import parallel
    sink = parallel.sink_for(dec.query_embed.weight)
    in_place = sink.early_tail(out) if sink is not None and sink.early_enabled() else set()
    # ---- LoRA blocks ----
    for j in reversed(range(len(st.enc))):
        lt, sv = st.enc[j]
        dx = lt.bwd(dx, sv, b, n, st.heads, grads, need_dx_in=(j > 0))
    for lt, _ in st.enc:
        lt.collect(grads, out)
    for pid in in_place:
        out[pid] = _IN_PLACE
    return out


_IN_PLACE = object()     # marker: the gradient was accumulated into p.grad from inside the backward


class _DetectorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, pixel_values, *params):
        with torch.no_grad():
            st = train_forward(model, pixel_values)
        ctx.model, ctx.st, ctx.n_params = model, st, len(params)
        ctx.param_ids = [id(p) for p in params]
        ctx.param_shapes = [tuple(p.shape) for p in params]
        ctx.param_dtypes = [p.dtype for p in params]
        return st.logits, st.boxes

    @staticmethod
    def backward(ctx, dlogits, dboxes):
        st = ctx.st
        dev = st.hs.device
        if dlogits is None:
            dlogits = torch.zeros(st.logits.shape, dtype=F32, device=dev)
        if dboxes is None:
            dboxes = torch.zeros(st.boxes.shape, dtype=F32, device=dev)
        with torch.no_grad():
            g = train_backward(ctx.model, st, dlogits.float().contiguous(), dboxes.float().contiguous())
        outs = []
        for pid, shape, dt in zip(ctx.param_ids, ctx.param_shapes, ctx.param_dtypes):
            t = g.get(pid)
            if t is _IN_PLACE:
                outs.append(None)
                continue
            # every parameter handed to this node gets a gradient (DDP waits for each hook)
            outs.append(torch.zeros(shape, dtype=dt, device=dev) if t is None
                        else t.reshape(shape).to(dt).contiguous())
        ctx.st = None
        return (None, None, *outs)


def detector_forward_train(model, pixel_values):
    params = []
    seen = set()
    for name, p in model.named_parameters():
        # decoder.reference_points is never used in the forward (reference detr_decoder.py:44-45): it is
        # not an input of the autograd node, so it stays grad-less / "unused" exactly like the reference
        if name.startswith("decoder.reference_points."):
            continue
        if p.requires_grad and id(p) not in seen:
            seen.add(id(p))
            params.append(p)
    logits, boxes = _DetectorFn.apply(model, pixel_values, *params)
    return {"pred_logits": logits, "pred_boxes": boxes}
