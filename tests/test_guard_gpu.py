"""GPU: out-of-bounds WRITE detection with guard bands (compute-sanitizer is closed on this pool -- gpurun answers
"compute-sanitizer is closed on this pool and stays closed", profiles/r02_summary.md -- so the TMA-store / direct-store
epilogues are checked with canaries instead): every output lives in the middle of a larger buffer pre-filled with a
sentinel, rows have a padded pitch, and after the op nothing outside the logical [rows, cols] may have changed.
Ragged shapes on purpose: the last tile of every kernel is partial in both dimensions."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

SENT_F32 = -12345.678
PAD_ROWS, PAD_COLS = 40, 24


@pytest.fixture(scope="module")
def ops():
    from dino_detector import ops as _ops
    return _ops


def guarded(rows, cols, dtype):
    """-> (view [rows, cols] with pitch cols + PAD_COLS inside a sentinel-filled buffer, checker)."""
    pitch = (cols + PAD_COLS + 7) // 8 * 8
    buf = torch.full((rows + 2 * PAD_ROWS, pitch), SENT_F32, dtype=dtype, device="cuda")
    view = buf[PAD_ROWS:PAD_ROWS + rows, :cols]
    sent = torch.tensor(SENT_F32, dtype=dtype).item()

    def check(name):
        torch.cuda.synchronize()
        assert (buf[:PAD_ROWS] == sent).all(), f"{name}: rows before the output were written"
        assert (buf[PAD_ROWS + rows:] == sent).all(), f"{name}: rows after the output were written"
        assert (buf[PAD_ROWS:PAD_ROWS + rows, cols:] == sent).all(), f"{name}: columns past n were written"
        assert not (view == sent).any(), f"{name}: part of the output was not written"
    return view, check


def _rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("m,n,k", [(515, 776, 200), (1370, 2304, 768), (300, 96, 72), (2741, 520, 136), (131, 56, 64)])
@pytest.mark.parametrize("mode", ["bf16", "f32", "residual", "residual_ln"])
def test_gemm_writes_stay_inside(ops, m, n, k, mode):
    if mode == "residual_ln" and (n < 256 or n % 16):
        pytest.skip("the folded-LayerNorm producer needs n >= 256, n % 16 == 0")
    a = _rnd((m, k), 1).bfloat16()
    w = _rnd((n, k), 2, 1 / math.sqrt(k)).bfloat16()
    bias = _rnd((n,), 3)
    if mode in ("bf16", "f32"):
        out, check = guarded(m, n, torch.bfloat16 if mode == "bf16" else torch.float32)
        ops.gemm(a, w, bias, act=ops.ACT_GELU_ERF, out=out)
        check(f"gemm {mode}")
        return
    res, _ = guarded(m, n, torch.float32)
    res.copy_(_rnd((m, n), 4))
    out, check = guarded(m, n, torch.float32)
    scale = torch.ones(n, device="cuda")
    if mode == "residual":
        ops.gemm(a, w, bias, scale=scale, residual=res, out=out)
        check("gemm residual")
    else:
        h16, check16 = guarded(m, n, torch.bfloat16)
        stats = torch.full((2 * ((n + 255) // 256) * m * 2 + 64,), SENT_F32, device="cuda")
        sv = stats[:2 * ((n + 255) // 256) * m * 2].view(2 * ((n + 255) // 256), m, 2)
        ops.gemm(a, w, bias, scale=scale, residual=res, out=out, ln_out=(h16, sv))
        check("gemm residual + ln_out (fp32)")
        check16("gemm residual + ln_out (bf16 copy)")
        assert (stats[-64:] == SENT_F32).all() and not (sv == SENT_F32).any()


@pytest.mark.parametrize("b,s,h", [(2, 257, 6), (1, 1370, 12), (3, 100, 2), (1, 129, 1)])
def test_fmha_writes_stay_inside(ops, b, s, h):
    d = h * 64
    qkv = _rnd((b * s, 3 * d), 5, 0.5).bfloat16()
    out, check = guarded(b * s, d, torch.bfloat16)
    lse = torch.full((b * h * s + 32,), SENT_F32, device="cuda")
    ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125, out=out, lse=lse[:b * h * s])
    check("fmha ctx")
    assert (lse[-32:] == SENT_F32).all() and not (lse[:-32] == SENT_F32).any()


@pytest.mark.parametrize("rows,d", [(1001, 384), (257, 768), (33, 1536)])
def test_layernorm_writes_stay_inside(ops, rows, d):
    x = _rnd((rows, d), 6) * 3 + 1
    out, check = guarded(rows, d, torch.bfloat16)
    ops.layernorm(x, _rnd((d,), 7), _rnd((d,), 8), 1e-6, out=out)
    check("layernorm")
