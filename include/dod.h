/* libdod — C ABI of the B200-native (sm_100a) detector hot path.
 *
 * Drop-in boundary for mudit1729/dinov2-od: the reference has no FFI, its
 * boundary is the Python module API `dino_detector.models` / `.matching`
 * (reference dino_detector/models/detector.py:9-69, matching.py:9-122).  The
 * Python host package in dinov2-od_b200/dino_detector mirrors that API and calls
 * the entry points below through ctypes with raw device pointers.  Each entry
 * point cites the reference (or third-party) code it replaces.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types.
 *   - every op is  int32_t dod_<op>(const dod_<op>_args*, dod_stream_t)  and
 *     returns 0 on success or a negative dod_status; dod_last_error() gives a
 *     thread-local message.
 *   - ops never allocate, never synchronise and never touch the host copy of
 *     the data: the caller owns all device memory and keeps it alive until the
 *     stream has passed the op (PyTorch caching allocator + same stream).
 *   - "bf16" buffers are raw 16-bit bfloat16, "f32" raw IEEE binary32.
 *   - leading dimensions (ld*) are in ELEMENTS.
 */
#ifndef DOD_H_
#define DOD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dod_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define DOD_API __attribute__((visibility("default")))
#else
#define DOD_API
#endif

typedef enum {
  DOD_OK = 0,
  DOD_ERR_INVALID = -1,     /* bad argument / unsupported shape            */
  DOD_ERR_CUDA = -2,        /* CUDA runtime / driver error                 */
  DOD_ERR_DEVICE = -3,      /* not an sm_100 device                        */
  DOD_ERR_UNSUPPORTED = -4  /* valid but not implemented                   */
} dod_status;

/* ---- library ---------------------------------------------------------- */
DOD_API int32_t dod_version(void);              /* major*10000 + minor*100 + patch */
DOD_API const char* dod_last_error(void);       /* thread-local, never NULL        */
DOD_API int32_t dod_device_check(int32_t dev);  /* DOD_OK iff device is sm_100     */
/* number of libdod kernels launched by this process since load / reset     */
DOD_API int64_t dod_launch_count(void);
DOD_API void dod_launch_count_reset(void);

/* ---- dense contraction (tcgen05 / TMEM / TMA) ---------------------------
 * out = residual + scale[n] * act( A[M,K] . W[N,K]^T  (+ A2[M,K2] . W2[N,K2]^T)  + bias[n] )
 *
 * Replaces every nn.Linear on the path: HF Dinov2 q/k/v, output.dense, fc1/fc2,
 * weights_in/out (transformers modeling_dinov2.py:199-201,246-251,317-345),
 * LoraLinear (reference utils.py:68-70: the rank-r update enters as the second
 * K segment A2 = x.A^T (padded to a multiple of 64 columns), W2 = alpha*B),
 * the patch-embedding conv-as-GEMM (modeling_dinov2.py:139,148), the backbone
 * projection (dinov2_backbone.py:64-65) and the decoder projections.
 */
typedef enum { DOD_ACT_NONE = 0, DOD_ACT_GELU_ERF = 1, DOD_ACT_RELU = 2, DOD_ACT_SWIGLU = 3 } dod_act;
typedef enum { DOD_BF16 = 0, DOD_F32 = 1 } dod_dtype;

typedef struct {
  const void* a;   /* bf16 [M, K], row stride lda (lda % 8 == 0, 16-B aligned) */
  const void* w;   /* bf16 [N, K], row stride ldw (nn.Linear layout)           */
  int64_t m, n, k, lda, ldw;
  const void* a2;  /* optional second K segment (NULL: none)                  */
  const void* w2;
  int64_t k2, lda2, ldw2;
  const float* bias;     /* f32 [N] or NULL                                   */
  int32_t act;           /* dod_act.  SWIGLU: W rows interleaved in blocks of
                            128 (gate) + 128 (linear); out has N/2 columns     */
  const float* scale;    /* f32 [N] or NULL (LayerScale lambda)               */
  const void* residual;  /* f32 [*, ldr] or NULL                              */
  int64_t ldr;
  void* out;             /* [M(.), N] bf16 or f32, row stride ldo             */
  int64_t ldo;
  int32_t out_dtype;     /* dod_dtype                                         */
  /* patch-embedding row map: 0 = identity.  P > 0: GEMM row m is patch
   * (m / P, m % P); it is stored at out row  m + m/P + 1  (token row, CLS
   * skipped) and the residual row is  1 + m % P  (position embedding).       */
  int32_t patch_rows;
  /* batched problems (attention backward / training attention): batch > 1 runs `batch`
   * independent GEMMs whose A / W / out start batch_stride_* ELEMENTS apart (multiples of 8).
   * batch <= 1 (or 0) is the plain 2-D problem.  No residual / a2 / patch_rows when batched -- except
   * batch_stride_w == 0: ONE W and (optionally) one fp32 residual [M, N] shared by every batch entry
   * (m >= 512, n >= 256, fp32 output with a residual).  The patch embedding runs this way, one GEMM per
   * image, so that an image's token rows are a TMA box (out starts at the image's first patch row,
   * batch_stride_out = tokens per image * D, the residual is the position embedding).               */
  int64_t batch, batch_stride_a, batch_stride_w, batch_stride_out;
  /* second (inner) batch level, e.g. attention heads inside an image: batch_inner > 1 runs
   * batch * batch_inner problems; problem (bo, bi) starts at bo * batch_stride_* + bi * inner_stride_*.
   * Strides are in elements and multiples of 8 (16 bytes); 0 / 1 = no inner level.                    */
  int64_t batch_inner, inner_stride_a, inner_stride_w, inner_stride_out;
  /* transposed operand views (the backward pass: dW = dY^T X, dX = dY W, dK = dS^T Q, ... without
   * materialising a transpose).  a_trans != 0: `a` holds A^T, i.e. [K, M] row-major with row stride
   * lda >= M (A[m, k] = a[k * lda + m]).  w_trans != 0: `w` holds W^T, [K, N] row-major with row
   * stride ldw >= N.  The tiles are then fed to tcgen05.mma as MN-major operands.  Not combined with
   * the second K segment (a2 / w2).                                                                 */
  int32_t a_trans, w_trans;
  /* LayerNorm folded across two GEMMs (the frozen, LoRA-free encoder blocks in bf16 mode; replaces the
   * standalone nn.LayerNorm pass of modeling_dinov2.py:354,359 between a residual GEMM and the next
   * projection).  With W'' = gamma (.) W minus its row means (every row of W'' sums to zero, so the
   * product removes the row mean of h by itself):
   *     LN(h) . W^T + b  =  rstd_m * ( h . W''^T )  +  ( b + W . beta )_n
   * PRODUCER side (a residual GEMM with fp32 output, any m, n >= 256, n % 16 == 0): besides `out`
   * (the fp32 residual stream h) it writes h rounded to bf16 to `out_bf16` (row stride ldo_bf16) and
   * per-row partial sums of h and h^2 to `row_stats_out` [2 * ceil(n / 256)][M][2] f32 (one slot per
   * 128 columns, slot-major so that a warp's 32 rows are contiguous; written without atomics so the result is run-to-run reproducible).
   * dod_ln_rstd then reduces the slots to rstd[M] (a 4 MB pass), and on the CONSUMER side (bf16 h as `a`,
   * W'' as `w`, the folded bias as `bias`) `row_scale` = rstd: the epilogue multiplies the accumulator of
   * row m by row_scale[m] before bias / activation.                                                   */
  void* out_bf16;
  int64_t ldo_bf16;
  float* row_stats_out;
  const float* row_scale; /* f32 [M] or NULL */
} dod_gemm_args;
DOD_API int32_t dod_gemm_bf16(const dod_gemm_args* a, dod_stream_t stream);

/* rstd[m] = rsqrt(var_m + eps) from the per-row partial sums a producer GEMM wrote (folded LayerNorm above;
 * nn.LayerNorm statistics of modeling_dinov2.py:354,359 in fp32).
 * max_mean_ratio (optional): running maximum over all rows seen of |mean_m| * rstd_m.  The folded form feeds
 * bf16(h) instead of bf16(LN(h)) to the projection, so its rounding noise relative to the normalised signal grows
 * with this ratio (equal at 0, ~sqrt(1 + ratio^2) x otherwise); a monitor, never read on the hot path.        */
typedef struct {
  const float* row_stats; /* f32 [slots][rows][2]: partial sums of h and h^2 */
  float* rstd;            /* f32 [rows]                                      */
  int64_t slots, rows, dim;
  float eps;
  float* max_mean_ratio;  /* f32 [1] (>= 0, atomically maximised) or NULL    */
} dod_ln_rstd_args;
DOD_API int32_t dod_ln_rstd(const dod_ln_rstd_args* a, dod_stream_t stream);

/* ---- LayerNorm (HBM-bound) ----------------------------------------------
 * y = (x - mean) * rsqrt(var + eps) * gamma + beta, fp32 statistics.
 * Replaces nn.LayerNorm at modeling_dinov2.py:354,359,449 (eps 1e-6) and the
 * decoder's post-norms (deformable_attention.py:197-209, eps 1e-5).          */
typedef struct {
  const void* x;      /* [rows, d] f32 or bf16, row stride ldx                */
  int32_t x_dtype;
  const float* gamma; /* f32 [d]                                              */
  const float* beta;  /* f32 [d]                                              */
  void* y;            /* [rows, d] bf16 or f32, row stride ldy                */
  int32_t y_dtype;
  void* y2;           /* optional second copy of y in the other dtype or NULL */
  int64_t rows, d, ldx, ldy;
  float eps;
} dod_layernorm_args;
DOD_API int32_t dod_layernorm(const dod_layernorm_args* a, dod_stream_t stream);

/* ---- patch embedding front end -------------------------------------------
 * im2col of non-overlapping 14x14 patches + cast (modeling_dinov2.py:148:
 * Conv2d(3, D, 14, 14) == GEMM over k = c*196 + i*14 + j).  Also writes the
 * CLS rows  x[b*N + 0, :] = cls + pos[0]  (modeling_dinov2.py:108-112).       */
typedef struct {
  const void* pixels;   /* pixel_format 0: f32 [B, 3, H, W] in [0, 1]
                           pixel_format 1: u8  [B, H, W, 3] in 0..255 (ToTensor's /255 fused) */
  void* patches;        /* bf16 [B*P, kpad]  (kpad >= 588, kpad % 8 == 0)      */
  int64_t batch, height, width, kpad;
  const float* cls;     /* f32 [D]                                            */
  const float* pos;     /* f32 [N, D] (already resized to this H, W)          */
  float* tokens;        /* f32 [B*N, D] residual stream; only CLS rows written */
  int64_t d;
  int32_t pixel_format;
} dod_patchify_args;
DOD_API int32_t dod_patchify14(const dod_patchify_args* a, dod_stream_t stream);

/* bicubic (A=-0.75, align_corners=False) resize of the [G0,G0,D] position grid
 * to [GH,GW,D] in fp32 — F.interpolate call at modeling_dinov2.py:84-89.      */
typedef struct {
  const float* src; /* f32 [1 + g0*g0, D] (row 0 = CLS position)              */
  float* dst;       /* f32 [1 + gh*gw, D]                                     */
  int64_t g0, gh, gw, d;
} dod_pos_resize_args;
DOD_API int32_t dod_pos_resize_bicubic(const dod_pos_resize_args* a, dod_stream_t stream);

/* ---- fused multi-head self-attention (tcgen05, flash-style) --------------
 * ctx = softmax(Q K^T * scale) V per (batch, head); non-causal, no mask.
 * Replaces SDPA at modeling_dinov2.py:215-229.  q/k/v are column slices of one
 * fused projection buffer qkv[B*S, ld]: head h of q at columns q_off + h*64.
 * Two kernels: 128-key tiles with two CTAs per SM (default), 64-key tiles with four CTAs per SM
 * (environment DOD_FMHA64=1: faster below ~300 tokens); same contract, both tested on every shape.  */
typedef struct {
  const void* qkv; /* bf16 [B*S, ld]                                          */
  void* ctx;       /* bf16 [B*S, ldo], head h at columns h*64                  */
  int64_t batch, seq, heads, ld, ldo;
  int64_t q_off, k_off, v_off; /* column offsets (elements)                    */
  float scale;     /* 1/sqrt(64)                                              */
  float* lse;      /* optional f32 [B, heads, S]: log2-domain log-sum-exp of the scaled
                      scores of every query row (saved for dod_fmha_bwd), or NULL */
} dod_fmha_args;
DOD_API int32_t dod_fmha_fwd(const dod_fmha_args* a, dod_stream_t stream);

/* fused backward of dod_fmha_fwd (autograd of SDPA, modeling_dinov2.py:215-229, reached from
 * loss.backward() at train.py:1101 for the LoRA-wrapped encoder layers).  Probabilities are
 * recomputed from q, k and `lse`; nothing of size [S, S] is written to memory.
 * dK and dV are written (bf16) into dqkv at k_off / v_off; dQ is ACCUMULATED (fp32, TMA
 * reduce-add over the key tiles) into dq_acc, which the caller zero-fills and converts.   */
typedef struct {
  const void* qkv;   /* bf16 [B*S, ld]    forward input                                */
  const void* ctx;   /* bf16 [B*S, ldo]   forward output O                             */
  const void* dctx;  /* bf16 [B*S, lddo]  dO                                           */
  const float* lse;  /* f32 [B, heads, S] from dod_fmha_fwd                            */
  float* dsum;       /* f32 [B, heads, S] workspace: rowsum(dO o O)                    */
  float* dq_acc;     /* f32 [B*S, ld_dq]  zero-initialised; head h at columns h*64      */
  void* dqkv;        /* bf16 [B*S, ld_dqkv]; dK at k_off + h*64, dV at v_off + h*64     */
  int64_t batch, seq, heads, ld, ldo, lddo, ld_dq, ld_dqkv;
  int64_t q_off, k_off, v_off;
  float scale;
} dod_fmha_bwd_args;
DOD_API int32_t dod_fmha_bwd(const dod_fmha_bwd_args* a, dod_stream_t stream);

/* ---- decoder attention with few queries (generic head dim) ---------------
 * nn.MultiheadAttention core (torch) used at deformable_attention.py:232-233
 * and inside nn.TransformerDecoderLayer (detr_decoder.py:29-35):
 * out[b, q, h*dh:(h+1)*dh] = softmax(Qh Kh^T / sqrt(dh)) Vh.
 * q: [B*Lq, ldq], k/v: [B*Lk, ldk/ldv], out [B*Lq, ldo]; fp32 math.  Also used
 * for the backbone attention in fp32 mode (any sequence that fits the
 * shared-memory score tile: Lk <= ~3000).                                     */
typedef struct {
  const void* q; const void* k; const void* v; void* out;
  int64_t batch, lq, lk, heads, head_dim;
  int64_t ldq, ldk, ldv, ldo;
  float scale;
  int32_t dtype;   /* dod_dtype of q, k, v and out (f32 = fp32 mode)          */
} dod_mha_small_args;
DOD_API int32_t dod_mha_small(const dod_mha_small_args* a, dod_stream_t stream);

/* ---- "deformable" sampling (reference deformable_attention.py:100-178) ---
 * For each (b, q, head, point): loc = clamp(ref + off, 0, 1); bilinear blend of
 * 4 value rows (memory index y*w + x, CLS included), weighted by softmax over
 * points of attention logits.                                                 */
typedef struct {
  const void* value;   /* [B*hw, ldv]  value_proj(memory), bf16 or f32         */
  const float* ref;    /* f32  [B*Q, ldref]  reference point (x, y) in cols 0,1 */
  const float* offs;   /* f32  [B*Q, ldoffs] sampling_offsets(query): H*P*2    */
  const float* logits; /* f32  [B*Q, ldlog]  attention_weights(query): H*P     */
  void* out;           /* [B*Q, ldo] bf16 or f32                               */
  int64_t batch, queries, heads, points, head_dim, grid_h, grid_w;
  int64_t ldv, ldref, ldoffs, ldlog, ldo;
  int32_t value_dtype, out_dtype; /* dod_dtype                                 */
  int32_t ref_is_logit; /* 1: apply sigmoid to ref (reference_points_proj output,
                           deformable_attention.py:238) inside the kernel      */
} dod_deform_sample_args;
DOD_API int32_t dod_deform_sample(const dod_deform_sample_args* a, dod_stream_t stream);

/* ---- small elementwise helpers of the decoder ---------------------------- */
/* y = LayerNorm(x + r) (post-norm residual, deformable_attention.py:234-235)  */
typedef struct {
  const float* x; const void* r; int32_t r_dtype;
  const float* gamma; const float* beta;
  float* y; void* y_bf16; /* y_bf16 optional                                  */
  int64_t rows, d; float eps;
} dod_add_layernorm_args;
DOD_API int32_t dod_add_layernorm(const dod_add_layernorm_args* a, dod_stream_t stream);

/* out[r, :n] = act(in[r, :n]) with f32 -> f32: act 0 copy, 1 sigmoid          */
typedef struct {
  const float* in; float* out; int64_t rows, n, ld_in, ld_out; int32_t act;
} dod_rowcopy_args;
DOD_API int32_t dod_rowcopy(const dod_rowcopy_args* a, dod_stream_t stream);

/* broadcast rows: out[b*rows + r, :] = src[r, :] (query_embed repeat,
 * detr_decoder.py:59), f32 and bf16 copies                                    */
typedef struct {
  const float* src; float* out; void* out_bf16; int64_t batch, rows, d;
} dod_broadcast_rows_args;
DOD_API int32_t dod_broadcast_rows(const dod_broadcast_rows_args* a, dod_stream_t stream);

/* f32 -> bf16 cast of a [rows, cols] matrix into a (possibly wider, zero
 * padded) destination: dst[r, c] = c < cols ? scale*src[r, c] : 0             */
typedef struct {
  const float* src; void* dst; int64_t rows, cols, ld_src, ld_dst, dst_cols; float scale;
} dod_cast_pad_args;
DOD_API int32_t dod_cast_pad_bf16(const dod_cast_pad_args* a, dod_stream_t stream);

/* fp32 mode: split an f32 matrix into three bf16 terms (x = hi + mid + lo) and
 * lay the six K segments out so that ONE dod_gemm_bf16 call with fp32
 * accumulation reproduces the fp32 product to ~2^-22 relative:
 *   activation side (w_side 0): [hi | hi  | mid | hi | lo | mid]
 *   weight side     (w_side 1): [hi | mid | hi  | lo | hi | mid]
 * dst is [rows, 6*kseg] bf16 (kseg >= cols, kseg % 8 == 0, zero padded).      */
typedef struct {
  const float* src; void* dst; int64_t rows, cols, ld_src, ld_dst, kseg; int32_t w_side;
} dod_split3_args;
DOD_API int32_t dod_split3_bf16(const dod_split3_args* a, dod_stream_t stream);

/* ---- backward pass / training helpers (csrc/train.cu) --------------------
 * The reference trains through torch autograd (train.py:1101 loss.backward()); these are the
 * hand-written counterparts used by dino_detector/_train.py.  All dense contractions of the
 * backward pass go through dod_gemm_bf16 (dgrad with transposed weight copies, wgrad / attention
 * backward with dod_transpose_bf16'd operands and the batch dimension).                         */
typedef struct {
  const void* in; void* out;             /* bf16 [batch, rows, cols] -> [batch, cols, rows]     */
  int64_t rows, cols, ld_in, ld_out, batch, batch_stride_in, batch_stride_out;
  /* optional inner batch level (heads): input (bo, bi) at bo*batch_stride_in + bi*inner_stride_in,
   * output index bo*batch_inner + bi with stride batch_stride_out                               */
  int64_t batch_inner, inner_stride_in;
} dod_transpose_args;
DOD_API int32_t dod_transpose_bf16(const dod_transpose_args* a, dod_stream_t stream);

/* out[c, j] += alpha * sum_m big[m, c] * small[m, j]  (transposed: out[j, c]); r <= 64.
 * LoRA gradients (utils.py:68-70): dB = alpha * dY^T (x A^T), dA = (alpha * dY B)^T x.          */
typedef struct {
  const void* big; const void* small; float* out;   /* bf16, bf16, f32 (accumulated)             */
  int64_t m, cols, r, ld_big, ld_small, ldo;
  int32_t transposed; float alpha;
} dod_lowrank_wgrad_args;
DOD_API int32_t dod_lowrank_wgrad(const dod_lowrank_wgrad_args* a, dod_stream_t stream);

typedef struct {                                     /* out[c] += sum_m x[m, c] (bias gradients)  */
  const void* x; int32_t x_dtype; float* out; int64_t m, cols, ld;
} dod_colsum_args;
DOD_API int32_t dod_colsum(const dod_colsum_args* a, dod_stream_t stream);

/* LayerNorm backward: dx = dres + d/dx LN(x) . dy; dgamma/dbeta accumulated when non-NULL.     */
typedef struct {
  const void* dy; int32_t dy_dtype;      /* [rows, d] bf16 or f32                               */
  const float* x;                        /* f32 [rows, d] input of the forward LayerNorm        */
  const float* gamma;
  const float* dres;                     /* optional f32 [rows, d] added to dx                  */
  float* dx;                             /* f32 [rows, d]                                       */
  float* dgamma; float* dbeta;           /* optional f32 [d], accumulated                       */
  int64_t rows, d; float eps;
} dod_layernorm_bwd_args;
DOD_API int32_t dod_layernorm_bwd(const dod_layernorm_bwd_args* a, dod_stream_t stream);

typedef enum {
  DOD_ELT_CAST = 0,         /* out = a                                   */
  DOD_ELT_SCALE_COLS = 1,   /* out = a * vec[c]      (LayerScale)        */
  DOD_ELT_ADD = 2,          /* out = a + b                               */
  DOD_ELT_GELU_FWD = 3,     /* out = gelu_erf(a)                         */
  DOD_ELT_GELU_BWD = 4,     /* out = a * gelu'(b)    b = pre-activation  */
  DOD_ELT_RELU_BWD = 5,     /* out = a * (b > 0)     b = relu output     */
  DOD_ELT_SIGMOID_BWD = 6,  /* out = a * b (1 - b)   b = sigmoid output  */
  DOD_ELT_SWIGLU_FWD = 7,   /* a [rows, 2 cols] -> silu(gate) * linear   */
  DOD_ELT_SWIGLU_BWD = 8,   /* a grad [rows, cols], b pre-act [rows, 2 cols] -> out [rows, 2 cols] */
  DOD_ELT_DROPOUT = 9,      /* out = a * keep(seed, i) / (1 - p0)        */
  DOD_ELT_AXPBY = 10        /* out = vec[0] * a + vec[1] * b (b optional); vec = device scalars */
} dod_eltwise_mode;
typedef struct {
  int32_t mode;
  const void* a; int32_t a_dtype;
  const void* b; int32_t b_dtype;
  const float* vec;
  void* out; int32_t out_dtype;
  void* out2; int32_t out2_dtype;        /* DROPOUT: optional second copy (other dtype)         */
  int64_t rows, cols, ld_a, ld_b, ld_out;
  float p0; int64_t seed;
  /* optional device counter (CUDA-graph replay: the mask must change between replays although the
   * launch arguments are frozen): the effective seed is  seed + (*seed_ptr << 44).            */
  const int64_t* seed_ptr;
} dod_eltwise_args;
DOD_API int32_t dod_eltwise(const dod_eltwise_args* a, dod_stream_t stream);

/* P[row, :n] = softmax(scale * S[row, :n]) as bf16, zero padded to ldp columns.                */
typedef struct {
  const void* s; int32_t s_dtype; void* p;
  int64_t rows, n, lds, ldp; float scale; float drop_p; int64_t seed;
  const int64_t* seed_ptr;   /* optional device counter, see dod_eltwise_args */
} dod_softmax_rows_args;
DOD_API int32_t dod_softmax_rows(const dod_softmax_rows_args* a, dod_stream_t stream);
/* dS = scale * P * (dP' - rowsum(P * dP')) as bf16 (zero padded to ldds); P is the probability
 * before dropout and dP' = dP * keep / (1 - drop_p) with the mask dod_softmax_rows drew for the
 * same (seed, ldp) -- nn.MultiheadAttention's attention dropout in train mode.                    */
typedef struct {
  const void* p; const void* dp; int32_t dp_dtype; void* ds;
  int64_t rows, n, ldp, lddp, ldds; float scale; float drop_p; int64_t seed;
  const int64_t* seed_ptr;   /* optional device counter, see dod_eltwise_args */
} dod_softmax_bwd_rows_args;
DOD_API int32_t dod_softmax_bwd_rows(const dod_softmax_bwd_rows_args* a, dod_stream_t stream);

/* Backward of dod_deform_sample.  dvalue (f32, [B*hw, lddv]) is accumulated with atomics and must
 * be zeroed by the caller; dqproj rows use the fused query-projection layout
 * [d offsets 2*H*P | d logits H*P | d reference logits 2].  With bf16 values, head_dim % 8 == 0 and
 * 16-byte-aligned rows the 8-channels-per-thread kernel runs: dvalue through vector reductions, dqproj
 * from ordered sums (no atomics: identical bits on every run).                                      */
typedef struct {
  const void* value; int32_t value_dtype;
  const float* ref; const float* offs; const float* logits;
  const void* dout; int32_t dout_dtype;
  float* dvalue; float* dqproj;
  int64_t batch, queries, heads, points, head_dim, grid_h, grid_w;
  int64_t ldv, ldref, ldoffs, ldlog, lddo, lddv, lddq;
  int32_t ref_is_logit;
} dod_deform_sample_bwd_args;
DOD_API int32_t dod_deform_sample_bwd(const dod_deform_sample_bwd_args* a, dod_stream_t stream);

/* ---- Hungarian matcher ----------------------------------------------------
 * Cost matrix (reference matching.py:63,80-98):
 *   C[q, j] = (wc*(pos[q, lab_j] - neg[q, lab_j]) + wb*L1(box_q, box_j)) + wg*(-GIoU)
 * for every image b over its own n_b targets; targets are packed (CSR offsets).
 * pred_image_stride == 0 reproduces the reference quirk (matching.py:102) that
 * the rows of image 0 are used for every image.                               */
typedef struct {
  const float* logits;   /* f32 [B, Q, C]                                     */
  const float* boxes;    /* f32 [B, Q, 4] cxcywh                              */
  const int64_t* tgt_labels; /* i64 [T]                                       */
  const float* tgt_boxes;    /* f32 [T, 4]                                    */
  const int32_t* tgt_offsets; /* i32 [B+1] CSR                                */
  float* cost;           /* f32 [B, Q, max_t] (row stride max_t)              */
  int64_t batch, queries, classes, max_t;
  float w_class, w_bbox, w_giou, alpha, gamma;
  int32_t use_image0_rows; /* 1 = reference_compat                            */
} dod_match_cost_args;
DOD_API int32_t dod_match_cost(const dod_match_cost_args* a, dod_stream_t stream);

/* Rectangular linear sum assignment, bit-identical to
 * scipy.optimize.linear_sum_assignment (scipy 1.18.1 _lsap, Crouse 2016
 * shortest augmenting path in float64; call site matching.py:105) on the same
 * fp32 cost.  One problem per image: rows = queries (Q), cols = n_b targets;
 * like scipy the solver transposes when n_b < Q.  Output per image b:
 * k_b = min(Q, n_b) pairs (out_q[b, i], out_t[b, i]) sorted by query index,
 * exactly the (row_ind, col_ind) arrays scipy returns.
 * status[b]: 0 ok, 1 = cost has NaN / -inf entries or is infeasible (scipy
 * raises ValueError there).                                                    */
typedef struct {
  const float* cost;          /* f32 [B, Q, max_t]                            */
  const int32_t* tgt_offsets; /* i32 [B+1]                                    */
  int32_t* out_q;             /* i32 [B, max_k]                               */
  int32_t* out_t;             /* i32 [B, max_k]                               */
  int32_t* status;            /* i32 [B]                                      */
  int64_t batch, queries, max_t, max_k;
} dod_lsap_args;
DOD_API int32_t dod_lsap_jv(const dod_lsap_args* a, dod_stream_t stream);

/* ---- fused SetCriterion forward + backward (reference losses.py:96-241) ---------------
 * losses[0..2] = weighted loss_ce, loss_bbox, loss_giou (accumulated: zero them first);
 * dlogits = d losses[0] / d pred_logits, dboxes = d losses[1] / d pred_boxes, dboxes_giou =
 * d losses[2] / d pred_boxes (both box buffers must be zeroed first: unmatched queries get none).  The assignment is the
 * device output of dod_lsap_jv; num_boxes is a device float (already all-reduced and clamped).  */
typedef struct {
  const float* logits;        /* f32 [B, Q, C]                                */
  const float* boxes;         /* f32 [B, Q, 4] cxcywh                         */
  const int64_t* tgt_labels;  /* i64 [T]                                      */
  const float* tgt_boxes;     /* f32 [T, 4] (NULL when T == 0)                */
  const int32_t* tgt_offsets; /* i32 [B+1]                                    */
  const int32_t* out_q;       /* i32 [B, max_k] matched query per pair        */
  const int32_t* out_t;       /* i32 [B, max_k] matched target per pair       */
  const float* num_boxes;     /* f32 [1]                                      */
  int32_t* tclass;            /* i32 [B, Q] scratch: target class per query   */
  float* losses;              /* f32 [3]                                      */
  float* dlogits;             /* f32 [B, Q, C]                                */
  float* dboxes;              /* f32 [B, Q, 4]  L1 part                       */
  float* dboxes_giou;         /* f32 [B, Q, 4]  GIoU part                     */
  int64_t batch, queries, classes, max_k;
  float focal_alpha, focal_gamma, w_ce, w_bbox, w_giou;
} dod_criterion_args;
DOD_API int32_t dod_criterion(const dod_criterion_args* a, dod_stream_t stream);

/* ---- COCO post-processing (reference utils.py:195-233) ------------------------------------
 * Per image b: every (class c >= 1, query q) with sigmoid(logit) > threshold, in (c, q) order:
 * out_score[b, i], out_class[b, i], out_box[b, i] = [x1, y1, x2 - x1, y2 - y1]; counts[b] entries.  */
typedef struct {
  const float* logits; const float* boxes;      /* f32 [B, Q, C], f32 [B, Q, 4] cxcywh            */
  float* out_score; float* out_box; int32_t* out_class; int32_t* counts;   /* [B, cap], [B, cap, 4], [B, cap], [B] */
  int64_t batch, queries, classes, capacity; float threshold;
} dod_postprocess_args;
DOD_API int32_t dod_postprocess(const dod_postprocess_args* a, dod_stream_t stream);

/* ---- fused clip + Adam over flat buffers (reference train.py:1000-1004, 1104-1110) -------
 * out[0] += sum x^2;  then  coef = min(1, max_norm / (sqrt(sumsq) + 1e-6)),
 * g = coef*g + wd*p, m = b1 m + (1-b1) g, v = b2 v + (1-b2) g^2,
 * p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)      (torch.optim.Adam, L2 weight decay)  */
typedef struct { const float* x; int64_t n; float* out; } dod_sumsq_args;
DOD_API int32_t dod_sumsq(const dod_sumsq_args* a, dod_stream_t stream);
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  int64_t n, step;               /* step >= 1                                   */
  const float* grad_sumsq;       /* device float: sum of squared gradients      */
  float max_grad_norm;           /* <= 0: no clipping                           */
  float lr, beta1, beta2, eps, weight_decay;
  const int64_t* step_ptr;       /* optional device step counter (>= 1) used instead of `step`
                                    (CUDA-graph replay of the train step)        */
} dod_adam_args;
DOD_API int32_t dod_adam_step(const dod_adam_args* a, dod_stream_t stream);
/* counters[i] += delta for i < n: the device-resident step / seed counters a captured train step
 * advances once per replay.                                                                    */
typedef struct { int64_t* counters; int64_t n; int64_t delta; } dod_counter_add_args;
DOD_API int32_t dod_counter_add(const dod_counter_add_args* a, dod_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DOD_H_ */
