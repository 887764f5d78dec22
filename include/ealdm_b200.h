/*
 * ealdm_b200.h -- C ABI of libealdm_b200.so: hand-written sm_100a kernels for the
 * latent-diffusion denoising hot path of EALDM
 * (NasrinKalanat/Environment-Aware_Latent_Diffusion_Model).
 *
 * The reference has no FFI for this path: every operator below is, in the reference, a chain of
 * eager PyTorch ATen calls inside an nn.Module.forward.  Each entry point cites the reference
 * file:line whose arithmetic it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *  - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *  - activations are NHWC ("channels-last"): element (n,h,w,c) of a tensor with row pitch `ld`
 *    (elements between consecutive pixels, ld >= c) lives at ((n*H+h)*W+w)*ld + c.  A channel
 *    concatenation is therefore two producers writing disjoint column ranges of one buffer;
 *  - `dtype` selects the storage type of activations and weights: EALDM_F32 (parity mode, FFMA
 *    kernels) or EALDM_BF16 (tcgen05 tensor-core kernels, fp32 accumulation).  Biases, norm
 *    affine parameters, per-image row vectors and statistics are always fp32/fp64;
 *  - tensors, weights and workspaces are the caller's and are passed in; the library itself keeps only three small
 *    zero-initialised scratch pools per device, created by cudaMalloc on first use OUTSIDE a stream capture and handed
 *    out per stream: the counters of the GroupNorm epilogue (ealdm_conv_args::gn_gamma), the ticket counters of
 *    ealdm_colsum, and the parking slots of the opt-in stream-K schedule.  Kernels leave them zeroed.  Launches on
 *    different streams never share a slot; CUDA graphs captured on the same stream do and must not replay concurrently;
 *  - every function is asynchronous on `stream` (a cudaStream_t), re-entrant per stream, and
 *    returns 0 on success or a negative EALDM_E* code; ealdm_last_error() returns the message of
 *    the last failure on the calling thread.
 */
#ifndef EALDM_B200_H
#define EALDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EALDM_ABI_VERSION 2

typedef void* ealdm_stream_t; /* cudaStream_t */

enum { EALDM_F32 = 0, EALDM_BF16 = 1 };
enum { EALDM_ACT_NONE = 0, EALDM_ACT_SILU = 1, EALDM_ACT_GEGLU = 2, EALDM_ACT_RELU = 3,
       EALDM_ACT_SOFTMAX4 = 4 /* softmax over every 4 consecutive output columns of base-2 logits (wi_* below) */ };
enum { EALDM_IMPL_AUTO = 0, EALDM_IMPL_SIMT = 1, EALDM_IMPL_TCGEN05 = 2 };
enum {
  EALDM_OK = 0,
  EALDM_EINVAL = -1,   /* bad argument (shape, alignment, unsupported combination) */
  EALDM_ECUDA = -2,    /* CUDA runtime / driver error */
  EALDM_EUNSUPPORTED = -3
};

/* ---- library ------------------------------------------------------------------------------ */
int ealdm_abi_version(void);
const char* ealdm_last_error(void);
/* 0 if the current device is compute capability 10.x (sm_100a code is loadable), else <0. */
int ealdm_device_check(void);
/* number of kernel launches issued through this library by the calling process so far */
int64_t ealdm_launch_count(void);
/*
 * Schedule switches of the tcgen05 conv / linear kernel (results are identical in every setting; they exist for
 * A/B measurements and for the parity tests that pin one schedule against the other).  The defaults come from the
 * environment variables named below, read once.  Returns the previous value, or EALDM_EINVAL for an unknown option.
 */
enum {
  EALDM_TC_OPT_CTA2 = 0,         /* EALDM_TC_CTA2: 0 one CTA per tile, 1 (default) CTA pairs (cta_group::2) where
                                    the reduction is long enough (K >= 1024), 2 pairs for every even tile count */
  EALDM_TC_OPT_WIDE = 1,         /* EALDM_TC_WIDE: 1 (default) 128-byte-row epilogue passes when there is no
                                    residual / shadow output, 0 one 32-column unit per TMA store */
  EALDM_TC_OPT_RELAXED_WAIT = 2, /* EALDM_TC_RELAXED_WAIT: 1 (default) one TMA store group may stay in flight */
  EALDM_TC_OPT_BN = 3,           /* EALDM_TC_BN: 0 (default) N tile chosen by wave count, 128 / 256 forced (for
                                    n_out > 128), so that small test problems reach the 256-wide paths */
  EALDM_TC_OPT_GELU_ERF = 4,     /* EALDM_TC_GELU_ERF: the ONE switch that changes results (by <= 4.7e-4 absolute per
                                    gated activation): 0 (default) GEGLU epilogues evaluate GELU in its tanh form on the
                                    hardware tanh, 1 the exact-erf form (rational approximation, error 1.7e-6) */
  EALDM_TC_OPT_STREAMK = 5       /* EALDM_TC_STREAMK: 0 (default) off; 1 CTA-pair launches of 3x3 convolutions whose last
                                    wave would leave more than 1/8 of the clusters idle deal the k-blocks of their last
                                    (clusters + remainder) tiles evenly to the clusters (a split tile is resumed from
                                    its parked fp32 accumulator: results bit-identical to the unsplit schedule);
                                    measured +3-4 % on such a launch timed alone and neutral inside the power-capped
                                    forward, hence opt-in; 2 whenever there is a remainder (tests) */
};
int ealdm_tc_set_option(int option, int value);
/*
 * Programmatic dependent launch between consecutive kernels of one stream (default off; environment EALDM_PDL=1
 * switches it on): the kernels of the forward path are then launched with the programmatic-stream-serialization
 * attribute and execute griddepcontrol.wait before their first global-memory access, so that a prologue overlaps the
 * tail of its predecessor.  Results are identical either way (measured: no gain under CUDA-graph replay).  Returns
 * the previous setting.
 */
int ealdm_set_pdl(int enabled);

/* ---- convolution / linear as implicit GEMM ------------------------------------------------- */
/*
 * out[m, j] = epilogue( sum_s sum_{kh,kw,c} src_s[n, oh*stride+kh-pad, ow*stride+kw-pad, c]
 *                                   * weight[j, koff_s + (kh*ksize+kw)*c_s + c] )
 * with m = (n*h_out+oh)*w_out+ow.  Out-of-range taps read zero.  `upsample` = 1 reads the source
 * through a nearest-neighbour 2x upsampling (reference: F.interpolate(scale_factor=2)).
 * Up to two sources are accumulated into one output (a 3x3 conv plus the 1x1 skip conv of a
 * ResBlock); their weights are concatenated along K in `weight`.
 * A Linear layer is the case n=1, h=1, w=rows, ksize=1.
 *
 * epilogue: v = acc + bias[j] + rowvec[n*ld_rowvec + j]; v = act(v); v += residual[m*ld_res+j];
 * act == GEGLU: weight rows are pre-interleaved in blocks of 32 (16 value rows then their 16 gate
 * rows); the output has n_out/2 columns: value * gelu_erf(gate).
 *
 * Replaces: nn.Conv2d / nn.Linear and the elementwise ops around them in
 *   ldm/modules/diffusionmodules/openaimodel.py:255-275 (ResBlock._forward), :109-119, :158-160
 *   ldm/modules/attention.py:37-64 (GEGLU / FeedForward), :170-193 (to_q/k/v/out), :250-261
 *   ldm/modules/diffusionmodules/model.py:42-79, :116-141, :175-202
 */
typedef struct {
  const void* x;
  int64_t n, h, w, c;
  int64_t ld;
  int32_t ksize;    /* 1 or 3 */
  int32_t stride;   /* 1 or 2 */
  int32_t pad;      /* zero padding on the top/left edge (bottom/right is implied by h_out/w_out) */
  int32_t upsample; /* 0 or 1 */
} ealdm_conv_src;

typedef struct {
  int32_t dtype;
  int32_t impl;
  int32_t n_src;
  int32_t act;
  ealdm_conv_src src[2];
  const void* weight; /* [n_out, k_total], row-major, `dtype` */
  int64_t n_out;
  int64_t k_total;
  int64_t h_out, w_out;
  const float* bias;   /* [n_out] or NULL */
  const float* rowvec; /* [n, ld_rowvec] or NULL */
  int64_t ld_rowvec;
  const void* residual; /* [M, ld_res] `dtype`, or NULL */
  int64_t ld_res;
  void* out;
  int64_t ld_out;
  int32_t out_f32; /* 1: `out` is float regardless of dtype */
  int32_t res_f32; /* 1: `residual` is float regardless of dtype (fp32 residual stream) */
  void* out2;      /* optional second copy of the result in `dtype` (the bf16 operand shadow of an
                      fp32 residual-stream tensor), pitch ld_out2; NULL if unused */
  int64_t ld_out2;
  /* optional GroupNorm partial statistics of the result, produced by the epilogue (tcgen05 path only, h_out*w_out a
     multiple of 32, w_out a power of two): gn_partial[((img*(hw/32) + chunk)*gn_ld + j/8)*2 + {0,1}] = {sum, sum of
     squares} of output channels [8*(j/8), 8*(j/8)+8) over the 32 pixels [32*chunk, 32*chunk+32) of image img.
     gn_partial points at the octet of output column 0; a later ealdm_group_norm that is given the same buffer
     skips its statistics pass.  Replaces the first half of GroupNorm32 (util.py:214-216). */
  float* gn_partial;
  int64_t gn_ld; /* octets (float2 entries) per (image, chunk) row of the partial buffer */
  /* weight_adjoint = 1 (tcgen05 path, n_src = 1): the call computes the DATA GRADIENT of a forward layer whose packed
     matrix is given unchanged: `weight` is the forward [src.c rows, ksize^2 * n_out columns] matrix (row pitch
     ld_weight, so a column window of a fused 3x3 + 1x1 matrix works) and
       out[m, j] = sum_{tap, c} src[..tap.., c] * weight[c, (ksize^2 - 1 - tap) * n_out + j]
     i.e. the transposed, tap-flipped weights are never materialised (the B operand is read MN-major). */
  int32_t weight_adjoint;
  /* upsample_phases = 1 (tcgen05 path, n_src = 1, src[0].upsample = 1, ksize 3 / stride 1 / pad 1): the convolution over
     the nearest-2x upsampled source (Upsample, openaimodel.py:109-119) is evaluated as its four OUTPUT PHASES: output
     pixel (2y + py, 2x + px) only sees source rows {y + py - 1, y + py} and columns {x + px - 1, x + px}, so each phase
     is a 2x2 convolution over the source itself with the 3x3 taps that coincide pre-summed.  `weight` is then
     [n_out, 4 phases (py, px) x 4 taps (row, column) x src.c] (k_total = 16 * src.c); 4/9 of the FLOPs of the 3x3
     form, and the upsampled tensor never exists.  Bias, row vector, shadow and GroupNorm partials as usual. */
  int32_t upsample_phases;
  int64_t ld_weight;
  /* LayerNorm FOLDED into the two GEMMs around it (tcgen05 path, linear layers: h_out = 1, src[0].n = 1).  Replaces the
     stand-alone nn.LayerNorm passes of BasicTransformerBlock._forward (ldm/modules/attention.py:203-205,211-215):
       LayerNorm(x) W^T + b  ==  rstd (x (W . gamma)^T) - rstd mu ((W . gamma) 1) + (W beta + b)
     PRODUCER (the GEMM that writes the fp32 token stream x, typically with its bf16 shadow in out2): ln_partial_out
     [rows, ealdm_conv_ln_parts(args)] float2 receives {sum, sum of squares} of disjoint column sets of every row.
     CONSUMER (the GEMM that followed the LayerNorm): src = the RAW stream (bf16), weight = W . gamma (packed by the
     caller), bias = W beta + b, ln_c1[n_out] = row sums of the packed weight, ln_partial_in / ln_parts_in = the
     producer's buffer, ln_channels = the LayerNorm width, ln_eps its epsilon.  Works with act NONE and GEGLU. */
  float* ln_partial_out;
  const float* ln_partial_in;
  int64_t ln_parts_in;
  const float* ln_c1;
  int64_t ln_channels;
  float ln_eps;
  /* PER-IMAGE B operand (wi_tokens = T > 0; tcgen05 path, n_src = 1, ksize 1, h*w a multiple of 128 so that a 128-row
     tile lies in one image).  Cross-attention on a short context collapses onto the context (DESIGN.md section 4;
     replaces to_q -> einsum/softmax/einsum -> to_out of CrossAttention.forward, ldm/modules/attention.py:170-193):
       scores[m, (h, j)] = x[m, :] . U_n[(h, j), :],     U_n[(h, j), c]  = sum_d Wq[h d, c] k_n[j, h d]  (scale folded)
       out[m, :]         = P[m, :] Zt_n + b,             Zt_n[(h, j), c] = sum_d Wo[c, h d] v_n[j, h d]
     with U and Zt rows of ONE projection of the context, `weight`[(n T + j) wi_ld + h wi_head_stride + c]:
       weight_adjoint = 0:  B_n = U_n   (n_out = wi_heads * T logits: 32, 64, 96 or 128; k_total = src.c; act SOFTMAX4)
       weight_adjoint = 1:  B_n = Zt_n  (src.c = wi_heads * T probabilities, n_out <= wi_head_stride)
     The 4-D TMA box (c, j, h, n) gathers the (h, j) rows of image n straight from that projection. */
  int32_t wi_tokens;
  int32_t wi_heads;
  int32_t gn_unit; /* channels per gn_partial entry: 0 or 8 = octets (above), 4 = quads (GroupNorm groups of 4 channels:
                      the 128-channel level of the autoencoder); gn_partial then points at the entry of column 0 and
                      gn_ld counts quads */
  int64_t wi_ld;
  int64_t wi_head_stride;
  /* LayerNorm APPLIED by the producing epilogue (ln_gamma != NULL; tcgen05 path, n_out == 256 so that one CTA holds whole
     rows, fp32 `out`, act NONE): `out` receives the fp32 result as usual and `out2` (bf16, pitch ld_out2) receives
     LayerNorm(result) * ln_gamma + ln_beta over the 256 columns (eps = ln_eps) instead of a copy -- the stand-alone
     nn.LayerNorm pass of BasicTransformerBlock._forward (ldm/modules/attention.py:211-215) that would re-read the fp32
     stream disappears.  Two passes over the accumulator in TMEM; the statistics are E[x], E[x^2] in fp32. */
  const float* ln_gamma;
  const float* ln_beta;
  /* GroupNorm (+ SiLU) APPLIED by the producing epilogue (gn_gamma != NULL; tcgen05 path, n_out >= 256 in 256-column
     tiles, gn_groups groups of 8 / 16 / 32 channels, h_out * w_out a multiple of 32 and at most 2048 -- the tiles of an
     image wait for each other and must be in flight together --, gn_partial given, act NONE).
     Replaces the stand-alone GroupNorm32 + SiLU pass in front of a ResBlock's second convolution and the GroupNorm in
     front of a SpatialTransformer (ldm/modules/diffusionmodules/openaimodel.py:255-275, ldm/modules/attention.py:254):
       gn_only = 0:  `out` (fp32) receives the result as usual, `out2` (bf16) act(GroupNorm(result) * gamma + beta)
       gn_only = 1:  `out` (bf16) receives act(GroupNorm(result) * gamma + beta); the result itself is never written
     with act = SiLU if gn_silu.  The statistics are the fp64 fold of the epilogue's own fp32 {sum, sum of squares}
     partials (gn_partial); tiles of one image wait for each other through counters in library-owned scratch. */
  const float* gn_gamma;
  const float* gn_beta;
  float gn_eps;
  int32_t gn_groups;
  int32_t gn_silu;
  int32_t gn_only;
} ealdm_conv_args;

int ealdm_conv(const ealdm_conv_args* a, ealdm_stream_t stream);
/* partials per row that a launch with these args writes to ln_partial_out: one per 32 output columns, whatever the
   schedule, so that every N tile / CTA-pair choice hands the consumer bit-identical sums */
int64_t ealdm_conv_ln_parts(const ealdm_conv_args* a);

/*
 * Fused GEGLU FeedForward + residual of one transformer block at model width c = 256 (hidden = 4c = 1024):
 *   out = (value * gelu_erf(gate)) * w2^T + b2 + residual,   [value | gate] = x * w1^T + b1
 * in ONE kernel: the 8c-wide projection and the 4c-wide gated activation stay in TMEM / shared memory.
 * `w1` / `b1` are the ROW-INTERLEAVED GEGLU projection (blocks of 32 rows: 16 value rows, then their 16 gate rows --
 * the layout ealdm_conv's EALDM_ACT_GEGLU epilogue uses), [2*hidden, c] bf16 / [2*hidden] fp32; `w2` is [c, hidden]
 * bf16, `b2` [c] fp32; `x` bf16 [rows, ld_x]; `residual` fp32 [rows, ld_res]; `out` fp32 or bf16 [rows, ld_out].
 * Bit-identical to ealdm_conv(act = GEGLU) followed by ealdm_conv(bias, fp32 residual).
 * Replaces: FeedForward.forward / GEGLU.forward (ldm/modules/attention.py:37-64) and the `+ x` of
 *   BasicTransformerBlock._forward (attention.py:214).
 */
typedef struct {
  const void* x;
  int64_t rows, ld_x;
  int64_t c, hidden;
  const void* w1;
  const float* b1;
  const void* w2;
  const float* b2;
  const float* residual;
  int64_t ld_res;
  void* out;
  int64_t ld_out;
  int32_t out_f32;
  int32_t reserved;
} ealdm_ff_fused_args;

int ealdm_ff_geglu_fused(const ealdm_ff_fused_args* a, ealdm_stream_t stream);

/* ---- normalisation -------------------------------------------------------------------------- */
/*
 * GroupNorm over (c/groups, hw) per image with biased variance, affine, optional SiLU.
 * `workspace` is caller-provided scratch of ealdm_group_norm_workspace_bytes(n, hw, c) bytes holding
 * per-(image, pixel chunk, 4-channel vector) partial sums; they are combined in a fixed order (fp64),
 * so the result is bit-reproducible run to run (no atomics).
 * Replaces: GroupNorm32 + SiLU, ldm/modules/diffusionmodules/util.py:214-216,
 *   openaimodel.py:201-205,225-232,682-686; Normalize (eps 1e-6) attention.py:76-77, model.py:38-39.
 */
typedef struct {
  int32_t dtype; /* type of y (and of x unless x_f32) */
  int32_t act;   /* EALDM_ACT_NONE or EALDM_ACT_SILU */
  const void* x;
  int64_t n, hw, c, ld_x;
  int32_t groups; /* bit 31 clear; see x_f32 below */
  float eps;
  const float* gamma;
  const float* beta;
  void* y;
  int64_t ld_y;
  void* workspace;
  int32_t x_f32; /* 1: x is float regardless of dtype (fp32 residual stream -> bf16 GEMM operand) */
  int32_t partial_unit; /* channels per `partial` entry: 0 or 8 = octets, 4 = quads (ealdm_conv_args::gn_unit) */
  float* stats_out; /* optional [n, groups, 2] = (mean, rstd) per (image, group), saved for the backward */
  const float* partial; /* optional: partial statistics of x written by the producing ealdm_conv (gn_partial) or by
                           ealdm_gn_partial; the kernel then only streams x once (no reduction pass) */
  int64_t partial_ld;
} ealdm_group_norm_args;

int64_t ealdm_group_norm_workspace_bytes(int64_t n, int64_t hw, int64_t c);
/* stand-alone producer of the partial statistics (for tensors that did not come out of the tcgen05 epilogue) */
int ealdm_gn_partial(const void* x, int64_t ld_x, int32_t dtype, int64_t n, int64_t hw, int64_t c, float* partial,
                     int64_t partial_ld, ealdm_stream_t stream);
int ealdm_group_norm(const ealdm_group_norm_args* a, ealdm_stream_t stream);

/* LayerNorm over the last dimension. Replaces nn.LayerNorm in attention.py:203-205,211-215. */
typedef struct {
  int32_t dtype;   /* type of y (and of x unless x_f32) */
  int32_t x_f32;   /* 1: x is float regardless of dtype (fp32 residual stream -> bf16 GEMM operand) */
  const void* x;
  int64_t rows, c, ld_x;
  float eps;
  int32_t reserved2;
  const float* gamma;
  const float* beta;
  void* y;
  int64_t ld_y;
} ealdm_layer_norm_args;

int ealdm_layer_norm(const ealdm_layer_norm_args* a, ealdm_stream_t stream);

/* ---- attention -------------------------------------------------------------------------------- */
/*
 * out[b,i,h,:] = softmax_j(scale * q[b,i,h,:].k[b,j,h,:]) . v[b,j,h,:]
 * element (b,i,h,d) of q is at q[(b*n_q+i)*ld_q + h*head_stride_q + d]; k and v likewise with
 * ld_kv / head_stride_kv; out is [(b*n_q+i)*ld_out + h*head_dim + d].
 * Replaces: CrossAttention.forward attention.py:170-193 (self- and cross-attention),
 *   QKVAttentionLegacy.forward openaimodel.py:356-372, AttnBlock model.py:175-199.
 */
typedef struct {
  int32_t dtype;
  int32_t impl;
  const void* q;
  const void* k;
  const void* v;
  int64_t ld_q, ld_kv;
  int64_t head_stride_q, head_stride_kv;
  int64_t batch, heads, n_q, n_kv, head_dim;
  float scale;
  int32_t reserved;
  void* out;
  int64_t ld_out;
  float* lse; /* optional [batch, heads, n_q]: log2 of sum_j 2^(scale*log2(e)*q_i.k_j), saved for the backward
                 (bf16 tensor-core path with n_kv % 64 == 0 only; NULL otherwise) */
} ealdm_attention_args;

int ealdm_attention(const ealdm_attention_args* a, ealdm_stream_t stream);

/* ---- timestep embedding, layout ---------------------------------------------------------------- */
/*
 * out[i, :] = [cos(t_i*f) | sin(t_i*f)] with f = freqs[0:dim/2] (device, fp32), util.py:151-171.
 * The host computes freqs = exp(-ln(max_period)*arange(half)/half) once, exactly as the reference
 * does, so that the arguments t*f are bit-identical to the reference's.  out is [n, dim] `dtype`.
 */
int ealdm_timestep_embedding(const int64_t* t, int64_t n, int32_t dim, const float* freqs,
                             int32_t dtype, void* out, ealdm_stream_t stream);

/* fp32 NCHW <-> `dtype` NHWC (pitch ld).  The reference keeps NCHW fp32 at every module boundary. */
int ealdm_nchw_to_nhwc(const float* x, int64_t n, int64_t c, int64_t h, int64_t w, int32_t dtype,
                       void* y, int64_t ld_y, ealdm_stream_t stream);
int ealdm_nhwc_to_nchw(const void* x, int64_t ld_x, int32_t dtype, int64_t n, int64_t c, int64_t h,
                       int64_t w, float* y, ealdm_stream_t stream);
/* nearest-neighbour 2x upsampling NHWC -> NHWC (openaimodel.py:116, model.py:54) */
int ealdm_upsample_nearest2x(const void* x, int64_t ld_x, int32_t dtype, int64_t n, int64_t h,
                             int64_t w, int64_t c, void* y, int64_t ld_y, ealdm_stream_t stream);
/* strided 2-D copy / cast between dtypes: y[r, 0:c] = x[r, 0:c] */
int ealdm_copy2d(const void* x, int64_t ld_x, int32_t dtype_x, void* y, int64_t ld_y,
                 int32_t dtype_y, int64_t rows, int64_t c, ealdm_stream_t stream);
/* row softmax in place over [rows, c] (pitch ld) after multiplying by `scale` (model.py:190-192) */
int ealdm_softmax_rows(void* x, int64_t ld, int32_t dtype, int64_t rows, int64_t c, float scale,
                       ealdm_stream_t stream);

/* ---- EALDM conditioner `UnetCond` (STDiff/models.py:411-539), fp32 ----------------------------------- */
/*
 * ConditioningTransform.forward + CondScale.forward (models.py:203-236, 298-309): features[t] = [cos, sin] pairs of
 * 2 pi freqs[i] time[t] (include_lin: pair 0 = (1, lin_lr * time[t])), out[t, j] = sum_k features[t, k] * weight[j, k]
 * * gain.  `features` ([t_count, 2 * n_freq]) may be NULL.
 */
int ealdm_fourier_style(const float* time, int64_t t_count, const float* freqs, int32_t n_freq, int32_t include_lin,
                        float lin_lr, const float* weight, int64_t n_out, float gain, float* features, float* out,
                        ealdm_stream_t stream);
/*
 * One torch.nn.LSTM step (WeatherLSTM, models.py:312-336) for `batch` sequences: gates = w_ih x + b_ih + rec + b_hh with
 * gate order i, f, g, o; rec = w_hh h_prev [batch, 4 * hidden] (from ealdm_conv as a linear layer) or NULL at step 0,
 * c_prev NULL = zero state.  n_in <= 64.
 */
int ealdm_lstm_cell(const float* x, int64_t ld_x, int64_t batch, int64_t n_in, const float* w_ih, const float* b_ih,
                    const float* b_hh, const float* rec, const float* c_prev, int64_t hidden, float* h_out,
                    int64_t ld_h, float* c_out, ealdm_stream_t stream);
/* AdaIN.forward (models.py:362-377) on an NHWC map [n, hw, c]: InstanceNorm2d (biased variance, eps) then
 * x * (1 + gamma) + beta, style[n] = [gamma(c) | beta(c)]; c <= 32. */
int ealdm_adain(const float* x, int64_t ld_x, int64_t n, int64_t hw, int64_t c, const float* style, int64_t ld_style,
                float eps, float* y, int64_t ld_y, ealdm_stream_t stream);
/* nn.BatchNorm2d (+ ReLU) of conv_cat (models.py:480-483) over [rows, c], c <= 32: training != 0 normalises with the
 * statistics of the batch and writes {mean(c), biased var(c)} to batch_stats (may be NULL), else the running buffers. */
int ealdm_batch_norm_relu(const float* x, int64_t ld_x, int64_t rows, int64_t c, const float* gamma, const float* beta,
                          const float* running_mean, const float* running_var, int32_t training, float eps,
                          int32_t relu, float* y, int64_t ld_y, float* batch_stats, ealdm_stream_t stream);

/* ---- adjoints of the conditioner pieces: the reference trains UnetCond together with the UNet
 * (`cond_stage_trainable: true`, ldm/models/diffusion/ddpm.py:1409-1415).  The encoder that feeds AdaIN is frozen, so
 * ealdm_adain_bwd returns the STYLE gradient only: dstyle[n, 2c] = [dgamma | dbeta] (STDiff/models.py:369-377). */
int ealdm_adain_bwd(const float* x, int64_t ld_x, int64_t n, int64_t hw, int64_t c, const float* dy, int64_t ld_dy,
                    float eps, float* dstyle, int64_t ld_dstyle, ealdm_stream_t stream);
/* BatchNorm2d + ReLU adjoint (nn.BatchNorm2d of conv_cat, STDiff/models.py:480-483): `y` is the forward output; dgamma /
 * dbeta ([c], optional) are ACCUMULATED into. */
int ealdm_batch_norm_relu_bwd(const float* x, int64_t ld_x, int64_t rows, int64_t c, const float* gamma,
                              const float* running_mean, const float* running_var, int32_t training, float eps,
                              int32_t relu, const float* y, int64_t ld_y, const float* dy, int64_t ld_dy, float* dx,
                              int64_t ld_dx, float* dgamma, float* dbeta, ealdm_stream_t stream);
/* one nn.LSTM step backwards (gates recomputed from the forward's inputs): dgates [batch, 4 hidden] (pre-activation,
 * order i, f, g, o) and dc_prev [batch, hidden] (optional) from dh, dc_next (optional); STDiff/models.py:323-336. */
int ealdm_lstm_cell_bwd(const float* x, int64_t ld_x, int64_t batch, int64_t n_in, const float* w_ih, const float* b_ih,
                        const float* b_hh, const float* rec, const float* c_prev, int64_t hidden, const float* dh,
                        int64_t ld_dh, const float* dc_next, float* dgates, float* dc_prev, ealdm_stream_t stream);
/* dx = dy * (y > 0) [* mask]: ReLU (+ inverted-dropout mask) adjoint of the conditioner's MLPs */
int ealdm_relu_bwd(const float* y, int64_t ld_y, const float* dy, int64_t ld_dy, const float* mask, int64_t ld_mask,
                   int64_t rows, int64_t c, float* dx, int64_t ld_dx, ealdm_stream_t stream);

/* ---- sampler / diffusion elementwise (fp32, bit-exact with the reference's op sequence) -------- */
/*
 * DDIMSampler.p_sample_ddim, ldm/models/diffusion/ddim.py:173-203, on fp32 tensors of `numel`
 * elements:
 *   e      = e_uncond ? e_uncond + cfg_scale*(e_cond - e_uncond) : e_cond
 *   pred   = (x - sqrt_one_minus_at*e) / sqrt_at
 *   x_prev = sqrt_a_prev*pred + dir_coef*e + (sigma_t*noise)*temperature
 * The scalars are computed by the host exactly as the reference does (fp32 torch.full tensors).
 * noise may be NULL only when sigma_t == 0.  e_out (optional) receives the guided eps.
 */
typedef struct {
  const float* x;
  const float* e_uncond;
  const float* e_cond;
  const float* noise;
  float* x_prev;
  float* pred_x0;
  float* e_out;
  int64_t numel;
  float cfg_scale;
  float sqrt_one_minus_at;
  float sqrt_at;
  float sqrt_a_prev;
  float dir_coef;
  float sigma_t;
  float temperature;
  int32_t reserved;
} ealdm_ddim_step_args;

int ealdm_ddim_step(const ealdm_ddim_step_args* a, ealdm_stream_t stream);

/* DDPM.q_sample, ldm/models/diffusion/ddpm.py:276-279: out = sa[t[b]]*x0 + s1a[t[b]]*noise */
int ealdm_q_sample(const float* x0, const float* noise, const int64_t* t,
                   const float* sqrt_alphas_cumprod, const float* sqrt_one_minus_alphas_cumprod,
                   int64_t batch, int64_t per_sample, float* out, ealdm_stream_t stream);

/*
 * LatentDiffusion.p_losses tail, ddpm.py:1040-1060: guided = e_uncond + s*(e_cond-e_uncond)
 * (or e_cond if e_uncond is NULL); loss_simple[b] = mean((guided - target)^2) over the sample.
 */
int ealdm_cfg_mse(const float* e_uncond, const float* e_cond, const float* target, float cfg_scale,
                  int64_t batch, int64_t per_sample, float* loss_simple, ealdm_stream_t stream);

/* ======================================================================================================
 * Backward pass (LatentDiffusion.p_losses -> loss.backward(), ldm/models/diffusion/ddpm.py:1036-1078;
 * the reference gets every adjoint below from torch.autograd over the modules cited at the forward
 * entry points).  Data gradients of conv / linear layers reuse ealdm_conv with transposed (and, for
 * 3x3, spatially flipped) weights; the entry points here are the remaining adjoints.
 * Parameter gradients are fp32 and, where stated, ACCUMULATE into the destination (.grad semantics).
 * ====================================================================================================== */

enum { EALDM_WGRAD_PACKED = 0, EALDM_WGRAD_OIHW = 1 };

/*
 * Weight gradient of ealdm_conv for one source:
 *   dw[j, (kh, kw, ci)] = sum_{n, oh, ow} dy[(n, oh, ow), j] * x[n, oh*stride+kh-pad, ow*stride+kw-pad, ci]
 * layout PACKED: dw[j*ld_dw + (kh*ksize+kw)*c + ci] (the K-major layout of ealdm_conv's `weight`, so the two
 *   sources of a fused 3x3 + 1x1-skip convolution write disjoint column windows of one matrix);
 * layout OIHW:   dw[j*ld_dw + (ci*ksize+kh)*ksize+kw] (nn.Conv2d.weight / nn.Linear.weight).
 * `workspace` holds the per-split partial sums (ealdm_conv_wgrad_workspace_bytes); the splits are added
 * in a fixed order (no atomics).  A Linear layer is n=1, h=1, w=rows, ksize=1.
 * Replaces: the weight gradient of nn.Conv2d / nn.Linear (see ealdm_conv for the module list).
 */
typedef struct {
  int32_t dtype; /* storage type of x and dy */
  int32_t impl;
  ealdm_conv_src src; /* the forward input with the forward conv's ksize / stride / pad; upsample = 0 */
  const void* dy;     /* [n*h_out*w_out, ld_dy] */
  int64_t ld_dy;
  int64_t n_out;
  int64_t h_out, w_out;
  float* dw;
  int64_t ld_dw;
  int32_t layout;
  int32_t accumulate; /* 1: dw += result */
  void* workspace;
  int64_t workspace_bytes;
} ealdm_conv_wgrad_args;

int64_t ealdm_conv_wgrad_workspace_bytes(const ealdm_conv_wgrad_args* a);
int ealdm_conv_wgrad(const ealdm_conv_wgrad_args* a, ealdm_stream_t stream);

/*
 * GroupNorm(+SiLU) backward.  dx = d(loss)/dx (+ add + add2, fp32 gradients of parallel branches that join
 * at x: the residual connection and the UNet skip concatenation); dx2 = optional `dtype` shadow of dx.
 * dgamma / dbeta accumulate.  `stats` is the forward's stats_out.
 */
typedef struct {
  int32_t dtype; /* type of dy and dx2; of x unless x_f32; of dx unless dx_f32 */
  int32_t act;
  const void* x;
  int64_t n, hw, c, ld_x;
  int32_t groups;
  int32_t x_f32;
  const float* stats;
  const float* gamma;
  const float* beta;
  const void* dy;
  int64_t ld_dy;
  const float* add;
  int64_t ld_add;
  const float* add2;
  int64_t ld_add2;
  void* dx;
  int64_t ld_dx;
  int32_t dx_f32;
  int32_t reserved;
  void* dx2;
  int64_t ld_dx2;
  float* dgamma;
  float* dbeta;
  void* workspace; /* ealdm_group_norm_bwd_workspace_bytes(n, hw, c) */
} ealdm_group_norm_bwd_args;

int64_t ealdm_group_norm_bwd_workspace_bytes(int64_t n, int64_t hw, int64_t c);
int ealdm_group_norm_bwd(const ealdm_group_norm_bwd_args* a, ealdm_stream_t stream);

/* LayerNorm backward (mean / rstd recomputed from x); same add / dx2 / accumulate conventions. */
typedef struct {
  int32_t dtype;
  int32_t x_f32;
  const void* x;
  int64_t rows, c, ld_x;
  float eps;
  int32_t dx_f32;
  const float* gamma;
  const void* dy;
  int64_t ld_dy;
  const float* add;
  int64_t ld_add;
  void* dx;
  int64_t ld_dx;
  void* dx2;
  int64_t ld_dx2;
  float* dgamma;
  float* dbeta;
  void* workspace; /* ealdm_layer_norm_bwd_workspace_bytes(rows, c) */
} ealdm_layer_norm_bwd_args;

int64_t ealdm_layer_norm_bwd_workspace_bytes(int64_t rows, int64_t c);
int ealdm_layer_norm_bwd(const ealdm_layer_norm_bwd_args* a, ealdm_stream_t stream);

/*
 * Attention backward: dq, dk, dv of ealdm_attention given out and dout (same addressing as the forward;
 * dq is laid out like q with head_stride_dq, dk / dv like k / v with head_stride_dkv).
 */
typedef struct {
  int32_t dtype;
  int32_t impl;
  const void* q;
  const void* k;
  const void* v;
  const void* out;
  const void* dout;
  int64_t ld_q, ld_kv, ld_out, ld_dout;
  int64_t head_stride_q, head_stride_kv;
  int64_t batch, heads, n_q, n_kv, head_dim;
  float scale;
  int32_t reserved;
  void* dq;
  int64_t ld_dq, head_stride_dq;
  void* dk;
  void* dv;
  int64_t ld_dkv, head_stride_dkv;
  void* workspace;
  int64_t workspace_bytes;
  const float* lse; /* optional: the forward's lse; enables the tensor-core backward (bf16, head_dim 32,
                       n_q and n_kv multiples of 64); without it the log-sum-exp is recomputed (SIMT path) */
} ealdm_attention_bwd_args;

int64_t ealdm_attention_bwd_workspace_bytes(const ealdm_attention_bwd_args* a);
int ealdm_attention_bwd(const ealdm_attention_bwd_args* a, ealdm_stream_t stream);

/* GEGLU on the natural [value | gate] layout (attention.py:37-44): out = value * gelu_erf(gate), and its
 * adjoint dpre = [dout * gelu(gate) | dout * value * gelu'(gate)].  pre / dpre are [rows, 2*inner]. */
int ealdm_geglu(const void* pre, int64_t ld_pre, int32_t dtype, int64_t rows, int64_t inner, void* out,
                int64_t ld_out, ealdm_stream_t stream);
int ealdm_geglu_bwd(const void* pre, int64_t ld_pre, const void* dout, int64_t ld_dout, int32_t dtype,
                    int64_t rows, int64_t inner, void* dpre, int64_t ld_dpre, ealdm_stream_t stream);

/* y = silu(x) and dx = dy * silu'(x) for an fp32 pre-activation x (time_embed / emb_layers MLP). */
int ealdm_silu(const float* x, int64_t ld_x, int64_t rows, int64_t c, int32_t dtype, void* y, int64_t ld_y,
               ealdm_stream_t stream);
int ealdm_silu_bwd(const float* x, int64_t ld_x, const void* dy, int64_t ld_dy, int32_t dtype, int64_t rows,
                   int64_t c, void* dx, int64_t ld_dx, ealdm_stream_t stream);

/* out[s, 0:c] (+)= column sums of rows [s*rows_per_seg, (s+1)*rows_per_seg) of x: bias gradients (segs = 1)
 * and the per-image timestep-embedding gradient of a ResBlock (segs = n). */
int64_t ealdm_colsum_workspace_bytes(int64_t segs, int64_t rows_per_seg, int64_t c);
int ealdm_colsum(const void* x, int64_t ld_x, int32_t dtype, int64_t segs, int64_t rows_per_seg, int64_t c,
                 float* out, int64_t ld_out, int32_t accumulate, void* workspace, ealdm_stream_t stream);

/* z[n, 2*oh, 2*ow, :] = dy[n, oh, ow, :], zero elsewhere: the data gradient of a stride-2 convolution is a
 * stride-1 convolution of z with the flipped weights (openaimodel.py:134-160 backward). */
int ealdm_zero_insert2x(const void* dy, int64_t ld_dy, int32_t dtype, int64_t n, int64_t h, int64_t w,
                        int64_t c, void* z, int64_t ld_z, ealdm_stream_t stream);
/* dx = 2x2 block sums of dup (+ add): adjoint of nearest-2x upsampling (openaimodel.py:116 backward). */
int ealdm_sumpool2x2(const void* dup, int64_t ld_dup, int32_t dtype, int64_t n, int64_t h, int64_t w,
                     int64_t c, const float* add, int64_t ld_add, float* dx, int64_t ld_dx, void* dx2,
                     int64_t ld_dx2, ealdm_stream_t stream);

/* d(loss)/d(e_uncond), d(loss)/d(e_cond) of the p_losses tail (ealdm_cfg_mse) for
 * loss = sum_b w[b] * loss_simple[b]:  g = 2*w[b]/per_sample * (guided - target);
 * de_cond = s*g, de_uncond = (1-s)*g (de_uncond NULL when there is no guidance). */
int ealdm_cfg_mse_bwd(const float* e_uncond, const float* e_cond, const float* target, const float* w,
                      float cfg_scale, int64_t batch, int64_t per_sample, float* de_uncond, float* de_cond,
                      ealdm_stream_t stream);

/*
 * Nearest-codebook quantisation of an NCHW fp32 latent z [n, e_dim, hw] against codebook [n_e, e_dim]:
 * indices[(n, p)] = argmin_j (|z|^2 + |e_j|^2 - 2 z.e_j) (first minimum), zq = codebook[indices] (NCHW).
 * Replaces VQModelInterface.decode's `self.quantize(h)` (ldm/models/autoencoder.py:274-279), i.e. taming-transformers
 * VectorQuantizer2.forward (un-vendored dependency, environment.yaml:25; its inference arithmetic is restated in
 * oracle/vq.py).
 */
int ealdm_vq_nearest(const float* z, int64_t n, int64_t e_dim, int64_t hw, const float* codebook, int64_t n_e,
                     float* zq, int64_t* indices, ealdm_stream_t stream);

/*
 * DDPM ancestral sampling step LatentDiffusion.p_sample (ldm/models/diffusion/ddpm.py:1081-1140, eps parameterisation):
 * predict_start_from_noise (:218-222), optional clamp, q_posterior mean (:224-231), x_prev = mean + (t != 0) *
 * exp(0.5 * posterior_log_variance_clipped[t]) * noise * temperature, with the schedule buffers gathered per sample.
 * x0_out (optional) receives the x0 prediction (return_x0).  fp32, the reference's operation order.
 */
int ealdm_ddpm_step(const float* x, const float* eps, const float* noise, const int64_t* t,
                    const float* sqrt_recip_alphas_cumprod, const float* sqrt_recipm1_alphas_cumprod,
                    const float* posterior_mean_coef1, const float* posterior_mean_coef2,
                    const float* posterior_log_variance_clipped, int32_t clip_denoised, float temperature,
                    int64_t batch, int64_t per_sample, float* x_prev, float* x0_out, ealdm_stream_t stream);

/*
 * PLMSSampler.p_sample_plms eps arithmetic (ldm/models/diffusion/plms.py:178-231), fp32, the reference's operation order:
 * e_t = e_uncond + cfg_scale*(e_cond - e_uncond) (e_cond when e_uncond is NULL) -> e_t_out (the history entry);
 * e_prime_out by `mode`: 0 = e_t, 1 = (e_t + old1)/2 (pseudo improved Euler, old1 = eps at the next timestep),
 * 2 = (3 e_t - old1)/2, 3 = (23 e_t - 16 old1 + 5 old2)/12, 4 = (55 e_t - 59 old1 + 37 old2 - 9 old3)/24.
 * The x_prev / pred_x0 update that follows is ealdm_ddim_step with e_cond = e_prime and sigma = 0.
 */
int ealdm_plms_eps(const float* e_uncond, const float* e_cond, float cfg_scale, const float* old1, const float* old2,
                   const float* old3, int32_t mode, float* e_t_out, float* e_prime_out, int64_t numel,
                   ealdm_stream_t stream);

/*
 * One optimizer step over flat fp32 buffers of `numel` elements: AdamW exactly as torch.optim.AdamW (decoupled weight
 * decay, bias correction with the 1-based `step`; reference: configure_optimizers, ddpm.py:1409-1431), then the EMA
 * shadow update shadow -= (1 - ema_decay) * (shadow - param) (LitEma.forward, ldm/modules/ema.py:25-44; the caller
 * computes ema_decay = min(decay, (1 + num_updates) / (10 + num_updates))), then the bf16 copy of the new weights.
 * `grad` is multiplied by grad_scale first (1 / world size when the buffer holds the all-reduced SUM).
 * ema and param_bf16 may be NULL.
 */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  float* ema;
  void* param_bf16;
  int64_t numel;
  int64_t step;
  float lr, beta1, beta2, eps, weight_decay, grad_scale, ema_decay;
  int32_t reserved;
} ealdm_adamw_args;

int ealdm_adamw_ema_step(const ealdm_adamw_args* a, ealdm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EALDM_B200_H */
