/* viewneti.h — C-ABI of libviewneti_sm100a.so (hand-written sm_100a CUDA for the ViewNeTI hot path).
 *
 * The reference (jmhb0/view_neti) is pure Python and has NO FFI: every entry point below is new and
 * replaces a library call the reference makes through torch/diffusers on the path
 *     training/coach.py:197-214          unet(noisy_latents, timesteps, _hs).sample ; mse ; backward
 *     models/xti_attention_processor.py  XTIAttenProc.__call__ (32x per UNet forward)
 *     sd_pipeline_call.py:71-101         two-pass CFG denoise loop
 * Each declaration cites the reference/diffusers operation it stands in for.  INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller
 *     (PyTorch caching allocator); the library allocates nothing on the device.
 *   - activations are NHWC / token-major [rows, C] bf16 with an explicit row stride `ld*` in ELEMENTS
 *     (so channel-slices of a concat buffer are first-class); norm/bias parameters are fp32.
 *   - all launches are asynchronous on the passed stream, no hidden syncs, CUDA-graph capturable.
 *   - return 0 on success, negative on error (message: vn_last_error()); never throws, never exits.
 *   - a process drives one device/stream at a time through this library (not thread-safe).
 */
#ifndef VIEWNETI_H_
#define VIEWNETI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vn_stream_t;          /* cudaStream_t */

#define VN_ABI_VERSION 1

int         vn_version(void);
const char* vn_last_error(void);
/* number of kernels launched by this library since load / since last reset (bench `gpu_launches`). */
int64_t     vn_launch_count(void);
void        vn_launch_count_reset(void);
/* Programmatic dependent launch (on by default): every kernel is launched so that its prologue overlaps the tail of
 * its predecessor on the stream.  Switch off to time kernels in isolation (per-kernel profiler durations otherwise
 * include the time a kernel spends waiting for its predecessor). */
void        vn_set_pdl(int enabled);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 / TMA GEMM and implicit-GEMM 3x3 convolution.
 *   D[M,N] = A[M,K] * B[N,K]^T  (+ bias[N]) (+ rowbias[row / rows_per_batch, N]) (+ R[M,N])
 * mode 0: A is [M,K] row-major bf16 (lda).          replaces torch Linear / 1x1 conv (cuBLAS) on
 *         xti_attention_processor.py:30,38-42,53 (to_q/to_k/to_v/to_out), diffusers proj_in/proj_out,
 *         FeedForward linears, conv_shortcut; with B = W^T it is the dgrad of the same op.
 * mode 1: A is NHWC [nb,H,W,C] bf16 (pixel stride lda), 3x3 stride 1 pad 1, K = 9*C with
 *         k = tap*C + c; M = nb*H*W.                 replaces cuDNN Conv2d in diffusers ResnetBlock2D /
 *         Upsample2D; with flipped+transposed weights it is the conv dgrad.
 * B is always [N,K] bf16 K-major (ldb).  D is bf16 (or fp32 when out_fp32) with row stride ldd.
 * B is treated as FROZEN WEIGHTS: its first tiles are fetched ahead of the programmatic dependency on the preceding
 * kernel of the stream, so B must not be the output of a preceding vn_* launch (A, R and D may be).
 * Constraints: K % 64 == 0 (mode 1: C % 64 == 0), N % 8 == 0, lda/ldb/ldd/ldr % 8 == 0, 16-byte aligned bases.
 * Split-K: `workspace` must hold >= vn_gemm_workspace_bytes() bytes, be all-zero before the first call
 * and is left all-zero by every call (self-cleaning), so one buffer serves a whole stream.
 * ------------------------------------------------------------------------------------------------ */
typedef struct vn_gemm_desc {
  int32_t mode;
  int32_t M, N, K;
  int32_t nb, H, W, C;              /* conv modes only: INPUT dims (mode 1: 3x3 s1 p1; 2: s2 p1, Downsample2D; 3: s2, zero beyond
                                       the far edge only - the VAE encoder's pad (0,1,0,1); M = nb*Ho*Wo) */
  const void* A;  int64_t lda;
  const void* B;  int64_t ldb;
  void*       D;  int64_t ldd;
  const float* bias;                /* [N] or NULL */
  const float* rowbias; int64_t ld_rowbias; int32_t rows_per_batch;   /* [nbatch,N] fp32 or NULL */
  const void* R;  int64_t ldr;      /* residual [M,N] bf16 (fp32 when r_fp32) or NULL (may alias D) */
  int32_t out_fp32;
  void*   workspace; size_t workspace_bytes;
  int32_t force_bn, force_split;    /* tuning/test overrides; 0 = auto */
  int32_t r_fp32;                   /* R holds fp32 (only with out_fp32: the fp32 residual stream of the text encoder) */
} vn_gemm_desc;

size_t vn_gemm_workspace_bytes(int max_M, int max_N);
/* diagnostics (library built with -DVN_TIMELINE only, see scripts/kernel_timeline.py): when non-NULL, every vn_gemm /
 * fused-GroupNorm CTA writes clock64 stamps of its phases into 16 int64 slots of this device buffer
 * (>= 16 * 2048 slots).  NULL (default) switches it off; the product build compiles the stamps out. */
void   vn_set_debug_buffer(void* dev_ptr);
int    vn_gemm(const vn_gemm_desc* d, vn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm (+ optional SiLU) over NHWC bf16 [nb, hw, C] — diffusers ResnetBlock2D.norm1/norm2,
 * Transformer2DModel.norm, conv_norm_out (torch native_group_norm + silu).
 * stats/red: fp64 [nb, groups, 2]; must be zero on entry of *_stats (they accumulate with atomics; fp64 keeps the
 * result independent of the accumulation order, i.e. reproducible run to run).
 *   fwd stats = (sum x, sum x^2);  bwd red = (sum dxhat, sum dxhat*xhat)
 * ------------------------------------------------------------------------------------------------ */
int vn_groupnorm_stats(const void* x, int64_t ldx, int nb, int hw, int C, int groups, double* stats, vn_stream_t s);
int vn_groupnorm_apply(const void* x, int64_t ldx, const double* stats, const float* gamma, const float* beta,
                       float eps, int silu, void* y, int64_t ldy, int nb, int hw, int C, int groups, vn_stream_t s);
int vn_groupnorm_bwd_stats(const void* x, int64_t ldx, const void* dy, int64_t lddy, const double* stats,
                           const float* gamma, const float* beta, float eps, int silu, double* red,
                           int nb, int hw, int C, int groups, vn_stream_t s);
/* dx = GN^T(dy) (+ add1) (+ add2) */
int vn_groupnorm_bwd_apply(const void* x, int64_t ldx, const void* dy, int64_t lddy, const double* stats,
                           const double* red, const float* gamma, const float* beta, float eps, int silu,
                           const void* add1, int64_t ldadd1, const void* add2, int64_t ldadd2,
                           void* dx, int64_t lddx, int nb, int hw, int C, int groups, vn_stream_t s);

/* One-launch forms (statistics + apply fused; what the engine calls).  `partials`: vn_groupnorm_partial_floats(nb)
 * floats private to this call, EVERY BYTE 0xff on entry (the engine presets one arena for all GroupNorms of a pass
 * with one memset): the per-CTA partial sums through which the CTAs of the grid exchange their statistics - a word
 * that is no longer 0xffffffff has been written, so the data is its own flag (no atomics, no counters, deterministic).
 * stats / red are WRITTEN (not accumulated).  The grid never exceeds one CTA per SM, so all CTAs are co-resident.
 * partials == NULL, or a shape the fused kernel does not cover (C > 4096, groups > 32, nb > #SMs, < 64 threads), runs
 * the two-kernel form above (which needs stats / red zero on entry, so callers keep them zeroed); results agree to
 * fp32 rounding of the statistics.
 * vn_set_groupnorm_fused(0) (env VN_GN_FUSED=0) forces the two-kernel form. */
int vn_groupnorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, int silu,
                     void* y, int64_t ldy, int nb, int hw, int C, int groups, double* stats,
                     float* partials, vn_stream_t s);
int vn_groupnorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, const double* stats,
                     const float* gamma, const float* beta, float eps, int silu,
                     const void* add1, int64_t ldadd1, const void* add2, int64_t ldadd2,
                     void* dx, int64_t lddx, int nb, int hw, int C, int groups, double* red,
                     float* partials, vn_stream_t s);
void vn_set_groupnorm_fused(int enabled);
size_t vn_groupnorm_partial_floats(int nb);

/* LayerNorm over rows of [rows, C] bf16 — diffusers BasicTransformerBlock.norm1/2/3.
 * stats: fp32 [rows,2] = (mean, rstd), written by fwd, read by bwd.  bwd: dx = LN^T(dy) (+ add). */
int vn_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                     void* y, int64_t ldy, float* stats, int rows, int C, vn_stream_t s);
int vn_layernorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* gamma,
                     const float* stats, const void* add, int64_t ldadd, void* dx, int64_t lddx,
                     int rows, int C, vn_stream_t s);

/* fp32-stream forms (the CLIP text encoder keeps its residual stream in fp32, models/clip_encoder.py - transformers'
 * CLIPEncoderLayer under reference models/neti_clip_text_encoder.py:101-108 runs in fp32): x / add / dx are fp32 rows, the
 * normalised row y (a GEMM operand) stays bf16; bwd also writes a bf16 copy of dx (the A operand of the next dgrad GEMM)
 * when dx_bf16 != NULL. */
int vn_layernorm_fwd_f32(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                         void* y, int64_t ldy, float* stats, int rows, int C, vn_stream_t s);
int vn_layernorm_bwd_f32(const float* x, int64_t ldx, const void* dy, int64_t lddy, const float* gamma,
                         const float* stats, const float* add, int64_t ldadd, float* dx, int64_t lddx,
                         void* dx_bf16, int64_t lddxb, int rows, int C, vn_stream_t s);

/* GEGLU — diffusers GEGLU: h = [a | g] ([rows, 2F]);  y = a * gelu_erf(g). */
int vn_geglu_fwd(const void* h, int64_t ldh, void* y, int64_t ldy, int rows, int F, vn_stream_t s);
int vn_geglu_bwd(const void* h, int64_t ldh, const void* dy, int64_t lddy, void* dh, int64_t lddh,
                 int rows, int F, vn_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * Attention core, head_dim 64 — xti_attention_processor.py:44-50:
 *   head_to_batch_dim, get_attention_scores (fp32 logits: baddbmm alpha=scale, softmax), bmm, batch_to_head_dim.
 * q/k/v/o are token-major with heads side by side: element (b, n, h, d) at  base + b*bs + n*ld + h*64 + d
 * so head split/merge copies disappear.  K and V come from DIFFERENT tensors (XTI: K from
 * CONTEXT_TENSOR_i, V from CONTEXT_TENSOR_BYPASS_i).  lse: fp32 [nb, heads, nq] (natural log), delta same shape.
 * fwd: one flash kernel (logits never leave the SM).
 * bwd: dq, dk, dv.  dk/dv may be accumulated in fp64 scratch `dkv_acc` ([2, nb, nk, heads*64] fp64, zero on
 * entry, left zero) when nk is too small to parallelise over keys (cross-attention, nk = 77); fp64 keeps the sum
 * independent of the order in which the query splits arrive.
 * ------------------------------------------------------------------------------------------------ */
typedef struct vn_attn_desc {
  int32_t nb, heads, nq, nk;
  float   scale;
  const void* q; int64_t ldq, bsq;
  const void* k; int64_t ldk, bsk;
  const void* v; int64_t ldv, bsv;
  void*       o; int64_t ldo, bso;        /* fwd: out; bwd: in */
  float*      lse;                        /* fwd: out; bwd: in */
  /* backward only */
  const void* d_o; int64_t lddo, bsdo;
  float*      delta;                      /* scratch [nb, heads, nq] */
  void* dq; int64_t lddq, bsdq;           /* may be NULL (pruned) */
  void* dk; int64_t lddk, bsdk;
  void* dv; int64_t lddv, bsdv;
  double* dkv_acc;                        /* NULL or zeroed scratch, see above */
  int32_t causal;                         /* 1: key j is visible to query i only if j <= i (CLIP text encoder) */
  void*   ws; int64_t ws_bytes;           /* fwd: scratch of vn_attention_fwd_workspace_bytes() bytes, or NULL (see below) */
  int32_t defer_dkv_finish;               /* bwd with dkv_acc: 1 = leave dK / dV in the fp64 accumulator; the caller runs
                                             vn_attention_dkv_finish later, e.g. on another stream (see below) */
} vn_attn_desc;
/* Forward work balancing: when the (query tile, head, image) items do not fill whole waves of SMs (64x64 latents at B = 1:
 * 160 items on 148 SMs), the leftover items are split along the keys over all SMs and merged by a second small launch;
 * the partial results live in `ws` (caller-owned, contents irrelevant on entry).  ws == NULL, or a shape for which
 * vn_attention_fwd_workspace_bytes returns 0, runs one CTA per item.  Results agree to fp32 rounding. */
size_t vn_attention_fwd_workspace_bytes(int nb, int heads, int nq, int nk);
/* Same for the backward (dK/dV + dQ items; has_dq = 0 when dq is pruned): persistent CTAs run whole items back to back and
 * the leftover items are split along their loops; `ws` holds plain fp32 partial tiles that a second small launch adds up in a
 * fixed order. */
size_t vn_attention_bwd_workspace_bytes(int nb, int heads, int nq, int nk, int has_dq);
int vn_attention_fwd(const vn_attn_desc* d, vn_stream_t s);
int vn_attention_bwd(const vn_attn_desc* d, vn_stream_t s);
/* Second half of a backward launched with defer_dkv_finish = 1 (same descriptor): fp64 accumulator -> bf16 dK / dV, accumulator
 * zeroed again.  Nothing downstream of the cross-attention's dQ needs dK / dV (they only feed the context gradients,
 * reference xti_attention_processor.py:38-42 under coach.py:214), so the UNet plan runs this on its side stream instead of on
 * the chain of the backward; a later backward that uses the SAME accumulator must be ordered after it.  A no-op when the
 * backward wrote dK / dV directly (no query split). */
int vn_attention_dkv_finish(const vn_attn_desc* d, vn_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * Resampling / small convolutions / glue
 * ------------------------------------------------------------------------------------------------ */
/* diffusers Upsample2D: nearest x2.  y [nb,2H,2W,C];  bwd: dx = 2x2 sum of dy (+ nothing) */
int vn_upsample2x_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int nb, int H, int W, int C, vn_stream_t s);
int vn_upsample2x_bwd(const void* dy, int64_t lddy, void* dx, int64_t lddx, int nb, int H, int W, int C, vn_stream_t s);
/* diffusers Downsample2D (Conv 3x3 stride 2 pad 1) = im2col + vn_gemm.  col [nb*Ho*Wo, 9*C], k = tap*C + c */
int vn_im2col_s2(const void* x, int64_t ldx, void* col, int nb, int H, int W, int C, vn_stream_t s);
/* the VAE encoder's Downsample2D(padding=0): F.pad(x, (0,1,0,1)) + Conv 3x3 stride 2 pad 0 (replaces diffusers
 * AutoencoderKL.encode called at reference training/coach.py:167).  Ho = (H-2)/2 + 1; same col layout */
int vn_im2col_s2_pad0(const void* x, int64_t ldx, void* col, int nb, int H, int W, int C, vn_stream_t s);
/* dgrad: dx[nb,H,W,C] = col2im(dcol) (+ add) */
int vn_col2im_s2(const void* dcol, const void* add, int64_t ldadd, void* dx, int64_t lddx,
                 int nb, int H, int W, int C, vn_stream_t s);
/* few-channel 3x3 / stride 1 / pad 1 conv as a GEMM (the VAE's conv_in layers, reference training/coach.py:167 and
 * sd_pipeline_call.py:115 -> diffusers Encoder / Decoder.conv_in): x NCHW fp32 [nb,Ct,H,W] -> col bf16 [nb*H*W, ldc],
 * k = tap*Ct + ct, zero for 9*Ct <= k < ldc; ldc % 64 == 0 */
int vn_im2col_thin(const float* x, void* col, int64_t ldc, int nb, int Ct, int H, int W, vn_stream_t s);
/* conv_in: NCHW fp32 latents [nb,Cin,H,W] -> NHWC bf16 [nb,H,W,Cout]; w fp32 [Cout,Cin,3,3] */
int vn_conv_in_fwd(const float* x, const float* w, const float* bias, void* y, int64_t ldy,
                   int nb, int Cin, int H, int W, int Cout, vn_stream_t s);
/* conv_out: NHWC bf16 [nb,H,W,Cin] -> NCHW fp32 [nb,Cout,H,W]; w fp32 [Cout,Cin,3,3] */
int vn_conv_out_fwd(const void* x, int64_t ldx, const float* w, const float* bias, float* y,
                    int nb, int Cin, int H, int W, int Cout, vn_stream_t s);
/* dgrad of conv_out: dy NCHW fp32 -> dx NHWC bf16 */
int vn_conv_out_bwd(const float* dy, const float* w, void* dx, int64_t lddx,
                    int nb, int Cin, int H, int W, int Cout, vn_stream_t s);
/* diffusers Timesteps(flip_sin_to_cos=True, freq_shift=0): out fp32 [nb, dim] = [cos | sin] */
int vn_timestep_sinusoid(const int64_t* t, float* out, int nb, int dim, vn_stream_t s);
/* small-M linear for the time-embedding MLP and the 22 per-ResBlock time_emb_proj layers (one launch):
 * y[b,n] = bias[n] + sum_k act(x[b,k]) * W[n,k];  W bf16 [N,K]; x,y fp32; silu_in applies SiLU to x. */
int vn_gemv(const float* x, int64_t ldx, const void* W, const float* bias, float* y, int64_t ldy,
            int nb, int N, int K, int silu_in, vn_stream_t s);
/* fp32 -> bf16 cast of context tensors etc. (n elements) and bf16 -> fp32 */
int vn_cast_f32_bf16(const float* x, void* y, int64_t n, vn_stream_t s);
int vn_cast_bf16_f32(const void* x, float* y, int64_t n, vn_stream_t s);
/* strided 2-D bf16 copy, dst[r, 0:cols] = src[r, 0:cols] (+ add[r, 0:cols]): torch.cat of the UNet skip
 * connections (diffusers up blocks) and the gradient fan-in of the residual stream in the backward pass. */
int vn_copy2d(const void* src, int64_t lds, const void* add, int64_t ldadd, void* dst, int64_t ldd,
              int64_t rows, int cols, vn_stream_t s);
/* coach.py:211-213  loss = mean((pred-target)^2) in fp32; also dpred = 2*(pred-target)/n * loss_scale */
int vn_mse_loss(const float* pred, const float* target, int64_t n, float loss_scale, float* loss, float* dpred, vn_stream_t s);
/* sd_pipeline_call.py:98,101 fused: eps = u + g*(c-u); DDIM (eta=0) step for epsilon / v prediction.
 * latents fp32 [n] updated in place.  acp_t / acp_prev = alphas_cumprod at t / t_prev; vpred: 0 eps, 1 v. */
int vn_cfg_ddim_step(float* latents, const float* eps_uncond, const float* eps_cond, int64_t n,
                     float guidance, float acp_t, float acp_prev, int vpred, vn_stream_t s);
/* [nb, hw, ldin] fp32 (the first Ct channels of every pixel) -> [nb, Ct, hw] fp32: NCHW view of the result of a thin conv_out
 * computed as an N-padded implicit GEMM (diffusers UNet2DConditionModel.conv_out, 320 -> 4, under reference coach.py:197). */
int vn_nhwc_to_nchw_thin(const float* in, int64_t ldin, float* out, int nb, int Ct, int64_t hw, vn_stream_t s);
/* the same fused step for DPM-Solver++(2M), the scheduler the reference's inference scripts install (reference
 * training/validate.py:568, training/inference_dtu.py:304): m = u + g (c - u); x0 = p x + q m;
 * latents = A x + B0 x0 + B1 x0_prev;  x0_prev = x0.  (p, q, A, B0, B1) per step from the host scheduler; B1 = 0 on
 * first-order steps, when x0_prev is not read */
int vn_cfg_dpmpp_step(float* latents, const float* eps_uncond, const float* eps_cond, float* x0_prev, int64_t n,
                      float guidance, float p, float q, float A, float B0, float B1, vn_stream_t s);
/* ------------------------------------------------------------------------------------------------
 * CLIP text transformer pieces (the batched conditioning path: reference
 * models/neti_clip_text_encoder.py:57-225 runs transformers' CLIPEncoder once per UNet layer; here the 16 passes are one batch).
 * Projections / LayerNorms use vn_gemm / vn_layernorm_*.
 *   vn_gelu_*            CLIPMLP activation (erf GELU): y = gelu(h);  dh = dy * gelu'(h).
 *   vn_seq_attention_*   CLIPAttention core for short sequences (nq == nk <= 128, head_dim 64), optional causal mask
 *                        (CLIPTextTransformer._build_causal_attention_mask) on CUDA cores; same descriptor as
 *                        vn_attention_*; lse is the natural-log sum-exp of the scaled logits; bwd needs q,k,v,o,lse,d_o and
 *                        writes dq,dk,dv.  The encoder uses vn_attention_* with desc.causal = 1 (tensor cores); these
 *                        kernels are the independent cross-check of that path.
 * ------------------------------------------------------------------------------------------------ */
int vn_gelu_fwd(const void* h, int64_t ldh, void* y, int64_t ldy, int rows, int F, vn_stream_t s);
int vn_gelu_bwd(const void* h, int64_t ldh, const void* dy, int64_t lddy, void* dh, int64_t lddh, int rows, int F,
                vn_stream_t s);
/* VAE mid-block attention (one head over all channels, reference training/coach.py:167 / sd_pipeline_call.py:115 ->
 * diffusers AttentionBlock): P[r,:] = softmax(scale * S[r,:]), S fp32 [rows, cols] (vn_gemm with out_fp32), P bf16.
 * cols % 4 == 0, cols <= 8192, scale > 0 */
int vn_softmax_rows(const float* S, int64_t lds, void* P, int64_t ldp, int rows, int cols, float scale, vn_stream_t s);
int vn_seq_attention_fwd(const struct vn_attn_desc* d, int causal, vn_stream_t s);
int vn_seq_attention_bwd(const struct vn_attn_desc* d, int causal, vn_stream_t s);

int vn_memset_zero(void* p, size_t bytes, vn_stream_t s);
int vn_memset(void* p, int byte_value, size_t bytes, vn_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * NeTI mapper (SURVEY.md 8f "next" #2) - the module the context gradients finally land in.
 * reference models/neti_mapper.py:165-197,416-438,542-611 (arch_view_net 15) + models/positional_encoding.py:174-195:
 *   enc = [sin(Wf x) | cos(Wf x)];  a1 = LeakyReLU(LN(W1 enc + b1));  a2 = LeakyReLU(LN(W2 a1 + b2));  y = W3 a2 + b3
 *   word = normalize(y[:, :dim]) * norm_scale (norm_scale <= 0: no normalisation),  bypass = y[:, dim:]
 * x [B, nfeat] fp32 = (t, l[, view parameters]) already scaled to [-1, 1];  Wf [32, nfeat] fp32 (not trained).
 * params / d_params: flat fp32, state_dict order  net.0.{weight[64,64],bias} net.1.{weight,bias} net.3.{weight,bias}
 * net.4.{weight,bias} output_layer.0.{weight[2*dim,64],bias[2*dim]}  (vn_mapper_param_count(dim) floats).
 * saved: [B, vn_mapper_saved_floats(dim)] fp32 written by fwd, read by bwd; scratch: [B, 2*dim + 64] fp32.
 * The backward sums over samples in a fixed order (deterministic) and OVERWRITES d_params.
 * ------------------------------------------------------------------------------------------------ */
int vn_mapper_param_count(int dim);
int vn_mapper_saved_floats(int dim);
int vn_mapper_fwd(const float* x, const float* Wf, const float* params, float norm_scale, float* word, float* bypass,
                  float* saved, int B, int nfeat, int dim, vn_stream_t s);
int vn_mapper_bwd(const float* d_word, const float* d_bypass, const float* params, const float* saved, float norm_scale,
                  float* d_params, float* scratch, int B, int dim, vn_stream_t s);
/* torch.optim.AdamW step (coach.py:216, :750-757) on a flat fp32 buffer; grads are multiplied by grad_scale first
 * (1/world after the all-reduce).  step counts from 1. */
int vn_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, float grad_scale, vn_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* VIEWNETI_H_ */
