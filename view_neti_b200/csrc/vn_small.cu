// vn_small.cu — resampling, the two 4-channel edge convolutions, time embedding, loss and sampler-step kernels.
// All of these are HBM/L2- or latency-bound (<0.1% of the FLOPs of the path); they exist so the whole
// coach.py:197-214 / sd_pipeline_call.py:71-101 step runs on the device without torch eager ops in between.
#include "vn_common.cuh"

namespace {

int grid_for(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__device__ __forceinline__ void acc8(float (&a)[8], const uint4& v) {
  float2 t;
  t = unpack_bf162(v.x); a[0] += t.x; a[1] += t.y;
  t = unpack_bf162(v.y); a[2] += t.x; a[3] += t.y;
  t = unpack_bf162(v.z); a[4] += t.x; a[5] += t.y;
  t = unpack_bf162(v.w); a[6] += t.x; a[7] += t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&a)[8]) {
  uint4 o;
  o.x = pack_bf162(a[0], a[1]); o.y = pack_bf162(a[2], a[3]);
  o.z = pack_bf162(a[4], a[5]); o.w = pack_bf162(a[6], a[7]);
  return o;
}

// ---------------------------------------------------------------------------------------------
// nearest x2 upsample (diffusers Upsample2D / F.interpolate(scale_factor=2, mode="nearest")) and its adjoint
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const bf16* __restrict__ x, long long ldx,
                                                             bf16* __restrict__ y, long long ldy, int nb, int H, int W,
                                                             int vecs) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)nb * (2 * H) * (2 * W) * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % vecs) * 8;
    long long p = i / vecs;
    const int wo = (int)(p % (2 * W)); p /= (2 * W);
    const int ho = (int)(p % (2 * H));
    const int b = (int)(p / (2 * H));
    const long long src = ((long long)b * H + (ho >> 1)) * W + (wo >> 1);
    const long long dst = ((long long)b * 2 * H + ho) * (2 * W) + wo;
    *reinterpret_cast<uint4*>(y + dst * ldy + c) = *reinterpret_cast<const uint4*>(x + src * ldx + c);
  }
}

__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const bf16* __restrict__ dy, long long lddy,
                                                             bf16* __restrict__ dx, long long lddx, int nb, int H,
                                                             int W, int vecs) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)nb * H * W * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % vecs) * 8;
    long long p = i / vecs;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int b = (int)(p / H);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dyy = 0; dyy < 2; ++dyy)
#pragma unroll
      for (int dxx = 0; dxx < 2; ++dxx) {
        const long long src = ((long long)b * 2 * H + 2 * h + dyy) * (2 * W) + 2 * w + dxx;
        acc8(a, *reinterpret_cast<const uint4*>(dy + src * lddy + c));
      }
    const long long dst = ((long long)b * H + h) * W + w;
    *reinterpret_cast<uint4*>(dx + dst * lddx + c) = pack8(a);
  }
}

// ---------------------------------------------------------------------------------------------
// stride-2 3x3 convolution support (diffusers Downsample2D): explicit im2col and its adjoint.
// col row = output pixel, k = tap*C + c.  `pad` zero rows/columns in front: 1 = the UNet's symmetric pad-1 form
// (Ho = (H-1)/2 + 1); 0 = the VAE encoder's F.pad(x, (0,1,0,1)) + pad-0 conv (Ho = (H-2)/2 + 1).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_s2_kernel(const bf16* __restrict__ x, long long ldx,
                                                        bf16* __restrict__ col, int nb, int H, int W, int Ho, int Wo,
                                                        int C, int pad) {
  pdl_trigger();
  pdl_wait();
  const int vecs = C >> 3;
  const long long total = (long long)nb * Ho * Wo * 9 * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % vecs) * 8;
    long long p = i / vecs;
    const int tap = (int)(p % 9); p /= 9;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const int h = 2 * ho + tap / 3 - pad, w = 2 * wo + tap % 3 - pad;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (h >= 0 && h < H && w >= 0 && w < W) v = *reinterpret_cast<const uint4*>(x + (((long long)b * H + h) * W + w) * ldx + c);
    const long long row = ((long long)b * Ho + ho) * Wo + wo;
    *reinterpret_cast<uint4*>(col + row * (9LL * C) + (long long)tap * C + c) = v;
  }
}

__global__ void __launch_bounds__(256) col2im_s2_kernel(const bf16* __restrict__ dcol, const bf16* __restrict__ add,
                                                        long long ldadd, bf16* __restrict__ dx, long long lddx, int nb,
                                                        int H, int W, int Ho, int Wo, int C) {
  pdl_trigger();
  pdl_wait();
  const int vecs = C >> 3;
  const long long total = (long long)nb * H * W * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % vecs) * 8;
    long long p = i / vecs;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int b = (int)(p / H);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hn = h + 1 - ky;
      if (hn < 0 || (hn & 1)) continue;
      const int ho = hn >> 1;
      if (ho >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int wn = w + 1 - kx;
        if (wn < 0 || (wn & 1)) continue;
        const int wo = wn >> 1;
        if (wo >= Wo) continue;
        const long long row = ((long long)b * Ho + ho) * Wo + wo;
        acc8(a, *reinterpret_cast<const uint4*>(dcol + row * (9LL * C) + (long long)(ky * 3 + kx) * C + c));
      }
    }
    const long long dst = ((long long)b * H + h) * W + w;
    if (add) acc8(a, *reinterpret_cast<const uint4*>(add + dst * ldadd + c));
    *reinterpret_cast<uint4*>(dx + dst * lddx + c) = pack8(a);
  }
}

// ---------------------------------------------------------------------------------------------
// few-channel 3x3 convolutions at the UNet edges.
//   thin -> wide: in NCHW fp32 [nb,Ct,H,W] (Ct small, <= 8), out NHWC bf16 [nb,H,W,Cw].
//     FLIP=0  conv_in forward      wgt[(cw*Ct + ct)*9 + tap]           (weight [Cw,Ct,3,3])
//     FLIP=1  conv_out dgrad       wgt[(ct*Cw + cw)*9 + (8 - tap)]     (weight [Ct,Cw,3,3])
// ---------------------------------------------------------------------------------------------
constexpr int kThinMax = 8;

template <int FLIP>
__global__ void __launch_bounds__(256) conv_thin_to_wide_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ wgt,
                                                                const float* __restrict__ bias, bf16* __restrict__ y,
                                                                long long ldy, int nb, int Ct, int H, int W, int Cw) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_w[];   // [9 taps][Ct][Cw]: a thread reads its 8 output channels as two 16-byte vectors
  for (int i = threadIdx.x; i < 9 * Ct * Cw; i += blockDim.x) {
    const int cw = i % Cw;
    const int ct = (i / Cw) % Ct;
    const int tap = i / (Cw * Ct);
    s_w[i] = FLIP ? wgt[((long long)ct * Cw + cw) * 9 + (8 - tap)] : wgt[((long long)cw * Ct + ct) * 9 + tap];
  }
  __syncthreads();
  const int vecs = Cw >> 3;
  const long long total = (long long)nb * H * W * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % vecs) * 8;
    long long p = i / vecs;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int b = (int)(p / H);
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = bias ? bias[c0 + j] : 0.f;
    for (int ct = 0; ct < Ct; ++ct) {
      const float* xp = x + ((long long)b * Ct + ct) * H * W;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
        const float v = __ldg(xp + hh * W + ww);
        const float4 w0 = *reinterpret_cast<const float4*>(s_w + ((long long)tap * Ct + ct) * Cw + c0);
        const float4 w1 = *reinterpret_cast<const float4*>(s_w + ((long long)tap * Ct + ct) * Cw + c0 + 4);
        a[0] = fmaf(v, w0.x, a[0]); a[1] = fmaf(v, w0.y, a[1]); a[2] = fmaf(v, w0.z, a[2]); a[3] = fmaf(v, w0.w, a[3]);
        a[4] = fmaf(v, w1.x, a[4]); a[5] = fmaf(v, w1.y, a[5]); a[6] = fmaf(v, w1.z, a[6]); a[7] = fmaf(v, w1.w, a[7]);
      }
    }
    const long long dst = ((long long)b * H + h) * W + w;
    *reinterpret_cast<uint4*>(y + dst * ldy + c0) = pack8(a);
  }
}

// wide -> thin (conv_out forward): x NHWC bf16 [nb,H,W,Cw], weight fp32 [Ct,Cw,3,3], y NCHW fp32 [nb,Ct,H,W].
// One warp per output pixel; lanes stride over channel pairs; weights staged in smem as [tap][ct][cw].
__global__ void __launch_bounds__(256) conv_wide_to_thin_kernel(const bf16* __restrict__ x, long long ldx,
                                                                const float* __restrict__ wgt,
                                                                const float* __restrict__ bias, float* __restrict__ y,
                                                                int nb, int Cw, int H, int W, int Ct) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_w[];   // [9][Ct][Cw]
  for (int i = threadIdx.x; i < 9 * Ct * Cw; i += blockDim.x) {
    const int cw = i % Cw;
    const int ct = (i / Cw) % Ct;
    const int tap = i / (Cw * Ct);
    s_w[i] = wgt[((long long)ct * Cw + cw) * 9 + tap];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long npix = (long long)nb * H * W;
  for (long long pix = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); pix < npix;
       pix += (long long)gridDim.x * warps_per_block) {
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    float acc[kThinMax];
#pragma unroll
    for (int t = 0; t < kThinMax; ++t) acc[t] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const bf16* xp = x + (((long long)b * H + hh) * W + ww) * ldx;
      for (int c = lane * 2; c < Cw; c += 64) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const bf162*>(xp + c));
#pragma unroll
        for (int t = 0; t < kThinMax; ++t) {
          if (t < Ct) {
            const float* wp = s_w + ((long long)tap * Ct + t) * Cw + c;
            acc[t] = fmaf(v.x, wp[0], fmaf(v.y, wp[1], acc[t]));
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < kThinMax; ++t) {
      if (t < Ct) {
        const float s = warp_sum(acc[t]);
        if (lane == 0) y[(((long long)b * Ct + t) * H + h) * W + w] = s + (bias ? bias[t] : 0.f);
      }
    }
  }
}

// im2col of a FEW-channel NCHW fp32 image for a 3x3 / stride 1 / pad 1 convolution on vn_gemm: col [nb*H*W, ldc] bf16,
// k = tap*Ct + ct for k < 9*Ct, zero up to ldc (a multiple of 64: one or two k-blocks).  The VAE's conv_in layers
// (3 -> 128 at the image resolution, 4 -> 512 at the latent resolution) become memory-bound GEMMs this way.
__global__ void __launch_bounds__(256) im2col_thin_kernel(const float* __restrict__ x, bf16* __restrict__ col,
                                                          long long ldc, int nb, int Ct, int H, int W) {
  pdl_trigger();
  pdl_wait();
  const int vecs = (int)(ldc >> 3);
  const int kmax = 9 * Ct;
  const long long total = (long long)nb * H * W * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long pix = i / vecs;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = v * 8 + j;
      a[j] = 0.f;
      if (k < kmax) {
        const int tap = k / Ct, ct = k - tap * Ct;
        const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) a[j] = __ldg(x + (((long long)b * Ct + ct) * H + hh) * W + ww);
      }
    }
    *reinterpret_cast<uint4*>(col + pix * ldc + v * 8) = pack8(a);
  }
}

// ---------------------------------------------------------------------------------------------
// time embedding pieces
// ---------------------------------------------------------------------------------------------
__global__ void timestep_sinusoid_kernel(const long long* __restrict__ t, float* __restrict__ out, int nb, int dim) {
  pdl_trigger();
  pdl_wait();
  const int half = dim >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb * half) return;
  const int b = i / half, j = i - b * half;
  const float freq = expf(-logf(10000.f) * (float)j / (float)half);
  const float arg = (float)t[b] * freq;
  out[(long long)b * dim + j] = cosf(arg);          // flip_sin_to_cos=True: [cos | sin]
  out[(long long)b * dim + half + j] = sinf(arg);
}

// y[b,n] = bias[n] + sum_k act(x[b,k]) * W[n,k]; one warp per output n, all (<= 8) batch rows at once.
// The activations (SiLU applied ONCE per CTA, not once per output) are staged in shared memory; every lane has all of its
// 16-byte weight loads of a row in flight before the first one is consumed, so the launch streams W at HBM rate (the
// per-ResBlock time-embedding projection reads 38 MB of weights: 39 us -> ~8 us on B200).
constexpr int kGemvMaxB = 8;
constexpr int kGemvMaxK = 2048;
__global__ void __launch_bounds__(256) gemv_kernel(const float* __restrict__ x, long long ldx,
                                                   const bf16* __restrict__ Wt, const float* __restrict__ bias,
                                                   float* __restrict__ y, long long ldy, int nb, int N, int K,
                                                   int silu_in) {
  pdl_trigger();
  extern __shared__ float sx[];                  // [nb][K]
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  // the weights are frozen: this lane's slice of row n is requested ahead of the grid dependency
  constexpr int kMaxIter = kGemvMaxK / 256;
  uint4 wv[kMaxIter];
  const int iters = (K + 255) >> 8;
  const bf16* wr = Wt + (long long)min(n, N - 1) * K;
#pragma unroll
  for (int i = 0; i < kMaxIter; ++i) {
    const int k = lane * 8 + i * 256;
    wv[i] = (i < iters && k < K) ? __ldg(reinterpret_cast<const uint4*>(wr + k)) : make_uint4(0u, 0u, 0u, 0u);
  }
  pdl_wait();
  for (int i = threadIdx.x; i < nb * K; i += blockDim.x) {
    const int b = i / K, k = i - b * K;
    float v = x[(long long)b * ldx + k];
    if (silu_in) v = silu_f(v);
    sx[i] = v;
  }
  __syncthreads();
  if (n >= N) return;
  float acc[kGemvMaxB];
#pragma unroll
  for (int b = 0; b < kGemvMaxB; ++b) acc[b] = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxIter; ++i) {
    const int k = lane * 8 + i * 256;
    if (i < iters && k < K) {
      float wf[8];
      float2 t;
      t = unpack_bf162(wv[i].x); wf[0] = t.x; wf[1] = t.y;
      t = unpack_bf162(wv[i].y); wf[2] = t.x; wf[3] = t.y;
      t = unpack_bf162(wv[i].z); wf[4] = t.x; wf[5] = t.y;
      t = unpack_bf162(wv[i].w); wf[6] = t.x; wf[7] = t.y;
#pragma unroll
      for (int b = 0; b < kGemvMaxB; ++b) {
        if (b < nb) {
          const float4 x0 = *reinterpret_cast<const float4*>(sx + b * K + k);
          const float4 x1 = *reinterpret_cast<const float4*>(sx + b * K + k + 4);
          acc[b] = fmaf(x0.x, wf[0], acc[b]); acc[b] = fmaf(x0.y, wf[1], acc[b]);
          acc[b] = fmaf(x0.z, wf[2], acc[b]); acc[b] = fmaf(x0.w, wf[3], acc[b]);
          acc[b] = fmaf(x1.x, wf[4], acc[b]); acc[b] = fmaf(x1.y, wf[5], acc[b]);
          acc[b] = fmaf(x1.z, wf[6], acc[b]); acc[b] = fmaf(x1.w, wf[7], acc[b]);
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < kGemvMaxB; ++b) {
    if (b < nb) {
      const float s = warp_sum(acc[b]);
      if (lane == 0) y[(long long)b * ldy + n] = s + (bias ? bias[n] : 0.f);
    }
  }
}

// [nb, H*W, ldin] fp32 (first Ct channels of every pixel) -> [nb, Ct, H*W] fp32: the NCHW view of a thin conv_out result
__global__ void __launch_bounds__(256) nhwc_to_nchw_thin_kernel(const float* __restrict__ in, int ldin, float* __restrict__ out,
                                                                int Ct, long long hw, long long total) {
  pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % hw;
    const long long nc = i / hw;
    const int c = (int)(nc % Ct);
    const long long n = nc / Ct;
    out[i] = in[(n * hw + p) * ldin + c];
  }
}

// ---------------------------------------------------------------------------------------------
// coach.py:211-213 fp32 MSE (+ gradient w.r.t. the prediction) — single CTA, deterministic.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) mse_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                   long long n, float loss_scale, float* __restrict__ loss,
                                                   float* __restrict__ dpred) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_part[32];
  float acc = 0.f;
  const float g = 2.f * loss_scale / (float)n;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = pred[i] - target[i];
    acc = fmaf(d, d, acc);
    if (dpred) dpred[i] = g * d;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0 && loss) *loss = v / (float)n;
  }
}

// sd_pipeline_call.py:98 + :101 (DDIM eta = 0): eps = u + g (c - u); x0 / eps from the model output; x_prev.
__global__ void __launch_bounds__(256) cfg_ddim_kernel(float* __restrict__ latents, const float* __restrict__ eu,
                                                       const float* __restrict__ ec, long long n, float guidance,
                                                       float acp_t, float acp_prev, int vpred) {
  pdl_trigger();
  pdl_wait();
  const float sa = sqrtf(acp_t), sb = sqrtf(1.f - acp_t);
  const float pa = sqrtf(acp_prev), pb = sqrtf(1.f - acp_prev);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float u = eu[i];
    const float m = u + guidance * (ec[i] - u);
    const float x = latents[i];
    float x0, eps;
    if (vpred) {
      x0 = sa * x - sb * m;
      eps = sa * m + sb * x;
    } else {
      x0 = (x - sb * m) / sa;
      eps = m;
    }
    latents[i] = pa * x0 + pb * eps;
  }
}

// sd_pipeline_call.py:98 + :101 with the DPM-Solver++(2M) scheduler the reference's inference scripts install: every
// update is linear - x0 = p x + q m (m = guided model output), x_prev = A x + B0 x0 + B1 x0_prev - so one kernel does
// guidance, conversion and update and keeps x0 for the next step.  Coefficients come from the host scheduler.
__global__ void __launch_bounds__(256) cfg_dpmpp_kernel(float* __restrict__ latents, const float* __restrict__ eu,
                                                        const float* __restrict__ ec, float* __restrict__ x0_prev,
                                                        long long n, float guidance, float p, float q, float A,
                                                        float B0, float B1) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float u = eu[i];
    const float m = u + guidance * (ec[i] - u);
    const float x = latents[i];
    const float x0 = p * x + q * m;
    float o = A * x + B0 * x0;
    if (B1 != 0.f) o += B1 * x0_prev[i];          // first step: the slot holds nothing yet
    latents[i] = o;
    x0_prev[i] = x0;
  }
}

template <typename K>
int thin_smem_config(K kernel, size_t smem) {
  VN_CHECK(smem <= 200 * 1024, "edge conv: weights (%zu B) do not fit in shared memory", smem);
  if (smem > 48 * 1024) VN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

}  // namespace

extern "C" int vn_upsample2x_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int nb, int H, int W, int C,
                                 vn_stream_t s) {
  VN_CHECK(C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "upsample2x: C and strides must be multiples of 8");
  const long long total = (long long)nb * 4 * H * W * (C / 8);
  VN_LAUNCH(upsample2x_fwd_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)s, (const bf16*)x, ldx, (bf16*)y, ldy, nb, H, W,
                                                                            C / 8);
  return 0;
}

extern "C" int vn_upsample2x_bwd(const void* dy, int64_t lddy, void* dx, int64_t lddx, int nb, int H, int W, int C,
                                 vn_stream_t s) {
  VN_CHECK(C % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0, "upsample2x: C and strides must be multiples of 8");
  const long long total = (long long)nb * H * W * (C / 8);
  VN_LAUNCH(upsample2x_bwd_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)s, (const bf16*)dy, lddy, (bf16*)dx, lddx, nb, H,
                                                                            W, C / 8);
  return 0;
}

extern "C" int vn_im2col_s2(const void* x, int64_t ldx, void* col, int nb, int H, int W, int C, vn_stream_t s) {
  VN_CHECK(C % 8 == 0 && ldx % 8 == 0, "im2col_s2: C and stride must be multiples of 8");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)nb * Ho * Wo * 9 * (C / 8);
  VN_LAUNCH(im2col_s2_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)s, (const bf16*)x, ldx, (bf16*)col, nb, H, W, Ho, Wo,
                                                                       C, 1);
  return 0;
}

extern "C" int vn_im2col_s2_pad0(const void* x, int64_t ldx, void* col, int nb, int H, int W, int C, vn_stream_t s) {
  VN_CHECK(C % 8 == 0 && ldx % 8 == 0, "im2col_s2_pad0: C and stride must be multiples of 8");
  VN_CHECK(H >= 2 && W >= 2, "im2col_s2_pad0: image smaller than 2x2");
  const int Ho = (H - 2) / 2 + 1, Wo = (W - 2) / 2 + 1;
  const long long total = (long long)nb * Ho * Wo * 9 * (C / 8);
  VN_LAUNCH(im2col_s2_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)s, (const bf16*)x, ldx, (bf16*)col, nb, H, W, Ho, Wo,
                                                                       C, 0);
  return 0;
}

extern "C" int vn_col2im_s2(const void* dcol, const void* add, int64_t ldadd, void* dx, int64_t lddx, int nb, int H,
                            int W, int C, vn_stream_t s) {
  VN_CHECK(C % 8 == 0 && lddx % 8 == 0 && (!add || ldadd % 8 == 0), "col2im_s2: C and strides must be multiples of 8");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)nb * H * W * (C / 8);
  VN_LAUNCH(col2im_s2_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)s, (const bf16*)dcol, (const bf16*)add, ldadd,
                                                                       (bf16*)dx, lddx, nb, H, W, Ho, Wo, C);
  return 0;
}

extern "C" int vn_im2col_thin(const float* x, void* col, int64_t ldc, int nb, int Ct, int H, int W, vn_stream_t s) {
  VN_CHECK(Ct >= 1 && ldc % 64 == 0 && ldc >= 9 * Ct, "im2col_thin: need ldc %% 64 == 0 and ldc >= 9*Ct (Ct=%d ldc=%lld)", Ct,
           (long long)ldc);
  const long long total = (long long)nb * H * W * (ldc / 8);
  VN_LAUNCH(im2col_thin_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)s, x, (bf16*)col, (long long)ldc, nb, Ct, H, W);
  return 0;
}

extern "C" int vn_conv_in_fwd(const float* x, const float* w, const float* bias, void* y, int64_t ldy, int nb, int Cin,
                              int H, int W, int Cout, vn_stream_t s) {
  VN_CHECK(Cout % 8 == 0 && ldy % 8 == 0 && Cin >= 1, "conv_in: Cout and ldy must be multiples of 8");
  const long long total = (long long)nb * H * W * (Cout / 8);
  const size_t smem = (size_t)9 * Cin * Cout * sizeof(float);
  if (thin_smem_config(conv_thin_to_wide_kernel<0>, smem)) return -2;
  int blocks = grid_for(total, 256);
  if (blocks > 148 * 2) blocks = 148 * 2;
  VN_LAUNCH(conv_thin_to_wide_kernel<0>, blocks, 256, smem, (cudaStream_t)s, x, w, bias, (bf16*)y, ldy, nb, Cin, H, W, Cout);
  return 0;
}

extern "C" int vn_conv_out_bwd(const float* dy, const float* w, void* dx, int64_t lddx, int nb, int Cin, int H, int W,
                               int Cout, vn_stream_t s) {
  VN_CHECK(Cin % 8 == 0 && lddx % 8 == 0 && Cout >= 1, "conv_out_bwd: Cin and lddx must be multiples of 8");
  const long long total = (long long)nb * H * W * (Cin / 8);
  const size_t smem = (size_t)9 * Cin * Cout * sizeof(float);
  if (thin_smem_config(conv_thin_to_wide_kernel<1>, smem)) return -2;
  int blocks = grid_for(total, 256);
  if (blocks > 148 * 2) blocks = 148 * 2;
  VN_LAUNCH(conv_thin_to_wide_kernel<1>, blocks, 256, smem, (cudaStream_t)s, dy, w, nullptr, (bf16*)dx, lddx, nb, Cout, H, W, Cin);
  return 0;
}

extern "C" int vn_conv_out_fwd(const void* x, int64_t ldx, const float* w, const float* bias, float* y, int nb, int Cin,
                               int H, int W, int Cout, vn_stream_t s) {
  VN_CHECK(Cin % 2 == 0 && ldx % 2 == 0 && Cout >= 1 && Cout <= kThinMax, "conv_out: need even Cin and Cout <= %d",
           kThinMax);
  const size_t smem = (size_t)9 * Cout * Cin * sizeof(float);
  VN_CHECK(smem <= 200 * 1024, "conv_out: weights (%zu B) do not fit in shared memory", smem);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    VN_CUDA(cudaFuncSetAttribute(conv_wide_to_thin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long npix = (long long)nb * H * W;
  int blocks = (int)((npix + 7) / 8);
  if (blocks > 148 * 2) blocks = 148 * 2;
  VN_LAUNCH(conv_wide_to_thin_kernel, blocks, 256, smem, (cudaStream_t)s, (const bf16*)x, ldx, w, bias, y, nb, Cin, H, W, Cout);
  return 0;
}

extern "C" int vn_timestep_sinusoid(const int64_t* t, float* out, int nb, int dim, vn_stream_t s) {
  VN_CHECK(dim % 2 == 0, "timestep_sinusoid: dim must be even");
  const int total = nb * (dim / 2);
  VN_LAUNCH(timestep_sinusoid_kernel, vn_cdiv(total, 128), 128, 0, (cudaStream_t)s, (const long long*)t, out, nb, dim);
  return 0;
}

extern "C" int vn_gemv(const float* x, int64_t ldx, const void* W, const float* bias, float* y, int64_t ldy, int nb,
                       int N, int K, int silu_in, vn_stream_t s) {
  VN_CHECK(nb >= 1 && nb <= kGemvMaxB, "gemv: batch %d not in [1,%d]", nb, kGemvMaxB);
  VN_CHECK(K % 8 == 0 && K <= kGemvMaxK, "gemv: K must be a multiple of 8 and <= %d", kGemvMaxK);
  const int smem = nb * K * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    VN_CUDA(cudaFuncSetAttribute(gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemvMaxB * kGemvMaxK * (int)sizeof(float)));
    configured = true;
  }
  VN_LAUNCH(gemv_kernel, vn_cdiv(N, 8), 256, smem, (cudaStream_t)s, x, ldx, (const bf16*)W, bias, y, ldy, nb, N, K, silu_in);
  return 0;
}

extern "C" int vn_nhwc_to_nchw_thin(const float* in, int64_t ldin, float* out, int nb, int Ct, int64_t hw, vn_stream_t s) {
  VN_CHECK(nb > 0 && Ct > 0 && hw > 0 && ldin >= Ct, "nhwc_to_nchw_thin: bad arguments");
  const long long total = (long long)nb * Ct * hw;
  VN_LAUNCH(nhwc_to_nchw_thin_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)s, in, (int)ldin, out, Ct, (long long)hw, total);
  return 0;
}

extern "C" int vn_mse_loss(const float* pred, const float* target, int64_t n, float loss_scale, float* loss,
                           float* dpred, vn_stream_t s) {
  VN_CHECK(n > 0, "mse: empty input");
  VN_LAUNCH(mse_kernel, 1, 1024, 0, (cudaStream_t)s, pred, target, n, loss_scale, loss, dpred);
  return 0;
}

extern "C" int vn_cfg_dpmpp_step(float* latents, const float* eps_uncond, const float* eps_cond, float* x0_prev,
                                 int64_t n, float guidance, float p, float q, float A, float B0, float B1,
                                 vn_stream_t s) {
  VN_CHECK(n > 0 && latents && eps_uncond && eps_cond && x0_prev, "cfg_dpmpp_step: bad arguments");
  VN_LAUNCH(cfg_dpmpp_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)s, latents, eps_uncond, eps_cond, x0_prev, n, guidance,
                                                                   p, q, A, B0, B1);
  return 0;
}

extern "C" int vn_cfg_ddim_step(float* latents, const float* eps_uncond, const float* eps_cond, int64_t n,
                                float guidance, float acp_t, float acp_prev, int vpred, vn_stream_t s) {
  VN_CHECK(n > 0 && acp_t > 0.f, "cfg_ddim_step: bad arguments");
  VN_LAUNCH(cfg_ddim_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)s, latents, eps_uncond, eps_cond, n, guidance, acp_t,
                                                                  acp_prev, vpred);
  return 0;
}
