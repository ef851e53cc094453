// vn_mapper.cu — the NeTI mapper (gradient sink of the hot path) and the optimiser step on its flat parameter buffer.
//
// Reference: models/neti_mapper.py:165-197 (forward), :416-438 (get_output: split word / bypass, normalize * norm_scale),
// :542-572 (do_positional_encoding), :601-608 (arch_view_net 15 network); models/positional_encoding.py:174-195
// (Fourier features); training/coach.py:216-218,750-757 (AdamW on the mapper parameters only).
//
//   enc = [sin(Wf x) | cos(Wf x)]                      Wf [32, nfeat], x [B, nfeat] = (t, l[, view params]) scaled to [-1, 1]
//   a1 = LeakyReLU(LN(W1 enc + b1)),  a2 = LeakyReLU(LN(W2 a1 + b2)),  y = W3 a2 + b3        [B, 2*dim]
//   word = normalize(y[:, :dim]) * norm_scale,  bypass = y[:, dim:]
// Tiny, latency-bound work (141 696 parameters): one CTA per sample forward; the backward is three small launches whose
// every cross-sample sum runs in a fixed order (deterministic gradients).
#include "vn_common.cuh"

namespace {

constexpr int HID = 64;
constexpr float kLnEps = 1e-5f;
constexpr float kSlope = 0.01f;
// flat parameter layout (state_dict order): W1 b1 g1 be1 W2 b2 g2 be2 W3 b3
constexpr int OFF_W1 = 0, OFF_B1 = OFF_W1 + HID * HID, OFF_G1 = OFF_B1 + HID, OFF_BE1 = OFF_G1 + HID;
constexpr int OFF_W2 = OFF_BE1 + HID, OFF_B2 = OFF_W2 + HID * HID, OFF_G2 = OFF_B2 + HID, OFF_BE2 = OFF_G2 + HID;
constexpr int OFF_W3 = OFF_BE2 + HID;
// per-sample saved activations: enc[64] xh1[64] a1[64] xh2[64] a2[64] rstd1 rstd2 nrm pad what[dim]
constexpr int SV_ENC = 0, SV_XH1 = 64, SV_A1 = 128, SV_XH2 = 192, SV_A2 = 256, SV_SC = 320, SV_WH = 324;

__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += s_red[i];
  __syncthreads();
  return t;
}

// hidden layer: out = LeakyReLU(LN(W in + b) * g + be); 64 threads compute, everybody syncs
__device__ __forceinline__ void hidden_layer(const float* __restrict__ W, const float* __restrict__ b,
                                             const float* __restrict__ g, const float* __restrict__ be,
                                             const float* s_in, float* s_h, float* s_out, float* xh_out, float* rstd_out) {
  const int t = threadIdx.x;
  if (t < HID) {
    float acc = b[t];
#pragma unroll 8
    for (int k = 0; k < HID; ++k) acc = fmaf(W[t * HID + k], s_in[k], acc);
    s_h[t] = acc;
  }
  __syncthreads();
  if (t < HID) {
    float m = 0.f;
    for (int k = 0; k < HID; ++k) m += s_h[k];
    m *= (1.f / HID);
    float v = 0.f;
    for (int k = 0; k < HID; ++k) { const float d = s_h[k] - m; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(v * (1.f / HID) + kLnEps);
    const float xh = (s_h[t] - m) * rstd;
    const float z = xh * g[t] + be[t];
    s_out[t] = z > 0.f ? z : kSlope * z;
    xh_out[t] = xh;
    if (t == 0) *rstd_out = rstd;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) mapper_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Wf,
                                                         const float* __restrict__ prm, float norm_scale,
                                                         float* __restrict__ word, float* __restrict__ bypass,
                                                         float* __restrict__ saved, int nfeat, int dim) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_enc[HID], s_h[HID], s_a1[HID], s_a2[HID], s_red[8];
  const int b = blockIdx.x, t = threadIdx.x;
  float* sv = saved + (long long)b * (SV_WH + dim);
  if (t < HID / 2) {
    float arg = 0.f;
    for (int f = 0; f < nfeat; ++f) arg = fmaf(Wf[t * nfeat + f], x[(long long)b * nfeat + f], arg);
    s_enc[t] = sinf(arg);
    s_enc[HID / 2 + t] = cosf(arg);
  }
  __syncthreads();
  if (t < HID) sv[SV_ENC + t] = s_enc[t];
  hidden_layer(prm + OFF_W1, prm + OFF_B1, prm + OFF_G1, prm + OFF_BE1, s_enc, s_h, s_a1, sv + SV_XH1, sv + SV_SC);
  hidden_layer(prm + OFF_W2, prm + OFF_B2, prm + OFF_G2, prm + OFF_BE2, s_a1, s_h, s_a2, sv + SV_XH2, sv + SV_SC + 1);
  if (t < HID) { sv[SV_A1 + t] = s_a1[t]; sv[SV_A2 + t] = s_a2[t]; }
  const float* W3 = prm + OFF_W3;
  const float* b3 = W3 + (long long)2 * dim * HID;
  float ss = 0.f;
  for (int j = t; j < 2 * dim; j += 256) {
    float acc = b3[j];
    const float4* wr = reinterpret_cast<const float4*>(W3 + (long long)j * HID);
#pragma unroll
    for (int k = 0; k < HID / 4; ++k) {
      const float4 w = wr[k];
      acc = fmaf(w.x, s_a2[4 * k], fmaf(w.y, s_a2[4 * k + 1], fmaf(w.z, s_a2[4 * k + 2], fmaf(w.w, s_a2[4 * k + 3], acc))));
    }
    if (j < dim) { word[(long long)b * dim + j] = acc; ss = fmaf(acc, acc, ss); }    // normalised below
    else bypass[(long long)b * dim + (j - dim)] = acc;
  }
  const float tot = block_sum_256(ss, s_red);
  const float nrm = fmaxf(sqrtf(tot), 1e-12f);                 // F.normalize eps
  if (t == 0) sv[SV_SC + 2] = nrm;
  __syncthreads();
  for (int j = t; j < dim; j += 256) {
    const float y = word[(long long)b * dim + j];
    const float wh = y / nrm;
    sv[SV_WH + j] = wh;
    word[(long long)b * dim + j] = norm_scale > 0.f ? wh * norm_scale : y;
  }
}

// per sample: dy [2*dim] (through the normalisation) and da2 [64] = W3^T dy
__global__ void __launch_bounds__(256) mapper_bwd_dy_kernel(const float* __restrict__ d_word,
                                                            const float* __restrict__ d_bypass,
                                                            const float* __restrict__ prm,
                                                            const float* __restrict__ saved, float norm_scale,
                                                            float* __restrict__ scratch, int dim) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_red[8], s_part[4][HID];
  const int b = blockIdx.x, t = threadIdx.x;
  const float* sv = saved + (long long)b * (SV_WH + dim);
  float* dy = scratch + (long long)b * (2 * dim + HID);
  float dot = 0.f;
  if (norm_scale > 0.f)
    for (int j = t; j < dim; j += 256) dot = fmaf(sv[SV_WH + j], d_word[(long long)b * dim + j], dot);
  dot = block_sum_256(dot, s_red);
  const float nrm = sv[SV_SC + 2];
  for (int j = t; j < 2 * dim; j += 256) {
    float g;
    if (j < dim) {
      const float dw = d_word[(long long)b * dim + j];
      g = norm_scale > 0.f ? norm_scale / nrm * (dw - sv[SV_WH + j] * dot) : dw;
    } else {
      g = d_bypass[(long long)b * dim + (j - dim)];
    }
    dy[j] = g;
  }
  __syncthreads();
  const float* W3 = prm + OFF_W3;
  const int k = t & (HID - 1), part = t >> 6;
  float acc = 0.f;
  for (int j = part; j < 2 * dim; j += 4) acc = fmaf(W3[(long long)j * HID + k], dy[j], acc);
  s_part[part][k] = acc;
  __syncthreads();
  if (t < HID) dy[2 * dim + t] = (s_part[0][t] + s_part[1][t]) + (s_part[2][t] + s_part[3][t]);
}

// dW3[j,k] = sum_b dy[b,j] a2[b,k], db3[j] = sum_b dy[b,j]; 4 rows per CTA, samples in order
__global__ void __launch_bounds__(256) mapper_bwd_w3_kernel(const float* __restrict__ saved,
                                                            const float* __restrict__ scratch,
                                                            float* __restrict__ d_prm, int B, int dim) {
  pdl_trigger();
  pdl_wait();
  const int j = blockIdx.x * 4 + (threadIdx.x >> 6), k = threadIdx.x & (HID - 1);
  if (j >= 2 * dim) return;
  float acc = 0.f, accb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float g = scratch[(long long)b * (2 * dim + HID) + j];
    acc = fmaf(g, saved[(long long)b * (SV_WH + dim) + SV_A2 + k], acc);
    accb += g;
  }
  d_prm[OFF_W3 + (long long)j * HID + k] = acc;
  if (k == 0) d_prm[OFF_W3 + (long long)2 * dim * HID + j] = accb;
}

// the two hidden layers, one CTA, samples in order; thread (r, c4) owns 4 elements of each 64 x 64 weight gradient
__global__ void __launch_bounds__(1024) mapper_bwd_hidden_kernel(const float* __restrict__ prm,
                                                                 const float* __restrict__ saved,
                                                                 const float* __restrict__ scratch,
                                                                 float* __restrict__ d_prm, int B, int dim) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_da[HID], s_dh[HID], s_in[HID], s_vec[6][HID];   // accumulators: db2 dg2 dbe2 db1 dg1 dbe1
  const int t = threadIdx.x, r = t >> 4, c0 = (t & 15) * 4;
  float dW2[4] = {0.f, 0.f, 0.f, 0.f}, dW1[4] = {0.f, 0.f, 0.f, 0.f};
  if (t < 6 * HID) s_vec[t / HID][t % HID] = 0.f;
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    const float* sv = saved + (long long)b * (SV_WH + dim);
    if (t < HID) s_da[t] = scratch[(long long)b * (2 * dim + HID) + 2 * dim + t];
    __syncthreads();
    for (int layer = 1; layer >= 0; --layer) {
      const float* g = prm + (layer ? OFF_G2 : OFF_G1);
      const float* be = prm + (layer ? OFF_BE2 : OFF_BE1);
      const float* W = prm + (layer ? OFF_W2 : OFF_W1);
      const float* xh = sv + (layer ? SV_XH2 : SV_XH1);
      const float* in = sv + (layer ? SV_A1 : SV_ENC);
      const float rstd = sv[SV_SC + layer];
      float* acc_b = s_vec[layer ? 0 : 3];
      float* acc_g = s_vec[layer ? 1 : 4];
      float* acc_be = s_vec[layer ? 2 : 5];
      if (t < HID) {
        const float z = xh[t] * g[t] + be[t];
        const float dz = s_da[t] * (z > 0.f ? 1.f : kSlope);
        acc_g[t] += dz * xh[t];
        acc_be[t] += dz;
        s_dh[t] = dz * g[t];                 // d xhat
        s_in[t] = in[t];
      }
      __syncthreads();
      if (t < HID) {
        float m1 = 0.f, m2 = 0.f;
        for (int k = 0; k < HID; ++k) { m1 += s_dh[k]; m2 = fmaf(s_dh[k], xh[k], m2); }
        m1 *= (1.f / HID); m2 *= (1.f / HID);
        const float dh = rstd * (s_dh[t] - m1 - xh[t] * m2);
        __syncwarp();
        s_da[t] = dh;                        // reuse: d(pre-LN)
        acc_b[t] += dh;
      }
      __syncthreads();
      // weight gradient: dW[r, c] += dh[r] * in[c]
      {
        const float dh = s_da[r];
        float* dW = layer ? dW2 : dW1;
#pragma unroll
        for (int c = 0; c < 4; ++c) dW[c] = fmaf(dh, s_in[c0 + c], dW[c]);
      }
      // d(input)[k] = sum_r W[r, k] dh[r]   (only needed below layer 2)
      if (layer == 1) {
        float v = 0.f;
        if (t < HID) {
          for (int rr = 0; rr < HID; ++rr) v = fmaf(W[rr * HID + t], s_da[rr], v);
        }
        __syncthreads();
        if (t < HID) s_da[t] = v;
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    d_prm[OFF_W2 + r * HID + c0 + c] = dW2[c];
    d_prm[OFF_W1 + r * HID + c0 + c] = dW1[c];
  }
  if (t < HID) {
    d_prm[OFF_B2 + t] = s_vec[0][t]; d_prm[OFF_G2 + t] = s_vec[1][t]; d_prm[OFF_BE2 + t] = s_vec[2][t];
    d_prm[OFF_B1 + t] = s_vec[3][t]; d_prm[OFF_G1 + t] = s_vec[4][t]; d_prm[OFF_BE1 + t] = s_vec[5][t];
  }
}

// torch.optim.AdamW (decoupled weight decay), one fused pass over the flat buffer
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                                    float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                    float grad_scale) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi;
  }
}

}  // namespace

extern "C" int vn_mapper_param_count(int dim) { return OFF_W3 + 2 * dim * HID + 2 * dim; }
extern "C" int vn_mapper_saved_floats(int dim) { return SV_WH + dim; }

extern "C" int vn_mapper_fwd(const float* x, const float* Wf, const float* params, float norm_scale, float* word,
                             float* bypass, float* saved, int B, int nfeat, int dim, vn_stream_t s) {
  VN_CHECK(B > 0 && nfeat > 0 && dim > 0 && dim % 4 == 0, "mapper: bad sizes B=%d nfeat=%d dim=%d", B, nfeat, dim);
  VN_LAUNCH(mapper_fwd_kernel, B, 256, 0, (cudaStream_t)s, x, Wf, params, norm_scale, word, bypass, saved, nfeat, dim);
  return 0;
}

extern "C" int vn_mapper_bwd(const float* d_word, const float* d_bypass, const float* params, const float* saved,
                             float norm_scale, float* d_params, float* scratch, int B, int dim, vn_stream_t s) {
  VN_CHECK(B > 0 && dim > 0 && dim % 4 == 0, "mapper bwd: bad sizes B=%d dim=%d", B, dim);
  cudaStream_t st = (cudaStream_t)s;
  VN_LAUNCH(mapper_bwd_dy_kernel, B, 256, 0, st, d_word, d_bypass, params, saved, norm_scale, scratch, dim);
  VN_LAUNCH(mapper_bwd_w3_kernel, vn_cdiv(2 * dim, 4), 256, 0, st, saved, (const float*)scratch, d_params, B, dim);
  VN_LAUNCH(mapper_bwd_hidden_kernel, 1, 1024, 0, st, params, saved, (const float*)scratch, d_params, B, dim);
  return 0;
}

extern "C" int vn_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                             vn_stream_t s) {
  VN_CHECK(n > 0 && step >= 1, "adamw: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = sqrtf(1.f - powf(beta2, (float)step));
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  VN_LAUNCH(adamw_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)s, params, grads, exp_avg, exp_avg_sq, (long long)n, lr,
            beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale);
  return 0;
}
