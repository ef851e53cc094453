// vn_norm.cu — GroupNorm(+SiLU), LayerNorm and GEGLU, forward and dgrad, over NHWC / token-major bf16.
// HBM/L2-bound row-wise kernels: 128-bit or 32-bit coalesced accesses along the channel axis, fp32 statistics,
// warp-shuffle + shared-memory reductions.  Replaces torch native_group_norm / native_layer_norm / gelu kernels
// under diffusers ResnetBlock2D, Transformer2DModel, BasicTransformerBlock, GEGLU (all on the coach.py:197 path).
#include "vn_common.cuh"

namespace {

constexpr int kMaxC = 2560;       // widest GroupNorm input in SD-2.1 (concat 1280+1280)
constexpr int kGNThreads = 256;

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics: grid (chunks, nb); each CTA walks `ppc` pixels over all channels.
// Threads own fixed bf16x2 channel pairs (pair p -> group (2p)/cpg; cpg is even), so loads are coalesced.
// MODE 0: (sum x, sum x^2).   MODE 1 (backward): (sum dxhat, sum dxhat*xhat).
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kGNThreads) gn_reduce_kernel(const bf16* __restrict__ x, long long ldx,
                                                              const bf16* __restrict__ dy, long long lddy,
                                                              const float* __restrict__ stats,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, int silu,
                                                              float* __restrict__ out, int hw, int C, int groups,
                                                              int ppc) {
  __shared__ float s_acc[64][2];
  __shared__ float s_mean[64], s_rstd[64];
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const int pairs = C >> 1;
  const int p0 = blockIdx.x * ppc;
  const int p1 = min(hw, p0 + ppc);
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    s_acc[g][0] = 0.f; s_acc[g][1] = 0.f;
    if (MODE == 1) {
      const float n = (float)cpg * (float)hw;
      const float m = stats[(b * groups + g) * 2] / n;
      const float var = fmaxf(stats[(b * groups + g) * 2 + 1] / n - m * m, 0.f);
      s_mean[g] = m; s_rstd[g] = rsqrtf(var + eps);
    }
  }
  __syncthreads();
  for (int pr = threadIdx.x; pr < pairs; pr += blockDim.x) {
    const int c = pr * 2;
    const int g = c / cpg;
    float a0 = 0.f, a1 = 0.f;
    float ga0 = 0.f, ga1 = 0.f, be0 = 0.f, be1 = 0.f, mean = 0.f, rstd = 0.f;
    if (MODE == 1) {
      ga0 = gamma[c]; ga1 = gamma[c + 1]; be0 = beta[c]; be1 = beta[c + 1];
      mean = s_mean[g]; rstd = s_rstd[g];
    }
    for (int p = p0; p < p1; ++p) {
      const long long row = (long long)b * hw + p;
      const float2 v = __bfloat1622float2(*reinterpret_cast<const bf162*>(x + row * ldx + c));
      if (MODE == 0) {
        a0 += v.x + v.y;
        a1 += v.x * v.x + v.y * v.y;
      } else {
        const float2 d = __bfloat1622float2(*reinterpret_cast<const bf162*>(dy + row * lddy + c));
        const float xh0 = (v.x - mean) * rstd, xh1 = (v.y - mean) * rstd;
        float dz0 = d.x, dz1 = d.y;
        if (silu) { dz0 *= dsilu_f(xh0 * ga0 + be0); dz1 *= dsilu_f(xh1 * ga1 + be1); }
        const float dh0 = dz0 * ga0, dh1 = dz1 * ga1;
        a0 += dh0 + dh1;
        a1 += dh0 * xh0 + dh1 * xh1;
      }
    }
    atomicAdd(&s_acc[g][0], a0);
    atomicAdd(&s_acc[g][1], a1);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    atomicAdd(&out[(b * groups + g) * 2], s_acc[g][0]);
    atomicAdd(&out[(b * groups + g) * 2 + 1], s_acc[g][1]);
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm apply: y = act(x * a[c] + b[c]) with per-(image, channel) a, b staged in shared memory.
// MODE 0 forward.  MODE 1 backward: dx = rstd*(dxhat - S1/n - xhat*S2/n) (+add1) (+add2).
// grid (chunks, nb); 8-channel (16 B) vectors.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kGNThreads) gn_apply_kernel(const bf16* __restrict__ x, long long ldx,
                                                             const bf16* __restrict__ dy, long long lddy,
                                                             const float* __restrict__ stats,
                                                             const float* __restrict__ red,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, int silu,
                                                             const bf16* __restrict__ add1, long long ld1,
                                                             const bf16* __restrict__ add2, long long ld2,
                                                             bf16* __restrict__ y, long long ldy, int hw, int C,
                                                             int groups, int ppc) {
  extern __shared__ float s_gn[];
  float* s_a = s_gn;            // fwd: gamma*rstd          bwd: gamma
  float* s_b = s_gn + C;        // fwd: beta - mean*a       bwd: beta
  float* s_m = s_gn + 2 * C;    // bwd only: mean, rstd, S1/n, S2/n per group (4*groups)
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const float n = (float)cpg * (float)hw;
  if (MODE == 1) {
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      const float m = stats[(b * groups + g) * 2] / n;
      const float var = fmaxf(stats[(b * groups + g) * 2 + 1] / n - m * m, 0.f);
      s_m[g * 4] = m;
      s_m[g * 4 + 1] = rsqrtf(var + eps);
      s_m[g * 4 + 2] = red[(b * groups + g) * 2] / n;
      s_m[g * 4 + 3] = red[(b * groups + g) * 2 + 1] / n;
    }
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (MODE == 0) {
      const int g = c / cpg;
      const float m = stats[(b * groups + g) * 2] / n;
      const float var = fmaxf(stats[(b * groups + g) * 2 + 1] / n - m * m, 0.f);
      const float a = gamma[c] * rsqrtf(var + eps);
      s_a[c] = a;
      s_b[c] = beta[c] - m * a;
    } else {
      s_a[c] = gamma[c];
      s_b[c] = beta[c];
    }
  }
  __syncthreads();
  const int vecs = C >> 3;
  const int p0 = blockIdx.x * ppc;
  const int p1 = min(hw, p0 + ppc);
  const int total = (p1 - p0) * vecs;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int p = p0 + i / vecs;
    const int c = (i % vecs) * 8;
    const long long row = (long long)b * hw + p;
    const uint4 raw = *reinterpret_cast<const uint4*>(x + row * ldx + c);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    float o[8];
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 v = unpack_bf162(w[j]);
        float z0 = v.x * s_a[c + 2 * j] + s_b[c + 2 * j];
        float z1 = v.y * s_a[c + 2 * j + 1] + s_b[c + 2 * j + 1];
        if (silu) { z0 = silu_f(z0); z1 = silu_f(z1); }
        o[2 * j] = z0; o[2 * j + 1] = z1;
      }
    } else {
      const uint4 draw = *reinterpret_cast<const uint4*>(dy + row * lddy + c);
      const uint32_t dw[4] = {draw.x, draw.y, draw.z, draw.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 v = unpack_bf162(w[j]);
        const float2 d = unpack_bf162(dw[j]);
        const float xv[2] = {v.x, v.y};
        const float dv[2] = {d.x, d.y};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int cc = c + 2 * j + e;
          const int g = cc / cpg;
          const float mean = s_m[g * 4], rstd = s_m[g * 4 + 1];
          const float xh = (xv[e] - mean) * rstd;
          float dz = dv[e];
          if (silu) dz *= dsilu_f(xh * s_a[cc] + s_b[cc]);
          const float dh = dz * s_a[cc];
          o[2 * j + e] = rstd * (dh - s_m[g * 4 + 2] - xh * s_m[g * 4 + 3]);
        }
      }
      if (add1) {
        const uint4 r = *reinterpret_cast<const uint4*>(add1 + row * ld1 + c);
        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 t = unpack_bf162(rw[j]); o[2 * j] += t.x; o[2 * j + 1] += t.y; }
      }
      if (add2) {
        const uint4 r = *reinterpret_cast<const uint4*>(add2 + row * ld2 + c);
        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 t = unpack_bf162(rw[j]); o[2 * j] += t.x; o[2 * j + 1] += t.y; }
      }
    }
    uint4 out;
    out.x = pack_bf162(o[0], o[1]); out.y = pack_bf162(o[2], o[3]);
    out.z = pack_bf162(o[4], o[5]); out.w = pack_bf162(o[6], o[7]);
    *reinterpret_cast<uint4*>(y + row * ldy + c) = out;
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (C <= 2048), two-pass variance in fp32.
// ---------------------------------------------------------------------------------------------
constexpr int kLNMaxPairs = 32;   // per lane: C/2/32 <= 32  -> C <= 2048

template <int MODE>
__global__ void __launch_bounds__(256) ln_kernel(const bf16* __restrict__ x, long long ldx,
                                                 const bf16* __restrict__ dy, long long lddy,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 float eps, float* __restrict__ stats, const bf16* __restrict__ add,
                                                 long long ldadd, bf16* __restrict__ y, long long ldy, int rows,
                                                 int C) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int pairs = C >> 1;
  const bf16* xr = x + (long long)row * ldx;
  float2 v[kLNMaxPairs];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLNMaxPairs; ++i) {
    const int p = lane + i * 32;
    if (p < pairs) {
      v[i] = __bfloat1622float2(*reinterpret_cast<const bf162*>(xr + 2 * p));
      s += v[i].x + v[i].y;
    }
  }
  float mean, rstd;
  if (MODE == 0) {
    mean = warp_sum(s) / (float)C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kLNMaxPairs; ++i) {
      const int p = lane + i * 32;
      if (p < pairs) { const float a = v[i].x - mean, c = v[i].y - mean; ss += a * a + c * c; }
    }
    rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
    if (lane == 0 && stats) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
    bf16* yr = y + (long long)row * ldy;
#pragma unroll
    for (int i = 0; i < kLNMaxPairs; ++i) {
      const int p = lane + i * 32;
      if (p < pairs) {
        const float2 g = *reinterpret_cast<const float2*>(gamma + 2 * p);
        const float2 bb = *reinterpret_cast<const float2*>(beta + 2 * p);
        *reinterpret_cast<bf162*>(yr + 2 * p) =
            __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
      }
    }
  } else {
    mean = stats[row * 2]; rstd = stats[row * 2 + 1];
    const bf16* dr = dy + (long long)row * lddy;
    float2 dh[kLNMaxPairs];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < kLNMaxPairs; ++i) {
      const int p = lane + i * 32;
      if (p < pairs) {
        const float2 d = __bfloat1622float2(*reinterpret_cast<const bf162*>(dr + 2 * p));
        const float2 g = *reinterpret_cast<const float2*>(gamma + 2 * p);
        dh[i] = make_float2(d.x * g.x, d.y * g.y);
        v[i].x = (v[i].x - mean) * rstd; v[i].y = (v[i].y - mean) * rstd;
        c1 += dh[i].x + dh[i].y;
        c2 += dh[i].x * v[i].x + dh[i].y * v[i].y;
      }
    }
    c1 = warp_sum(c1) / (float)C;
    c2 = warp_sum(c2) / (float)C;
    bf16* yr = y + (long long)row * ldy;
    const bf16* ar = add ? add + (long long)row * ldadd : nullptr;
#pragma unroll
    for (int i = 0; i < kLNMaxPairs; ++i) {
      const int p = lane + i * 32;
      if (p < pairs) {
        float o0 = rstd * (dh[i].x - c1 - v[i].x * c2);
        float o1 = rstd * (dh[i].y - c1 - v[i].y * c2);
        if (ar) { const float2 a = __bfloat1622float2(*reinterpret_cast<const bf162*>(ar + 2 * p)); o0 += a.x; o1 += a.y; }
        *reinterpret_cast<bf162*>(yr + 2 * p) = __floats2bfloat162_rn(o0, o1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GEGLU: h = [a | g], y = a * gelu_erf(g)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float g) { return 0.5f * g * (1.f + erff(g * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_f(float g) {
  const float cdf = 0.5f * (1.f + erff(g * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * g * g);
  return cdf + g * pdf;
}

template <int MODE>
__global__ void __launch_bounds__(256) geglu_kernel(const bf16* __restrict__ h, long long ldh,
                                                    const bf16* __restrict__ dy, long long lddy,
                                                    bf16* __restrict__ out, long long ldo, long long total_vecs,
                                                    int F) {
  const int vecs = F >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vecs;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vecs;
    const int c = (int)(i % vecs) * 8;
    const uint4 ar = *reinterpret_cast<const uint4*>(h + row * ldh + c);
    const uint4 gr = *reinterpret_cast<const uint4*>(h + row * ldh + F + c);
    const uint32_t aw[4] = {ar.x, ar.y, ar.z, ar.w};
    const uint32_t gw[4] = {gr.x, gr.y, gr.z, gr.w};
    if (MODE == 0) {
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf162(aw[j]), g = unpack_bf162(gw[j]);
        o[j] = pack_bf162(a.x * gelu_f(g.x), a.y * gelu_f(g.y));
      }
      *reinterpret_cast<uint4*>(out + row * ldo + c) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
      const uint4 dr = *reinterpret_cast<const uint4*>(dy + row * lddy + c);
      const uint32_t dw[4] = {dr.x, dr.y, dr.z, dr.w};
      uint32_t oa[4], og[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf162(aw[j]), g = unpack_bf162(gw[j]), d = unpack_bf162(dw[j]);
        oa[j] = pack_bf162(d.x * gelu_f(g.x), d.y * gelu_f(g.y));
        og[j] = pack_bf162(d.x * a.x * dgelu_f(g.x), d.y * a.y * dgelu_f(g.y));
      }
      *reinterpret_cast<uint4*>(out + row * ldo + c) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
      *reinterpret_cast<uint4*>(out + row * ldo + F + c) = make_uint4(og[0], og[1], og[2], og[3]);
    }
  }
}

int gn_ppc(int hw) {  // pixels per CTA: aim for a few hundred CTAs
  int chunks = hw < 296 ? hw : 296;
  return vn_cdiv(hw, chunks);
}

int gn_check(int C, int groups, long long ldx) {
  VN_CHECK(groups > 0 && groups <= 64 && C % groups == 0, "groupnorm: C=%d groups=%d", C, groups);
  VN_CHECK((C / groups) % 2 == 0 && C % 8 == 0 && C <= kMaxC, "groupnorm: need even channels/group, C%%8==0, C<=%d (C=%d)", kMaxC, C);
  VN_CHECK(ldx % 8 == 0, "groupnorm: row stride must be a multiple of 8");
  return 0;
}

}  // namespace

extern "C" int vn_groupnorm_stats(const void* x, int64_t ldx, int nb, int hw, int C, int groups, float* stats,
                                  vn_stream_t s) {
  if (gn_check(C, groups, ldx)) return -1;
  const int ppc = gn_ppc(hw);
  dim3 grid(vn_cdiv(hw, ppc), nb);
  gn_reduce_kernel<0><<<grid, kGNThreads, 0, (cudaStream_t)s>>>((const bf16*)x, ldx, nullptr, 0, nullptr, nullptr,
                                                                  nullptr, 0.f, 0, stats, hw, C, groups, ppc);
  VN_LAUNCH_OK();
  return 0;
}

extern "C" int vn_groupnorm_apply(const void* x, int64_t ldx, const float* stats, const float* gamma,
                                  const float* beta, float eps, int silu, void* y, int64_t ldy, int nb, int hw, int C,
                                  int groups, vn_stream_t s) {
  if (gn_check(C, groups, ldx)) return -1;
  VN_CHECK(ldy % 8 == 0, "groupnorm: ldy must be a multiple of 8");
  int ppc = gn_ppc(hw);
  if (ppc < 4) ppc = hw < 4 ? hw : 4;       // amortise the per-CTA a/b staging
  dim3 grid(vn_cdiv(hw, ppc), nb);
  gn_apply_kernel<0><<<grid, kGNThreads, 2 * C * sizeof(float), (cudaStream_t)s>>>(
      (const bf16*)x, ldx, nullptr, 0, stats, nullptr, gamma, beta, eps, silu, nullptr, 0, nullptr, 0, (bf16*)y, ldy,
      hw, C, groups, ppc);
  VN_LAUNCH_OK();
  return 0;
}

extern "C" int vn_groupnorm_bwd_stats(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* stats,
                                      const float* gamma, const float* beta, float eps, int silu, float* red, int nb,
                                      int hw, int C, int groups, vn_stream_t s) {
  if (gn_check(C, groups, ldx)) return -1;
  const int ppc = gn_ppc(hw);
  dim3 grid(vn_cdiv(hw, ppc), nb);
  gn_reduce_kernel<1><<<grid, kGNThreads, 0, (cudaStream_t)s>>>((const bf16*)x, ldx, (const bf16*)dy, lddy, stats,
                                                                  gamma, beta, eps, silu, red, hw, C, groups, ppc);
  VN_LAUNCH_OK();
  return 0;
}

extern "C" int vn_groupnorm_bwd_apply(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* stats,
                                      const float* red, const float* gamma, const float* beta, float eps, int silu,
                                      const void* add1, int64_t ldadd1, const void* add2, int64_t ldadd2, void* dx,
                                      int64_t lddx, int nb, int hw, int C, int groups, vn_stream_t s) {
  if (gn_check(C, groups, ldx)) return -1;
  VN_CHECK(lddy % 8 == 0 && lddx % 8 == 0 && ldadd1 % 8 == 0 && ldadd2 % 8 == 0, "groupnorm bwd: strides must be multiples of 8");
  int ppc = gn_ppc(hw);
  if (ppc < 4) ppc = hw < 4 ? hw : 4;
  dim3 grid(vn_cdiv(hw, ppc), nb);
  gn_apply_kernel<1><<<grid, kGNThreads, (2 * C + 4 * groups) * sizeof(float), (cudaStream_t)s>>>(
      (const bf16*)x, ldx, (const bf16*)dy, lddy, stats, red, gamma, beta, eps, silu, (const bf16*)add1, ldadd1,
      (const bf16*)add2, ldadd2, (bf16*)dx, lddx, hw, C, groups, ppc);
  VN_LAUNCH_OK();
  return 0;
}

extern "C" int vn_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y,
                                int64_t ldy, float* stats, int rows, int C, vn_stream_t s) {
  VN_CHECK(C % 2 == 0 && C <= 64 * kLNMaxPairs, "layernorm: C=%d unsupported", C);
  ln_kernel<0><<<vn_cdiv(rows, 8), 256, 0, (cudaStream_t)s>>>((const bf16*)x, ldx, nullptr, 0, gamma, beta, eps, stats,
                                                                nullptr, 0, (bf16*)y, ldy, rows, C);
  VN_LAUNCH_OK();
  return 0;
}

extern "C" int vn_layernorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* gamma,
                                const float* stats, const void* add, int64_t ldadd, void* dx, int64_t lddx, int rows,
                                int C, vn_stream_t s) {
  VN_CHECK(C % 2 == 0 && C <= 64 * kLNMaxPairs, "layernorm: C=%d unsupported", C);
  ln_kernel<1><<<vn_cdiv(rows, 8), 256, 0, (cudaStream_t)s>>>((const bf16*)x, ldx, (const bf16*)dy, lddy, gamma, nullptr,
                                                                0.f, const_cast<float*>(stats), (const bf16*)add, ldadd,
                                                                (bf16*)dx, lddx, rows, C);
  VN_LAUNCH_OK();
  return 0;
}

extern "C" int vn_geglu_fwd(const void* h, int64_t ldh, void* y, int64_t ldy, int rows, int F, vn_stream_t s) {
  VN_CHECK(F % 8 == 0 && ldh % 8 == 0 && ldy % 8 == 0, "geglu: F and strides must be multiples of 8");
  const long long total = (long long)rows * (F >> 3);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  geglu_kernel<0><<<blocks, 256, 0, (cudaStream_t)s>>>((const bf16*)h, ldh, nullptr, 0, (bf16*)y, ldy, total, F);
  VN_LAUNCH_OK();
  return 0;
}

extern "C" int vn_geglu_bwd(const void* h, int64_t ldh, const void* dy, int64_t lddy, void* dh, int64_t lddh, int rows,
                            int F, vn_stream_t s) {
  VN_CHECK(F % 8 == 0 && ldh % 8 == 0 && lddy % 8 == 0 && lddh % 8 == 0, "geglu: F and strides must be multiples of 8");
  const long long total = (long long)rows * (F >> 3);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  geglu_kernel<1><<<blocks, 256, 0, (cudaStream_t)s>>>((const bf16*)h, ldh, (const bf16*)dy, lddy, (bf16*)dh, lddh, total,
                                                         F);
  VN_LAUNCH_OK();
  return 0;
}
