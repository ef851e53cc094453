// vn_norm.cu — GroupNorm(+SiLU), LayerNorm and GEGLU, forward and dgrad, over NHWC / token-major bf16.
// HBM/L2-bound streaming kernels: 128-bit coalesced accesses along the channel axis, fp32 statistics, several
// independent loads in flight per thread, grids sized for >= 2 CTAs per SM.  Replaces torch native_group_norm /
// native_layer_norm / gelu kernels (and their autograd backward) under diffusers ResnetBlock2D, Transformer2DModel,
// BasicTransformerBlock, GEGLU — all on the coach.py:197-214 path.
#include "vn_common.cuh"

namespace {

constexpr int kMaxC = 2560;       // widest GroupNorm input in SD-2.1 (concat 1280+1280)
constexpr int kMaxGroups = 64;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  float2 t;
  t = unpack_bf162(v.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf162(v.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf162(v.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf162(v.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&a)[8]) {
  uint4 o;
  o.x = pack_bf162(a[0], a[1]); o.y = pack_bf162(a[2], a[3]);
  o.z = pack_bf162(a[4], a[5]); o.w = pack_bf162(a[6], a[7]);
  return o;
}
__device__ __forceinline__ uint4 ld16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// ---------------------------------------------------------------------------------------------
// GroupNorm geometry: a CTA has vecs * k threads (vecs = C/8 16-byte vectors per pixel); thread (slot, v) owns the
// fixed channel vector v and walks pixels slot, slot + k, ... of the CTA's pixel range, so per-channel constants and
// partial sums live in registers and every warp access is a contiguous run of one pixel row.
// grid = (pixel chunks, nb).
// ---------------------------------------------------------------------------------------------
struct GNGeom {
  int threads, k, ppc, chunks;
};
GNGeom gn_geom(int nb, int hw, int C) {
  GNGeom g;
  const int vecs = C / 8;
  g.k = 256 / vecs;
  if (g.k < 1) g.k = 1;
  if (g.k > hw) g.k = hw;
  g.threads = vecs * g.k;
  int target = (148 * 3 + nb - 1) / nb;                 // CTAs per image for ~3 CTAs per SM
  int ppc = (hw + target - 1) / target;
  ppc = ((ppc + g.k - 1) / g.k) * g.k;                  // whole slots
  if (ppc < 2 * g.k && hw >= 2 * g.k) ppc = 2 * g.k;    // at least two pixels per thread to amortise the prologue
  g.ppc = ppc;
  g.chunks = (hw + ppc - 1) / ppc;
  return g;
}

// MODE 0: stats (sum x, sum x^2) per (image, group).   MODE 1 (backward): (sum dxhat, sum dxhat*xhat).
template <int MODE>
__global__ void __launch_bounds__(512) gn_reduce_kernel(const bf16* __restrict__ x, long long ldx,
                                                         const bf16* __restrict__ dy, long long lddy,
                                                         const double* __restrict__ stats,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, int silu,
                                                         double* __restrict__ out, int hw, int C, int groups, int k,
                                                         int ppc) {
  pdl_trigger();
  pdl_wait();
  // Per-thread partial sums are fp32 in a fixed order; everything that is combined in a scheduling-dependent order
  // (threads of a CTA, CTAs of the grid) is accumulated in fp64, so the statistics are reproducible run to run.
  __shared__ double s_acc[kMaxGroups][2];
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs, slot = threadIdx.x / vecs;
  const int c0 = v * 8;
  const int p0 = blockIdx.x * ppc;
  const int p1 = min(hw, p0 + ppc);
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups];
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    s_acc[g][0] = 0.0; s_acc[g][1] = 0.0;
    if (MODE == 1) {                       // the fp64 -> fp32 group constants are computed once per CTA
      const double n = (double)cpg * (double)hw;
      const double m = stats[(b * groups + g) * 2] / n;
      const double var = fmax(stats[(b * groups + g) * 2 + 1] / n - m * m, 0.0);
      s_mean[g] = (float)m; s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  __syncthreads();
  float ga[8], be[8], mean[8], rstd[8];
  if (MODE == 1) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c0 + j) / cpg;
      mean[j] = s_mean[g]; rstd[j] = s_rstd[g];
    }
  }
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
  const bf16* xb = x + (long long)b * hw * ldx + c0;
  const bf16* db = MODE == 1 ? dy + (long long)b * hw * lddy + c0 : nullptr;
  int p = p0 + slot;
  for (; p + 3 * k < p1; p += 4 * k) {
    uint4 xv[4], dv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      xv[u] = ld16(xb + (long long)(p + u * k) * ldx);
      if (MODE == 1) dv[u] = ld16(db + (long long)(p + u * k) * lddy);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float xf[8];
      unpack8(xv[u], xf);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += xf[j]; a1[j] = fmaf(xf[j], xf[j], a1[j]); }
      } else {
        float df[8];
        unpack8(dv[u], df);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xf[j] - mean[j]) * rstd[j];
          float dz = df[j];
          if (silu) dz *= dsilu_f(xh * ga[j] + be[j]);
          const float dh = dz * ga[j];
          a0[j] += dh; a1[j] = fmaf(dh, xh, a1[j]);
        }
      }
    }
  }
  for (; p < p1; p += k) {
    float xf[8];
    unpack8(ld16(xb + (long long)p * ldx), xf);
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { a0[j] += xf[j]; a1[j] = fmaf(xf[j], xf[j], a1[j]); }
    } else {
      float df[8];
      unpack8(ld16(db + (long long)p * lddy), df);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xf[j] - mean[j]) * rstd[j];
        float dz = df[j];
        if (silu) dz *= dsilu_f(xh * ga[j] + be[j]);
        const float dh = dz * ga[j];
        a0[j] += dh; a1[j] = fmaf(dh, xh, a1[j]);
      }
    }
  }
  // channels -> groups (a vector spans at most 1 + 8/cpg groups); merge equal groups before the shared atomics
  int gprev = c0 / cpg;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c0 + j) / cpg;
    if (g != gprev) {
      atomicAdd(&s_acc[gprev][0], (double)s0); atomicAdd(&s_acc[gprev][1], (double)s1);
      s0 = 0.f; s1 = 0.f; gprev = g;
    }
    s0 += a0[j]; s1 += a1[j];
  }
  atomicAdd(&s_acc[gprev][0], (double)s0); atomicAdd(&s_acc[gprev][1], (double)s1);
  __syncthreads();
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    atomicAdd(&out[(b * groups + g) * 2], s_acc[g][0]);
    atomicAdd(&out[(b * groups + g) * 2 + 1], s_acc[g][1]);
  }
}

// MODE 0 forward: y = act(x * a[c] + b[c]).  MODE 1 backward: dx = rstd*(dxhat - S1/n - xhat*S2/n) (+add1) (+add2).
template <int MODE>
__global__ void __launch_bounds__(512) gn_apply_kernel(const bf16* __restrict__ x, long long ldx,
                                                        const bf16* __restrict__ dy, long long lddy,
                                                        const double* __restrict__ stats,
                                                        const double* __restrict__ red,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int silu,
                                                        const bf16* __restrict__ add1, long long ld1,
                                                        const bf16* __restrict__ add2, long long ld2,
                                                        bf16* __restrict__ y, long long ldy, int hw, int C,
                                                        int groups, int k, int ppc) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs, slot = threadIdx.x / vecs;
  const int c0 = v * 8;
  const int p0 = blockIdx.x * ppc;
  const int p1 = min(hw, p0 + ppc);
  // group constants once per CTA (fp64 statistics -> fp32), then per-channel constants from shared memory:
  //   fwd  y = x*A + B            (A = gamma*rstd, B = beta - mean*A)
  //   bwd  xh = x*R + M (R = rstd, M = -mean*rstd), z = xh*G + Bt, dx = R*(dz*G - S1 - xh*S2)
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups], s_s1[kMaxGroups], s_s2[kMaxGroups];
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    const double n = (double)cpg * (double)hw;
    const double md = stats[(b * groups + g) * 2] / n;
    const double var = fmax(stats[(b * groups + g) * 2 + 1] / n - md * md, 0.0);
    s_mean[g] = (float)md;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    if (MODE == 1) {
      s_s1[g] = (float)(red[(b * groups + g) * 2] / n);
      s_s2[g] = (float)(red[(b * groups + g) * 2 + 1] / n);
    }
  }
  __syncthreads();
  float A[8], Bc[8], G[8], S1[8], S2[8], Bt[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c0 + j) / cpg;
      const float m = s_mean[g], rs = s_rstd[g];
      if (MODE == 0) {
        A[j] = gam[j] * rs;
        Bc[j] = bet[j] - m * A[j];
      } else {
        A[j] = rs; Bc[j] = -m * rs;
        G[j] = gam[j]; Bt[j] = bet[j];
        S1[j] = s_s1[g]; S2[j] = s_s2[g];
      }
    }
  }
  const bf16* xb = x + (long long)b * hw * ldx + c0;
  const bf16* db = MODE == 1 ? dy + (long long)b * hw * lddy + c0 : nullptr;
  const bf16* a1b = add1 ? add1 + (long long)b * hw * ld1 + c0 : nullptr;
  const bf16* a2b = add2 ? add2 + (long long)b * hw * ld2 + c0 : nullptr;
  bf16* yb = y + (long long)b * hw * ldy + c0;
  constexpr int U = MODE == 0 ? 4 : 2;
  int p = p0 + slot;
  for (; p + (U - 1) * k < p1; p += U * k) {
    uint4 xv[U], dv[U], r1[U], r2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pp = p + u * k;
      xv[u] = ld16(xb + pp * ldx);
      if (MODE == 1) {
        dv[u] = ld16(db + pp * lddy);
        if (a1b) r1[u] = ld16(a1b + pp * ld1);
        if (a2b) r2[u] = ld16(a2b + pp * ld2);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float xf[8], o[8];
      unpack8(xv[u], xf);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(xf[j], A[j], Bc[j]);
          o[j] = silu ? silu_f(z) : z;
        }
      } else {
        float df[8];
        unpack8(dv[u], df);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = fmaf(xf[j], A[j], Bc[j]);
          float dz = df[j];
          if (silu) dz *= dsilu_f(fmaf(xh, G[j], Bt[j]));
          o[j] = A[j] * (dz * G[j] - S1[j] - xh * S2[j]);
        }
        if (a1b) { float t[8]; unpack8(r1[u], t);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += t[j]; }
        if (a2b) { float t[8]; unpack8(r2[u], t);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += t[j]; }
      }
      *reinterpret_cast<uint4*>(yb + (long long)(p + u * k) * ldy) = pack8(o);
    }
  }
  for (; p < p1; p += k) {
    float xf[8], o[8];
    unpack8(ld16(xb + (long long)p * ldx), xf);
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(xf[j], A[j], Bc[j]);
        o[j] = silu ? silu_f(z) : z;
      }
    } else {
      float df[8];
      unpack8(ld16(db + (long long)p * lddy), df);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(xf[j], A[j], Bc[j]);
        float dz = df[j];
        if (silu) dz *= dsilu_f(fmaf(xh, G[j], Bt[j]));
        o[j] = A[j] * (dz * G[j] - S1[j] - xh * S2[j]);
      }
      if (a1b) { float t[8]; unpack8(ld16(a1b + (long long)p * ld1), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += t[j]; }
      if (a2b) { float t[8]; unpack8(ld16(a2b + (long long)p * ld2), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += t[j]; }
    }
    *reinterpret_cast<uint4*>(yb + (long long)p * ldy) = pack8(o);
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers as NV 16-byte vectors per lane (C <= 256*NV).
// ---------------------------------------------------------------------------------------------
template <int MODE, int NV>
__global__ void __launch_bounds__(256) ln_kernel(const bf16* __restrict__ x, long long ldx,
                                                 const bf16* __restrict__ dy, long long lddy,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 float eps, float* __restrict__ stats, const bf16* __restrict__ add,
                                                 long long ldadd, bf16* __restrict__ y, long long ldy, int rows,
                                                 int C) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int vecs = C >> 3;
  const bf16* xr = x + (long long)row * ldx;
  float xf[NV][8];
  uint4 raw[NV], draw[NV], araw[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < vecs) {
      raw[i] = ld16(xr + v * 8);
      if (MODE == 1) {
        draw[i] = ld16(dy + (long long)row * lddy + v * 8);
        if (add) araw[i] = ld16(add + (long long)row * ldadd + v * 8);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + i * 32 < vecs) {
      unpack8(raw[i], xf[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += xf[i][j];
    }
  }
  bf16* yr = y + (long long)row * ldy;
  if (MODE == 0) {
    const float mean = warp_sum(s) / (float)C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < vecs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float a = xf[i][j] - mean; ss = fmaf(a, a, ss); }
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
    if (lane == 0 && stats) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < vecs) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (xf[i][j] - mean) * rstd * g[j] + bb[j];
        *reinterpret_cast<uint4*>(yr + v * 8) = pack8(o);
      }
    }
  } else {
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float dh[NV][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < vecs) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float df[8];
        unpack8(draw[i], df);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dh[i][j] = df[j] * g[j];
          xf[i][j] = (xf[i][j] - mean) * rstd;
          c1 += dh[i][j];
          c2 = fmaf(dh[i][j], xf[i][j], c2);
        }
      }
    }
    c1 = warp_sum(c1) / (float)C;
    c2 = warp_sum(c2) / (float)C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < vecs) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (dh[i][j] - c1 - xf[i][j] * c2);
        if (add) {
          float t[8];
          unpack8(araw[i], t);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += t[j];
        }
        *reinterpret_cast<uint4*>(yr + v * 8) = pack8(o);
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
  pdl_trigger();
  pdl_wait();
  const int vecs = F >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vecs;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vecs;
    const int c = (int)(i % vecs) * 8;
    float a[8], g[8];
    unpack8(ld16(h + row * ldh + c), a);
    unpack8(ld16(h + row * ldh + F + c), g);
    if (MODE == 0) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = a[j] * gelu_f(g[j]);
      *reinterpret_cast<uint4*>(out + row * ldo + c) = pack8(o);
    } else {
      float d[8], oa[8], og[8];
      unpack8(ld16(dy + row * lddy + c), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        oa[j] = d[j] * gelu_f(g[j]);
        og[j] = d[j] * a[j] * dgelu_f(g[j]);
      }
      *reinterpret_cast<uint4*>(out + row * ldo + c) = pack8(oa);
      *reinterpret_cast<uint4*>(out + row * ldo + F + c) = pack8(og);
    }
  }
}

int gn_check(int C, int groups, long long ldx) {
  VN_CHECK(groups > 0 && groups <= kMaxGroups && C % groups == 0, "groupnorm: C=%d groups=%d", C, groups);
  VN_CHECK(C % 8 == 0 && C <= kMaxC * 4, "groupnorm: need C %% 8 == 0 (C=%d)", C);
  VN_CHECK(C / 8 <= 512, "groupnorm: C=%d too wide (C <= 4096)", C);
  VN_CHECK(ldx % 8 == 0, "groupnorm: row stride must be a multiple of 8");
  return 0;
}

template <int MODE>
int ln_launch(const bf16* x, long long ldx, const bf16* dy, long long lddy, const float* gamma, const float* beta,
              float eps, float* stats, const bf16* add, long long ldadd, bf16* y, long long ldy, int rows, int C,
              cudaStream_t s) {
  VN_CHECK(C % 8 == 0 && C <= 2048, "layernorm: C=%d unsupported (C %% 8 == 0, C <= 2048)", C);
  VN_CHECK(ldx % 8 == 0 && ldy % 8 == 0 && lddy % 8 == 0 && ldadd % 8 == 0, "layernorm: strides must be multiples of 8");
  const int nv = (C / 8 + 31) / 32;
  const int grid = vn_cdiv(rows, 8);
#define VN_LN_CASE(NV)                                                                                              \
  case NV:                                                                                                          \
    VN_LAUNCH((ln_kernel<MODE, NV>), grid, 256, 0, s, x, ldx, dy, lddy, gamma, beta, eps, stats, add, ldadd, y, ldy, rows, C); \
    break;
  switch (nv) {
    VN_LN_CASE(1) VN_LN_CASE(2) VN_LN_CASE(3) VN_LN_CASE(4) VN_LN_CASE(5) VN_LN_CASE(6) VN_LN_CASE(7) VN_LN_CASE(8)
    default: VN_CHECK(false, "layernorm: C=%d unsupported", C);
  }
#undef VN_LN_CASE
  return 0;
}

}  // namespace

extern "C" int vn_groupnorm_stats(const void* x, int64_t ldx, int nb, int hw, int C, int groups, double* stats,
                                  vn_stream_t s) {
  if (gn_check(C, groups, ldx)) return -1;
  const GNGeom g = gn_geom(nb, hw, C);
  dim3 grid(g.chunks, nb);
  VN_LAUNCH(gn_reduce_kernel<0>, grid, g.threads, 0, (cudaStream_t)s, (const bf16*)x, ldx, nullptr, 0, nullptr, nullptr,
                                                                 nullptr, 0.f, 0, stats, hw, C, groups, g.k, g.ppc);
  return 0;
}

extern "C" int vn_groupnorm_apply(const void* x, int64_t ldx, const double* stats, const float* gamma,
                                  const float* beta, float eps, int silu, void* y, int64_t ldy, int nb, int hw, int C,
                                  int groups, vn_stream_t s) {
  if (gn_check(C, groups, ldx)) return -1;
  VN_CHECK(ldy % 8 == 0, "groupnorm: ldy must be a multiple of 8");
  const GNGeom g = gn_geom(nb, hw, C);
  dim3 grid(g.chunks, nb);
  VN_LAUNCH(gn_apply_kernel<0>, grid, g.threads, 0, (cudaStream_t)s, (const bf16*)x, ldx, nullptr, 0, stats, nullptr, gamma,
                                                                beta, eps, silu, nullptr, 0, nullptr, 0, (bf16*)y, ldy,
                                                                hw, C, groups, g.k, g.ppc);
  return 0;
}

extern "C" int vn_groupnorm_bwd_stats(const void* x, int64_t ldx, const void* dy, int64_t lddy, const double* stats,
                                      const float* gamma, const float* beta, float eps, int silu, double* red, int nb,
                                      int hw, int C, int groups, vn_stream_t s) {
  if (gn_check(C, groups, ldx)) return -1;
  VN_CHECK(lddy % 8 == 0, "groupnorm bwd: strides must be multiples of 8");
  const GNGeom g = gn_geom(nb, hw, C);
  dim3 grid(g.chunks, nb);
  VN_LAUNCH(gn_reduce_kernel<1>, grid, g.threads, 0, (cudaStream_t)s, (const bf16*)x, ldx, (const bf16*)dy, lddy, stats, gamma,
                                                                 beta, eps, silu, red, hw, C, groups, g.k, g.ppc);
  return 0;
}

extern "C" int vn_groupnorm_bwd_apply(const void* x, int64_t ldx, const void* dy, int64_t lddy, const double* stats,
                                      const double* red, const float* gamma, const float* beta, float eps, int silu,
                                      const void* add1, int64_t ldadd1, const void* add2, int64_t ldadd2, void* dx,
                                      int64_t lddx, int nb, int hw, int C, int groups, vn_stream_t s) {
  if (gn_check(C, groups, ldx)) return -1;
  VN_CHECK(lddy % 8 == 0 && lddx % 8 == 0 && ldadd1 % 8 == 0 && ldadd2 % 8 == 0, "groupnorm bwd: strides must be multiples of 8");
  const GNGeom g = gn_geom(nb, hw, C);
  dim3 grid(g.chunks, nb);
  VN_LAUNCH(gn_apply_kernel<1>, grid, g.threads, 0, (cudaStream_t)s, (const bf16*)x, ldx, (const bf16*)dy, lddy, stats, red,
                                                                gamma, beta, eps, silu, (const bf16*)add1, ldadd1,
                                                                (const bf16*)add2, ldadd2, (bf16*)dx, lddx, hw, C, groups,
                                                                g.k, g.ppc);
  return 0;
}

extern "C" int vn_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y,
                                int64_t ldy, float* stats, int rows, int C, vn_stream_t s) {
  return ln_launch<0>((const bf16*)x, ldx, nullptr, 0, gamma, beta, eps, stats, nullptr, 0, (bf16*)y, ldy, rows, C,
                      (cudaStream_t)s);
}

extern "C" int vn_layernorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* gamma,
                                const float* stats, const void* add, int64_t ldadd, void* dx, int64_t lddx, int rows,
                                int C, vn_stream_t s) {
  return ln_launch<1>((const bf16*)x, ldx, (const bf16*)dy, lddy, gamma, nullptr, 0.f, const_cast<float*>(stats),
                      (const bf16*)add, ldadd, (bf16*)dx, lddx, rows, C, (cudaStream_t)s);
}

extern "C" int vn_geglu_fwd(const void* h, int64_t ldh, void* y, int64_t ldy, int rows, int F, vn_stream_t s) {
  VN_CHECK(F % 8 == 0 && ldh % 8 == 0 && ldy % 8 == 0, "geglu: F and strides must be multiples of 8");
  const long long total = (long long)rows * (F >> 3);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  VN_LAUNCH(geglu_kernel<0>, blocks, 256, 0, (cudaStream_t)s, (const bf16*)h, ldh, nullptr, 0, (bf16*)y, ldy, total, F);
  return 0;
}

extern "C" int vn_geglu_bwd(const void* h, int64_t ldh, const void* dy, int64_t lddy, void* dh, int64_t lddh, int rows,
                            int F, vn_stream_t s) {
  VN_CHECK(F % 8 == 0 && ldh % 8 == 0 && lddy % 8 == 0 && lddh % 8 == 0, "geglu: F and strides must be multiples of 8");
  const long long total = (long long)rows * (F >> 3);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  VN_LAUNCH(geglu_kernel<1>, blocks, 256, 0, (cudaStream_t)s, (const bf16*)h, ldh, (const bf16*)dy, lddy, (bf16*)dh, lddh, total,
                                                         F);
  return 0;
}
