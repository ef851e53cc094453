// vn_norm.cu — GroupNorm(+SiLU), LayerNorm and GEGLU, forward and dgrad, over NHWC / token-major bf16.
// HBM/L2-bound streaming kernels: 128-bit coalesced accesses along the channel axis, fp32 statistics, several
// independent loads in flight per thread, grids sized for >= 2 CTAs per SM.  Replaces torch native_group_norm /
// native_layer_norm / gelu kernels (and their autograd backward) under diffusers ResnetBlock2D, Transformer2DModel,
// BasicTransformerBlock, GEGLU — all on the coach.py:197-214 path.
#include "vn_common.cuh"

#include <stdlib.h>

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
// Fused GroupNorm: statistics + apply in ONE launch.  At per-GPU batch 1 a GroupNorm tensor is 0.2 - 8 MB and both
// phases are bound by launch / load latency, not bandwidth, so the two-kernel form pays that latency twice (119 x per
// train step).  Here the grid is at most one CTA per SM and every CTA
//   1. loads its slab of pixels into REGISTERS (NP vectors per thread; NP == 0: too large, phase 3 re-reads it from L2),
//   2. reduces it to 64 (group, statistic) fp32 partials - shared memory + T/64 lanes per pair, no atomics,
//   3. stores them to its slot of `partials` (every word preset to 0xffffffff by the host once per pass) and polls ALL
//      CTAs' slots until no word is 0xffffffff: the data is its own flag, so there is no fence, no arrival counter
//      (same-address atomics serialise at ~20 ns each on B200: 148 arrivals cost more than the whole reduction) and
//      no second hop; every CTA adds the partials up in the same fixed order in fp64,
//   4. normalises its slab from registers.
// All CTAs are co-resident by construction (grid <= #SMs, and nothing that waits on this kernel can hold an SM it
// needs: PDL successors start only after every CTA of this grid has started), which the polling needs; a protocol bug
// traps instead of hanging.  Shared / global fp64 atomics and in-kernel fp64 div / sqrt, measured at ~2 us per phase
// (scripts/kernel_timeline.py), are gone; the sums are deterministic by construction.
//   MODE 0: y = act(GN(x));  acc = (sum x, sum x^2) kept for the backward.
//   MODE 1: dx = GN^T(dy) (+add1) (+add2);  stats = forward sums, acc = (sum dxhat, sum dxhat*xhat).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4* p) {          // L2 read, never served from a stale L1 line
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

struct GNConst {                                  // per-thread constants of its 8 channels
  float gam[8], bet[8];
  float R[8], M[8];                               // MODE 1: xhat = x*R + M
  float A[8], Bc[8];                              // MODE 0 phase 2: y = x*A + Bc
  float S1[8], S2[8];                             // MODE 1 phase 2
};

template <int MODE>
__device__ __forceinline__ void gn_accumulate(const GNConst& c, int silu, const uint4& xraw, const uint4& draw,
                                              float (&a0)[8], float (&a1)[8]) {
  float xf[8];
  unpack8(xraw, xf);
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { a0[j] += xf[j]; a1[j] = fmaf(xf[j], xf[j], a1[j]); }
  } else {
    float df[8];
    unpack8(draw, df);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = fmaf(xf[j], c.R[j], c.M[j]);
      float dz = df[j];
      if (silu) dz *= dsilu_f(fmaf(xh, c.gam[j], c.bet[j]));
      const float dh = dz * c.gam[j];
      a0[j] += dh; a1[j] = fmaf(dh, xh, a1[j]);
    }
  }
}

template <int MODE>
__device__ __forceinline__ uint4 gn_finish(const GNConst& c, int silu, const uint4& xraw, const uint4& draw, bool h1,
                                           const uint4& r1, bool h2, const uint4& r2) {
  float xf[8], o[8];
  unpack8(xraw, xf);
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(xf[j], c.A[j], c.Bc[j]);
      o[j] = silu ? silu_f(z) : z;
    }
  } else {
    float df[8];
    unpack8(draw, df);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = fmaf(xf[j], c.R[j], c.M[j]);
      float dz = df[j];
      if (silu) dz *= dsilu_f(fmaf(xh, c.gam[j], c.bet[j]));
      o[j] = c.R[j] * (dz * c.gam[j] - c.S1[j] - xh * c.S2[j]);
    }
    if (h1) { float t[8]; unpack8(r1, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += t[j]; }
    if (h2) { float t[8]; unpack8(r2, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += t[j]; }
  }
  return pack8(o);
}

template <int MODE, int NP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) gn_fused_kernel(const bf16* __restrict__ x, long long ldx,
                                                           const bf16* __restrict__ dy, long long lddy,
                                                           const double* __restrict__ stats,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, int silu,
                                                           double* acc, float* part,
                                                           const bf16* __restrict__ add1, long long ld1,
                                                           const bf16* __restrict__ add2, long long ld2,
                                                           bf16* __restrict__ y, long long ldy, int hw, int C,
                                                           int groups, int k, int ppc, long long* dbg) {
  pdl_trigger();
#ifdef VN_TIMELINE
#define GN_STAMP(slot)                                                                                     \
  do {                                                                                                     \
    if (dbg && threadIdx.x == 0) dbg[((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = clock64(); \
  } while (0)
#else
#define GN_STAMP(slot) do { } while (0)
#endif
  GN_STAMP(1);
  constexpr int NR = NP > 0 ? NP : 1;
  constexpr int kPairs = 64;                       // (group, statistic) pairs per image: groups <= 32
  __shared__ __align__(16) float s_part[2][MAXT * 8];   // [statistic][thread's 8 channels]
  __shared__ double s_tot[kPairs];
  __shared__ double s_wq[(MAXT / 32) * 16 * 4];
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups], s_s1[kMaxGroups], s_s2[kMaxGroups];
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs, slot = threadIdx.x / vecs;
  const int c0 = v * 8;
  const int p0 = blockIdx.x * ppc;
  const int p1 = min(hw, p0 + ppc);
  const double inv_n = 1.0 / ((double)cpg * (double)hw);
  GNConst c;
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    c.gam[0] = g0.x; c.gam[1] = g0.y; c.gam[2] = g0.z; c.gam[3] = g0.w; c.gam[4] = g1.x; c.gam[5] = g1.y; c.gam[6] = g1.z; c.gam[7] = g1.w;
    c.bet[0] = b0.x; c.bet[1] = b0.y; c.bet[2] = b0.z; c.bet[3] = b0.w; c.bet[4] = b1.x; c.bet[5] = b1.y; c.bet[6] = b1.z; c.bet[7] = b1.w;
  }
  pdl_wait();          // parameters above are frozen; activations / statistics only from here on
  GN_STAMP(2);
  if (MODE == 1) {
    // E[x^2] - mean^2 in fp64 (cancellation), the reciprocal square root in fp32 (fp64 div / sqrt are long sequences)
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      const double m = stats[(b * groups + g) * 2] * inv_n;
      const double var = fmax(fma(-m, m, stats[(b * groups + g) * 2 + 1] * inv_n), 0.0);
      s_mean[g] = (float)m; s_rstd[g] = 1.0f / sqrtf((float)var + eps);
    }
    __syncthreads();
  }
  // group of each of the thread's 8 channels: one division, then a running remainder
  int gj[8];
  {
    int g = c0 / cpg, rem = c0 - g * cpg;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (rem == cpg) { rem = 0; ++g; }
      gj[j] = g;
      ++rem;
    }
  }
  if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c.R[j] = s_rstd[gj[j]]; c.M[j] = -s_mean[gj[j]] * s_rstd[gj[j]];
    }
  }
  const bf16* xb = x + (long long)b * hw * ldx + c0;
  const bf16* db = MODE == 1 ? dy + (long long)b * hw * lddy + c0 : xb;

  // ---- phase 1: partial sums of the CTA's slab ----
  uint4 xv[NR], dv[NR];
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
  if (NP > 0) {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int p = p0 + slot + i * k;
      xv[i] = make_uint4(0u, 0u, 0u, 0u);
      dv[i] = make_uint4(0u, 0u, 0u, 0u);
      if (p < p1) {
        xv[i] = ld16(xb + (long long)p * ldx);
        if (MODE == 1) dv[i] = ld16(db + (long long)p * lddy);
      }
    }
#pragma unroll
    for (int i = 0; i < NR; ++i)
      if (p0 + slot + i * k < p1) gn_accumulate<MODE>(c, silu, xv[i], dv[i], a0, a1);
  } else {
    int p = p0 + slot;
    for (; p + 3 * k < p1; p += 4 * k) {
      uint4 xq[4], dq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xq[u] = ld16(xb + (long long)(p + u * k) * ldx);
        dq[u] = xq[u];
        if (MODE == 1) dq[u] = ld16(db + (long long)(p + u * k) * lddy);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) gn_accumulate<MODE>(c, silu, xq[u], dq[u], a0, a1);
    }
    for (; p < p1; p += k) {
      uint4 xq = ld16(xb + (long long)p * ldx), dq = xq;
      if (MODE == 1) dq = ld16(db + (long long)p * lddy);
      gn_accumulate<MODE>(c, silu, xq, dq, a0, a1);
    }
  }
  GN_STAMP(3);
  // ---- CTA reduction without atomics: per-thread channel sums -> shared memory -> T/64 lanes per (group, statistic) ----
  {
    float* q0 = &s_part[0][threadIdx.x * 8];
    float* q1 = &s_part[1][threadIdx.x * 8];
    *reinterpret_cast<float4*>(q0) = make_float4(a0[0], a0[1], a0[2], a0[3]);
    *reinterpret_cast<float4*>(q0 + 4) = make_float4(a0[4], a0[5], a0[6], a0[7]);
    *reinterpret_cast<float4*>(q1) = make_float4(a1[0], a1[1], a1[2], a1[3]);
    *reinterpret_cast<float4*>(q1 + 4) = make_float4(a1[4], a1[5], a1[6], a1[7]);
  }
  __syncthreads();
  const int T = blockDim.x;
  // L = 4 / 2 / 1 adjacent lanes per (group, statistic) pair: the first 64*L threads (whole warps) do the reduction
  const int lsh = T >= 256 ? 2 : T >= 128 ? 1 : 0;
  const int L = 1 << lsh;
  if ((int)threadIdx.x < kPairs * L) {
    // thread (slot, v) wrote channels [v*8, v*8+8) at slot*C: pair (g, st) = k runs of cpg contiguous floats.  A lane
    // takes whole runs (k >= L) or a segment of one run (k < L); loads are independent, so the loop pipelines.
    const int pr = threadIdx.x >> lsh, sub = threadIdx.x & (L - 1);
    float sum = 0.f;
    if (pr < 2 * groups) {
      const int g = pr >> 1, st = pr & 1;
      int sl = sub, sl_step = L, seg = 0, psh = 0;
      if (k < L) {
        psh = k == 1 ? lsh : (k == 2 && L == 4) ? 1 : 0;      // parts = L / k rounded down to a power of two
        while (sl >= k) { sl -= k; ++seg; }
        sl_step = k;                                           // one run per lane
        if (seg >= (1 << psh)) sl = k;                         // surplus lane
      }
      const int c_lo = (seg * cpg) >> psh, c_hi = ((seg + 1) * cpg) >> psh;
      for (; sl < k; sl += sl_step) {
        const float* src = &s_part[st][sl * C + g * cpg];
#pragma unroll 4
        for (int ch = c_lo; ch < c_hi; ++ch) sum += src[ch];
      }
    }
    for (int o = L >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (sub == 0) {
      unsigned bits = __float_as_uint(sum);
      if (bits == 0xffffffffu) bits = 0x7fffffffu;             // 0xffffffff is the "not written yet" pattern
      __stcg(reinterpret_cast<unsigned*>(part) + ((long long)b * gridDim.x + blockIdx.x) * kPairs + pr, bits);
    }
  }
  GN_STAMP(4);
  // ---- exchange: every CTA polls ALL CTAs' partials (preset to 0xffffffff by the host once per pass) until they are
  // written, and adds them up in the same fixed order (fp64).  No atomics, no arrival counter (same-address atomics
  // serialise at ~20 ns each: 148 arrivals cost more than the whole reduction), no second hop, deterministic. ----
  {
    const uint4* pv = reinterpret_cast<const uint4*>(part + (long long)b * gridDim.x * kPairs);
    const int nvec = gridDim.x * (kPairs / 4);     // blockDim.x % 16 == 0: a thread always meets the same four pairs
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    for (int base = 0; base < nvec; base += 8 * T) {
      uint4 f[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + threadIdx.x + u * T;
        f[u] = make_uint4(0u, 0u, 0u, 0u);
        if (i < nvec) f[u] = ld_relaxed_v4(pv + i);
      }
      long long tstart = 0;
      for (unsigned spins = 0;; ++spins) {
        bool ok = true;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = base + threadIdx.x + u * T;
          if (i < nvec && (f[u].x == 0xffffffffu || f[u].y == 0xffffffffu || f[u].z == 0xffffffffu || f[u].w == 0xffffffffu)) {
            f[u] = ld_relaxed_v4(pv + i);
            ok = false;
          }
        }
        if (ok) break;
        if (spins == 0) tstart = clock64();
        if ((spins & 0xffu) == 0xffu && (clock64() - tstart) > 4000000000LL) {
          printf("viewneti: groupnorm partials never arrived (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
          __trap();
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        t0 += (double)__uint_as_float(f[u].x); t1 += (double)__uint_as_float(f[u].y);
        t2 += (double)__uint_as_float(f[u].z); t3 += (double)__uint_as_float(f[u].w);
      }
    }
    GN_STAMP(6);
    // thread t holds the sums of pairs 4*(t%16) .. +3 over its CTAs: lanes l and l+16 of a (whole) warp share a quad
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwhole = T >> 5, nw = (T + 31) >> 5;
    if (wid < nwhole) {
      t0 += __shfl_xor_sync(0xffffffffu, t0, 16); t1 += __shfl_xor_sync(0xffffffffu, t1, 16);
      t2 += __shfl_xor_sync(0xffffffffu, t2, 16); t3 += __shfl_xor_sync(0xffffffffu, t3, 16);
    }
    double* wq = s_wq;        // [warp][16 quads][4]; NOT aliased on s_part: its readers are not synchronised with us
    if (lane < 16) {
      double* d = wq + (wid * 16 + lane) * 4;
      d[0] = t0; d[1] = t1; d[2] = t2; d[3] = t3;
    }
    __syncthreads();
    if ((int)threadIdx.x < kPairs * L) {
      const int pr = threadIdx.x >> lsh, sub = threadIdx.x & (L - 1);
      double tot = 0.0;
      for (int w = sub; w < nw; w += L) tot += wq[(w * 16 + (pr >> 2)) * 4 + (pr & 3)];
      for (int o = L >> 1; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (sub == 0 && pr < 2 * groups) {
        s_tot[pr] = tot;
        if (blockIdx.x == 0) acc[(long long)b * groups * 2 + pr] = tot;     // kept for the backward / the caller
      }
    }
    __syncthreads();
  }
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    const double t0 = s_tot[2 * g], t1 = s_tot[2 * g + 1];
    if (MODE == 0) {
      const double m = t0 * inv_n;
      const double var = fmax(fma(-m, m, t1 * inv_n), 0.0);
      s_mean[g] = (float)m; s_rstd[g] = 1.0f / sqrtf((float)var + eps);
    } else {
      s_s1[g] = (float)(t0 * inv_n); s_s2[g] = (float)(t1 * inv_n);
    }
  }
  __syncthreads();
  GN_STAMP(8);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = gj[j];
    if (MODE == 0) {
      c.A[j] = c.gam[j] * s_rstd[g];
      c.Bc[j] = c.bet[j] - s_mean[g] * c.A[j];
    } else {
      c.S1[j] = s_s1[g]; c.S2[j] = s_s2[g];
    }
  }
  const bool h1 = MODE == 1 && add1 != nullptr, h2 = MODE == 1 && add2 != nullptr;
  const bf16* a1b = h1 ? add1 + (long long)b * hw * ld1 + c0 : xb;
  const bf16* a2b = h2 ? add2 + (long long)b * hw * ld2 + c0 : xb;
  bf16* yb = y + (long long)b * hw * ldy + c0;
  if (NP > 0) {
    constexpr int U = NR < 4 ? NR : 4;
#pragma unroll
    for (int i0 = 0; i0 < NR; i0 += U) {
      uint4 r1[U], r2[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = p0 + slot + (i0 + u) * k;
        r1[u] = make_uint4(0u, 0u, 0u, 0u);
        r2[u] = make_uint4(0u, 0u, 0u, 0u);
        if (MODE == 1 && p < p1) {
          if (h1) r1[u] = ld16(a1b + (long long)p * ld1);
          if (h2) r2[u] = ld16(a2b + (long long)p * ld2);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = p0 + slot + (i0 + u) * k;
        if (p < p1)
          *reinterpret_cast<uint4*>(yb + (long long)p * ldy) = gn_finish<MODE>(c, silu, xv[i0 + u], dv[i0 + u], h1, r1[u], h2, r2[u]);
      }
    }
  } else {
    constexpr int U = MODE == 0 ? 4 : 2;
    int p = p0 + slot;
    for (; p + (U - 1) * k < p1; p += U * k) {
      uint4 xq[U], dq[U], r1[U], r2[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long pp = p + u * k;
        xq[u] = ld16(xb + pp * ldx);
        dq[u] = xq[u]; r1[u] = xq[u]; r2[u] = xq[u];
        if (MODE == 1) {
          dq[u] = ld16(db + pp * lddy);
          if (h1) r1[u] = ld16(a1b + pp * ld1);
          if (h2) r2[u] = ld16(a2b + pp * ld2);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        *reinterpret_cast<uint4*>(yb + (long long)(p + u * k) * ldy) = gn_finish<MODE>(c, silu, xq[u], dq[u], h1, r1[u], h2, r2[u]);
    }
    for (; p < p1; p += k) {
      uint4 xq = ld16(xb + (long long)p * ldx), dq = xq, r1 = xq, r2 = xq;
      if (MODE == 1) {
        dq = ld16(db + (long long)p * lddy);
        if (h1) r1 = ld16(a1b + (long long)p * ld1);
        if (h2) r2 = ld16(a2b + (long long)p * ld2);
      }
      *reinterpret_cast<uint4*>(yb + (long long)p * ldy) = gn_finish<MODE>(c, silu, xq, dq, h1, r1, h2, r2);
    }
  }
  GN_STAMP(7);
#undef GN_STAMP
}

// Geometry of the fused form: at most one CTA per SM over the whole grid (all co-resident, the exchange needs that),
// up to maxt threads per CTA with threads % 16 == 0 (exchange indexing) and >= 64 (one lane per pair at least).
bool gn_fused_geom(int nb, int hw, int C, int maxt, GNGeom* out) {
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int vecs = C / 8;
  if (vecs > maxt || nb > sms) return false;
  GNGeom g;
  g.k = maxt / vecs;
  if (g.k > hw) g.k = hw;
  while (g.k > 1 && (vecs * g.k) % 16 != 0) --g.k;
  g.threads = vecs * g.k;
  const int target = sms / nb;                            // CTAs per image
  g.ppc = (hw + target - 1) / target;
  g.chunks = (hw + g.ppc - 1) / g.ppc;
  *out = g;
  return g.threads >= 64 && g.threads % 16 == 0 && g.chunks * nb <= sms;
}

static int g_gn_fused = -1;
bool gn_fused_enabled() {
  if (g_gn_fused < 0) {
    const char* e = getenv("VN_GN_FUSED");
    g_gn_fused = e ? (atoi(e) != 0) : 1;
  }
  return g_gn_fused != 0;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers as NV 16-byte vectors per lane (C <= 256*NV).
// ---------------------------------------------------------------------------------------------
// F32 (the CLIP text encoder's fp32 residual stream, models/clip_encoder.py): x and add are fp32 rows; MODE 0 still
// writes the normalised row in bf16 (it is a GEMM operand), MODE 1 writes dx in fp32 to y32 AND a bf16 copy to y (the
// gradient stream is both the next residual input and the A operand of the next dgrad GEMM).
template <int MODE, int NV, bool F32 = false>
__global__ void __launch_bounds__(256) ln_kernel(const void* __restrict__ x_, long long ldx,
                                                 const bf16* __restrict__ dy, long long lddy,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 float eps, float* __restrict__ stats, const void* __restrict__ add_,
                                                 long long ldadd, bf16* __restrict__ y, long long ldy,
                                                 float* __restrict__ y32, long long ldy32, int rows, int C) {
  const bf16* x = reinterpret_cast<const bf16*>(x_);
  const bf16* add = reinterpret_cast<const bf16*>(add_);
  const float* x32 = reinterpret_cast<const float*>(x_);
  const float* add32 = reinterpret_cast<const float*>(add_);
  pdl_trigger();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int vecs = C >> 3;
  // the affine parameters are frozen: fetched ahead of the grid dependency, so their latency overlaps the predecessor
  float gm[NV][8], bt[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < vecs) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
      gm[i][0] = g0.x; gm[i][1] = g0.y; gm[i][2] = g0.z; gm[i][3] = g0.w;
      gm[i][4] = g1.x; gm[i][5] = g1.y; gm[i][6] = g1.z; gm[i][7] = g1.w;
      if (MODE == 0) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
        bt[i][0] = b0.x; bt[i][1] = b0.y; bt[i][2] = b0.z; bt[i][3] = b0.w;
        bt[i][4] = b1.x; bt[i][5] = b1.y; bt[i][6] = b1.z; bt[i][7] = b1.w;
      }
    }
  }
  pdl_wait();
  if (row >= rows) return;
  const bf16* xr = x + (long long)row * ldx;
  float xf[NV][8];
  uint4 raw[NV], draw[NV], araw[NV];
  float af[F32 ? NV : 1][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < vecs) {
      if (F32) {
        const float4 a0 = *reinterpret_cast<const float4*>(x32 + (long long)row * ldx + v * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(x32 + (long long)row * ldx + v * 8 + 4);
        xf[i][0] = a0.x; xf[i][1] = a0.y; xf[i][2] = a0.z; xf[i][3] = a0.w;
        xf[i][4] = a1.x; xf[i][5] = a1.y; xf[i][6] = a1.z; xf[i][7] = a1.w;
      } else {
        raw[i] = ld16(xr + v * 8);
      }
      if (MODE == 1) {
        draw[i] = ld16(dy + (long long)row * lddy + v * 8);
        if (add_) {
          if (F32) {
            const float4 a0 = *reinterpret_cast<const float4*>(add32 + (long long)row * ldadd + v * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(add32 + (long long)row * ldadd + v * 8 + 4);
            af[F32 ? i : 0][0] = a0.x; af[F32 ? i : 0][1] = a0.y; af[F32 ? i : 0][2] = a0.z; af[F32 ? i : 0][3] = a0.w;
            af[F32 ? i : 0][4] = a1.x; af[F32 ? i : 0][5] = a1.y; af[F32 ? i : 0][6] = a1.z; af[F32 ? i : 0][7] = a1.w;
          } else {
            araw[i] = ld16(add + (long long)row * ldadd + v * 8);
          }
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + i * 32 < vecs) {
      if (!F32) unpack8(raw[i], xf[i]);
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
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (xf[i][j] - mean) * rstd * gm[i][j] + bt[i][j];
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
        float df[8];
        unpack8(draw[i], df);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dh[i][j] = df[j] * gm[i][j];
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
        if (add_) {
          if (F32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += af[F32 ? i : 0][j];
          } else {
            float t[8];
            unpack8(araw[i], t);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += t[j];
          }
        }
        if (F32) {
          float* d32 = y32 + (long long)row * ldy32 + v * 8;
          *reinterpret_cast<float4*>(d32) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(d32 + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
        if (y) *reinterpret_cast<uint4*>(yr + v * 8) = pack8(o);
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

// ---------------------------------------------------------------------------------------------
// GroupNorm on thread-block CLUSTERS (round 2).  The one-launch kernel above slices the tensor by PIXELS, so every CTA
// needs the statistics of all 32 groups and all <= 148 CTAs exchange partials through global memory (each one reads all
// the others' slots: 5.6 MB of L2 traffic for a 2.6 MB tensor, and ~5 us of arrival / totals phases).  Here the tensor
// is sliced by GROUPS first: a cluster of CS <= 8 CTAs owns G2 adjacent groups of one image and splits the pixels, so
// statistics only travel between the CTAs of a cluster - through distributed shared memory, two cluster barriers, no
// global memory, no flags to preset, fixed summation order (bit-reproducible).  The slab of a CTA ((hw / CS) pixels x
// G2 * cpg channels) stays in registers between the two phases: thread t owns 4-byte words t, t + T, ... of the slab
// (word = 2 channels; consecutive threads read consecutive words of a pixel's G2 * cpg * 2-byte run).
// Clusters are independent of each other, so the grid may exceed the SM count (any batch size).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gn_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void gn_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void gn_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ double gn_ld_remote_f64(const double* local, uint32_t rank) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(local)), "r"(rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}

// MODE 0: y = GN(x) (+SiLU), stats_out = (sum x, sum x^2).   MODE 1: dx = GN^T(dy) (+add1) (+add2), stats_out = (sum dh, sum dh*xhat)
template <int MODE, int NI, int T>
__global__ void __launch_bounds__(T, 1) gn_cluster_kernel(const bf16* __restrict__ x, long long ldx,
                                                          const bf16* __restrict__ dy, long long lddy,
                                                          const double* __restrict__ stats_in,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float eps, int silu, double* __restrict__ stats_out,
                                                          const bf16* __restrict__ add1, long long ld1,
                                                          const bf16* __restrict__ add2, long long ld2,
                                                          bf16* __restrict__ y, long long ldy, int hw, int C, int groups,
                                                          int G2, int CS) {
  pdl_trigger();
  __shared__ float s_w[T / 32][4];                 // per-warp partial sums: [statistic 0 | 1] x [group 0 | 1]
  __shared__ double s_part[4];                     // this CTA's partials, read by the whole cluster
  __shared__ float s_c[2][4];                      // per group of the cluster: (mean, rstd, S1, S2); S1 / S2 in the backward only
  const uint32_t rank = gn_cluster_rank();
  const int cl = blockIdx.x / CS;                  // cluster index = image * (groups / G2) + group set
  const int nsets = groups / G2;
  const int b = cl / nsets, g0 = (cl - b * nsets) * G2;
  const int cpg = C / groups;
  const int Wd = (G2 * cpg) >> 1;                  // 4-byte words per pixel in this cluster's channel run
  const int wpg = cpg >> 1;                        // words per group
  const int ppc = (hw + CS - 1) / CS;
  const int p0 = (int)rank * ppc, p1 = min(hw, p0 + ppc);
  const int nitems = max(0, p1 - p0) * Wd;
  const int c0 = g0 * cpg;
  const double inv_n = 1.0 / ((double)cpg * (double)hw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // first item of this thread and the (pixel, word) step between its consecutive items
  const int dpx = T / Wd, dw = T - dpx * Wd;
  int px = t / Wd, w = t - px * Wd;
  pdl_wait();                                      // activations / statistics only from here on
  if (MODE == 1) {
    if (t < G2) {
      const double m = stats_in[((long long)b * groups + g0 + t) * 2] * inv_n;
      const double var = fmax(fma(-m, m, stats_in[((long long)b * groups + g0 + t) * 2 + 1] * inv_n), 0.0);
      s_c[t][0] = (float)m; s_c[t][1] = 1.0f / sqrtf((float)var + eps);
    }
    __syncthreads();
  }
  const bf16* xb = x + ((long long)b * hw + p0) * ldx + c0;
  const bf16* db = MODE == 1 ? dy + ((long long)b * hw + p0) * lddy + c0 : xb;
  uint32_t xr[NI], dr[MODE == 1 ? NI : 1];
  // ---- phase 1: loads (all issued before the first use) and partial sums ----
  {
    int ppx = px, ww = w;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const bool ok = t + i * T < nitems;
      xr[i] = 0u;
      if (MODE == 1) dr[i] = 0u;
      if (ok) {
        xr[i] = __ldg(reinterpret_cast<const uint32_t*>(xb + (long long)ppx * ldx) + ww);
        if (MODE == 1) dr[i] = __ldg(reinterpret_cast<const uint32_t*>(db + (long long)ppx * lddy) + ww);
      }
      ppx += dpx; ww += dw;
      if (ww >= Wd) { ww -= Wd; ++ppx; }
    }
  }
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;        // [statistic][group]
  {
    int ww = w;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      if (t + i * T < nitems) {
        const float2 xf = unpack_bf162(xr[i]);
        const int gi = ww >= wpg ? 1 : 0;
        float u0, u1;
        if (MODE == 0) {
          u0 = xf.x + xf.y;
          u1 = fmaf(xf.x, xf.x, xf.y * xf.y);
        } else {
          const float2 gm = __ldg(reinterpret_cast<const float2*>(gamma + c0) + ww);
          const float2 bt = __ldg(reinterpret_cast<const float2*>(beta + c0) + ww);
          const float2 df = unpack_bf162(dr[i]);
          const float mean = s_c[gi][0], rstd = s_c[gi][1];
          const float xh0 = (xf.x - mean) * rstd, xh1 = (xf.y - mean) * rstd;
          float dz0 = df.x, dz1 = df.y;
          if (silu) { dz0 *= dsilu_f(fmaf(xh0, gm.x, bt.x)); dz1 *= dsilu_f(fmaf(xh1, gm.y, bt.y)); }
          const float dh0 = dz0 * gm.x, dh1 = dz1 * gm.y;
          u0 = dh0 + dh1;
          u1 = fmaf(dh0, xh0, dh1 * xh1);
        }
        if (gi) { a01 += u0; a11 += u1; } else { a00 += u0; a10 += u1; }
      }
      ww += dw;
      if (ww >= Wd) ww -= Wd;
    }
  }
  a00 = warp_sum(a00); a01 = warp_sum(a01); a10 = warp_sum(a10); a11 = warp_sum(a11);
  if (lane == 0) { s_w[warp][0] = a00; s_w[warp][1] = a01; s_w[warp][2] = a10; s_w[warp][3] = a11; }
  __syncthreads();
  if (t < 4) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < T / 32; ++k) acc += (double)s_w[k][t];           // fixed order
    s_part[t] = acc;
  }
  gn_cluster_arrive();                             // release: s_part is visible to the cluster
  gn_cluster_wait();
  if (t < 2 * G2) {
    const int st = t / G2, gi = t - st * G2;       // statistic, group
    double tot = 0.0;
    for (int r = 0; r < CS; ++r) tot += gn_ld_remote_f64(&s_part[st * 2 + gi], (uint32_t)r);   // fixed order
    if (rank == 0) stats_out[((long long)b * groups + g0 + gi) * 2 + st] = tot;
    if (MODE == 0) {
      // mean needs statistic 0, the variance both: park the totals (s_w is dead by now) and finish below
      reinterpret_cast<double*>(s_w)[st * 2 + gi] = tot;
    } else {
      s_c[gi][2 + st] = (float)(tot * inv_n);      // S1 = sum dh / n, S2 = sum dh*xhat / n
    }
  }
  gn_cluster_arrive();                             // my remote reads are done: peers may exit / reuse s_part
  __syncthreads();
  if (MODE == 0 && t < G2) {
    const double* tot = reinterpret_cast<const double*>(s_w);
    const double m = tot[0 * 2 + t] * inv_n;
    const double var = fmax(fma(-m, m, tot[1 * 2 + t] * inv_n), 0.0);
    s_c[t][0] = (float)m; s_c[t][1] = 1.0f / sqrtf((float)var + eps);
  }
  __syncthreads();
  // ---- phase 2: normalise from registers ----
  {
    int ppx = px, ww = w;
    bf16* yb = y + ((long long)b * hw + p0) * ldy + c0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      if (t + i * T < nitems) {
        const float2 xf = unpack_bf162(xr[i]);
        const float2 gm = __ldg(reinterpret_cast<const float2*>(gamma + c0) + ww);
        const float2 bt = __ldg(reinterpret_cast<const float2*>(beta + c0) + ww);
        const int gi = ww >= wpg ? 1 : 0;
        const float mean = s_c[gi][0], rstd = s_c[gi][1];
        float o0, o1;
        if (MODE == 0) {
          const float z0 = fmaf((xf.x - mean) * rstd, gm.x, bt.x), z1 = fmaf((xf.y - mean) * rstd, gm.y, bt.y);
          o0 = silu ? silu_f(z0) : z0;
          o1 = silu ? silu_f(z1) : z1;
        } else {
          const float2 df = unpack_bf162(dr[i]);
          const float S1 = s_c[gi][2], S2 = s_c[gi][3];
          const float xh0 = (xf.x - mean) * rstd, xh1 = (xf.y - mean) * rstd;
          float dz0 = df.x, dz1 = df.y;
          if (silu) { dz0 *= dsilu_f(fmaf(xh0, gm.x, bt.x)); dz1 *= dsilu_f(fmaf(xh1, gm.y, bt.y)); }
          o0 = rstd * (dz0 * gm.x - S1 - xh0 * S2);
          o1 = rstd * (dz1 * gm.y - S1 - xh1 * S2);
          if (add1) {
            const float2 a = unpack_bf162(__ldg(reinterpret_cast<const uint32_t*>(add1 + ((long long)b * hw + p0 + ppx) * ld1 + c0) + ww));
            o0 += a.x; o1 += a.y;
          }
          if (add2) {
            const float2 a = unpack_bf162(__ldg(reinterpret_cast<const uint32_t*>(add2 + ((long long)b * hw + p0 + ppx) * ld2 + c0) + ww));
            o0 += a.x; o1 += a.y;
          }
        }
        reinterpret_cast<uint32_t*>(yb + (long long)ppx * ldy)[ww] = pack_bf162(o0, o1);
      }
      ppx += dpx; ww += dw;
      if (ww >= Wd) { ww -= Wd; ++ppx; }
    }
  }
  gn_cluster_wait();                               // no CTA leaves while a peer may still read its s_part
}

// 16-byte items: the same cluster scheme with uint4 (8-channel) items, for channel runs that are whole 16-byte vectors
// (G2 * cpg % 8 == 0 and cpg >= 8, i.e. every SD-2.1 width: cpg 10 -> 4 groups per cluster, 20 -> 2, 30 -> 4, 40 -> 1, 60 -> 2,
// 80 -> 1).  A vector spans at most two groups; address arithmetic, parameter loads and the group lookup are paid once per 8
// channels instead of once per 2 (the 4-byte kernel above spends ~190 instructions per item, profiles/r2_gn_cluster_notes.txt).
template <int MODE, int NI, int T>
__global__ void __launch_bounds__(T, 1) gn_cluster16_kernel(const bf16* __restrict__ x, long long ldx,
                                                            const bf16* __restrict__ dy, long long lddy,
                                                            const double* __restrict__ stats_in,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, int silu, double* __restrict__ stats_out,
                                                            const bf16* __restrict__ add1, long long ld1,
                                                            const bf16* __restrict__ add2, long long ld2,
                                                            bf16* __restrict__ y, long long ldy, int hw, int C, int groups,
                                                            int G2, int CS) {
  pdl_trigger();
  __shared__ float s_w[T / 32][8];                 // per-warp partial sums: [statistic][group 0..3]
  __shared__ double s_part[8];                     // this CTA's partials, read by the whole cluster
  __shared__ double s_tot[8];
  __shared__ float s_c[4][4];                      // per group of the cluster: (mean, rstd, S1, S2)
  const uint32_t rank = gn_cluster_rank();
  const int cl = blockIdx.x / CS;
  const int nsets = groups / G2;
  const int b = cl / nsets, g0 = (cl - b * nsets) * G2;
  const int cpg = C / groups;
  const int Wv = (G2 * cpg) >> 3;                  // 16-byte vectors per pixel in this cluster's channel run
  const int ppc = (hw + CS - 1) / CS;
  const int p0 = (int)rank * ppc, p1 = min(hw, p0 + ppc);
  const int nitems = max(0, p1 - p0) * Wv;
  const int c0 = g0 * cpg;
  const double inv_n = 1.0 / ((double)cpg * (double)hw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int dpx = T / Wv, dv = T - dpx * Wv;
  const int px = t / Wv, v0 = t - px * Wv;
  pdl_wait();
  if (MODE == 1) {
    if (t < G2) {
      const double m = stats_in[((long long)b * groups + g0 + t) * 2] * inv_n;
      const double var = fmax(fma(-m, m, stats_in[((long long)b * groups + g0 + t) * 2 + 1] * inv_n), 0.0);
      s_c[t][0] = (float)m; s_c[t][1] = 1.0f / sqrtf((float)var + eps);
    }
    __syncthreads();
  }
  const bf16* xb = x + ((long long)b * hw + p0) * ldx + c0;
  const bf16* db = MODE == 1 ? dy + ((long long)b * hw + p0) * lddy + c0 : xb;
  uint4 xr[NI], dr[MODE == 1 ? NI : 1];
  {
    int ppx = px, vv = v0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      xr[i] = make_uint4(0u, 0u, 0u, 0u);
      if (MODE == 1) dr[i] = make_uint4(0u, 0u, 0u, 0u);
      if (t + i * T < nitems) {
        xr[i] = ld16(xb + (long long)ppx * ldx + vv * 8);
        if (MODE == 1) dr[i] = ld16(db + (long long)ppx * lddy + vv * 8);
      }
      ppx += dpx; vv += dv;
      if (vv >= Wv) { vv -= Wv; ++ppx; }
    }
  }
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;     // statistic 0 / 1 per group
  {
    int vv = v0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      if (t + i * T < nitems) {
        const int ch = vv * 8;
        const int glo = ch / cpg;
        const int split = (glo + 1) * cpg - ch;        // channels of this vector that belong to group glo (>= 1)
        float xf[8];
        unpack8(xr[i], xf);
        float u0l = 0.f, u1l = 0.f, u0h = 0.f, u1h = 0.f;
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < split) { u0l += xf[j]; u1l = fmaf(xf[j], xf[j], u1l); }
            else { u0h += xf[j]; u1h = fmaf(xf[j], xf[j], u1h); }
          }
        } else {
          float df[8], gm[8], bt[8];
          unpack8(dr[i], df);
          *reinterpret_cast<float4*>(gm) = __ldg(reinterpret_cast<const float4*>(gamma + c0 + ch));
          *reinterpret_cast<float4*>(gm + 4) = __ldg(reinterpret_cast<const float4*>(gamma + c0 + ch + 4));
          *reinterpret_cast<float4*>(bt) = __ldg(reinterpret_cast<const float4*>(beta + c0 + ch));
          *reinterpret_cast<float4*>(bt + 4) = __ldg(reinterpret_cast<const float4*>(beta + c0 + ch + 4));
          const int ghi = min(glo + 1, 3);
          const float ml = s_c[glo][0], rl = s_c[glo][1], mh = s_c[ghi][0], rh = s_c[ghi][1];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool lo = j < split;
            const float xh = (xf[j] - (lo ? ml : mh)) * (lo ? rl : rh);
            float dz = df[j];
            if (silu) dz *= dsilu_f(fmaf(xh, gm[j], bt[j]));
            const float dh = dz * gm[j];
            if (lo) { u0l += dh; u1l = fmaf(dh, xh, u1l); } else { u0h += dh; u1h = fmaf(dh, xh, u1h); }
          }
        }
        if (glo == 0) { a0 += u0l; q0 += u1l; a1 += u0h; q1 += u1h; }
        else if (glo == 1) { a1 += u0l; q1 += u1l; a2 += u0h; q2 += u1h; }
        else if (glo == 2) { a2 += u0l; q2 += u1l; a3 += u0h; q3 += u1h; }
        else { a3 += u0l; q3 += u1l; }
      }
      vv += dv;
      if (vv >= Wv) vv -= Wv;
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  q0 = warp_sum(q0); q1 = warp_sum(q1); q2 = warp_sum(q2); q3 = warp_sum(q3);
  if (lane == 0) {
    s_w[warp][0] = a0; s_w[warp][1] = a1; s_w[warp][2] = a2; s_w[warp][3] = a3;
    s_w[warp][4] = q0; s_w[warp][5] = q1; s_w[warp][6] = q2; s_w[warp][7] = q3;
  }
  __syncthreads();
  if (t < 8) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < T / 32; ++k) acc += (double)s_w[k][t];           // fixed order
    s_part[t] = acc;
  }
  gn_cluster_arrive();
  gn_cluster_wait();
  if (t < 8) {
    const int st = t >> 2, gi = t & 3;
    double tot = 0.0;
    if (gi < G2) {
      for (int r = 0; r < CS; ++r) tot += gn_ld_remote_f64(&s_part[t], (uint32_t)r);   // fixed order
      if (rank == 0) stats_out[((long long)b * groups + g0 + gi) * 2 + st] = tot;
    }
    s_tot[t] = tot;
  }
  gn_cluster_arrive();                             // my remote reads are done
  __syncthreads();
  if (t < G2) {
    if (MODE == 0) {
      const double m = s_tot[t] * inv_n;
      const double var = fmax(fma(-m, m, s_tot[4 + t] * inv_n), 0.0);
      s_c[t][0] = (float)m; s_c[t][1] = 1.0f / sqrtf((float)var + eps);
    } else {
      s_c[t][2] = (float)(s_tot[t] * inv_n);       // S1 = sum dh / n
      s_c[t][3] = (float)(s_tot[4 + t] * inv_n);   // S2 = sum dh*xhat / n
    }
  }
  __syncthreads();
  {
    int ppx = px, vv = v0;
    bf16* yb = y + ((long long)b * hw + p0) * ldy + c0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      if (t + i * T < nitems) {
        const int ch = vv * 8;
        const int glo = ch / cpg;
        const int split = (glo + 1) * cpg - ch;
        const int ghi = min(glo + 1, 3);
        float xf[8], gm[8], bt[8], o[8];
        unpack8(xr[i], xf);
        *reinterpret_cast<float4*>(gm) = __ldg(reinterpret_cast<const float4*>(gamma + c0 + ch));
        *reinterpret_cast<float4*>(gm + 4) = __ldg(reinterpret_cast<const float4*>(gamma + c0 + ch + 4));
        *reinterpret_cast<float4*>(bt) = __ldg(reinterpret_cast<const float4*>(beta + c0 + ch));
        *reinterpret_cast<float4*>(bt + 4) = __ldg(reinterpret_cast<const float4*>(beta + c0 + ch + 4));
        const float ml = s_c[glo][0], rl = s_c[glo][1], mh = s_c[ghi][0], rh = s_c[ghi][1];
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool lo = j < split;
            const float z = fmaf((xf[j] - (lo ? ml : mh)) * (lo ? rl : rh), gm[j], bt[j]);
            o[j] = silu ? silu_f(z) : z;
          }
        } else {
          float df[8];
          unpack8(dr[i], df);
          const float S1l = s_c[glo][2], S2l = s_c[glo][3], S1h = s_c[ghi][2], S2h = s_c[ghi][3];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool lo = j < split;
            const float rs = lo ? rl : rh;
            const float xh = (xf[j] - (lo ? ml : mh)) * rs;
            float dz = df[j];
            if (silu) dz *= dsilu_f(fmaf(xh, gm[j], bt[j]));
            o[j] = rs * (dz * gm[j] - (lo ? S1l : S1h) - xh * (lo ? S2l : S2h));
          }
          if (add1) {
            float a[8];
            unpack8(ld16(add1 + ((long long)b * hw + p0 + ppx) * ld1 + c0 + ch), a);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += a[j];
          }
          if (add2) {
            float a[8];
            unpack8(ld16(add2 + ((long long)b * hw + p0 + ppx) * ld2 + c0 + ch), a);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += a[j];
          }
        }
        *reinterpret_cast<uint4*>(yb + (long long)ppx * ldy + ch) = pack8(o);
      }
      ppx += dpx; vv += dv;
      if (vv >= Wv) { vv -= Wv; ++ppx; }
    }
  }
  gn_cluster_wait();                               // no CTA leaves while a peer may still read its s_part
}

struct GNCluster {
  int G2, CS, T, NI;
  int vec16;                 // 1: gn_cluster16_kernel (uint4 items), 0: gn_cluster_kernel (4-byte items)
};
static int g_gn_cluster = -1;
// Shapes the cluster kernel covers: even channels per group, at most two groups per cluster, the CTA's slab in <= 32 words
// per thread.  Everything else (the VAE's 512 x 512 tensors) takes the pixel-sliced kernels.
bool gn_cluster_geom(int nb, int hw, int C, int groups, GNCluster* g) {
  if (g_gn_cluster < 0) {
    const char* e = getenv("VN_GN_CLUSTER");
    g_gn_cluster = e ? atoi(e) : 1;
  }
  if (!g_gn_cluster) return false;
  const int cpg = C / groups;
  if (cpg < 2 || (cpg & 1)) return false;
  g->CS = hw >= 512 ? 8 : hw >= 128 ? 4 : hw >= 32 ? 2 : 1;
  g->vec16 = 0;
  if (cpg >= 8 && g_gn_cluster != 4) {            // VN_GN_CLUSTER=4: 4-byte items only (A/B switch)
    // 16-byte items: the smallest G2 in {1, 2, 4} that makes the cluster's channel run whole vectors
    for (int g2 = 1; g2 <= 4; g2 *= 2) {
      if ((g2 * cpg) % 8 != 0 || groups % g2 != 0) continue;
      const int Wv = g2 * cpg / 8;
      const int items = vn_cdiv(hw, g->CS) * Wv;
      static const int lim16 = getenv("VN_GN_CLUSTER_ITEMS") ? atoi(getenv("VN_GN_CLUSTER_ITEMS")) : 512;
      if (items > (g_gn_cluster >= 2 ? 512 * 8 : lim16)) break;      // same 2048-word slab limit as below (VN_GN_CLUSTER=2 lifts it)
      g->G2 = g2;
      g->T = items > 1024 ? 512 : 256;
      if (Wv > g->T) break;
      const int per = vn_cdiv(items, g->T);
      g->NI = per <= 4 ? 4 : (per <= 5 && g->T == 512) ? 5 : 8;
      g->vec16 = 1;
      return true;
    }
  }
  // two groups per cluster when the slab still fits (longer contiguous runs per pixel, fewer clusters); a CTA keeps at most
  // 32 words per thread at 256 threads, 16 at 512 (register budget of the backward: x and dy words both stay live)
  for (int g2 = (cpg <= 20 && groups % 2 == 0) ? 2 : 1; g2 >= 1; --g2) {
    const int Wd = g2 * cpg / 2;
    const int items = vn_cdiv(hw, g->CS) * Wd;
    // Measured on B200 inside the step graph (profiles/r2_gn_cluster_notes.txt): slabs of <= 2048 words per CTA run in
    // 5.4 / 7.1 us (forward / backward) against 7.3 / 8.9 us of the pixel-sliced kernel; larger slabs lose (4-byte items cost
    // ~190 instructions each, 15 k-instruction unrolled bodies miss the instruction cache, 8 x 512-thread clusters start late),
    // so they stay on the pixel-sliced kernel.  VN_GN_CLUSTER=2 lifts the limit (experiments).
    if (items > (g_gn_cluster >= 2 ? 8192 : 2048)) continue;
    g->G2 = g2;
    g->T = items > 2048 ? 512 : 256;
    if (Wd > g->T) return false;
    const int ni = vn_cdiv(items, g->T);
    g->NI = ni <= 8 ? 8 : ni <= 16 ? 16 : 32;
    (void)nb;
    return true;
  }
  return false;
}

template <int MODE, int NI, int T, bool VEC16 = false>
int gn_cluster_launch(const GNCluster& g, const bf16* x, long long ldx, const bf16* dy, long long lddy, const double* stats_in,
                      const float* gamma, const float* beta, float eps, int silu, double* stats_out, const bf16* add1,
                      long long ld1, const bf16* add2, long long ld2, bf16* y, long long ldy, int nb, int hw, int C,
                      int groups, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(nb * (groups / g.G2) * g.CS), 1, 1);
  cfg.blockDim = dim3(T, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (vn_pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = (unsigned)g.CS; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = attr; cfg.numAttrs = na;
  if constexpr (VEC16) {
    VN_CUDA(cudaLaunchKernelEx(&cfg, gn_cluster16_kernel<MODE, NI, T>, x, ldx, dy, lddy, stats_in, gamma, beta, eps, silu,
                               stats_out, add1, ld1, add2, ld2, y, ldy, hw, C, groups, g.G2, g.CS));
  } else {
    VN_CUDA(cudaLaunchKernelEx(&cfg, gn_cluster_kernel<MODE, NI, T>, x, ldx, dy, lddy, stats_in, gamma, beta, eps, silu, stats_out,
                               add1, ld1, add2, ld2, y, ldy, hw, C, groups, g.G2, g.CS));
  }
  vn_count_launch();
  return 0;
}

template <int MODE>
int gn_cluster_dispatch(const GNCluster& g, const bf16* x, long long ldx, const bf16* dy, long long lddy, const double* stats_in,
                        const float* gamma, const float* beta, float eps, int silu, double* stats_out, const bf16* add1,
                        long long ld1, const bf16* add2, long long ld2, bf16* y, long long ldy, int nb, int hw, int C,
                        int groups, cudaStream_t st) {
#define VN_GNC(NI_, T_)                                                                                                  \
  return gn_cluster_launch<MODE, NI_, T_>(g, x, ldx, dy, lddy, stats_in, gamma, beta, eps, silu, stats_out, add1, ld1, add2, \
                                          ld2, y, ldy, nb, hw, C, groups, st)
#define VN_GNC16(NI_, T_)                                                                                                \
  return gn_cluster_launch<MODE, NI_, T_, true>(g, x, ldx, dy, lddy, stats_in, gamma, beta, eps, silu, stats_out, add1, ld1, \
                                                add2, ld2, y, ldy, nb, hw, C, groups, st)
  if (g.vec16) {
    if (g.T == 256) { if (g.NI == 4) VN_GNC16(4, 256); VN_GNC16(8, 256); }
    if (g.NI == 4) VN_GNC16(4, 512);
    if (g.NI == 5) VN_GNC16(5, 512);
    VN_GNC16(8, 512);
  }
  if (g.T == 256) {
    if (g.NI == 8) VN_GNC(8, 256);
    if (g.NI == 16) VN_GNC(16, 256);
    VN_GNC(32, 256);
  }
  if (g.NI == 8) VN_GNC(8, 512);
  VN_GNC(16, 512);
#undef VN_GNC
#undef VN_GNC16
}

int gn_check(int C, int groups, long long ldx) {
  VN_CHECK(groups > 0 && groups <= kMaxGroups && C % groups == 0, "groupnorm: C=%d groups=%d", C, groups);
  VN_CHECK(C % 8 == 0 && C <= kMaxC * 4, "groupnorm: need C %% 8 == 0 (C=%d)", C);
  VN_CHECK(C / 8 <= 512, "groupnorm: C=%d too wide (C <= 4096)", C);
  VN_CHECK(ldx % 8 == 0, "groupnorm: row stride must be a multiple of 8");
  return 0;
}

template <int MODE, bool F32 = false>
int ln_launch(const void* x, long long ldx, const bf16* dy, long long lddy, const float* gamma, const float* beta,
              float eps, float* stats, const void* add, long long ldadd, bf16* y, long long ldy, int rows, int C,
              cudaStream_t s, float* y32 = nullptr, long long ldy32 = 0) {
  VN_CHECK(C % 8 == 0 && C <= 2048, "layernorm: C=%d unsupported (C %% 8 == 0, C <= 2048)", C);
  VN_CHECK(ldx % (F32 ? 4 : 8) == 0 && ldy % 8 == 0 && lddy % 8 == 0 && ldadd % (F32 ? 4 : 8) == 0 && ldy32 % 4 == 0,
           "layernorm: strides must be multiples of 8 (bf16) / 4 (fp32)");
  const int nv = (C / 8 + 31) / 32;
  const int grid = vn_cdiv(rows, 8);
#define VN_LN_CASE(NV)                                                                                              \
  case NV:                                                                                                          \
    VN_LAUNCH((ln_kernel<MODE, NV, F32>), grid, 256, 0, s, x, ldx, dy, lddy, gamma, beta, eps, stats, add, ldadd, y, ldy, y32, ldy32, rows, C); \
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

extern "C" void vn_set_groupnorm_fused(int enabled) { g_gn_fused = enabled ? 1 : 0; }

// floats of `partials` one GroupNorm call needs for nb images: 64 per CTA, at most one CTA per SM
extern "C" size_t vn_groupnorm_partial_floats(int nb) {
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  (void)nb;                                       // grid <= #SMs in total, whatever nb is
  return (size_t)sms * 64;
}

extern "C" int vn_groupnorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, int silu,
                                void* y, int64_t ldy, int nb, int hw, int C, int groups, double* stats,
                                float* partials, vn_stream_t s) {
  if (gn_check(C, groups, ldx)) return -1;
  VN_CHECK(ldy % 8 == 0, "groupnorm: ldy must be a multiple of 8");
  GNCluster gc;
  if (gn_fused_enabled() && gn_cluster_geom(nb, hw, C, groups, &gc))
    return gn_cluster_dispatch<0>(gc, (const bf16*)x, ldx, nullptr, 0, nullptr, gamma, beta, eps, silu, stats, nullptr, 0,
                                  nullptr, 0, (bf16*)y, ldy, nb, hw, C, groups, (cudaStream_t)s);
  GNGeom g;
  if (partials && groups <= 32 && gn_fused_enabled() && gn_fused_geom(nb, hw, C, 512, &g)) {
    dim3 grid(g.chunks, nb);
    const int npt = vn_cdiv(g.ppc, g.k);
#define VN_GN_F(NP)                                                                                                   \
  VN_LAUNCH((gn_fused_kernel<0, NP, 512>), grid, g.threads, 0, (cudaStream_t)s, (const bf16*)x, ldx, nullptr, 0,      \
            nullptr, gamma, beta, eps, silu, stats, partials, nullptr, 0, nullptr, 0, (bf16*)y, ldy, hw, C, groups,   \
            g.k, g.ppc, vn_debug_buffer())
    if (npt <= 4) { VN_GN_F(4); }
    else if (npt <= 8) { VN_GN_F(8); }
    else { VN_GN_F(0); }
#undef VN_GN_F
    return 0;
  }
  if (vn_groupnorm_stats(x, ldx, nb, hw, C, groups, stats, s)) return -1;
  return vn_groupnorm_apply(x, ldx, stats, gamma, beta, eps, silu, y, ldy, nb, hw, C, groups, s);
}

extern "C" int vn_groupnorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, const double* stats,
                                const float* gamma, const float* beta, float eps, int silu, const void* add1,
                                int64_t ldadd1, const void* add2, int64_t ldadd2, void* dx, int64_t lddx, int nb, int hw,
                                int C, int groups, double* red, float* partials, vn_stream_t s) {
  if (gn_check(C, groups, ldx)) return -1;
  VN_CHECK(lddy % 8 == 0 && lddx % 8 == 0 && ldadd1 % 8 == 0 && ldadd2 % 8 == 0, "groupnorm bwd: strides must be multiples of 8");
  GNCluster gc;
  if (gn_fused_enabled() && gn_cluster_geom(nb, hw, C, groups, &gc))
    return gn_cluster_dispatch<1>(gc, (const bf16*)x, ldx, (const bf16*)dy, lddy, stats, gamma, beta, eps, silu, red,
                                  (const bf16*)add1, ldadd1, (const bf16*)add2, ldadd2, (bf16*)dx, lddx, nb, hw, C, groups,
                                  (cudaStream_t)s);
  GNGeom g;
  if (partials && groups <= 32 && gn_fused_enabled() && gn_fused_geom(nb, hw, C, 384, &g)) {
    dim3 grid(g.chunks, nb);
    const int npt = vn_cdiv(g.ppc, g.k);
#define VN_GN_B(NP)                                                                                                    \
  VN_LAUNCH((gn_fused_kernel<1, NP, 384>), grid, g.threads, 0, (cudaStream_t)s, (const bf16*)x, ldx, (const bf16*)dy,  \
            lddy, stats, gamma, beta, eps, silu, red, partials, (const bf16*)add1, ldadd1, (const bf16*)add2, ldadd2,  \
            (bf16*)dx, lddx, hw, C, groups, g.k, g.ppc, vn_debug_buffer())
    if (npt <= 4) { VN_GN_B(4); }
    else { VN_GN_B(0); }
#undef VN_GN_B
    return 0;
  }
  if (vn_groupnorm_bwd_stats(x, ldx, dy, lddy, stats, gamma, beta, eps, silu, red, nb, hw, C, groups, s)) return -1;
  return vn_groupnorm_bwd_apply(x, ldx, dy, lddy, stats, red, gamma, beta, eps, silu, add1, ldadd1, add2, ldadd2, dx,
                                lddx, nb, hw, C, groups, s);
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

extern "C" int vn_layernorm_fwd_f32(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y,
                                    int64_t ldy, float* stats, int rows, int C, vn_stream_t s) {
  VN_CHECK(ldx % 4 == 0, "layernorm f32: row stride must be a multiple of 4");
  return ln_launch<0, true>(x, ldx, nullptr, 0, gamma, beta, eps, stats, nullptr, 0, (bf16*)y, ldy, rows, C, (cudaStream_t)s);
}

extern "C" int vn_layernorm_bwd_f32(const float* x, int64_t ldx, const void* dy, int64_t lddy, const float* gamma,
                                    const float* stats, const float* add, int64_t ldadd, float* dx, int64_t lddx,
                                    void* dx_bf16, int64_t lddxb, int rows, int C, vn_stream_t s) {
  VN_CHECK(ldx % 4 == 0 && ldadd % 4 == 0 && lddx % 4 == 0 && dx != nullptr, "layernorm f32: strides must be multiples of 4");
  return ln_launch<1, true>(x, ldx, (const bf16*)dy, lddy, gamma, nullptr, 0.f, const_cast<float*>(stats), add, ldadd,
                            (bf16*)dx_bf16, lddxb, rows, C, (cudaStream_t)s, dx, lddx);
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
