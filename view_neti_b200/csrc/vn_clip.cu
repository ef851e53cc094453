// vn_clip.cu — the two ops the CLIP text transformer needs beyond the UNet kernels (the batched
// 16-layer conditioning path, reference models/neti_clip_text_encoder.py:57-225 -> transformers CLIPEncoder):
//   * erf-GELU forward / backward (CLIPMLP, hidden_act = "gelu" for the SD-2.1 OpenCLIP-H text model)
//   * causal self-attention over SHORT sequences (77 tokens, head_dim 64), forward and backward.
// The attention is ~0.4 GFLOP per layer for the 16 x B sequences of a step - three orders of magnitude below the
// projections around it (which run on vn_gemm) - and its 77 x 77 triangle does not map onto 128-wide tensor-core tiles,
// so it is a CUDA-core kernel: one CTA per (sequence, head), K / V (and Q / dO in the backward) staged in shared memory
// as fp32, one thread per query row streaming over its causal keys with an online softmax (keys are broadcast reads,
// conflict-free), and in the backward a second phase with one thread per KEY row, so dK / dV are plain per-thread sums:
// no atomics, deterministic.
#include "vn_common.cuh"

namespace {

constexpr int HD = 64;            // head_dim
constexpr int kMaxL = 128;        // longest sequence one CTA handles

__device__ __forceinline__ float gelu_erf(float g) { return 0.5f * g * (1.f + erff(g * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_erf(float g) {
  const float cdf = 0.5f * (1.f + erff(g * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * g * g);
  return cdf + g * pdf;
}
__device__ __forceinline__ void unpack8f(const uint4& v, float* f) {
  float2 t;
  t = unpack_bf162(v.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf162(v.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf162(v.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf162(v.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8f(const float* a) {
  uint4 o;
  o.x = pack_bf162(a[0], a[1]); o.y = pack_bf162(a[2], a[3]);
  o.z = pack_bf162(a[4], a[5]); o.w = pack_bf162(a[6], a[7]);
  return o;
}

// MODE 0: y = gelu(h).  MODE 1: dh = dy * gelu'(h).
template <int MODE>
__global__ void __launch_bounds__(256) gelu_kernel(const bf16* __restrict__ h, long long ldh, const bf16* __restrict__ dy,
                                                   long long lddy, bf16* __restrict__ out, long long ldo,
                                                   long long total_vecs, int F) {
  pdl_trigger();
  pdl_wait();
  const int vecs = F >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vecs;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vecs;
    const int c = (int)(i - row * vecs) * 8;
    float x[8], o[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(h + row * ldh + c)), x);
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = gelu_erf(x[j]);
    } else {
      float d[8];
      unpack8f(__ldg(reinterpret_cast<const uint4*>(dy + row * lddy + c)), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = d[j] * dgelu_erf(x[j]);
    }
    *reinterpret_cast<uint4*>(out + row * ldo + c) = pack8f(o);
  }
}

struct SeqAttnParams {
  int L, heads, causal;
  float scale;
  const bf16 *q, *k, *v;
  long long ldq, ldk, ldv, bsq, bsk, bsv;        // row / sequence strides in elements
  bf16* o; long long ldo, bso;
  float* lse;                                     // [nseq, heads, L] natural-log sum-exp of the scaled logits
  const bf16* d_o; long long lddo, bsdo;
  bf16 *dq, *dk, *dv; long long lddq, lddk, lddv, bsdq, bsdk, bsdv;
};

// cooperative load of a [L x 64] bf16 head slice into fp32 shared memory (row stride 64: broadcast reads only)
__device__ __forceinline__ void load_head(float* dst, const bf16* src, long long ld, int L) {
  for (int i = threadIdx.x; i < L * (HD / 8); i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    float f[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(src + (long long)r * ld + c)), f);
    float4* d = reinterpret_cast<float4*>(dst + r * HD + c);
    d[0] = make_float4(f[0], f[1], f[2], f[3]);
    d[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

__global__ void __launch_bounds__(128) seq_attn_fwd_kernel(const SeqAttnParams p) {
  pdl_trigger();
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y, L = p.L;
  float* sK = sm;
  float* sV = sm + L * HD;
  pdl_wait();
  load_head(sK, p.k + (long long)b * p.bsk + h * HD, p.ldk, L);
  load_head(sV, p.v + (long long)b * p.bsv + h * HD, p.ldv, L);
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= L) return;
  float q[HD], o[HD];
  {
    const bf16* qr = p.q + (long long)b * p.bsq + (long long)i * p.ldq + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 8) unpack8f(__ldg(reinterpret_cast<const uint4*>(qr + c)), q + c);
#pragma unroll
    for (int c = 0; c < HD; ++c) { q[c] *= p.scale; o[c] = 0.f; }
  }
  const int nk = p.causal ? i + 1 : L;
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j < nk; ++j) {
    const float4* kr = reinterpret_cast<const float4*>(sK + j * HD);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 kv = kr[c];
      s0 = fmaf(q[4 * c], kv.x, s0); s1 = fmaf(q[4 * c + 1], kv.y, s1);
      s2 = fmaf(q[4 * c + 2], kv.z, s2); s3 = fmaf(q[4 * c + 3], kv.w, s3);
    }
    const float s = (s0 + s1) + (s2 + s3);
    const float mn = fmaxf(m, s);
    const float a = __expf(m - mn), pj = __expf(s - mn);     // first key: exp(-inf) = 0
    l = l * a + pj;
    m = mn;
    const float4* vr = reinterpret_cast<const float4*>(sV + j * HD);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 vv = vr[c];
      o[4 * c] = fmaf(o[4 * c], a, pj * vv.x); o[4 * c + 1] = fmaf(o[4 * c + 1], a, pj * vv.y);
      o[4 * c + 2] = fmaf(o[4 * c + 2], a, pj * vv.z); o[4 * c + 3] = fmaf(o[4 * c + 3], a, pj * vv.w);
    }
  }
  const float inv = 1.f / l;
  bf16* orow = p.o + (long long)b * p.bso + (long long)i * p.ldo + h * HD;
#pragma unroll
  for (int c = 0; c < HD; c += 8) {
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = o[c + u] * inv;
    *reinterpret_cast<uint4*>(orow + c) = pack8f(t);
  }
  p.lse[((long long)b * p.heads + h) * L + i] = m + __logf(l);
}

// Backward: three independent jobs per (sequence, head), one kernel instantiation each (so each gets its own register
// budget / occupancy): JOB 0: dQ (thread = query row; K, V in shared memory), JOB 1: dV, JOB 2: dK (thread = key row;
// Q, dO in shared memory).  Every job recomputes the probabilities it needs from q.k and the saved lse; dK / dV are
// plain per-thread sums over the (causal) query rows: no atomics, deterministic.
template <int JOB>
__global__ void __launch_bounds__(128) seq_attn_bwd_kernel(const SeqAttnParams p) {
  pdl_trigger();
  extern __shared__ float sm[];
  constexpr int job = JOB;
  const int h = blockIdx.x, b = blockIdx.y, L = p.L;
  float* sA = sm;                                  // job 0: K      jobs 1, 2: Q
  float* sB = sA + L * HD;                         // job 0: V      jobs 1, 2: dO
  float* sLse = sB + L * HD;                       // [L]
  float* sDelta = sLse + L;                        // [L]  delta_i = dO_i . O_i
  pdl_wait();
  if (job == 0) {
    load_head(sA, p.k + (long long)b * p.bsk + h * HD, p.ldk, L);
    load_head(sB, p.v + (long long)b * p.bsv + h * HD, p.ldv, L);
  } else {
    load_head(sA, p.q + (long long)b * p.bsq + h * HD, p.ldq, L);
    load_head(sB, p.d_o + (long long)b * p.bsdo + h * HD, p.lddo, L);
  }
  const int t = threadIdx.x;
  if (t < L) {
    sLse[t] = p.lse[((long long)b * p.heads + h) * L + t];
    if (job != 1) {
      const bf16* orow = p.o + (long long)b * p.bso + (long long)t * p.ldo + h * HD;
      const bf16* drow = p.d_o + (long long)b * p.bsdo + (long long)t * p.lddo + h * HD;
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int c = 0; c < HD; c += 8) {
        float a[8], g[8];
        unpack8f(__ldg(reinterpret_cast<const uint4*>(orow + c)), a);
        unpack8f(__ldg(reinterpret_cast<const uint4*>(drow + c)), g);
#pragma unroll
        for (int u = 0; u < 8; u += 2) { d0 = fmaf(a[u], g[u], d0); d1 = fmaf(a[u + 1], g[u + 1], d1); }
      }
      sDelta[t] = d0 + d1;
    }
  }
  __syncthreads();
  if (t >= L) return;
  if (job == 0) {
    // ---- dq_i = scale * sum_j ds_ij k_j,  ds_ij = p_ij (dO_i . v_j - delta_i) ----
    const int i = t;
    float q[HD], g[HD], dq[HD];
    {
      const bf16* qr = p.q + (long long)b * p.bsq + (long long)i * p.ldq + h * HD;
      const bf16* gr = p.d_o + (long long)b * p.bsdo + (long long)i * p.lddo + h * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 8) {
        unpack8f(__ldg(reinterpret_cast<const uint4*>(qr + c)), q + c);
        unpack8f(__ldg(reinterpret_cast<const uint4*>(gr + c)), g + c);
      }
#pragma unroll
      for (int c = 0; c < HD; ++c) { q[c] *= p.scale; dq[c] = 0.f; }
    }
    const float lse = sLse[i], delta = sDelta[i];
    const int nk = p.causal ? i + 1 : L;
    for (int j = 0; j < nk; ++j) {
      const float4* kr = reinterpret_cast<const float4*>(sA + j * HD);
      const float4* vr = reinterpret_cast<const float4*>(sB + j * HD);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        const float4 kv = kr[c], vv = vr[c];
        s0 = fmaf(q[4 * c], kv.x, s0); s1 = fmaf(q[4 * c + 1], kv.y, s1);
        s2 = fmaf(q[4 * c + 2], kv.z, s2); s3 = fmaf(q[4 * c + 3], kv.w, s3);
        e0 = fmaf(g[4 * c], vv.x, e0); e1 = fmaf(g[4 * c + 1], vv.y, e1);
        e2 = fmaf(g[4 * c + 2], vv.z, e2); e3 = fmaf(g[4 * c + 3], vv.w, e3);
      }
      const float pij = __expf(((s0 + s1) + (s2 + s3)) - lse);
      const float ds = pij * (((e0 + e1) + (e2 + e3)) - delta) * p.scale;
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        const float4 kv = kr[c];
        dq[4 * c] = fmaf(ds, kv.x, dq[4 * c]); dq[4 * c + 1] = fmaf(ds, kv.y, dq[4 * c + 1]);
        dq[4 * c + 2] = fmaf(ds, kv.z, dq[4 * c + 2]); dq[4 * c + 3] = fmaf(ds, kv.w, dq[4 * c + 3]);
      }
    }
    bf16* dr = p.dq + (long long)b * p.bsdq + (long long)i * p.lddq + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 8) *reinterpret_cast<uint4*>(dr + c) = pack8f(dq + c);
    return;
  }
  // ---- thread = key row j: dv_j = sum_i p_ij dO_i (job 1),  dk_j = scale * sum_i ds_ij q_i (job 2);  i >= j if causal ----
  const int j = t;
  const int i0 = p.causal ? j : 0;
  float k[HD];
  {
    const bf16* kr = p.k + (long long)b * p.bsk + (long long)j * p.ldk + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 8) unpack8f(__ldg(reinterpret_cast<const uint4*>(kr + c)), k + c);
#pragma unroll
    for (int c = 0; c < HD; ++c) k[c] *= p.scale;
  }
  if (job == 1) {
    float dv[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) dv[c] = 0.f;
    for (int i = 0; i < L; ++i) {            // same row for every thread of the warp: broadcast reads, no bank conflicts
      if (i < i0) continue;
      const float4* qr = reinterpret_cast<const float4*>(sA + i * HD);
      const float4* gr = reinterpret_cast<const float4*>(sB + i * HD);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        const float4 qv = qr[c];
        s0 = fmaf(k[4 * c], qv.x, s0); s1 = fmaf(k[4 * c + 1], qv.y, s1);
        s2 = fmaf(k[4 * c + 2], qv.z, s2); s3 = fmaf(k[4 * c + 3], qv.w, s3);
      }
      const float pij = __expf(((s0 + s1) + (s2 + s3)) - sLse[i]);
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        const float4 gv = gr[c];
        dv[4 * c] = fmaf(pij, gv.x, dv[4 * c]); dv[4 * c + 1] = fmaf(pij, gv.y, dv[4 * c + 1]);
        dv[4 * c + 2] = fmaf(pij, gv.z, dv[4 * c + 2]); dv[4 * c + 3] = fmaf(pij, gv.w, dv[4 * c + 3]);
      }
    }
    bf16* vr = p.dv + (long long)b * p.bsdv + (long long)j * p.lddv + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 8) *reinterpret_cast<uint4*>(vr + c) = pack8f(dv + c);
  } else {
    float v[HD], dk[HD];
    {
      const bf16* vr = p.v + (long long)b * p.bsv + (long long)j * p.ldv + h * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 8) unpack8f(__ldg(reinterpret_cast<const uint4*>(vr + c)), v + c);
#pragma unroll
      for (int c = 0; c < HD; ++c) dk[c] = 0.f;
    }
    for (int i = 0; i < L; ++i) {            // same row for every thread of the warp: broadcast reads, no bank conflicts
      if (i < i0) continue;
      const float4* qr = reinterpret_cast<const float4*>(sA + i * HD);
      const float4* gr = reinterpret_cast<const float4*>(sB + i * HD);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        const float4 qv = qr[c], gv = gr[c];
        s0 = fmaf(k[4 * c], qv.x, s0); s1 = fmaf(k[4 * c + 1], qv.y, s1);
        s2 = fmaf(k[4 * c + 2], qv.z, s2); s3 = fmaf(k[4 * c + 3], qv.w, s3);
        e0 = fmaf(v[4 * c], gv.x, e0); e1 = fmaf(v[4 * c + 1], gv.y, e1);
        e2 = fmaf(v[4 * c + 2], gv.z, e2); e3 = fmaf(v[4 * c + 3], gv.w, e3);
      }
      const float pij = __expf(((s0 + s1) + (s2 + s3)) - sLse[i]);
      const float ds = pij * (((e0 + e1) + (e2 + e3)) - sDelta[i]) * p.scale;
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        const float4 qv = qr[c];
        dk[4 * c] = fmaf(ds, qv.x, dk[4 * c]); dk[4 * c + 1] = fmaf(ds, qv.y, dk[4 * c + 1]);
        dk[4 * c + 2] = fmaf(ds, qv.z, dk[4 * c + 2]); dk[4 * c + 3] = fmaf(ds, qv.w, dk[4 * c + 3]);
      }
    }
    bf16* kr = p.dk + (long long)b * p.bsdk + (long long)j * p.lddk + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 8) *reinterpret_cast<uint4*>(kr + c) = pack8f(dk + c);
  }
}

int seq_attn_check(const vn_attn_desc* d, bool bwd) {
  VN_CHECK(d != nullptr, "seq attention: null descriptor");
  VN_CHECK(d->nb > 0 && d->heads > 0 && d->nq > 0 && d->nq == d->nk && d->nq <= kMaxL,
           "seq attention: need nq == nk <= %d (got %d, %d)", kMaxL, d->nq, d->nk);
  VN_CHECK(d->ldq % 8 == 0 && d->ldk % 8 == 0 && d->ldv % 8 == 0 && d->ldo % 8 == 0 && d->bsq % 8 == 0 && d->bsk % 8 == 0 &&
               d->bsv % 8 == 0 && d->bso % 8 == 0, "seq attention: strides must be multiples of 8 elements");
  VN_CHECK(d->q && d->k && d->v && d->o && d->lse, "seq attention: q / k / v / o / lse are required");
  if (bwd) {
    VN_CHECK(d->d_o && d->dq && d->dk && d->dv, "seq attention bwd: d_o / dq / dk / dv are required");
    VN_CHECK(d->lddo % 8 == 0 && d->lddq % 8 == 0 && d->lddk % 8 == 0 && d->lddv % 8 == 0 && d->bsdo % 8 == 0 &&
                 d->bsdq % 8 == 0 && d->bsdk % 8 == 0 && d->bsdv % 8 == 0, "seq attention bwd: strides must be multiples of 8");
  }
  return 0;
}

SeqAttnParams seq_params(const vn_attn_desc* d, int causal) {
  SeqAttnParams p{};
  p.L = d->nq; p.heads = d->heads; p.causal = causal; p.scale = d->scale;
  p.q = (const bf16*)d->q; p.k = (const bf16*)d->k; p.v = (const bf16*)d->v;
  p.ldq = d->ldq; p.ldk = d->ldk; p.ldv = d->ldv; p.bsq = d->bsq; p.bsk = d->bsk; p.bsv = d->bsv;
  p.o = (bf16*)d->o; p.ldo = d->ldo; p.bso = d->bso;
  p.lse = d->lse;
  p.d_o = (const bf16*)d->d_o; p.lddo = d->lddo; p.bsdo = d->bsdo;
  p.dq = (bf16*)d->dq; p.dk = (bf16*)d->dk; p.dv = (bf16*)d->dv;
  p.lddq = d->lddq; p.lddk = d->lddk; p.lddv = d->lddv; p.bsdq = d->bsdq; p.bsdk = d->bsdk; p.bsdv = d->bsdv;
  return p;
}

}  // namespace

extern "C" int vn_gelu_fwd(const void* h, int64_t ldh, void* y, int64_t ldy, int rows, int F, vn_stream_t s) {
  VN_CHECK(F % 8 == 0 && ldh % 8 == 0 && ldy % 8 == 0, "gelu: F and strides must be multiples of 8");
  const long long total = (long long)rows * (F >> 3);
  if (total <= 0) return 0;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  VN_LAUNCH(gelu_kernel<0>, blocks, 256, 0, (cudaStream_t)s, (const bf16*)h, ldh, nullptr, 0, (bf16*)y, ldy, total, F);
  return 0;
}

extern "C" int vn_gelu_bwd(const void* h, int64_t ldh, const void* dy, int64_t lddy, void* dh, int64_t lddh, int rows,
                           int F, vn_stream_t s) {
  VN_CHECK(F % 8 == 0 && ldh % 8 == 0 && lddy % 8 == 0 && lddh % 8 == 0, "gelu: F and strides must be multiples of 8");
  const long long total = (long long)rows * (F >> 3);
  if (total <= 0) return 0;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  VN_LAUNCH(gelu_kernel<1>, blocks, 256, 0, (cudaStream_t)s, (const bf16*)h, ldh, (const bf16*)dy, lddy, (bf16*)dh, lddh,
            total, F);
  return 0;
}

extern "C" int vn_seq_attention_fwd(const vn_attn_desc* d, int causal, vn_stream_t s) {
  if (seq_attn_check(d, false)) return -1;
  const SeqAttnParams p = seq_params(d, causal);
  constexpr int smem_max = 2 * kMaxL * HD * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    VN_CUDA(cudaFuncSetAttribute(seq_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    configured = true;
  }
  const int smem = 2 * d->nq * HD * (int)sizeof(float);           // sized by the sequence: more CTAs per SM
  VN_LAUNCH(seq_attn_fwd_kernel, dim3(d->heads, d->nb), 128, smem, (cudaStream_t)s, p);
  return 0;
}

extern "C" int vn_seq_attention_bwd(const vn_attn_desc* d, int causal, vn_stream_t s) {
  if (seq_attn_check(d, true)) return -1;
  const SeqAttnParams p = seq_params(d, causal);
  constexpr int smem_max = (2 * kMaxL * HD + 2 * kMaxL) * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    VN_CUDA(cudaFuncSetAttribute(seq_attn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    VN_CUDA(cudaFuncSetAttribute(seq_attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    VN_CUDA(cudaFuncSetAttribute(seq_attn_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    configured = true;
  }
  const int smem = (2 * d->nq * HD + 2 * d->nq) * (int)sizeof(float);
  VN_LAUNCH(seq_attn_bwd_kernel<2>, dim3(d->heads, d->nb), 128, smem, (cudaStream_t)s, p);     // longest job first
  VN_LAUNCH(seq_attn_bwd_kernel<0>, dim3(d->heads, d->nb), 128, smem, (cudaStream_t)s, p);
  VN_LAUNCH(seq_attn_bwd_kernel<1>, dim3(d->heads, d->nb), 128, smem, (cudaStream_t)s, p);
  return 0;
}
