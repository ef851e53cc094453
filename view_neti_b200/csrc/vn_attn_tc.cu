// vn_attn_tc.cu — attention forward on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), head_dim 64.
//
// Same contract as the reference's attention core (models/xti_attention_processor.py:44-50: head split, fp32 logits
// with alpha = scale, softmax, bmm, head merge), with K and V taken from different tensors (XTI).
//
// CTA = one (128-query tile, head, image), one CTA per SM.
//   warp 0     TMA producer : Q tile once, then (K, V) tiles of 128 keys through a 3-stage ring (3-D tensor maps
//                             {64 d, rows, image}: rows beyond nq / nk are zero-filled, head picked by the column offset)
//   warp 1     MMA issuer   : S(j) = Q K(j)^T  (tcgen05.mma 128x128x16, operands K-major) -> TMEM S[j&1]   (2 x 128 columns)
//                             PV(j) = P(j) V(j) (A = P from shared memory, B = V used in place as an MN-major
//                                                operand: the [keys x d] tile needs no transpose)  -> TMEM PV[j&1] (2 x 64)
//                             issue order S(0) S(1) PV(0) S(2) PV(1) ...: the tensor pipe computes S(j+1) while the
//                             softmax warps work on S(j)
//   warps 2-9  softmax      : TWO threads per query row (= TMEM lane), one per 64-key half of every tile, each running an
//                             independent online softmax (own running max / sum / output) so the halves never exchange
//                             anything inside the loop: 64 logits pulled into registers with tcgen05.ld ONCE, ex2.approx,
//                             P packed to bf16 into the swizzled A-operand tile P[j&1]; PV is issued per half into its own
//                             TMEM columns and O = alpha * O + PV stays in registers (read back from TMEM while the next
//                             MMAs run; no tcgen05.st / correction pass).  The two halves are merged once at the end
//                             (log-sum-exp merge through shared memory).  Two softmax warps per scheduler hide the ALU /
//                             MUFU latencies that a single warp cannot.
#include "vn_tma.cuh"

namespace {

constexpr int D = 64;
constexpr int BQ = 128;           // queries per CTA
constexpr int BKV = 128;          // keys per iteration
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * 128;             // [128 rows x 64 bf16]
constexpr int kThreads = 320;             // producer warp, MMA warp, 8 softmax warps
constexpr int TMEM_COLS = 512;                    // S[2]: 2 x 128 fp32 columns, PV[2][half]: 4 x 64
constexpr int SMEM_BYTES = TILE_BYTES /*Q*/ + KV_STAGES * 2 * TILE_BYTES /*K,V*/ + 2 * 2 * TILE_BYTES /*P[2]: two 64-key blocks each*/ +
                           256 /*barriers*/ + 1024 /*alignment*/;
constexpr float kLog2e = 1.4426950408889634f;

struct FwdParams {
  int nq, nk, heads, causal;
  float scale;
  bf16* o; long long ldo, bso;
  float* lse;
};

// MN-major shared-memory operand descriptor, 128B swizzle: rows are K (128 B apart), 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;             // LBO: stride between 64-element MN blocks (single block here)
  d |= (uint64_t)(1024 >> 4) << 32;   // SBO: stride between 8-row K groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// bf16 x bf16 -> fp32, A K-major, B K-major (b_mn = 0) or MN-major (b_mn = 1)
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads, 1) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV,
                                                                   const FwdParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + TILE_BYTES;                          // stage s: K at s*2*TILE, V right after
  uint8_t* sP = sKV + KV_STAGES * 2 * TILE_BYTES;          // P[b]: two [128 x 64] blocks (keys 0-63, 64-127)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                            // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES;                // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;                 // [2]
  uint64_t* p_full = s_full + 2;                           // [2]
  uint64_t* pv_full = p_full + 2;                          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nt = (p.nk + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 256); mbar_init(&pv_full[s], 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, h * D, q0, b);
      for (int j = 0; j < nt; ++j) {
        const int s = j % KV_STAGES;
        mbar_wait(&kv_empty[s], ((j / KV_STAGES) & 1) ^ 1);
        mbar_expect_tx(&kv_full[s], 2 * TILE_BYTES);
        tma_load_3d(sKV + s * 2 * TILE_BYTES, &tmK, &kv_full[s], h * D, j * BKV, b);
        tma_load_3d(sKV + s * 2 * TILE_BYTES + TILE_BYTES, &tmV, &kv_full[s], h * D, j * BKV, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(BQ, BKV, 0);
      constexpr uint32_t idesc_pv = idesc_bf16(BQ, D, 1);
      const uint32_t aQ = smem_u32(sQ);
      auto issue_s = [&](int j) {
        // S[j&1] is free: the softmax threads arrived on p_full(j-2) before PV(j-2) was issued (program order below)
        const int s = j % KV_STAGES;
        mbar_wait(&kv_full[s], (j / KV_STAGES) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sKV + s * 2 * TILE_BYTES);
        const uint32_t tS = tmem_base + (uint32_t)((j & 1) * BKV);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_bf16(tS, umma_desc_k_sw128(aQ) + (uint64_t)(k * 2), umma_desc_k_sw128(aK) + (uint64_t)(k * 2), idesc_s,
                    k ? 1u : 0u);
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) issue_s(j + 1);
        // PV(j) = P(j) V(j)
        const int s = j % KV_STAGES;
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aP = smem_u32(sP + (j & 1) * 2 * TILE_BYTES);
        const uint32_t aV = smem_u32(sKV + s * 2 * TILE_BYTES + TILE_BYTES);
        const uint32_t tPV = tmem_base + 2 * BKV + (uint32_t)((j & 1) * 2 * D);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // keys [0,64) accumulate into PV[j&1][0], keys [64,128) into PV[j&1][1]
          const uint64_t adesc = umma_desc_k_sw128(aP + (k >> 2) * TILE_BYTES) + (uint64_t)((k & 3) * 2);
          const uint64_t bdesc = umma_desc_mn_sw128(aV + k * 2048);
          umma_bf16(tPV + (uint32_t)((k >> 2) * D), adesc, bdesc, idesc_pv, (k & 3) ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);           // K(j) (read by S(j), issued earlier) and V(j) are no longer needed
        umma_commit(&pv_full[j & 1]);
      }
    }
    __syncwarp();
  } else {
    // ---- softmax / output: thread = (query row, 64-key half) ----
    const int qd = warp & 3;                   // TMEM lane quarter this warp may access
    const int hf = (warp - 2) >> 2;            // which 64 keys of every tile
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const float sl2 = p.scale * kLog2e;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;

    auto accumulate_pv = [&](int jj, float alpha) {   // O = alpha * O + PV(jj)[hf]
      mbar_wait(&pv_full[jj & 1], (jj >> 1) & 1);
      tc_fence_after();
      const uint32_t tPV = tmem_base + 2 * BKV + (uint32_t)(((jj & 1) * 2 + hf) * D) + lane_addr;
      uint32_t r0[32], r1[32];
      tmem_ld32(tPV, r0);
      tmem_ld32(tPV + 32, r1);
      tmem_ld_wait();
      const float2 av = make_float2(alpha, alpha);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {                  // packed f32x2 arithmetic (sm_100)
        const float2 a = __ffma2_rn(make_float2(o[i], o[i + 1]), av, make_float2(__uint_as_float(r0[i]), __uint_as_float(r0[i + 1])));
        const float2 c = __ffma2_rn(make_float2(o[32 + i], o[33 + i]), av, make_float2(__uint_as_float(r1[i]), __uint_as_float(r1[i + 1])));
        o[i] = a.x; o[i + 1] = a.y; o[32 + i] = c.x; o[33 + i] = c.y;
      }
    };

    for (int j = 0; j < nt; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t tS = tmem_base + (uint32_t)((j & 1) * BKV + hf * 64) + lane_addr;
      uint32_t s[64];
      {
        uint32_t(&c0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&c1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        tmem_ld32(tS, c0); tmem_ld32(tS + 32, c1);
        tmem_ld_wait();
      }
      int kvalid = p.nk - j * BKV - hf * 64;           // keys of this half-tile that exist (may be <= 0 on the last tile)
      if (p.causal) kvalid = min(kvalid, q0 + r + 1 - j * BKV - hf * 64);   // ... and that query row q0 + r may see
      if (kvalid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kvalid) s[i] = 0xff800000u;   // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(s[i])); mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(s[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
      }
      float mx = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2);   // scale > 0
      // a half that has not seen a valid key yet keeps m = -inf; use 0 as the reference so that ex2(-inf - 0) = 0
      const float mref = (mx == -INFINITY) ? 0.f : mx;
      const float alpha = ex2_approx(m_run - mref);    // first tile: ex2(-inf) = 0
      m_run = mx;
      float2 rsa = make_float2(0.f, 0.f), rsb = make_float2(0.f, 0.f);
      const float2 sl2v = make_float2(sl2, sl2), nm = make_float2(-mref, -mref);
      uint8_t* prow = sP + ((j & 1) * 2 + hf) * TILE_BYTES + r * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[g * 8 + i]), __uint_as_float(s[g * 8 + i + 1])), sl2v, nm);
          e[i] = ex2_approx(x.x); e[i + 1] = ex2_approx(x.y);
        }
        rsa = __fadd2_rn(rsa, __fadd2_rn(make_float2(e[0], e[1]), make_float2(e[2], e[3])));
        rsb = __fadd2_rn(rsb, __fadd2_rn(make_float2(e[4], e[5]), make_float2(e[6], e[7])));
        uint4 w;
        w.x = pack_bf162(e[0], e[1]); w.y = pack_bf162(e[2], e[3]);
        w.z = pack_bf162(e[4], e[5]); w.w = pack_bf162(e[6], e[7]);
        *reinterpret_cast<uint4*>(prow + ((g ^ (r & 7)) << 4)) = w;
      }
      l_run = l_run * alpha + ((rsa.x + rsa.y) + (rsb.x + rsb.y));
      tc_fence_before();                 // my TMEM reads of S(j) are done
      fence_async_smem();                // my P writes are visible to the tensor core (async proxy)
      mbar_arrive(&p_full[j & 1]);
      if (j > 0) accumulate_pv(j - 1, alpha_prev);     // overlaps PV(j) / S(j+1) on the tensor pipe
      alpha_prev = alpha;
    }
    accumulate_pv(nt - 1, alpha_prev);
    tc_fence_before();
    // ---- merge the two key halves of each row: half 1 parks (O, m, l) in shared memory (the P tiles are dead now) ----
    float* xch = reinterpret_cast<float*>(sP);          // [128 rows][67] fp32, odd stride => conflict-free
    asm volatile("bar.sync 1, 256;" ::: "memory");      // every softmax thread has seen the last pv_full
    if (hf == 1) {
#pragma unroll
      for (int i = 0; i < D; ++i) xch[r * 67 + i] = o[i];
      xch[r * 67 + 64] = m_run;
      xch[r * 67 + 65] = l_run;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int row = q0 + r;
    if (hf == 0 && row < p.nq) {
      const float m1 = xch[r * 67 + 64], l1 = xch[r * 67 + 65];
      const float m = fmaxf(m_run, m1);                // half 0 always holds key 0, so m is finite
      const float a0 = ex2_approx(m_run - m), a1 = (m1 == -INFINITY) ? 0.f : ex2_approx(m1 - m);
      const float l = l_run * a0 + l1 * a1;
      const float inv = 1.f / l;
      const float w0 = a0 * inv, w1 = a1 * inv;
      bf16* dst = p.o + (long long)b * p.bso + (long long)row * p.ldo + h * D;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = o[g * 8 + i] * w0 + xch[r * 67 + g * 8 + i] * w1;
        uint4 w;
        w.x = pack_bf162(v[0], v[1]); w.y = pack_bf162(v[2], v[3]);
        w.z = pack_bf162(v[4], v[5]); w.w = pack_bf162(v[6], v[7]);
        *reinterpret_cast<uint4*>(dst + g * 8) = w;
      }
      if (p.lse) p.lse[((long long)b * p.heads + h) * p.nq + row] = (m + log2f(l)) / kLog2e;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// {64 d, rows, images} view of a [nb, rows, heads*64] tensor with row stride ld and image stride bs
int make_qkv_map(CUtensorMap* m, const void* base, int width, int rows, int nb, long long ld, long long bs) {
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)nb};
  cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(nb > 1 ? bs : ld * rows) * 2};
  cuuint32_t box[3] = {64, 128, 1};
  return vn_make_map(m, base, 3, dims, str, box);
}

}  // namespace

extern "C" int vn_attention_fwd(const vn_attn_desc* d, vn_stream_t s) {
  VN_CHECK(d != nullptr, "attention: null descriptor");
  VN_CHECK(d->nb > 0 && d->heads > 0 && d->nq > 0 && d->nk > 0, "attention: empty problem");
  VN_CHECK(d->ldq % 8 == 0 && d->ldk % 8 == 0 && d->ldv % 8 == 0 && d->ldo % 8 == 0 && d->bsq % 8 == 0 &&
               d->bsk % 8 == 0 && d->bsv % 8 == 0 && d->bso % 8 == 0,
           "attention: strides must be multiples of 8 elements");
  VN_CHECK(((reinterpret_cast<uintptr_t>(d->q) | reinterpret_cast<uintptr_t>(d->k) | reinterpret_cast<uintptr_t>(d->v) |
             reinterpret_cast<uintptr_t>(d->o)) & 15) == 0, "attention: q/k/v/o must be 16-byte aligned");
  CUtensorMap tq, tk, tv;
  const int width = d->heads * D;
  if (make_qkv_map(&tq, d->q, width, d->nq, d->nb, d->ldq, d->bsq)) return -1;
  if (make_qkv_map(&tk, d->k, width, d->nk, d->nb, d->ldk, d->bsk)) return -1;
  if (make_qkv_map(&tv, d->v, width, d->nk, d->nb, d->ldv, d->bsv)) return -1;
  FwdParams p{};
  p.nq = d->nq; p.nk = d->nk; p.heads = d->heads; p.scale = d->scale; p.causal = d->causal;
  p.o = (bf16*)d->o; p.ldo = d->ldo; p.bso = d->bso;
  p.lse = d->lse;
  static bool configured = false;
  if (!configured) {
    VN_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  dim3 grid(vn_cdiv(d->nq, BQ), d->heads, d->nb);
  VN_LAUNCH(attn_fwd_tc_kernel, grid, kThreads, SMEM_BYTES, (cudaStream_t)s, tq, tk, tv, p);
  return 0;
}
