// vn_attn_bwd_tc.cu — attention backward on tcgen05 / TMEM / TMA (head_dim 64): the autograd backward of
// reference models/xti_attention_processor.py:44-50 (training/coach.py:214), logits recomputed from the saved
// log-sum-exp, nothing of size nq x nk ever leaves the SM.
//
//   delta = rowsum(dO * O)                                        (small CUDA-core kernel)
//   dKV kernel, CTA = (128-key tile, head, image[, query split]), loop over 128-query tiles:
//       S^T = K Q^T, dP^T = V dO^T            tcgen05.mma  -> TMEM [0,128) and [128,256)
//       P^T = exp2(S^T*c - lse[q]),  dS^T = P^T * (dP^T - delta[q]) * scale        (8 warps, thread = key row x 64 queries)
//       dV += P^T dO,  dK += dS^T Q           tcgen05.mma, A = P^T / dS^T from shared memory, B = the dO / Q tile used in
//                                             place as an MN-major operand        -> TMEM [256,320) and [320,384)
//   dQ kernel, CTA = (128-query tile, head, image), loop over 128-key tiles:
//       S = Q K^T, dP = dO V^T -> TMEM; dS = P * (dP - delta) * scale -> smem; dQ += dS K (K tile in place, MN-major).
// Two kernels (S recomputed twice) instead of one with global atomics on dQ: results stay deterministic.
// Cross-attention (nk = 77) has a single key tile: the dKV kernel then splits the query range over CTAs and adds its
// partial dK / dV into an fp64 scratch (order-independent), finished by a small conversion kernel.
// PERSISTENT schedule (self-attention whose 2 * tiles * heads * images work items do not fill whole waves of SMs - 64x64
// latents at B = 1: 160 dK/dV + 160 dQ items on 148 SMs = 2.16 waves, i.e. three): grid = #SMs, CTA c runs the whole items
// c, c + #SMs, ... one after the other and then one PART of a leftover item (the leftover items' loops are cut into equal
// ranges spread over all CTAs); a part leaves a plain fp32 partial tile in a caller-provided scratch and
// attn_bwd_fixup_kernel adds the parts of each leftover item up in a fixed order (deterministic, no atomics).
#include "vn_tma.cuh"

#include <stdlib.h>

namespace {

constexpr int D = 64;
constexpr int T = 128;                            // tile edge (queries and keys)
constexpr int TILE_BYTES = T * 128;               // [128 rows x 64 bf16], 128B-swizzled
// Compute warps per TMEM lane quarter.  P / dS are elementwise in the backward (row constants lse / delta are known), so
// the 64 columns of a half tile split freely over warps: with 4 per quarter (16 compute warps, 16 columns per thread and
// half) each scheduler has 4 warps to hide the tcgen05.ld -> ex2 -> pack -> st.shared -> fence -> arrive chain behind,
// instead of 2 (ncu: issue utilisation 0.35 per scheduler, tensor pipe 27 % of active cycles).
#ifndef VN_ATTN_BWD_CW
#define VN_ATTN_BWD_CW 4
#endif
constexpr int kCW = VN_ATTN_BWD_CW;                // 2 or 4
constexpr int kComputeThreads = 128 * kCW;
constexpr int CWCOLS = 64 / kCW;                   // columns per thread and half tile
constexpr int kThreads = 64 + kComputeThreads;     // producer warp, MMA warp, compute warps
static_assert(kCW == 2 || kCW == 4, "compute warps per lane quarter");
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[N]) {
  if constexpr (N == 32) tmem_ld32(taddr, v);
  else tmem_ld16(taddr, v);
}
// P^T / dS^T (dK/dV body) and dS (dQ body) go back into TENSOR memory as packed bf16 pairs and feed the second product as
// its A operand from there: with A in shared memory a 128 x 64 x 16 MMA reads 4 KB (A) + 2 KB (B) per 32 tensor cycles -
// 192 B/clk against the 128 B/clk shared-memory port - and the compute warps pay 64 KB of swizzled stores + a proxy fence
// per tile pair.  VN_ATTN_BWD_TS=0 keeps the shared-memory operand (cross-check).
#ifndef VN_ATTN_BWD_TS
#define VN_ATTN_BWD_TS 1
#endif
constexpr bool kTS = VN_ATTN_BWD_TS != 0;
constexpr int T_PT = 384, T_DST = 448;             // TMEM columns of the packed operands: 2 halves x 32 columns each
template <int N>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&v)[N]) {
  if constexpr (N == 16) tmem_st16(taddr, v);
  else tmem_st8(taddr, v);
}
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory"); }
constexpr int TMEM_COLS = 512;
constexpr float kLog2e = 1.4426950408889634f;

struct BwdParams {
  int nb, heads, nq, nk;
  float scale;
  const float* lse; const float* delta;
  bf16* dq; long long lddq, bsdq;
  bf16* dk; long long lddk, bsdk;
  bf16* dv; long long lddv, bsdv;
  double* dkv_acc;
  int qsplits, qtiles_per_split;
  int n_dkv, ktiles, qtiles;        // fused-grid decomposition
  int causal;                       // key j visible to query i only if j <= i
  // persistent schedule: items [0, n_dkv) are dK/dV items, [n_dkv, items) dQ items; the last n_left items are split
  int persistent, items, rounds, n_left, parts;
  int left_dkv;                     // how many of the n_left split items are dK/dV items (the rest are dQ items)
  float* part;                      // [n_left * parts][2][128][64] fp32 partial tiles (dQ: [0]; dK/dV: [0] = dV, [1] = dK)
  long long* dbg;                   // -DVN_TIMELINE builds: clock stamps of CTA 0 (scripts/attn_timeline.py), else unused
};

// In-kernel timeline (instrumented builds): CTA 0, first body (a dK/dV item), query tiles [8, 16): dbg[1024 + (i - 8) * 16 + e];
//   e = 0/1 S^T,dP^T(i, half 0/1) issued, 2/3 dV,dK(i, half 0/1) issued (p_full seen), 4/5 s_full(half 0/1) seen by compute
//   thread 64, 6/7 its arrive on p_full(half 0/1)
#ifdef VN_TIMELINE
#define VN_BSTAMP(i, e)                                                                                              \
  do {                                                                                                               \
    if (p.dbg && blockIdx.x == 0 && seg == 0 && (i) >= 8 && (i) < 16) p.dbg[1024 + ((i) - 8) * 16 + (e)] = clock64(); \
  } while (0)
#else
#define VN_BSTAMP(i, e) do { } while (0)
#endif
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int m, int n, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// a persistent CTA runs several bodies with different shared-memory layouts: a body retires its barrier words before the next
// one may reuse the bytes for data
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// half-tile variants (64 of the 128 columns / k-rows), used to pipeline the tensor pipe against the compute warps
// D[128 x 64] (+)= A[block h: 128 x 64 k, K-major] * B[k-rows h*64.. of the MN-major tile]
__device__ __forceinline__ void mma_kmn_half(uint32_t tacc, uint32_t a, uint32_t b, int h, uint32_t id, bool accumulate) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16(tacc, umma_desc_k_sw128(a + h * TILE_BYTES) + (uint64_t)(k * 2), desc_mn_sw128(b + (h * 4 + k) * 2048), id,
              (accumulate || k) ? 1u : 0u);
}
// The same products from PRECOMPUTED tile descriptors.  All operand tiles are 1024-byte aligned and below 256 KB, so the
// descriptor of (tile + off) is the tile's descriptor plus (off >> 4) in the 14-bit start-address field: one 64-bit add per
// MMA in the single issuing thread instead of rebuilding the descriptor (shift, two masks, merge) in front of every
// tcgen05.mma - the issue thread is on the critical path between the compute warps' arrive and the next S / dP.
__device__ __forceinline__ void mma_kk_half_d(uint32_t tacc, uint64_t da, uint64_t db, int h, uint32_t id64) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16(tacc + (uint32_t)(h * 64), da + (uint64_t)(k * 2), db + (uint64_t)(h * 512 + k * 2), id64, k ? 1u : 0u);
}
__device__ __forceinline__ void mma_tmn_half_d(uint32_t tacc, uint32_t ta, uint64_t dbmn, int h, uint32_t id, bool accumulate) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16_ts(tacc, ta + (uint32_t)(k * 8), dbmn + (uint64_t)((h * 4 + k) * 128), id, (accumulate || k) ? 1u : 0u);
}
__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t (&v)[64]) {
  uint32_t(&c0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
  uint32_t(&c1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
  tmem_ld32(taddr, c0);
  tmem_ld32(taddr + 32, c1);
}
[[maybe_unused]] __device__ __forceinline__ void ld_acc(uint32_t taddr, uint32_t (&v)[64]) { ld64(taddr, v); }   // kCW == 2
__device__ __forceinline__ void ld_acc(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld32(taddr, v); }
__device__ __forceinline__ void store_row8(uint8_t* row, int chunk, int r, const float (&e)[8]) {
  uint4 w;
  w.x = pack_bf162(e[0], e[1]); w.y = pack_bf162(e[2], e[3]);
  w.z = pack_bf162(e[4], e[5]); w.w = pack_bf162(e[6], e[7]);
  *reinterpret_cast<uint4*>(row + ((chunk ^ (r & 7)) << 4)) = w;
}

// =================================================================================================
// dK, dV
// =================================================================================================
constexpr int DKV_STAGE_BYTES = 2 * TILE_BYTES + 2 * T * 4;        // Q, dO, lse, delta
constexpr int kMaxSegments = 4;                    // bodies a persistent CTA may run (each on its own block of 16 barrier words)
constexpr int DKV_SMEM = 2 * TILE_BYTES /*K,V*/ + 2 * DKV_STAGE_BYTES + 2 * 2 * TILE_BYTES /*P^T, dS^T*/ + kMaxSegments * 128 + 1024;

// seg: which barrier block of the CTA to use (a persistent CTA runs several bodies one after the other, each on fresh
// mbarriers); t_count > 0: loop over query tiles [t_begin, t_begin + t_count) only; slot >= 0: leave fp32 partial tiles
template <bool ATOMIC>
__device__ __forceinline__ void dkv_body(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                         const CUtensorMap& tmdO, const BwdParams& p, uint32_t tmem_base, int bx, int by, int bz,
                                         int seg = 0, int t_begin = 0, int t_count = 0, int slot = -1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sP = sV + TILE_BYTES;                     // P^T  : two [128 keys x 64 queries] blocks
  uint8_t* sdS = sP + 2 * TILE_BYTES;                // dS^T : same
  uint8_t* sStage = sdS + 2 * TILE_BYTES;            // 2 x {Q tile, dO tile, lse[128], delta[128]}
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + 2 * DKV_STAGE_BYTES) + seg * 16;
  uint64_t* kv_full = bars;
  uint64_t* st_full = bars + 1;                      // [2]
  uint64_t* st_empty = bars + 3;                     // [2]
  uint64_t* s_full = bars + 5;                       // [2] per 64-query half
  uint64_t* p_full = bars + 7;                       // [2]
  uint64_t* acc_full = bars + 9;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = bx * T, h = by;
  const int b = bz / p.qsplits, split = bz % p.qsplits;
  const int total_qt = (p.nq + T - 1) / T;
  const int qt_begin = t_count > 0 ? t_begin : split * p.qtiles_per_split;
  const int nt = t_count > 0 ? t_count : min(total_qt, qt_begin + p.qtiles_per_split) - qt_begin;   // host guarantees >= 1

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&st_full[s], 1); mbar_init(&st_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], kComputeThreads / 32);     // one arrival per compute WARP
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();                                   // barrier words initialised (TMEM is allocated once per CTA by the kernel)
  tc_fence_after();
  const uint32_t tST = tmem_base, tdPT = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320;
  const long long sidx = ((long long)b * p.heads + h) * p.nq;
  pdl_wait();

  if (warp == 0) {
    {                                                  // warp-uniform control flow, one elected lane issues
      const bool leader = elect_one();
      if (leader) {
        mbar_expect_tx(kv_full, 2 * TILE_BYTES);
        tma_load_3d(sK, &tmK, kv_full, h * D, k0, b);
        tma_load_3d(sV, &tmV, kv_full, h * D, k0, b);
      }
      for (int i = 0; i < nt; ++i) {
        const int s = i & 1;
        const int q0 = (qt_begin + i) * T;
        uint8_t* st = sStage + s * DKV_STAGE_BYTES;
        mbar_wait(&st_empty[s], ((i >> 1) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&st_full[s], 2 * TILE_BYTES);
          tma_load_3d(st, &tmQ, &st_full[s], h * D, q0, b);
          tma_load_3d(st + TILE_BYTES, &tmdO, &st_full[s], h * D, q0, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {                                                  // warp-uniform control flow; one elected lane issues
      const bool leader = elect_one();
      constexpr uint32_t id_s = idesc(T, 64, 0);       // 128 keys x 64 queries, both K-major
      constexpr uint32_t id_acc = idesc(T, D, 1);      // 128 x 64, B MN-major
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP), adS = smem_u32(sdS);
      // Software pipeline over 64-query halves: while the compute warps turn S^T/dP^T of one half into P^T/dS^T, the
      // tensor pipe accumulates dV/dK of the other half and produces the next S^T/dP^T.
      const uint64_t dKk = umma_desc_k_sw128(aK), dVk = umma_desc_k_sw128(aV);
      const uint32_t aSt0 = smem_u32(sStage), aSt1 = aSt0 + DKV_STAGE_BYTES;
      const uint64_t dQk[2] = {umma_desc_k_sw128(aSt0), umma_desc_k_sw128(aSt1)};
      const uint64_t ddOk[2] = {umma_desc_k_sw128(aSt0 + TILE_BYTES), umma_desc_k_sw128(aSt1 + TILE_BYTES)};
      const uint64_t dQmn[2] = {desc_mn_sw128(aSt0), desc_mn_sw128(aSt1)};
      const uint64_t ddOmn[2] = {desc_mn_sw128(aSt0 + TILE_BYTES), desc_mn_sw128(aSt1 + TILE_BYTES)};
      auto issue_sdp = [&](int i, int hh) {            // S^T_h = K Q_h^T, dP^T_h = V dO_h^T of query tile i
        if (leader) {
          mma_kk_half_d(tST, dKk, dQk[i & 1], hh, id_s);
          mma_kk_half_d(tdPT, dVk, ddOk[i & 1], hh, id_s);
          umma_commit(&s_full[hh]);                    // in-order retirement: also covers dV/dK(i-1, h) -> operand block h free
          VN_BSTAMP(i, hh);
        }
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&st_full[0], 0);
      tc_fence_after();
      issue_sdp(0, 0);
      issue_sdp(0, 1);
      for (int i = 0; i < nt; ++i) {
        const uint32_t aQ = smem_u32(sStage + (i & 1) * DKV_STAGE_BYTES), adO = aQ + TILE_BYTES;
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(&p_full[hh], i & 1);               // P^T_h / dS^T_h in smem; TMEM half h has been read
          tc_fence_after();
          if (leader) {
            if (kTS) {
              mma_tmn_half_d(tdV, tmem_base + T_PT + hh * 32, ddOmn[i & 1], hh, id_acc, i > 0 || hh > 0);   // dV += P^T_h dO_h
              mma_tmn_half_d(tdK, tmem_base + T_DST + hh * 32, dQmn[i & 1], hh, id_acc, i > 0 || hh > 0);   // dK += dS^T_h Q_h
            } else {
              mma_kmn_half(tdV, aP, adO, hh, id_acc, i > 0 || hh > 0);
              mma_kmn_half(tdK, adS, aQ, hh, id_acc, i > 0 || hh > 0);
            }
            if (hh == 1) umma_commit(&st_empty[i & 1]);
            VN_BSTAMP(i, 2 + hh);
          }
          if (i + 1 < nt) {
            if (hh == 0) { mbar_wait(&st_full[(i + 1) & 1], ((i + 1) >> 1) & 1); tc_fence_after(); }
            issue_sdp(i + 1, hh);
          }
        }
      }
      if (leader) umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    const int qd = warp & 3, cs = (warp - 2) >> 2;     // cs: which CWCOLS of the 64 queries of a half (and dV / dK columns at the end)
    const int r = qd * 32 + lane;                      // key row of the tile
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const float sl2 = p.scale * kLog2e;
    // per-query lse (pre-scaled by log2 e) and delta of a query tile live in the stage's [2][128] fp32 vector; the 256
    // compute threads fetch the NEXT tile's values into a register while they work on the current one
    const int te = threadIdx.x - 64;                   // [0,128) -> lse, [128,256) -> delta (threads beyond 256 fetch nothing)
    auto fetch = [&](int i) -> float {
      const int q = (qt_begin + i) * T + (te & 127);
      if (i >= nt || q >= p.nq || te >= 256) return 0.f;
      return te < 128 ? -p.lse[sidx + q] * kLog2e : -p.delta[sidx + q] * p.scale;     // pre-negated / pre-scaled
    };
    if (te < 256) reinterpret_cast<float*>(sStage + 2 * TILE_BYTES)[te] = fetch(0);
    bar_compute();
    for (int i = 0; i < nt; ++i) {
      const int s = i & 1;
      const int q0 = (qt_begin + i) * T;
      const float next_val = fetch(i + 1);
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int cbase = hh * 64 + cs * CWCOLS;       // first query column of this thread in the tile
        const float* lse = reinterpret_cast<const float*>(sStage + s * DKV_STAGE_BYTES + 2 * TILE_BYTES) + cbase;
        const float* dl = lse + T;
        mbar_wait(&s_full[hh], i & 1);
        tc_fence_after();
        if (threadIdx.x == 64) VN_BSTAMP(i, 4 + hh);
        uint32_t sv[CWCOLS], dp[CWCOLS];
        tmem_ld_cols(tST + lane_addr + cbase, sv);
        tmem_ld_cols(tdPT + lane_addr + cbase, dp);
        tmem_ld_wait();
        const int qvalid = p.nq - q0 - cbase;          // queries of this slice that exist
        uint8_t* prow = sP + hh * TILE_BYTES + r * 128;
        uint8_t* drow = sdS + hh * TILE_BYTES + r * 128;
        uint32_t pw[CWCOLS / 2], dw[CWCOLS / 2];      // packed bf16 pairs for the tensor-memory operand
        const float2 sl2v = make_float2(sl2, sl2), scv = make_float2(p.scale, p.scale);
#pragma unroll
        for (int g = 0; g < CWCOLS / 8; ++g) {
          // nl = -lse * log2(e), nd = -delta * scale for 8 consecutive queries (packed f32x2 arithmetic, sm_100)
          const float4 l0 = *reinterpret_cast<const float4*>(lse + g * 8), l1 = *reinterpret_cast<const float4*>(lse + g * 8 + 4);
          const float4 d0 = *reinterpret_cast<const float4*>(dl + g * 8), d1 = *reinterpret_cast<const float4*>(dl + g * 8 + 4);
          const float2 nl[4] = {make_float2(l0.x, l0.y), make_float2(l0.z, l0.w), make_float2(l1.x, l1.y), make_float2(l1.z, l1.w)};
          const float2 nd[4] = {make_float2(d0.x, d0.y), make_float2(d0.z, d0.w), make_float2(d1.x, d1.y), make_float2(d1.z, d1.w)};
          float pe[8], de[8];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[g * 8 + 2 * c]), __uint_as_float(sv[g * 8 + 2 * c + 1])), sl2v, nl[c]);
            const float2 pp = make_float2(ex2_approx(x.x), ex2_approx(x.y));
            const float2 tt = __ffma2_rn(make_float2(__uint_as_float(dp[g * 8 + 2 * c]), __uint_as_float(dp[g * 8 + 2 * c + 1])), scv, nd[c]);
            const float2 dd = __fmul2_rn(pp, tt);
            pe[2 * c] = pp.x; pe[2 * c + 1] = pp.y;
            de[2 * c] = dd.x; de[2 * c + 1] = dd.y;
          }
          if (qvalid < CWCOLS) {                       // ragged last query tile: padded queries contribute nothing
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (g * 8 + c >= qvalid) { pe[c] = 0.f; de[c] = 0.f; }
          }
          if (p.causal) {                              // queries before this key row do not see it
            const int qfirst = k0 + r - (q0 + cbase + g * 8);
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < qfirst) { pe[c] = 0.f; de[c] = 0.f; }
          }
          if (kTS) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              pw[g * 4 + c] = pack_bf162(pe[2 * c], pe[2 * c + 1]);
              dw[g * 4 + c] = pack_bf162(de[2 * c], de[2 * c + 1]);
            }
          } else {
            store_row8(prow, cs * (CWCOLS / 8) + g, r, pe);
            store_row8(drow, cs * (CWCOLS / 8) + g, r, de);
          }
        }
        if (kTS) {
          tmem_st_cols(tmem_base + lane_addr + T_PT + (uint32_t)(cbase >> 1), pw);
          tmem_st_cols(tmem_base + lane_addr + T_DST + (uint32_t)(cbase >> 1), dw);
          tmem_st_wait();
          tc_fence_before();
        } else {
          tc_fence_before();
          fence_async_smem();
        }
        __syncwarp();                              // every lane's stores are complete and fenced
        if (lane == 0) mbar_arrive(&p_full[hh]);   // one arrival per warp (512 arrivals on one shared-memory word serialise)
        __syncwarp();                              // reconverge: the next tcgen05.ld is .sync.aligned
        if (threadIdx.x == 64) VN_BSTAMP(i, 6 + hh);
      }
      if (te < 256) reinterpret_cast<float*>(sStage + ((i + 1) & 1) * DKV_STAGE_BYTES + 2 * TILE_BYTES)[te] = next_val;
      bar_compute();
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // dV | dK sit side by side in TMEM columns [256, 384): compute warp group cs finishes 128 / kCW of those columns
    constexpr int EW = 128 / kCW;                      // 64 (one whole accumulator) or 32 (half of one)
    const int sel = (cs * EW) / 64, col0 = (cs * EW) % 64;            // sel 0: dV, 1: dK
    uint32_t acc[EW];
    ld_acc(tdV + lane_addr + cs * EW, acc);
    tmem_ld_wait();
    tc_fence_before();
    const int krow = k0 + r;
    if (slot >= 0) {
      float* dst = p.part + (((long long)slot * 2 + sel) * T + r) * D + col0;
#pragma unroll
      for (int c = 0; c < EW; c += 4)
        *reinterpret_cast<float4*>(dst + c) = make_float4(__uint_as_float(acc[c]), __uint_as_float(acc[c + 1]),
                                                          __uint_as_float(acc[c + 2]), __uint_as_float(acc[c + 3]));
    } else if (krow < p.nk) {
      if (ATOMIC) {
        const long long C = (long long)p.heads * D;
        double* dst = p.dkv_acc + (sel == 0 ? (long long)p.nb * p.nk * C : 0) + ((long long)b * p.nk + krow) * C + h * D + col0;
#pragma unroll
        for (int c = 0; c < EW; ++c) atomicAdd(dst + c, (double)__uint_as_float(acc[c]));
      } else {
        bf16* dst = (sel == 0 ? p.dv + (long long)b * p.bsdv + (long long)krow * p.lddv
                              : p.dk + (long long)b * p.bsdk + (long long)krow * p.lddk) + h * D + col0;
#pragma unroll
        for (int g = 0; g < EW / 8; ++g) {
          uint4 w;
          w.x = pack_bf162(__uint_as_float(acc[g * 8 + 0]), __uint_as_float(acc[g * 8 + 1]));
          w.y = pack_bf162(__uint_as_float(acc[g * 8 + 2]), __uint_as_float(acc[g * 8 + 3]));
          w.z = pack_bf162(__uint_as_float(acc[g * 8 + 4]), __uint_as_float(acc[g * 8 + 5]));
          w.w = pack_bf162(__uint_as_float(acc[g * 8 + 6]), __uint_as_float(acc[g * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + g * 8) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.persistent) {
    if (threadIdx.x == 0)
      for (int i = 0; i < 10; ++i) mbar_inval(bars + i);
    __syncthreads();
  }
}

// =================================================================================================
// dQ
// =================================================================================================
constexpr int DQ_SMEM = 2 * TILE_BYTES /*Q,dO*/ + 2 * 2 * TILE_BYTES /*K,V x 2 stages*/ + 2 * TILE_BYTES /*dS*/ + kMaxSegments * 128 + 1024;

__device__ __forceinline__ void dq_body(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                        const CUtensorMap& tmdO, const BwdParams& p, uint32_t tmem_base, int bx, int by, int bz,
                                        int seg = 0, int t_begin = 0, int t_count = 0, int slot = -1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + TILE_BYTES;
  uint8_t* sdS = sdO + TILE_BYTES;                   // two [128 queries x 64 keys] blocks
  uint8_t* sKV = sdS + 2 * TILE_BYTES;               // stage s: K at s*2*TILE, V right after
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + 4 * TILE_BYTES) + seg * 16;
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                      // [2]
  uint64_t* kv_empty = bars + 3;                     // [2]
  uint64_t* s_full = bars + 5;                       // [2] per 64-key half
  uint64_t* p_full = bars + 7;                       // [2]
  uint64_t* acc_full = bars + 9;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = bx * T, h = by, b = bz;
  const int kt0 = t_count > 0 ? t_begin : 0;                                   // first key tile of this body's range
  const int nt = t_count > 0 ? t_count : (p.nk + T - 1) / T;                   // key tiles in the range

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], kComputeThreads / 32);     // one arrival per compute WARP
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();                                   // barrier words initialised (TMEM is allocated once per CTA by the kernel)
  tc_fence_after();
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdQ = tmem_base + 256;
  pdl_wait();

  if (warp == 0) {
    {                                                  // warp-uniform control flow, one elected lane issues
      const bool leader = elect_one();
      if (leader) {
        mbar_expect_tx(q_full, 2 * TILE_BYTES);
        tma_load_3d(sQ, &tmQ, q_full, h * D, q0, b);
        tma_load_3d(sdO, &tmdO, q_full, h * D, q0, b);
      }
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&kv_full[s], 2 * TILE_BYTES);
          tma_load_3d(sKV + s * 2 * TILE_BYTES, &tmK, &kv_full[s], h * D, (kt0 + j) * T, b);
          tma_load_3d(sKV + s * 2 * TILE_BYTES + TILE_BYTES, &tmV, &kv_full[s], h * D, (kt0 + j) * T, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {                                                  // warp-uniform control flow; one elected lane issues
      const bool leader = elect_one();
      constexpr uint32_t id_s = idesc(T, 64, 0);       // 128 queries x 64 keys
      constexpr uint32_t id_acc = idesc(T, D, 1);
      const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), adS = smem_u32(sdS);
      const uint64_t dQk = umma_desc_k_sw128(aQ), ddOk = umma_desc_k_sw128(adO);
      const uint32_t aKV0 = smem_u32(sKV), aKV1 = aKV0 + 2 * TILE_BYTES;
      const uint64_t dKk[2] = {umma_desc_k_sw128(aKV0), umma_desc_k_sw128(aKV1)};
      const uint64_t dVk[2] = {umma_desc_k_sw128(aKV0 + TILE_BYTES), umma_desc_k_sw128(aKV1 + TILE_BYTES)};
      const uint64_t dKmn[2] = {desc_mn_sw128(aKV0), desc_mn_sw128(aKV1)};
      auto issue_sdp = [&](int j, int hh) {            // S_h = Q K_h^T, dP_h = dO V_h^T of key tile j
        if (leader) {
          mma_kk_half_d(tS, dQk, dKk[j & 1], hh, id_s);
          mma_kk_half_d(tdP, ddOk, dVk[j & 1], hh, id_s);
          umma_commit(&s_full[hh]);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_sdp(0, 0);
      issue_sdp(0, 1);
      for (int j = 0; j < nt; ++j) {
        const uint32_t aK = smem_u32(sKV + (j & 1) * 2 * TILE_BYTES);
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(&p_full[hh], j & 1);
          tc_fence_after();
          if (leader) {
            if (kTS) mma_tmn_half_d(tdQ, tmem_base + T_PT + hh * 32, dKmn[j & 1], hh, id_acc, j > 0 || hh > 0);     // dQ += dS_h K_h
            else mma_kmn_half(tdQ, adS, aK, hh, id_acc, j > 0 || hh > 0);
            if (hh == 1) umma_commit(&kv_empty[j & 1]);
          }
          if (j + 1 < nt) {
            if (hh == 0) { mbar_wait(&kv_full[(j + 1) & 1], ((j + 1) >> 1) & 1); tc_fence_after(); }
            issue_sdp(j + 1, hh);
          }
        }
      }
      if (leader) umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    const int qd = warp & 3, cs = (warp - 2) >> 2;
    const int r = qd * 32 + lane;                      // query row of the tile
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const float sl2 = p.scale * kLog2e;
    const int row = q0 + r;
    const long long sidx = ((long long)b * p.heads + h) * p.nq;
    const float nlse2 = row < p.nq ? -p.lse[sidx + row] * kLog2e : -INFINITY;  // padded query rows: P = 0
    const float ndl = row < p.nq ? -p.delta[sidx + row] * p.scale : 0.f;
    for (int j = 0; j < nt; ++j) {
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int cbase = hh * 64 + cs * CWCOLS;       // first key column of this thread in the tile
        mbar_wait(&s_full[hh], j & 1);
        tc_fence_after();
        uint32_t sv[CWCOLS], dp[CWCOLS];
        tmem_ld_cols(tS + lane_addr + cbase, sv);
        tmem_ld_cols(tdP + lane_addr + cbase, dp);
        tmem_ld_wait();
        const int kvalid = p.nk - (kt0 + j) * T - cbase;
        uint8_t* drow = sdS + hh * TILE_BYTES + r * 128;
        uint32_t dw[CWCOLS / 2];
        const float2 sl2v = make_float2(sl2, sl2), scv = make_float2(p.scale, p.scale);
        const float2 nlv = make_float2(nlse2, nlse2), ndv = make_float2(ndl, ndl);
#pragma unroll
        for (int g = 0; g < CWCOLS / 8; ++g) {
          float de[8];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[g * 8 + 2 * c]), __uint_as_float(sv[g * 8 + 2 * c + 1])), sl2v, nlv);
            const float2 pp = make_float2(ex2_approx(x.x), ex2_approx(x.y));
            const float2 tt = __ffma2_rn(make_float2(__uint_as_float(dp[g * 8 + 2 * c]), __uint_as_float(dp[g * 8 + 2 * c + 1])), scv, ndv);
            const float2 dd = __fmul2_rn(pp, tt);
            de[2 * c] = dd.x; de[2 * c + 1] = dd.y;
          }
          if (kvalid < CWCOLS) {                       // ragged last key tile
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (g * 8 + c >= kvalid) de[c] = 0.f;
          }
          if (p.causal) {                              // keys after this query row are masked
            const int klast = row - ((kt0 + j) * T + cbase + g * 8);
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c > klast) de[c] = 0.f;
          }
          if (kTS) {
#pragma unroll
            for (int c = 0; c < 4; ++c) dw[g * 4 + c] = pack_bf162(de[2 * c], de[2 * c + 1]);
          } else {
            store_row8(drow, cs * (CWCOLS / 8) + g, r, de);
          }
        }
        if (kTS) {
          tmem_st_cols(tmem_base + lane_addr + T_PT + (uint32_t)(cbase >> 1), dw);
          tmem_st_wait();
          tc_fence_before();
        } else {
          tc_fence_before();
          fence_async_smem();
        }
        __syncwarp();                              // every lane's stores are complete and fenced
        if (lane == 0) mbar_arrive(&p_full[hh]);   // one arrival per warp (512 arrivals on one shared-memory word serialise)
        __syncwarp();                              // reconverge: the next tcgen05.ld is .sync.aligned
      }
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    constexpr int QW = 64 / kCW;                       // dQ columns finished per thread
    uint32_t acc[QW];
    tmem_ld_cols(tdQ + lane_addr + cs * QW, acc);
    tmem_ld_wait();
    tc_fence_before();
    if (slot >= 0) {
      float* dst = p.part + (((long long)slot * 2) * T + r) * D + cs * QW;
#pragma unroll
      for (int c = 0; c < QW; c += 4)
        *reinterpret_cast<float4*>(dst + c) = make_float4(__uint_as_float(acc[c]), __uint_as_float(acc[c + 1]),
                                                          __uint_as_float(acc[c + 2]), __uint_as_float(acc[c + 3]));
    } else if (row < p.nq) {
      bf16* dst = p.dq + (long long)b * p.bsdq + (long long)row * p.lddq + h * D + cs * QW;
#pragma unroll
      for (int g = 0; g < QW / 8; ++g) {
        uint4 w;
        w.x = pack_bf162(__uint_as_float(acc[g * 8 + 0]), __uint_as_float(acc[g * 8 + 1]));
        w.y = pack_bf162(__uint_as_float(acc[g * 8 + 2]), __uint_as_float(acc[g * 8 + 3]));
        w.z = pack_bf162(__uint_as_float(acc[g * 8 + 4]), __uint_as_float(acc[g * 8 + 5]));
        w.w = pack_bf162(__uint_as_float(acc[g * 8 + 6]), __uint_as_float(acc[g * 8 + 7]));
        *reinterpret_cast<uint4*>(dst + g * 8) = w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.persistent) {
    if (threadIdx.x == 0)
      for (int i = 0; i < 10; ++i) mbar_inval(bars + i);
    __syncthreads();
  }
}

// Persistent schedule, position in the work order -> item id (id < n_dkv: dK/dV item, else dQ item id - n_dkv).  Order:
// whole dK/dV items, whole dQ items, then the split items (left_dkv dK/dV items, the rest dQ items): a dK/dV item takes
// ~1.5x as long as a dQ item (four products per tile pair against three), so every CTA should get one of each kind in its
// whole rounds and the split items should come from both kinds.
__device__ __forceinline__ int item_at(const BwdParams& p, int pos) {
  const int n_dq = p.items - p.n_dkv;
  const int A = p.n_dkv - p.left_dkv, B = n_dq - (p.n_left - p.left_dkv);
  if (pos < A) return pos;
  if (pos < A + B) return p.n_dkv + (pos - A);
  if (pos < A + B + p.left_dkv) return A + (pos - A - B);
  return p.n_dkv + B + (pos - A - B - p.left_dkv);
}

// One launch for both directions: CTAs [0, n_dkv) run the dK/dV body, the rest the dQ body.  With ~160 work items of each
// kind on 148 SMs, two separate launches cost two partial waves each; fused, the hardware scheduler back-fills SMs as
// CTAs retire (the longer dK/dV items are scheduled first).
template <bool ATOMIC>
__global__ void __launch_bounds__(kThreads, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV,
                                                                   const __grid_constant__ CUtensorMap tmdO,
                                                                   const BwdParams p) {
  pdl_trigger();
  // TMEM is allocated ONCE per CTA (the permit is relinquished right after, so a second tcgen05.alloc would be illegal);
  // both bodies lay their accumulators out inside the same 512 columns
  __shared__ uint32_t s_tmem;
  if ((threadIdx.x >> 5) == 1) tmem_alloc<TMEM_COLS>(&s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  auto release_tmem = [&]() {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) {
      tc_fence_after();
      tmem_dealloc<TMEM_COLS>(tmem_base);
    }
  };
  if (!p.persistent) {
    int id = blockIdx.x;
    if (id < p.n_dkv) {
      const int bx = id % p.ktiles; id /= p.ktiles;
      const int by = id % p.heads;
      dkv_body<ATOMIC>(tmQ, tmK, tmV, tmdO, p, tmem_base, bx, by, id / p.heads);
    } else {
      id -= p.n_dkv;
      const int bx = id % p.qtiles; id /= p.qtiles;
      const int by = id % p.heads;
      dq_body(tmQ, tmK, tmV, tmdO, p, tmem_base, bx, by, id / p.heads);
    }
    release_tmem();
    return;
  }
  // persistent: whole items blockIdx.x + seg * gridDim.x, then (CTAs below n_left * parts) one part of a leftover item
  const int nseg = p.rounds + ((int)blockIdx.x < p.n_left * p.parts ? 1 : 0);
  for (int seg = 0; seg < nseg; ++seg) {
    int id, t_begin = 0, t_count = 0, slot = -1;
    if (seg < p.rounds) {
      id = item_at(p, (int)blockIdx.x + seg * (int)gridDim.x);
    } else {
      const int li = (int)blockIdx.x / p.parts, part = (int)blockIdx.x - li * p.parts;
      id = item_at(p, p.items - p.n_left + li);
      const int tiles = id < p.n_dkv ? p.qtiles : p.ktiles;          // the loop of a dK/dV item runs over query tiles
      t_begin = part * tiles / p.parts;
      t_count = (part + 1) * tiles / p.parts - t_begin;
      slot = (int)blockIdx.x;
    }
    if (id < p.n_dkv) {
      const int bx = id % p.ktiles; id /= p.ktiles;
      const int by = id % p.heads;
      dkv_body<false>(tmQ, tmK, tmV, tmdO, p, tmem_base, bx, by, id / p.heads, seg, t_begin, t_count, slot);
    } else {
      id -= p.n_dkv;
      const int bx = id % p.qtiles; id /= p.qtiles;
      const int by = id % p.heads;
      dq_body(tmQ, tmK, tmV, tmdO, p, tmem_base, bx, by, id / p.heads, seg, t_begin, t_count, slot);
    }
  }
  release_tmem();
}

// Adds the parts of every leftover item of the persistent schedule (fixed order) and writes the bf16 rows:
// thread = (tile row, 8 columns), 16 rows per CTA, 8 CTAs per leftover item.
__global__ void __launch_bounds__(128) attn_bwd_fixup_kernel(const BwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int li = (int)blockIdx.x >> 3;
  const int r = (((int)blockIdx.x & 7) << 4) + ((int)threadIdx.x >> 3), cg = threadIdx.x & 7;
  int id = item_at(p, p.items - p.n_left + li);
  const bool is_dkv = id < p.n_dkv;
  if (!is_dkv) id -= p.n_dkv;
  const int tiles = is_dkv ? p.ktiles : p.qtiles;
  const int bx = id % tiles; id /= tiles;
  const int h = id % p.heads, b = id / p.heads;
  const int row = bx * T + r;
  if (row >= (is_dkv ? p.nk : p.nq)) return;
  const int nplanes = is_dkv ? 2 : 1;
  for (int pl = 0; pl < nplanes; ++pl) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int q = 0; q < p.parts; ++q) {
      const float* src = p.part + ((((long long)li * p.parts + q) * 2 + pl) * T + r) * D + cg * 8;
      const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
      acc[0] += v0.x; acc[1] += v0.y; acc[2] += v0.z; acc[3] += v0.w;
      acc[4] += v1.x; acc[5] += v1.y; acc[6] += v1.z; acc[7] += v1.w;
    }
    uint4 w;
    w.x = pack_bf162(acc[0], acc[1]); w.y = pack_bf162(acc[2], acc[3]);
    w.z = pack_bf162(acc[4], acc[5]); w.w = pack_bf162(acc[6], acc[7]);
    bf16* dst = !is_dkv ? p.dq + (long long)b * p.bsdq + (long long)row * p.lddq
                        : (pl == 0 ? p.dv + (long long)b * p.bsdv + (long long)row * p.lddv
                                   : p.dk + (long long)b * p.bsdk + (long long)row * p.lddk);
    *reinterpret_cast<uint4*>(dst + h * D + cg * 8) = w;
  }
}

// Persistent-schedule decision shared by vn_attention_bwd and vn_attention_bwd_workspace_bytes.
void bwd_split(int nb, int heads, int nq, int nk, int has_dq, int* n_left, int* parts) {
  *n_left = 0; *parts = 0;
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("VN_ATTN_SPLIT"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int ktiles = vn_cdiv(nk, T), qtiles = vn_cdiv(nq, T);
  const int items = ktiles * heads * nb + (has_dq ? qtiles * heads * nb : 0);
  const int left = items % sms;
  const int tiles = ktiles < qtiles ? ktiles : qtiles;
  if (items <= sms || left == 0 || left > sms / 2 || tiles < 4 || items / sms + 1 > kMaxSegments) return;
  int q = sms / left;
  if (q > tiles / 2) q = tiles / 2;               // at least two tiles per part
  if (q < 2) return;
  *n_left = left; *parts = q;
}

// delta[b,h,n] = sum_d dO * O : one 16-byte vector per lane, 8 lanes per (row, head)
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, long long ldo, long long bso,
                                                         const bf16* __restrict__ d_o, long long lddo, long long bsdo,
                                                         float* __restrict__ delta, int nb, int heads, int nq) {
  pdl_trigger();
  pdl_wait();
  const long long nvec = (long long)nb * nq * heads * 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f;
  long long rh = 0;
  const bool ok = i < nvec;
  if (ok) {
    rh = i >> 3;
    const int h = (int)(rh % heads);
    const long long bn = rh / heads;
    const int n = (int)(bn % nq), b = (int)(bn / nq);
    const int c = h * D + (int)(i & 7) * 8;
    const uint4 ov = *reinterpret_cast<const uint4*>(o + (long long)b * bso + (long long)n * ldo + c);
    const uint4 dv = *reinterpret_cast<const uint4*>(d_o + (long long)b * bsdo + (long long)n * lddo + c);
    float2 a, g;
    a = unpack_bf162(ov.x); g = unpack_bf162(dv.x); s = fmaf(a.x, g.x, fmaf(a.y, g.y, s));
    a = unpack_bf162(ov.y); g = unpack_bf162(dv.y); s = fmaf(a.x, g.x, fmaf(a.y, g.y, s));
    a = unpack_bf162(ov.z); g = unpack_bf162(dv.z); s = fmaf(a.x, g.x, fmaf(a.y, g.y, s));
    a = unpack_bf162(ov.w); g = unpack_bf162(dv.w); s = fmaf(a.x, g.x, fmaf(a.y, g.y, s));
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && (threadIdx.x & 7) == 0) {
    const int h = (int)(rh % heads);
    const long long bn = rh / heads;
    const int n = (int)(bn % nq), b = (int)(bn / nq);
    delta[((long long)b * heads + h) * nq + n] = s;
  }
}

// fp64 scratch -> bf16 dk / dv, leaving the scratch zeroed for the next launch
__global__ void __launch_bounds__(256) attn_dkv_finish_kernel(const BwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int C = p.heads * D;
  const long long per = (long long)p.nb * p.nk * C;
  const long long total = per / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 2;
    const int c = (int)(e % C);
    const long long r = e / C;
    const int n = (int)(r % p.nk), b = (int)(r / p.nk);
    double2* ak = reinterpret_cast<double2*>(p.dkv_acc + e);
    double2* av = reinterpret_cast<double2*>(p.dkv_acc + per + e);
    const double2 a = *ak, v = *av;
    *reinterpret_cast<bf162*>(p.dk + (long long)b * p.bsdk + (long long)n * p.lddk + c) = __floats2bfloat162_rn((float)a.x, (float)a.y);
    *reinterpret_cast<bf162*>(p.dv + (long long)b * p.bsdv + (long long)n * p.lddv + c) = __floats2bfloat162_rn((float)v.x, (float)v.y);
    *ak = make_double2(0.0, 0.0);
    *av = make_double2(0.0, 0.0);
  }
}

int qkv_map(CUtensorMap* m, const void* base, int width, int rows, int nb, long long ld, long long bs) {
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)nb};
  cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(nb > 1 ? bs : ld * rows) * 2};
  cuuint32_t box[3] = {64, 128, 1};
  return vn_make_map(m, base, 3, dims, str, box);
}

}  // namespace

extern "C" int vn_attention_bwd(const vn_attn_desc* d, vn_stream_t s) {
  cudaStream_t st = (cudaStream_t)s;
  VN_CHECK(d != nullptr, "attention bwd: null descriptor");
  VN_CHECK(d->nb > 0 && d->heads > 0 && d->nq > 0 && d->nk > 0, "attention bwd: empty problem");
  VN_CHECK(d->lse && d->delta && d->d_o && d->dk && d->dv, "attention bwd: lse, delta, d_o, dk, dv are required");
  VN_CHECK(d->ldq % 8 == 0 && d->ldk % 8 == 0 && d->ldv % 8 == 0 && d->ldo % 8 == 0 && d->lddo % 8 == 0 && d->bsq % 8 == 0 &&
               d->bsk % 8 == 0 && d->bsv % 8 == 0 && d->bso % 8 == 0 && d->bsdo % 8 == 0 && d->lddk % 8 == 0 &&
               d->lddv % 8 == 0 && d->bsdk % 8 == 0 && d->bsdv % 8 == 0 && (!d->dq || (d->lddq % 8 == 0 && d->bsdq % 8 == 0)),
           "attention bwd: strides must be multiples of 8 elements");
  VN_CHECK(((reinterpret_cast<uintptr_t>(d->q) | reinterpret_cast<uintptr_t>(d->k) | reinterpret_cast<uintptr_t>(d->v) |
             reinterpret_cast<uintptr_t>(d->o) | reinterpret_cast<uintptr_t>(d->d_o) | reinterpret_cast<uintptr_t>(d->dq) |
             reinterpret_cast<uintptr_t>(d->dk) | reinterpret_cast<uintptr_t>(d->dv) | reinterpret_cast<uintptr_t>(d->lse) |
             reinterpret_cast<uintptr_t>(d->delta)) & 15) == 0, "attention bwd: tensors must be 16-byte aligned");
  const int width = d->heads * D;
  CUtensorMap tq, tk, tv, tdo;
  if (qkv_map(&tq, d->q, width, d->nq, d->nb, d->ldq, d->bsq)) return -1;
  if (qkv_map(&tk, d->k, width, d->nk, d->nb, d->ldk, d->bsk)) return -1;
  if (qkv_map(&tv, d->v, width, d->nk, d->nb, d->ldv, d->bsv)) return -1;
  if (qkv_map(&tdo, d->d_o, width, d->nq, d->nb, d->lddo, d->bsdo)) return -1;
  BwdParams p{};
  p.dbg = vn_debug_buffer();
  p.nb = d->nb; p.heads = d->heads; p.nq = d->nq; p.nk = d->nk; p.scale = d->scale;
  p.lse = d->lse; p.delta = d->delta;
  p.dq = (bf16*)d->dq; p.lddq = d->lddq; p.bsdq = d->bsdq;
  p.dk = (bf16*)d->dk; p.lddk = d->lddk; p.bsdk = d->bsdk;
  p.dv = (bf16*)d->dv; p.lddv = d->lddv; p.bsdv = d->bsdv;
  p.dkv_acc = d->dkv_acc;
  p.causal = d->causal;
  constexpr int SMEM = DKV_SMEM > DQ_SMEM ? DKV_SMEM : DQ_SMEM;
  static bool configured = false;
  if (!configured) {
    VN_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    VN_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  VN_LAUNCH(attn_delta_kernel, (unsigned)vn_cdiv64((long long)d->nb * d->nq * d->heads * 8, 256), 256, 0, st, 
      (const bf16*)d->o, d->ldo, d->bso, (const bf16*)d->d_o, d->lddo, d->bsdo, d->delta, d->nb, d->heads, d->nq);
  const int ktiles = vn_cdiv(d->nk, T), qtiles = vn_cdiv(d->nq, T);
  const long long base_ctas = (long long)ktiles * d->heads * d->nb;
  int splits = 1;
  if (d->dkv_acc && base_ctas < 148) {
    splits = (int)((148 + base_ctas - 1) / base_ctas);
    if (splits > qtiles) splits = qtiles;
    if (splits < 1) splits = 1;
  }
  p.qtiles_per_split = vn_cdiv(qtiles, splits);
  splits = vn_cdiv(qtiles, p.qtiles_per_split);
  p.qsplits = splits;
  p.ktiles = ktiles; p.qtiles = qtiles;
  p.n_dkv = ktiles * d->heads * d->nb * splits;
  const int n_dq = d->dq ? qtiles * d->heads * d->nb : 0;
  unsigned grid = (unsigned)(p.n_dkv + n_dq);
  if (splits == 1 && base_ctas >= 148) {
    int n_left = 0, parts = 0;
    bwd_split(d->nb, d->heads, d->nq, d->nk, d->dq != nullptr, &n_left, &parts);
    const size_t need = (size_t)n_left * parts * 2 * T * D * sizeof(float);
    if (parts > 0 && d->ws != nullptr && (size_t)d->ws_bytes >= need) {
      int sms = 0, dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      p.persistent = 1;
      p.items = (int)grid;
      p.rounds = p.items / sms;
      p.n_left = n_left; p.parts = parts;
      p.left_dkv = n_dq > 0 ? n_left / 2 : n_left;
      if (p.left_dkv > p.n_dkv) p.left_dkv = p.n_dkv;
      if (n_left - p.left_dkv > n_dq) p.left_dkv = n_left - n_dq;
      p.part = reinterpret_cast<float*>(d->ws);
      grid = (unsigned)sms;
    }
  }
  if (p.persistent) {
    VN_LAUNCH(attn_bwd_tc_kernel<false>, grid, kThreads, SMEM, st, tq, tk, tv, tdo, p);
    VN_LAUNCH(attn_bwd_fixup_kernel, p.n_left * 8, 128, 0, st, p);
  } else if (splits > 1) {
    VN_LAUNCH(attn_bwd_tc_kernel<true>, grid, kThreads, SMEM, st, tq, tk, tv, tdo, p);
    if (!d->defer_dkv_finish) {
      const long long pairs = (long long)d->nb * d->nk * d->heads * D / 2;
      int blocks = (int)vn_cdiv64(pairs, 256);
      if (blocks > 148 * 8) blocks = 148 * 8;
      VN_LAUNCH(attn_dkv_finish_kernel, blocks, 256, 0, st, p);
    }
  } else {
    VN_LAUNCH(attn_bwd_tc_kernel<false>, grid, kThreads, SMEM, st, tq, tk, tv, tdo, p);
  }
  return 0;
}

extern "C" size_t vn_attention_bwd_workspace_bytes(int nb, int heads, int nq, int nk, int has_dq) {
  int n_left = 0, parts = 0;
  bwd_split(nb, heads, nq, nk, has_dq, &n_left, &parts);
  return (size_t)n_left * parts * 2 * T * D * sizeof(float);
}

extern "C" int vn_attention_dkv_finish(const vn_attn_desc* d, vn_stream_t s) {
  VN_CHECK(d != nullptr, "attention dkv finish: null descriptor");
  VN_CHECK(d->dkv_acc != nullptr && d->dk != nullptr && d->dv != nullptr, "attention dkv finish: dkv_acc, dk and dv are required");
  cudaStream_t st = (cudaStream_t)s;
  // the same query-split decision as vn_attention_bwd: without a split dK / dV were written directly
  const int ktiles = vn_cdiv(d->nk, T), qtiles = vn_cdiv(d->nq, T);
  const long long base_ctas = (long long)ktiles * d->heads * d->nb;
  if (base_ctas >= 148) return 0;
  int splits = (int)((148 + base_ctas - 1) / base_ctas);
  if (splits > qtiles) splits = qtiles;
  if (splits < 1) splits = 1;
  splits = vn_cdiv(qtiles, vn_cdiv(qtiles, splits));
  if (splits <= 1) return 0;
  BwdParams p{};
  p.nb = d->nb; p.heads = d->heads; p.nq = d->nq; p.nk = d->nk;
  p.dk = (bf16*)d->dk; p.lddk = d->lddk; p.bsdk = d->bsdk;
  p.dv = (bf16*)d->dv; p.lddv = d->lddv; p.bsdv = d->bsdv;
  p.dkv_acc = d->dkv_acc;
  const long long pairs = (long long)d->nb * d->nk * d->heads * D / 2;
  int blocks = (int)vn_cdiv64(pairs, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  VN_LAUNCH(attn_dkv_finish_kernel, blocks, 256, 0, st, p);
  return 0;
}
