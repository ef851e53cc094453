// vn_attn_tc.cu — attention forward on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), head_dim 64.
//
// Same contract as the reference's attention core (models/xti_attention_processor.py:44-50: head split, fp32 logits
// with alpha = scale, softmax, bmm, head merge), with K and V taken from different tensors (XTI).
//
// Work item = one (128-query tile, head, image); one CTA per SM.  When the items do not fill whole waves (64x64 latents,
// B = 1: 160 items on 148 SMs = two waves, the second one 8 % full) the launch is PERSISTENT over #SMs CTAs: every CTA
// runs items / #SMs whole items and the items % #SMs leftover ones are split along the KEYS into equal parts that are
// spread over all CTAs; a part leaves an unnormalised partial (O, m, l) in a caller-provided scratch and
// attn_fwd_fixup_kernel merges the parts of each leftover item (log-sum-exp) - 2 waves become ~1.1.
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

#include <stdlib.h>

namespace {
// P goes back into TENSOR memory (packed bf16 pairs over the first half of the S columns it was computed from) and feeds
// P V as the A operand from there: no swizzled shared-memory stores / proxy fence in the softmax threads, and only the V tile
// crosses the shared-memory port during the product.  VN_ATTN_FWD_TS=0 keeps the shared-memory operand (cross-check).
#ifndef VN_ATTN_FWD_TS
#define VN_ATTN_FWD_TS 1
#endif
constexpr bool kFwdTS = VN_ATTN_FWD_TS != 0;
// OPTIONAL schedule (VN_ATTN_FWD_HALVES=1, off by default): the two 64-key halves of a tile as independent chains
// (S_h -> softmax_h -> P_h V) with their own S / P / PV barriers and S issued as two N = 64 products, so that the softmax warps of
// the two halves may drift apart.  Built to test the hypothesis that the two warps of a scheduler lose time by running the same
// phase at the same time; measured NEUTRAL on B200 (64x64: 70.8 vs 70.2 us eager), as were 4 K/V stages and one barrier arrival
// per warp instead of per thread.  Timing ablations (profiles/r2_ncu_attn_summary.txt) and the in-kernel timeline
// (profiles/r2_attn_timeline.txt): the two softmax warps of a scheduler need ~2 080 cycles per tile (1 024 of them MUFU) and are
// the bottleneck; with them idle the barrier round trips between the roles still take 31 of the 49 us.
#ifndef VN_ATTN_FWD_HALVES
#define VN_ATTN_FWD_HALVES 0
#endif
constexpr bool kHalves = (VN_ATTN_FWD_HALVES != 0) && kFwdTS;


constexpr int D = 64;
constexpr int BQ = 128;           // queries per CTA
constexpr int BKV = 128;          // keys per iteration
#ifndef VN_ATTN_KV_STAGES
#define VN_ATTN_KV_STAGES 3
#endif
constexpr int KV_STAGES = VN_ATTN_KV_STAGES;
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
  // work decomposition: CTA c runs the whole items c, c + grid, ... (`rounds` of them) and then, if c < n_left * parts,
  // key part c % parts of leftover item (items - n_left + c / parts)
  int nqt, items, rounds, n_left, parts;
  float* part_o;        // [n_left * parts][128][64] unnormalised partial outputs
  float* part_ml;       // [n_left * parts][128][2]  (running max in the log2 domain, running sum)
  long long* dbg;       // -DVN_TIMELINE builds: clock stamps of CTA 0 (scripts/attn_timeline.py), else unused
};

// In-kernel timeline (instrumented builds only): CTA 0 stamps clock64 for iterations [8, 16) of its first segment.
//   slot layout: dbg[(j - 8) * 16 + e]; e = 0 S(j) issued, 1 P V(j) issued (p_full(j) seen), 2 s_full(j) seen by a softmax
//   thread, 3 that thread's arrive on p_full(j), 4 pv_full(j-1) seen by it, 5 K/V stage of tile j free (producer), 6 its TMA issued,
//   7 kv_full(j) seen by the MMA warp
#ifdef VN_TIMELINE
#define VN_ASTAMP(j, e)                                                                                  \
  do {                                                                                                   \
    if (p.dbg && blockIdx.x == 0 && (j) >= 8 && (j) < 16) p.dbg[((j) - 8) * 16 + (e)] = clock64();       \
  } while (0)
#else
#define VN_ASTAMP(j, e) do { } while (0)
#endif

struct Segment {
  int q0, h, b;         // query-tile origin, head, image
  int kb0, kb1;         // key blocks [kb0, kb1)
  int slot;             // partial slot, -1 for a whole item
};
__device__ __forceinline__ int num_segments(const FwdParams& p) {
  return p.rounds + ((int)blockIdx.x < p.n_left * p.parts ? 1 : 0);
}
__device__ __forceinline__ Segment segment(const FwdParams& p, int seg, int nt) {
  Segment g;
  int item;
  if (seg < p.rounds) {
    item = (int)blockIdx.x + seg * (int)gridDim.x;
    g.kb0 = 0; g.kb1 = nt; g.slot = -1;
  } else {
    const int li = (int)blockIdx.x / p.parts, part = (int)blockIdx.x - li * p.parts;
    item = p.items - p.n_left + li;
    g.kb0 = part * nt / p.parts; g.kb1 = (part + 1) * nt / p.parts; g.slot = (int)blockIdx.x;
  }
  const int qt = item % p.nqt, hb = item / p.nqt;
  g.q0 = qt * 128; g.h = hb % p.heads; g.b = hb / p.heads;
  return g;
}

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
  uint64_t* s_full = kv_empty + KV_STAGES;                 // [2] (kHalves: [half][2], index h * 2 + buffer)
  uint64_t* p_full = s_full + 4;                           // [2] / [half][2]
  uint64_t* pv_full = p_full + 4;                          // [2] / [half][2]
  uint64_t* q_empty = pv_full + 4;                         // every S product of a segment has read the Q tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (p.nk + BKV - 1) / BKV;
  const int nseg = num_segments(p);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < (kHalves ? 4 : 2); ++s) {
      mbar_init(&s_full[s], 1); mbar_init(&pv_full[s], 1);
      mbar_init(&p_full[s], kHalves ? 4 : 8);      // one arrival per softmax WARP
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    {                                              // warp-uniform control flow, one elected lane issues
      const bool leader = elect_one();
      int it = 0;                                  // iterations issued so far, over all segments (barrier phases run on)
      for (int seg = 0; seg < nseg; ++seg) {
        const Segment g = segment(p, seg, nt);
        if (seg > 0) mbar_wait(q_empty, (seg - 1) & 1);      // the previous segment's S products are done with sQ
        if (leader) {
          mbar_expect_tx(q_full, TILE_BYTES);
          tma_load_3d(sQ, &tmQ, q_full, g.h * D, g.q0, g.b);
        }
        for (int kb = g.kb0; kb < g.kb1; ++kb, ++it) {
          const int s = it % KV_STAGES;
          mbar_wait(&kv_empty[s], ((it / KV_STAGES) & 1) ^ 1);
          if (leader) {
            VN_ASTAMP(it, 5);
            mbar_expect_tx(&kv_full[s], 2 * TILE_BYTES);
            tma_load_3d(sKV + s * 2 * TILE_BYTES, &tmK, &kv_full[s], g.h * D, kb * BKV, g.b);
            tma_load_3d(sKV + s * 2 * TILE_BYTES + TILE_BYTES, &tmV, &kv_full[s], g.h * D, kb * BKV, g.b);
            VN_ASTAMP(it, 6);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {                                              // warp-uniform control flow, one elected lane issues (vn_common.cuh: elect_one)
      const bool leader = elect_one();
      constexpr uint32_t idesc_s = idesc_bf16(BQ, BKV, 0);
      constexpr uint32_t idesc_pv = idesc_bf16(BQ, D, 1);
      const uint32_t aQ = smem_u32(sQ);
      auto issue_s = [&](int j, bool last_of_segment) {
        // S[j&1] is free: the softmax threads arrived on p_full(j-2) before PV(j-2) was issued (program order below)
        const int s = j % KV_STAGES;
        mbar_wait(&kv_full[s], (j / KV_STAGES) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sKV + s * 2 * TILE_BYTES);
        const uint32_t tS = tmem_base + (uint32_t)((j & 1) * BKV);
        if (leader) {
          VN_ASTAMP(j, 7);
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_bf16(tS, umma_desc_k_sw128(aQ) + (uint64_t)(k * 2), umma_desc_k_sw128(aK) + (uint64_t)(k * 2), idesc_s,
                      k ? 1u : 0u);
          umma_commit(&s_full[j & 1]);
          VN_ASTAMP(j, 0);
          if (last_of_segment) umma_commit(q_empty);
        }
      };
      if constexpr (kHalves) {
        // two independent chains per tile (key halves h = 0 / 1), served in the fixed order h0(j), h1(j), h0(j+1), ...
        constexpr uint32_t idesc_s64 = idesc_bf16(BQ, 64, 0);
        auto wait_kv = [&](int j) {
          mbar_wait(&kv_full[j % KV_STAGES], (j / KV_STAGES) & 1);
          tc_fence_after();
        };
        auto issue_s_half = [&](int j, int h) {      // S_h(j) = Q K_h(j)^T -> S[j & 1] columns [64 h, 64 h + 64)
          const uint32_t aK = smem_u32(sKV + (j % KV_STAGES) * 2 * TILE_BYTES) + (uint32_t)(h * 64 * 128);
          const uint32_t tS = tmem_base + (uint32_t)((j & 1) * BKV + h * 64);
          if (leader) {
#ifndef VN_ABL_NOSMMA
#pragma unroll
            for (int k = 0; k < D / 16; ++k)
              umma_bf16(tS, umma_desc_k_sw128(aQ) + (uint64_t)(k * 2), umma_desc_k_sw128(aK) + (uint64_t)(k * 2), idesc_s64,
                        k ? 1u : 0u);
#endif
            umma_commit(&s_full[h * 2 + (j & 1)]);
          }
        };
        int it0 = 0;
        for (int seg = 0; seg < nseg; ++seg) {
          const Segment g = segment(p, seg, nt);
          const int n = g.kb1 - g.kb0;
          mbar_wait(q_full, seg & 1);
          wait_kv(it0);
          issue_s_half(it0, 0); issue_s_half(it0, 1);
          if (n > 1) {
            wait_kv(it0 + 1);
            issue_s_half(it0 + 1, 0); issue_s_half(it0 + 1, 1);
          }
          if (n <= 2 && leader) umma_commit(q_empty);          // every S product of the segment has been issued
          for (int jj = 0; jj < n; ++jj) {
            const int j = it0 + jj;
            const int s = j % KV_STAGES;
            const uint32_t aV = smem_u32(sKV + s * 2 * TILE_BYTES + TILE_BYTES);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              mbar_wait(&p_full[h * 2 + (j & 1)], (j >> 1) & 1);
              tc_fence_after();
              if (leader) {
#ifndef VN_ABL_NOPVMMA
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const uint64_t bdesc = umma_desc_mn_sw128(aV + (h * 4 + kk) * 2048);
                  const uint32_t tP = tmem_base + (uint32_t)((j & 1) * BKV + h * 64 + kk * 8);
                  umma_bf16_ts(tmem_base + 2 * BKV + (uint32_t)(((j & 1) * 2 + h) * D), tP, bdesc, idesc_pv, kk ? 1u : 0u);
                }
#endif
                umma_commit(&pv_full[h * 2 + (j & 1)]);
                if (h == 1) umma_commit(&kv_empty[s]);         // K(j) and V(j) are no longer needed
              }
              if (jj + 2 < n) {                                // S_h(j+2) takes over the columns P_h(j) has just been read from
                if (h == 0) wait_kv(j + 2);
                issue_s_half(j + 2, h);
                if (h == 1 && jj + 3 == n && leader) umma_commit(q_empty);
              }
            }
          }
          it0 += n;
        }
      } else {
      int it0 = 0;
      for (int seg = 0; seg < nseg; ++seg) {
      const Segment g = segment(p, seg, nt);
      const int n = g.kb1 - g.kb0;
      mbar_wait(q_full, seg & 1);
      issue_s(it0, n == 1);
      for (int jj = 0; jj < n; ++jj) {
        const int j = it0 + jj;                    // iteration index over all segments: buffer parities and phases
        if (jj + 1 < n) issue_s(j + 1, jj + 2 == n);
        // PV(j) = P(j) V(j)
        const int s = j % KV_STAGES;
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aP = smem_u32(sP + (j & 1) * 2 * TILE_BYTES);
        const uint32_t aV = smem_u32(sKV + s * 2 * TILE_BYTES + TILE_BYTES);
        const uint32_t tPV = tmem_base + 2 * BKV + (uint32_t)((j & 1) * 2 * D);
        if (leader) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            // keys [0,64) accumulate into PV[j&1][0], keys [64,128) into PV[j&1][1]
            const uint64_t bdesc = umma_desc_mn_sw128(aV + k * 2048);
            if (kFwdTS) {
              // P of key block (k >> 2): 32 packed columns at the start of that block's 64 S columns, 8 columns per k-step
              const uint32_t tP = tmem_base + (uint32_t)((j & 1) * BKV + (k >> 2) * 64 + (k & 3) * 8);
              umma_bf16_ts(tPV + (uint32_t)((k >> 2) * D), tP, bdesc, idesc_pv, (k & 3) ? 1u : 0u);
            } else {
              const uint64_t adesc = umma_desc_k_sw128(aP + (k >> 2) * TILE_BYTES) + (uint64_t)((k & 3) * 2);
              umma_bf16(tPV + (uint32_t)((k >> 2) * D), adesc, bdesc, idesc_pv, (k & 3) ? 1u : 0u);
            }
          }
          umma_commit(&kv_empty[s]);         // K(j) (read by S(j), issued earlier) and V(j) are no longer needed
          umma_commit(&pv_full[j & 1]);
          VN_ASTAMP(j, 1);
        }
      }
      it0 += n;
      }
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
    int it0 = 0;
    for (int seg = 0; seg < nseg; ++seg) {
    const Segment g = segment(p, seg, nt);
    const int q0 = g.q0, h = g.h, b = g.b;
    const int nloc = g.kb1 - g.kb0;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;

    auto accumulate_pv = [&](int jj, float alpha) {   // O = alpha * O + PV(jj)[hf]
      mbar_wait(&pv_full[(kHalves ? hf * 2 : 0) + (jj & 1)], (jj >> 1) & 1);
      tc_fence_after();
      const uint32_t tPV = tmem_base + 2 * BKV + (uint32_t)(((jj & 1) * 2 + hf) * D) + lane_addr;
      uint32_t r0[32], r1[32];
#ifdef VN_ABL_NOPV
#pragma unroll
      for (int i = 0; i < 32; ++i) { r0[i] = 0u; r1[i] = 0u; }                  // timing ablation: no PV read-back
#else
      tmem_ld32(tPV, r0);
      tmem_ld32(tPV + 32, r1);
      tmem_ld_wait();
#endif
      const float2 av = make_float2(alpha, alpha);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {                  // packed f32x2 arithmetic (sm_100)
        const float2 a = __ffma2_rn(make_float2(o[i], o[i + 1]), av, make_float2(__uint_as_float(r0[i]), __uint_as_float(r0[i + 1])));
        const float2 c = __ffma2_rn(make_float2(o[32 + i], o[33 + i]), av, make_float2(__uint_as_float(r1[i]), __uint_as_float(r1[i + 1])));
        o[i] = a.x; o[i + 1] = a.y; o[32 + i] = c.x; o[33 + i] = c.y;
      }
    };

    for (int jj = 0; jj < nloc; ++jj) {
      const int j = it0 + jj;                      // iteration index over all segments (buffer parities, phases)
      const int kb = g.kb0 + jj;                   // key block
      mbar_wait(&s_full[(kHalves ? hf * 2 : 0) + (j & 1)], (j >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 64) VN_ASTAMP(j, 2);
      const uint32_t tS = tmem_base + (uint32_t)((j & 1) * BKV + hf * 64) + lane_addr;
#ifdef VN_ABL_NOSOFTMAX
      tc_fence_before();                           // timing ablation: the softmax threads only pass the barriers on
      __syncwarp();                      // every lane's tensor-memory stores are complete and fenced
      if (lane == 0) mbar_arrive(&p_full[(kHalves ? hf * 2 : 0) + (j & 1)]);      // one arrival per warp: 256 arrivals on one
                                                                                  // shared-memory word serialise
      __syncwarp();                      // reconverge: the next tcgen05.ld is .sync.aligned
      if (jj > 0) { mbar_wait(&pv_full[(kHalves ? hf * 2 : 0) + ((j - 1) & 1)], ((j - 1) >> 1) & 1); tc_fence_after(); }
      alpha_prev = 1.f;
      continue;
#endif
      uint32_t s[64];
      {
        uint32_t(&c0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&c1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        tmem_ld32(tS, c0); tmem_ld32(tS + 32, c1);
        tmem_ld_wait();
      }
      int kvalid = p.nk - kb * BKV - hf * 64;          // keys of this half-tile that exist (may be <= 0 on the last tile)
      if (p.causal) kvalid = min(kvalid, q0 + r + 1 - kb * BKV - hf * 64);  // ... and that query row q0 + r may see
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
#ifdef VN_ABL_NOMAX
      float mx = fmaxf(m_run, __uint_as_float(s[0]) * sl2);                     // timing ablation: no row maximum
#else
      float mx = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2);   // scale > 0
#endif
      // a half that has not seen a valid key yet keeps m = -inf; use 0 as the reference so that ex2(-inf - 0) = 0
      const float mref = (mx == -INFINITY) ? 0.f : mx;
      const float alpha = ex2_approx(m_run - mref);    // first tile: ex2(-inf) = 0
      m_run = mx;
      float2 rsa = make_float2(0.f, 0.f), rsb = make_float2(0.f, 0.f);
      const float2 sl2v = make_float2(sl2, sl2), nm = make_float2(-mref, -mref);
      uint8_t* prow = sP + ((j & 1) * 2 + hf) * TILE_BYTES + r * 128;
      uint32_t pw[32];                   // packed bf16 pairs of my 64 probabilities (tensor-memory operand)
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[g * 8 + i]), __uint_as_float(s[g * 8 + i + 1])), sl2v, nm);
#ifdef VN_ABL_NOEXP
          e[i] = fmaf(x.x, 0.01f, 1.f); e[i + 1] = fmaf(x.y, 0.01f, 1.f);      // timing ablation: no MUFU
#else
          e[i] = ex2_approx(x.x); e[i + 1] = ex2_approx(x.y);
#endif
        }
        rsa = __fadd2_rn(rsa, __fadd2_rn(make_float2(e[0], e[1]), make_float2(e[2], e[3])));
        rsb = __fadd2_rn(rsb, __fadd2_rn(make_float2(e[4], e[5]), make_float2(e[6], e[7])));
        if (kFwdTS) {
          pw[g * 4 + 0] = pack_bf162(e[0], e[1]); pw[g * 4 + 1] = pack_bf162(e[2], e[3]);
          pw[g * 4 + 2] = pack_bf162(e[4], e[5]); pw[g * 4 + 3] = pack_bf162(e[6], e[7]);
        } else {
          uint4 w;
          w.x = pack_bf162(e[0], e[1]); w.y = pack_bf162(e[2], e[3]);
          w.z = pack_bf162(e[4], e[5]); w.w = pack_bf162(e[6], e[7]);
          *reinterpret_cast<uint4*>(prow + ((g ^ (r & 7)) << 4)) = w;
        }
      }
      l_run = l_run * alpha + ((rsa.x + rsa.y) + (rsb.x + rsb.y));
      if (kFwdTS) {
        tmem_st32(tS, pw);               // over the first 32 of my own 64 S columns (all of them are in registers by now)
        tmem_st_wait();
        tc_fence_before();
      } else {
        tc_fence_before();               // my TMEM reads of S(j) are done
        fence_async_smem();              // my P writes are visible to the tensor core (async proxy)
      }
      __syncwarp();                      // every lane's tensor-memory stores are complete and fenced
      if (lane == 0) mbar_arrive(&p_full[(kHalves ? hf * 2 : 0) + (j & 1)]);      // one arrival per warp: 256 arrivals on one
                                                                                  // shared-memory word serialise
      __syncwarp();                      // reconverge: the next tcgen05.ld is .sync.aligned
      if (threadIdx.x == 64) VN_ASTAMP(j, 3);
      if (jj > 0) accumulate_pv(j - 1, alpha_prev);    // overlaps PV(j) / S(j+1) on the tensor pipe
      if (threadIdx.x == 64) VN_ASTAMP(j, 4);
      alpha_prev = alpha;
    }
    accumulate_pv(it0 + nloc - 1, alpha_prev);
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
    if (hf == 0 && g.slot >= 0) {
      // key part of a leftover item: park the merged, UNNORMALISED (O, m, l) of this row; attn_fwd_fixup_kernel finishes
      const float m1 = xch[r * 67 + 64], l1 = xch[r * 67 + 65];
      const float m = fmaxf(m_run, m1);                // -inf if this part holds no key the row may see (causal)
      const float a0 = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m), a1 = (m1 == -INFINITY) ? 0.f : ex2_approx(m1 - m);
      float* po = p.part_o + ((long long)g.slot * BQ + r) * D;
#pragma unroll
      for (int i = 0; i < D; i += 4)
        *reinterpret_cast<float4*>(po + i) = make_float4(o[i] * a0 + xch[r * 67 + i] * a1, o[i + 1] * a0 + xch[r * 67 + i + 1] * a1,
                                                         o[i + 2] * a0 + xch[r * 67 + i + 2] * a1, o[i + 3] * a0 + xch[r * 67 + i + 3] * a1);
      *reinterpret_cast<float2*>(p.part_ml + ((long long)g.slot * BQ + r) * 2) = make_float2(m, l_run * a0 + l1 * a1);
    } else if (hf == 0 && row < p.nq) {
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
    it0 += nloc;
    if (seg + 1 < nseg) asm volatile("bar.sync 1, 256;" ::: "memory");   // xch aliases the P tiles of the next segment
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// =================================================================================================
// Forward with SIXTEEN softmax warps (four per scheduler).  The in-kernel timeline of the kernel above shows its two softmax
// warps per scheduler as the bottleneck (profiles/r2_attn_timeline.txt: 2 080 cycles per 128-key tile, 1 024 of them MUFU time,
// the rest latencies two warps cannot overlap).  Here a thread owns (query row, 32-key QUARTER of every tile):
//   * O is accumulated by the tensor core itself: four accumulators O_c [128 x 64] in tensor memory, one per key quarter
//     (2 x 128 columns of S + 4 x 64 = all 512), P_c V_c issued with accumulate; no per-tile read-back, no 64 accumulator registers;
//   * every quarter keeps its OWN running maximum, and only moves it when a row maximum grows by more than 2^8 (probabilities then
//     stay below 256, exact in bf16's range; sums are fp32): the rescale of O_c - tcgen05.ld, multiply, tcgen05.st, warp-voted
//     because the tensor-memory accesses are warp-collective - is rare after the first tiles;
//   * the four quarters of a row are merged once at the end, like the two halves above.
// Same work decomposition (persistent CTAs, key parts of leftover items, attn_fwd_fixup_kernel) and the same results up to the
// rounding of P against a different reference maximum.
// =================================================================================================
constexpr int kThreads16 = 64 + 512;
constexpr int XCH16_FLOATS = 3 * BQ * 67;                         // quarters 1..3 park (O, m, l) per row, odd stride
constexpr int SMEM16_BYTES = TILE_BYTES + KV_STAGES * 2 * TILE_BYTES + XCH16_FLOATS * 4 + 256 + 1024;
static_assert(SMEM16_BYTES <= 232448, "shared memory budget");
constexpr float kRescaleThreshold = 8.f;                          // log2 domain

__global__ void __launch_bounds__(kThreads16, 1) attn_fwd16_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                       const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmV,
                                                                       const FwdParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + TILE_BYTES;                          // stage s: K at s*2*TILE, V right after
  float* xch = reinterpret_cast<float*>(sKV + KV_STAGES * 2 * TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + XCH16_FLOATS);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                            // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES;                // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;                 // [2]
  uint64_t* p_full = s_full + 2;                           // [2], one arrival per softmax warp
  uint64_t* pv_full = p_full + 2;                          // [2], alternating per tile
  uint64_t* q_empty = pv_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (p.nk + BKV - 1) / BKV;
  const int nseg = num_segments(p);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 16); mbar_init(&pv_full[s], 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    const bool leader = elect_one();
    int it = 0;
    for (int seg = 0; seg < nseg; ++seg) {
      const Segment g = segment(p, seg, nt);
      if (seg > 0) mbar_wait(q_empty, (seg - 1) & 1);
      if (leader) {
        mbar_expect_tx(q_full, TILE_BYTES);
        tma_load_3d(sQ, &tmQ, q_full, g.h * D, g.q0, g.b);
      }
      for (int kb = g.kb0; kb < g.kb1; ++kb, ++it) {
        const int s = it % KV_STAGES;
        mbar_wait(&kv_empty[s], ((it / KV_STAGES) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&kv_full[s], 2 * TILE_BYTES);
          tma_load_3d(sKV + s * 2 * TILE_BYTES, &tmK, &kv_full[s], g.h * D, kb * BKV, g.b);
          tma_load_3d(sKV + s * 2 * TILE_BYTES + TILE_BYTES, &tmV, &kv_full[s], g.h * D, kb * BKV, g.b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = idesc_bf16(BQ, BKV, 0);
    constexpr uint32_t idesc_pv = idesc_bf16(BQ, D, 1);
    const uint32_t aQ = smem_u32(sQ);
    auto issue_s = [&](int j, bool last_of_segment) {
      const int s = j % KV_STAGES;
      mbar_wait(&kv_full[s], (j / KV_STAGES) & 1);
      tc_fence_after();
      const uint32_t aK = smem_u32(sKV + s * 2 * TILE_BYTES);
      const uint32_t tS = tmem_base + (uint32_t)((j & 1) * BKV);
      if (leader) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_bf16(tS, umma_desc_k_sw128(aQ) + (uint64_t)(k * 2), umma_desc_k_sw128(aK) + (uint64_t)(k * 2), idesc_s,
                    k ? 1u : 0u);
        umma_commit(&s_full[j & 1]);
        if (last_of_segment) umma_commit(q_empty);
      }
    };
    int it0 = 0;
    for (int seg = 0; seg < nseg; ++seg) {
      const Segment g = segment(p, seg, nt);
      const int n = g.kb1 - g.kb0;
      mbar_wait(q_full, seg & 1);
      issue_s(it0, n == 1);
      for (int jj = 0; jj < n; ++jj) {
        const int j = it0 + jj;
        if (jj + 1 < n) issue_s(j + 1, jj + 2 == n);
        const int s = j % KV_STAGES;
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);       // P(j) in tensor memory, O rescaled where a maximum moved
        tc_fence_after();
        const uint32_t aV = smem_u32(sKV + s * 2 * TILE_BYTES + TILE_BYTES);
        if (leader) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              // O_c (+)= P_c[128 x 16 keys] V[16 keys x 64]: P_c = 16 packed columns at the start of quarter c's 32 S columns
              const uint64_t bdesc = umma_desc_mn_sw128(aV + (c * 2 + kk) * 2048);
              const uint32_t tP = tmem_base + (uint32_t)((j & 1) * BKV + c * 32 + kk * 8);
              umma_bf16_ts(tmem_base + 2 * BKV + (uint32_t)(c * D), tP, bdesc, idesc_pv, (jj > 0 || kk > 0) ? 1u : 0u);
            }
          }
          umma_commit(&kv_empty[s]);
          umma_commit(&pv_full[j & 1]);
        }
      }
      it0 += n;
    }
    __syncwarp();
  } else {
    // ---- softmax: thread = (query row, 32-key quarter) ----
    const int qd = warp & 3;                   // TMEM lane quarter this warp may access
    const int cs = (warp - 2) >> 2;            // key quarter of every tile
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const uint32_t tO = tmem_base + 2 * BKV + (uint32_t)(cs * D) + lane_addr;
    const float sl2 = p.scale * kLog2e;
    int it0 = 0;
    for (int seg = 0; seg < nseg; ++seg) {
      const Segment g = segment(p, seg, nt);
      const int q0 = g.q0, h = g.h, b = g.b;
      const int nloc = g.kb1 - g.kb0;
      float m_used = -INFINITY, l_run = 0.f;
      for (int jj = 0; jj < nloc; ++jj) {
        const int j = it0 + jj;
        const int kb = g.kb0 + jj;
        mbar_wait(&s_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t tS = tmem_base + (uint32_t)((j & 1) * BKV + cs * 32) + lane_addr;
        uint32_t s[32];
        int kvalid = p.nk - kb * BKV - cs * 32;
        if (p.causal) kvalid = min(kvalid, q0 + r + 1 - kb * BKV - cs * 32);
        auto load_s = [&]() {
          tmem_ld32(tS, s);
          tmem_ld_wait();
          if (kvalid < 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= kvalid) s[i] = 0xff800000u;   // -inf
          }
        };
        load_s();
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[i])); mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
        }
        const float m_new = fmaxf(m_used, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2);      // scale > 0
        const bool fresh = m_used == -INFINITY;                 // nothing accumulated against a finite reference yet
        const bool need = !fresh && m_new > m_used + kRescaleThreshold;
        if (fresh) m_used = m_new;
        if (__any_sync(0xffffffffu, need)) {                    // warp-voted: tcgen05.ld / st are warp-collective
          // (need implies jj > 0: the reference is finite only after a tile of this segment has been accumulated)
          mbar_wait(&pv_full[(j - 1) & 1], ((j - 1) >> 1) & 1);   // P V(j-1) has been added to O_c
          tc_fence_after();
          const float alpha = need ? ex2_approx(m_used - m_new) : 1.f;
          const float2 av = make_float2(alpha, alpha);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t oc[32];
            tmem_ld32(tO + half * 32, oc);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 t = __fmul2_rn(make_float2(__uint_as_float(oc[i]), __uint_as_float(oc[i + 1])), av);
              oc[i] = __float_as_uint(t.x); oc[i + 1] = __float_as_uint(t.y);
            }
            tmem_st32(tO + half * 32, oc);
          }
          tmem_st_wait();
          l_run *= alpha;
          if (need) m_used = m_new;
          load_s();          // the scores are read again instead of being kept across this (rare) block: with both s[32] and the
                             // O chunk live the 96-register budget spilled s on EVERY tile
        }
        const float mref = (m_used == -INFINITY) ? 0.f : m_used;
        float2 rsa = make_float2(0.f, 0.f), rsb = make_float2(0.f, 0.f);
        const float2 sl2v = make_float2(sl2, sl2), nm = make_float2(-mref, -mref);
        uint32_t pw[16];
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[gq * 8 + i]), __uint_as_float(s[gq * 8 + i + 1])), sl2v, nm);
            e[i] = ex2_approx(x.x); e[i + 1] = ex2_approx(x.y);
          }
          rsa = __fadd2_rn(rsa, __fadd2_rn(make_float2(e[0], e[1]), make_float2(e[2], e[3])));
          rsb = __fadd2_rn(rsb, __fadd2_rn(make_float2(e[4], e[5]), make_float2(e[6], e[7])));
          pw[gq * 4 + 0] = pack_bf162(e[0], e[1]); pw[gq * 4 + 1] = pack_bf162(e[2], e[3]);
          pw[gq * 4 + 2] = pack_bf162(e[4], e[5]); pw[gq * 4 + 3] = pack_bf162(e[6], e[7]);
        }
        l_run += (rsa.x + rsa.y) + (rsb.x + rsb.y);
        tmem_st16(tS, pw);                   // over the first 16 of my own 32 S columns (all of them are in registers)
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[j & 1]);
        __syncwarp();
      }
      // ---- the segment's O_c: read back once, merge the four key quarters of each row ----
      const int jl = it0 + nloc - 1;
      mbar_wait(&pv_full[jl & 1], (jl >> 1) & 1);
      tc_fence_after();
      float o[D];
      {
        uint32_t r0[32], r1[32];
        tmem_ld32(tO, r0);
        tmem_ld32(tO + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) { o[i] = __uint_as_float(r0[i]); o[32 + i] = __uint_as_float(r1[i]); }
      }
      tc_fence_before();
      if (cs > 0) {
        float* x = xch + ((cs - 1) * BQ + r) * 67;
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] = o[i];
        x[64] = m_used;
        x[65] = l_run;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      const int row = q0 + r;
      if (cs == 0) {
        float m = m_used;
#pragma unroll
        for (int c = 0; c < 3; ++c) m = fmaxf(m, xch[(c * BQ + r) * 67 + 64]);
        const float a0 = (m_used == -INFINITY) ? 0.f : ex2_approx(m_used - m);
        float l = l_run * a0;
#pragma unroll
        for (int i = 0; i < D; ++i) o[i] *= a0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* x = xch + (c * BQ + r) * 67;
          const float mc = x[64];
          const float ac = (mc == -INFINITY) ? 0.f : ex2_approx(mc - m);
          l += x[65] * ac;
#pragma unroll
          for (int i = 0; i < D; ++i) o[i] = fmaf(x[i], ac, o[i]);
        }
        if (g.slot >= 0) {
          // key part of a leftover item: merged, UNNORMALISED (O, m, l) of this row; attn_fwd_fixup_kernel finishes
          float* po = p.part_o + ((long long)g.slot * BQ + r) * D;
#pragma unroll
          for (int i = 0; i < D; i += 4) *reinterpret_cast<float4*>(po + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
          *reinterpret_cast<float2*>(p.part_ml + ((long long)g.slot * BQ + r) * 2) = make_float2(m, l);
        } else if (row < p.nq) {
          const float inv = 1.f / l;
          bf16* dst = p.o + (long long)b * p.bso + (long long)row * p.ldo + h * D;
#pragma unroll
          for (int gq = 0; gq < 8; ++gq) {
            uint4 w;
            w.x = pack_bf162(o[gq * 8 + 0] * inv, o[gq * 8 + 1] * inv); w.y = pack_bf162(o[gq * 8 + 2] * inv, o[gq * 8 + 3] * inv);
            w.z = pack_bf162(o[gq * 8 + 4] * inv, o[gq * 8 + 5] * inv); w.w = pack_bf162(o[gq * 8 + 6] * inv, o[gq * 8 + 7] * inv);
            *reinterpret_cast<uint4*>(dst + gq * 8) = w;
          }
          if (p.lse) p.lse[((long long)b * p.heads + h) * p.nq + row] = (m + log2f(l)) / kLog2e;
        }
      }
      it0 += nloc;
      if (seg + 1 < nseg) asm volatile("bar.sync 1, 512;" ::: "memory");   // xch is written again by the next segment
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// Merge the key parts of every leftover item: thread = (query row, 8 output columns), 16 rows per CTA (8 CTAs per item so
// that the launch spreads over the SMs); all loads of a thread are independent and issued up front - the kernel is pure
// latency otherwise (12 dependent round trips to L2 measured 15 us for what is 5 MB of traffic).
constexpr int kMaxParts = 16;
__global__ void __launch_bounds__(128) attn_fwd_fixup_kernel(const FwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int li = (int)blockIdx.x >> 3;
  const int r = (((int)blockIdx.x & 7) << 4) + ((int)threadIdx.x >> 3), cg = threadIdx.x & 7;
  const int item = p.items - p.n_left + li;
  const int qt = item % p.nqt, hb = item / p.nqt;
  const int h = hb % p.heads, b = hb / p.heads;
  const int row = qt * BQ + r;
  if (row >= p.nq) return;
  const long long slot0 = (long long)li * p.parts;
  float2 ml[kMaxParts];
  float4 v0[kMaxParts], v1[kMaxParts];
#pragma unroll
  for (int q = 0; q < kMaxParts; ++q) {
    ml[q] = make_float2(-INFINITY, 0.f);
    v0[q] = v1[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < p.parts) {
      ml[q] = *reinterpret_cast<const float2*>(p.part_ml + ((slot0 + q) * BQ + r) * 2);
      const float* po = p.part_o + ((slot0 + q) * BQ + r) * D + cg * 8;
      v0[q] = *reinterpret_cast<const float4*>(po);
      v1[q] = *reinterpret_cast<const float4*>(po + 4);
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int q = 0; q < kMaxParts; ++q) m = fmaxf(m, ml[q].x);
  float l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
  for (int q = 0; q < kMaxParts; ++q) {
    const float a = (ml[q].x == -INFINITY) ? 0.f : ex2_approx(ml[q].x - m);
    l = fmaf(ml[q].y, a, l);
    acc[0] = fmaf(v0[q].x, a, acc[0]); acc[1] = fmaf(v0[q].y, a, acc[1]); acc[2] = fmaf(v0[q].z, a, acc[2]); acc[3] = fmaf(v0[q].w, a, acc[3]);
    acc[4] = fmaf(v1[q].x, a, acc[4]); acc[5] = fmaf(v1[q].y, a, acc[5]); acc[6] = fmaf(v1[q].z, a, acc[6]); acc[7] = fmaf(v1[q].w, a, acc[7]);
  }
  const float inv = l > 0.f ? 1.f / l : 0.f;
  uint4 w;
  w.x = pack_bf162(acc[0] * inv, acc[1] * inv); w.y = pack_bf162(acc[2] * inv, acc[3] * inv);
  w.z = pack_bf162(acc[4] * inv, acc[5] * inv); w.w = pack_bf162(acc[6] * inv, acc[7] * inv);
  *reinterpret_cast<uint4*>(p.o + (long long)b * p.bso + (long long)row * p.ldo + h * D + cg * 8) = w;
  if (p.lse && cg == 0) p.lse[((long long)b * p.heads + h) * p.nq + row] = (m + log2f(l)) / kLog2e;
}

// Split decision shared by vn_attention_fwd and vn_attention_fwd_workspace_bytes: leftover items and key parts per item
// (0 parts = every item whole, one CTA per item).
void fwd_split(int nb, int heads, int nq, int nk, int* n_left, int* parts) {
  *n_left = 0; *parts = 0;
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("VN_ATTN_SPLIT"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int items = vn_cdiv(nq, BQ) * heads * nb, nt = vn_cdiv(nk, BKV);
  const int left = items % sms;
  if (items <= sms || left == 0 || left > sms / 2 || nt < 4) return;
  int q = sms / left;
  if (q > nt / 2) q = nt / 2;                     // at least two key blocks per part
  if (q > kMaxParts) q = kMaxParts;
  if (q < 2) return;
  *n_left = left; *parts = q;
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
  p.nqt = vn_cdiv(d->nq, BQ);
  p.items = p.nqt * d->heads * d->nb;
  p.rounds = 1;
  int grid = p.items;
  int n_left = 0, parts = 0;
  fwd_split(d->nb, d->heads, d->nq, d->nk, &n_left, &parts);
  const size_t need = (size_t)n_left * parts * BQ * (D + 2) * sizeof(float);
  if (parts > 0 && d->ws != nullptr && (size_t)d->ws_bytes >= need) {
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    grid = sms;
    p.rounds = p.items / sms;
    p.n_left = n_left; p.parts = parts;
    p.part_o = reinterpret_cast<float*>(d->ws);
    p.part_ml = p.part_o + (size_t)n_left * parts * BQ * D;
  }
  p.dbg = vn_debug_buffer();
  // VN_ATTN_FWD_W16: 0 = never, 1 = always, 2 (default) = long key ranges only (nk >= 2048: the 64 x 64 self-attention, where it
  // measures 4-8 % faster; it loses on items of a few tiles)
  static int w16_mode = -1;
  if (w16_mode < 0) {
    const char* e = getenv("VN_ATTN_FWD_W16");
    w16_mode = e ? atoi(e) : 2;
    if (w16_mode) VN_CUDA(cudaFuncSetAttribute(attn_fwd16_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM16_BYTES));
  }
  const bool w16 = w16_mode == 1 || (w16_mode == 2 && d->nk >= 2048);
  if (w16) VN_LAUNCH(attn_fwd16_tc_kernel, grid, kThreads16, SMEM16_BYTES, (cudaStream_t)s, tq, tk, tv, p);
  else
  VN_LAUNCH(attn_fwd_tc_kernel, grid, kThreads, SMEM_BYTES, (cudaStream_t)s, tq, tk, tv, p);
  if (p.n_left > 0) VN_LAUNCH(attn_fwd_fixup_kernel, p.n_left * 8, 128, 0, (cudaStream_t)s, p);
  return 0;
}

extern "C" size_t vn_attention_fwd_workspace_bytes(int nb, int heads, int nq, int nk) {
  int n_left = 0, parts = 0;
  fwd_split(nb, heads, nq, nk, &n_left, &parts);
  return (size_t)n_left * parts * BQ * (D + 2) * sizeof(float);
}
