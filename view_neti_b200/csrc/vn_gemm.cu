// vn_gemm.cu — tcgen05 / TMEM / TMA GEMM and implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T (+bias) (+rowbias) (+residual)      bf16 x bf16 -> fp32 (TMEM) -> bf16|fp32
//
// Replaces the cuBLAS / cuDNN calls under diffusers' Linear / Conv2d modules on the reference hot path
// (xti_attention_processor.py:30,38-42,53; ResnetBlock2D conv1/conv2/conv_shortcut; Transformer2DModel
// proj_in/out; FeedForward) and, with pre-transposed weights, their dgrads (coach.py:214).
//
// One kernel template, two schedules:
//   PERSISTENT (SPLIT = false): grid = min(tiles, #SMs), one CTA per SM walks 128 x BN output tiles.
//     warp 0    : TMA producer of the A tiles [128 x 64] (128B swizzle); conv mode takes A from a 4-D NHWC tensor map,
//                 one (tap, 64-channel) slab per k-block at (c0, w0+dx-1, h0+dy-1, img): TMA zero-fills the halo, so
//                 padding is free and no im2col buffer exists.  It also prefetches the residual tile into the output
//                 staging buffer.  The loop carries no division / modulo: one thread's issue latency per k-block
//                 bounds the whole main loop.
//     last warp : TMA producer of the B (weight) tiles [BN x 64]: same empty barriers, its own thread.
//     warp 1    : MMA issuer — one lane issues 4 x tcgen05.mma (K=16) per stage into one of TWO TMEM accumulator
//                 stages, so the epilogue of tile i overlaps the main loop of tile i+1.
//     warps 2-9 : epilogue, two warps per TMEM lane quarter (each half of the tile's columns) — tcgen05.ld pipelined one
//                 chunk ahead, + bias / time-embedding row-bias / residual (read from smem), bf16 pack into the swizzled
//                 staging tile, TMA store (clips the M / N tails).
//     CG = 2   : CTA PAIRS (tcgen05.mma.cta_group::2).  Two m-tiles of one n-tile form a 256 x BN tile; each CTA loads
//                its A rows and HALF of the B rows, every load completes on the LEADER's mbarrier, the leader issues the
//                MMAs for the pair and multicasts tcgen05.commit to both CTAs; TMEM is allocated with cta_group::2.
//     MC > 1   : thread-block clusters of MC CTAs along M share every B (weight) tile: each CTA loads 1/MC of it and
//                TMA-MULTICASTS the slice into the shared memory of all CTAs of the cluster (opt-in: measured no gain).
//   SPLIT-K (SPLIT = true): for few-tile / long-K problems (deep UNet levels, M = 64..1024) and for BN = 160.  A
//     thread-block CLUSTER of S in {2,4,8} CTAs shares one output tile, each CTA accumulates K/S in its own TMEM;
//     partials are exchanged through DISTRIBUTED SHARED MEMORY (16-byte st.shared::cluster, column-quad-packed), CTA r
//     reduces rows [r*128/S, (r+1)*128/S) and runs the epilogue for them.  No global atomics, no workspace,
//     deterministic.
// scripts/kernel_timeline.py (-DVN_TIMELINE build) shows where the microseconds of one launch go.
#include "vn_tma.cuh"

#include <type_traits>

#include <stdlib.h>
#include <string.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
#ifndef VN_EPI_HALVES
#define VN_EPI_HALVES 2              // epilogue warps per TMEM lane quarter in the persistent schedule (1 or 2)
#endif
#ifndef VN_PRODUCERS
#define VN_PRODUCERS 2               // 2: a second TMA warp (the last warp of the CTA) issues the B loads
#endif
constexpr int kEpiHalves = VN_EPI_HALVES;
constexpr int kProducers = VN_PRODUCERS;
constexpr int kThreadsSplit = 192 + 32 * (kProducers - 1);   // split-K: TMA warp, MMA warp, 4 epilogue warps (+ B warp)
constexpr int kThreadsPers = 64 + 128 * kEpiHalves + 32 * (kProducers - 1);   // persistent: TMA, MMA, 4 or 8 epilogue warps
constexpr int kEpiThreads = 128 * kEpiHalves;

struct GemmParams {
  int M, N;
  int kb_total, kb_per_split, splits;
  int mode;
  int m_tiles, n_tiles;
  int H, W, cblocks, tw, th, tiles_w, tiles_h, rows_a;   // conv: H, W = OUTPUT grid
  int cs, cpad;         // conv stride (1 | 2) and leading zero padding (1 | 0): tap (dx,dy) of output (w,h) reads input (w*cs + dx - cpad, ...)
  void* D; long long ldd;
  const float* bias;
  const float* rowbias; long long ld_rowbias; int rows_per_batch;
  const bf16* R; long long ldr;
  int r_fp32;           // R holds fp32 (direct epilogue only: fp32 outputs)
  int out_fp32;
  int use_tma_epilogue;
  int nbimg;            // images (conv mode); tiles of a padded cluster slot may decode to img >= nbimg
  int pf_dist;          // weight look-ahead of the B producer in k-blocks (0 = off), see vn_gemm()
  long long* dbg;       // optional in-kernel timeline (vn_set_debug_buffer): 16 slots per CTA, clock64 stamps
};

#ifdef VN_TIMELINE
__device__ __forceinline__ long long gtimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif
#ifdef VN_TIMELINE
#define VN_STAMP(slot)                                                          \
  do {                                                                          \
    if (p.dbg) p.dbg[(long long)blockIdx.x * 16 + (slot)] = clock64();          \
  } while (0)
#else
#define VN_STAMP(slot) do { } while (0)
#endif

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// arrive on the mbarrier at this offset in every CTA of `mask` once all previously issued tcgen05.mma have completed
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// ---- cta_group::2 (CTA pair) variants: loads of both CTAs complete on the LEADER's mbarrier (shared::cluster address),
// the pair's MMA is issued by the leader and its completion is multicast to the barriers of both CTAs ----
__device__ __forceinline__ void tma_load_2d_cg2(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// pull one B box into L2 without landing it anywhere (weight look-ahead of the B producer)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// Direct (register -> global) finish of 8 consecutive columns of one output row; used by the fp32 and split-K paths.
__device__ __forceinline__ void store8(const GemmParams& p, float (&o)[8], long long gm, int bidx, int n) {
  if (p.bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
    o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
    o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
  }
  if (p.rowbias) {
    const float* rb = p.rowbias + (long long)bidx * p.ld_rowbias + n;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(rb));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(rb + 4));
    o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
    o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
  }
  if (p.R && p.r_fp32) {
    const float* rf = reinterpret_cast<const float*>(p.R) + gm * p.ldr + n;
    const float4 r0 = *reinterpret_cast<const float4*>(rf);
    const float4 r1 = *reinterpret_cast<const float4*>(rf + 4);
    o[0] += r0.x; o[1] += r0.y; o[2] += r0.z; o[3] += r0.w;
    o[4] += r1.x; o[5] += r1.y; o[6] += r1.z; o[7] += r1.w;
  } else if (p.R) {
    const uint4 r = *reinterpret_cast<const uint4*>(p.R + gm * p.ldr + n);
    float2 t;
    t = unpack_bf162(r.x); o[0] += t.x; o[1] += t.y;
    t = unpack_bf162(r.y); o[2] += t.x; o[3] += t.y;
    t = unpack_bf162(r.z); o[4] += t.x; o[5] += t.y;
    t = unpack_bf162(r.w); o[6] += t.x; o[7] += t.y;
  }
  if (p.out_fp32) {
    float* dst = reinterpret_cast<float*>(p.D) + gm * p.ldd + n;
    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
  } else {
    uint4 w;
    w.x = pack_bf162(o[0], o[1]); w.y = pack_bf162(o[2], o[3]);
    w.z = pack_bf162(o[4], o[5]); w.w = pack_bf162(o[6], o[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.D) + gm * p.ldd + n) = w;
  }
}

// store8 for the split-K finishing pass: bias (and a conv's per-image row-bias) and a bf16 residual were already added
// from shared memory / registers by the caller
__device__ __forceinline__ void store8_rest(const GemmParams& p, float (&o)[8], long long gm, int bidx, int n,
                                            bool rowbias_done, bool r_done) {
  if (p.rowbias && !rowbias_done) {
    const float* rb = p.rowbias + (long long)bidx * p.ld_rowbias + n;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(rb));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(rb + 4));
    o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
    o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
  }
  if (p.R && !r_done) {                            // fp32 residual stream (the CLIP text encoder)
    const float* rf = reinterpret_cast<const float*>(p.R) + gm * p.ldr + n;
    const float4 r0 = *reinterpret_cast<const float4*>(rf);
    const float4 r1 = *reinterpret_cast<const float4*>(rf + 4);
    o[0] += r0.x; o[1] += r0.y; o[2] += r0.z; o[3] += r0.w;
    o[4] += r1.x; o[5] += r1.y; o[6] += r1.z; o[7] += r1.w;
  }
  if (p.out_fp32) {
    float* dst = reinterpret_cast<float*>(p.D) + gm * p.ldd + n;
    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
  } else {
    uint4 w;
    w.x = pack_bf162(o[0], o[1]); w.y = pack_bf162(o[2], o[3]);
    w.z = pack_bf162(o[4], o[5]); w.w = pack_bf162(o[6], o[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.D) + gm * p.ldd + n) = w;
  }
}

__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

struct TileCoord {
  int m0, n0, img, h0, w0;
  int ah0, aw0;         // conv: input coordinates of tap (0,0) of the tile's first output pixel
};
__device__ __forceinline__ TileCoord tile_coord(const GemmParams& p, int tile, int bn) {
  TileCoord c;
  const int mt = tile % p.m_tiles;
  c.n0 = (tile / p.m_tiles) * bn;
  c.m0 = 0; c.img = 0; c.h0 = 0; c.w0 = 0; c.ah0 = 0; c.aw0 = 0;
  if (p.mode == 0) {
    c.m0 = mt * BM;
  } else {
    const int tpi = p.tiles_w * p.tiles_h;
    c.img = mt / tpi;
    const int r = mt - c.img * tpi;
    c.h0 = (r / p.tiles_w) * p.th;
    c.w0 = (r % p.tiles_w) * p.tw;
    c.ah0 = c.h0 * p.cs - p.cpad;
    c.aw0 = c.w0 * p.cs - p.cpad;
  }
  return c;
}
// global row index, batch index and validity of tile row r
__device__ __forceinline__ bool tile_row(const GemmParams& p, const TileCoord& c, int r, long long* gm, int* bidx) {
  if (p.mode == 0) {
    *gm = (long long)c.m0 + r;
    *bidx = p.rows_per_batch > 0 ? (int)(*gm / p.rows_per_batch) : 0;
    return *gm < p.M;
  }
  const int ty = r / p.tw, tx = r - ty * p.tw;
  const int h = c.h0 + ty, w = c.w0 + tx;
  *gm = ((long long)c.img * p.H + h) * p.W + w;
  *bidx = c.img;
  return (r < p.rows_a) && (h < p.H) && (w < p.W);
}

// split-K: does a dedicated exchange buffer (128 x BN fp32) fit behind the pipeline stages?  (227 KB of dynamic shared memory)
template <int BN, int STAGES>
constexpr bool split_exch_own() {
  return STAGES * (BM * BK * 2 + BN * BK * 2) + BM * BN * 4 + BN * 4 + (2 * STAGES + 6) * 8 + 16 + 1024 <= 232448;
}
template <int BN, int STAGES, bool SPLIT, int MC, int CG>
__global__ void __launch_bounds__(SPLIT ? kThreadsSplit : kThreadsPers, 1) vn_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmD,
                                                               const __grid_constant__ CUtensorMap tmR,
                                                               const GemmParams p) {
  pdl_trigger();
#ifdef VN_TIMELINE
  if (p.dbg && threadIdx.x == 0) { p.dbg[(long long)blockIdx.x * 16 + 0] = gtimer_ns(); VN_STAMP(1); }
#endif
  constexpr int A_BYTES = BM * BK * 2;            // 16 KB
  constexpr int B_BYTES = (BN / CG) * BK * 2;     // CG == 2: each CTA of the pair holds half of the B tile
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // split-K: where it fits (BN 64 / 128 with 6 / 5 stages; BN 160 would have to drop to 4 stages, measured slower: 6.32 -> 6.36 ms) the
  // exchange buffer gets its OWN shared memory behind the
  // pipeline stages; wider tiles alias it with the stages and need a cluster barrier between the main loop and the exchange
  constexpr bool kExchOwn = SPLIT && split_exch_own<BN, STAGES>();
  constexpr int STAGING_BYTES = SPLIT ? (kExchOwn ? BM * BN * 4 : 0) : BM * BN * 2;     // BN/64 blocks of [128 rows x 128 B], 128B-swizzled
  constexpr int ACC_STAGES = SPLIT ? 1 : 2;
  constexpr int TMEM_COLS = (ACC_STAGES * BN) <= 32 ? 32 : (ACC_STAGES * BN) <= 64 ? 64 : (ACC_STAGES * BN) <= 128 ? 128
                          : (ACC_STAGES * BN) <= 256 ? 256 : 512;
  static_assert(STAGE_BYTES % 1024 == 0 && BN % 32 == 0 && (SPLIT || BN % 64 == 0) && ACC_STAGES * BN <= 512,
                "tile configuration (the staged epilogue works on 64-column blocks; split-K stores directly)");
  static_assert(!SPLIT || STAGES * STAGE_BYTES >= BM * BN * 4, "exchange buffer must fit in the pipeline stages");
  static_assert(MC == 1 || (!SPLIT && BN % MC == 0), "multicast clusters only in the persistent schedule");
  static_assert(CG == 1 || (CG == 2 && !SPLIT && MC == 1 && kProducers == 2), "CTA pairs only in the persistent schedule");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  float* sbias = reinterpret_cast<float*>(staging + STAGING_BYTES);      // [BN] bias (+ per-image row-bias) of the tile
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbias + BN);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;       // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;           // [2] accumulator drained
  uint64_t* rfull_bar = tempty_bar + 2;           // residual tile landed in staging
  uint64_t* sfree_bar = rfull_bar + 1;            // staging buffer reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sfree_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (!SPLIT && p.use_tma_epilogue) {
      tma_prefetch_desc(&tmD);
      if (p.R) tma_prefetch_desc(&tmR);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], MC);                 // every CTA of the multicast cluster must have consumed the stage
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], SPLIT ? 4 : 4 * kEpiHalves * CG);      // one arrival per epilogue warp (of the pair)
    }
    mbar_init(rfull_bar, 1);
    mbar_init(sfree_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_cg2<TMEM_COLS>(tmem_slot);      // one warp of EACH CTA of the pair, same slot offset
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if (MC > 1 || CG == 2) cluster_sync_all();       // peers' barriers must exist before a multicast can signal them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // (own exchange buffer: the only thing the first cluster barrier still has to guarantee is that every CTA of the cluster has
  //  started before a peer writes into its shared memory - arrive here, wait in front of the exchange, nobody ever blocks)
  if (kExchOwn) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  if (threadIdx.x == 0) VN_STAMP(2);

  // ---- work assignment ----
  // CG == 2: the work items are PAIR tiles (256 rows x BN): m-tiles 2q and 2q+1 of one n-tile; CTA `crank` of the pair
  // owns m-tile 2q + crank (its A rows, its half of the accumulator, its D tile) and loads half of the B rows
  const int num_tiles = (p.m_tiles / CG) * p.n_tiles;
  const int mh = p.m_tiles / CG;
  int tile_begin, tile_step, kb_begin, nkb;
  uint32_t crank = (SPLIT || MC > 1 || CG == 2) ? cluster_ctarank() : 0;
#define VN_TILE_OF(w) (CG == 2 ? (((w) / mh) * p.m_tiles + 2 * ((w) % mh) + (int)crank) : (w))
  if (SPLIT) {
    tile_begin = blockIdx.x / p.splits;
    tile_step = num_tiles;                       // exactly one tile per cluster
    kb_begin = (int)crank * p.kb_per_split;
    nkb = min(p.kb_total, kb_begin + p.kb_per_split) - kb_begin;   // host guarantees >= 1
  } else {
    tile_begin = blockIdx.x / CG;
    tile_step = gridDim.x / CG;
    kb_begin = 0;
    nkb = p.kb_total;
  }
  // Weight look-ahead (launches that stream a large, cold weight matrix through few CTAs: M <= 256).  With 3-6 stages of
  // shared memory a CTA keeps ~100 KB of weights in flight, one stage turn-over per DRAM round trip - a third of the
  // HBM rate over 80 CTAs.  The B producer warp, which has slack, pulls the k-blocks `pf_dist` ahead into L2 with
  // cp.async.bulk.prefetch.tensor (no shared memory, no barrier); the first ones even before griddepcontrol.wait,
  // because nothing on the device ever writes the frozen weights.
  constexpr int kBWarp = SPLIT ? 6 : 2 + 4 * kEpiHalves;
  // The B producer also puts the weight tiles of the first STAGES k-blocks in flight BEFORE the grid dependency is
  // resolved (all stages are free at start; the A producer arms the barriers with expect_tx afterwards - a complete_tx
  // that lands first only drives the transaction count negative until then, the phase cannot complete without the
  // pending arrival).  Costs the A producer nothing: it is a different thread.
  int b_pre = 0;
  if (kProducers == 2 && MC == 1 && warp == kBWarp && lane == 0 && tile_begin < num_tiles) {
    int n0 = (VN_TILE_OF(tile_begin) / p.m_tiles) * BN;
    if (CG == 2) n0 += (int)crank * (((min(BN, p.N - n0) + 15) & ~15) >> 1);
    b_pre = min(nkb, STAGES);
    const uint32_t lead_full0 = CG == 2 ? mapa_shared(smem_u32(full_bar), 0) : 0;
    for (int i = 0; i < b_pre; ++i) {
      if (CG == 2) tma_load_2d_cg2(smem + i * STAGE_BYTES + A_BYTES, &tmB, lead_full0 + (uint32_t)(i * 8), (kb_begin + i) * BK, n0);
      else tma_load_2d(smem + i * STAGE_BYTES + A_BYTES, &tmB, &full_bar[i], (kb_begin + i) * BK, n0);
    }
    if (p.pf_dist > 0) {
      const int pf_end = min(nkb, p.pf_dist);
      for (int i = b_pre; i < pf_end; ++i) tma_prefetch_l2_2d(&tmB, (kb_begin + i) * BK, n0);
    }
  }
  // coordinates of the CTA's first tile (a handful of integer divisions): computed here, where they overlap the
  // predecessor kernel's tail, instead of at the head of every role's loop behind griddepcontrol.wait
  const int tile_first = VN_TILE_OF(tile_begin);
  const TileCoord c_first = tile_coord(p, tile_first, BN);
  // griddepcontrol.wait is executed PER ROLE, as late as each role allows (it is a per-thread instruction): the A producer
  // right in front of its first activation load - its ~150 instructions of loop set-up used to sit behind the wait, 0.45 us
  // of every launch (profiles/r1_kernel_timeline_v4.txt: "pdl_wait passed" -> "first A load issued") -, the epilogue /
  // finishing warps before their first global access (residual, row-bias, D), the weight producer and the MMA warp never:
  // frozen weights and shared / tensor memory do not depend on the previous kernel.
  const bool tma_epi = !SPLIT && p.use_tma_epilogue;

  if (warp == 0) {
    // ================= TMA producer =================
    // One thread issues every load; its per-k-block instruction chain is on the critical path of the whole main loop
    // (a stage is re-armed only after this thread has seen it drained), so the loop carries no division or modulo:
    // stage index / phase and the conv (tap, channel-block) coordinates advance incrementally.  The warp's control flow is
    // uniform and one ELECTED lane issues (elect_one): coordinates, barrier addresses and the tensor-map pointer stay in
    // uniform registers instead of being broadcast lane by lane in front of every cp.async.bulk.tensor.
    {
      const bool leader = elect_one();
      const uint32_t tx_bytes = (uint32_t)(CG * (p.rows_a * BK * 2 + B_BYTES));      // both CTAs of a pair
      const uint32_t lead_full = CG == 2 ? mapa_shared(smem_u32(full_bar), 0) : 0;  // leader's full_bar[0]
      int s = 0, t = 0;
      uint32_t ph = 1;                             // parity to wait for on empty_bar[s]
      for (int w = tile_begin; w < num_tiles; w += tile_step, ++t) {
        const TileCoord c = w == tile_begin ? c_first : tile_coord(p, VN_TILE_OF(w), BN);
        int kx = kb_begin * BK;                    // K coordinate (elements) of the next k-block
        int cb = 0, dx = 0, dy = 0;                // conv: 64-channel block and 3x3 tap of the next k-block
        if (p.mode == 1) {
          const int tap = kb_begin / p.cblocks;
          cb = kb_begin - tap * p.cblocks;
          dy = tap / 3; dx = tap - dy * 3;
        }
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(&empty_bar[s], ph);
          uint8_t* sa = smem + s * STAGE_BYTES;
          if (i == 0 && t == 0) {
            pdl_wait();                            // activations (A, the residual tile) are read from here on
            if (leader) VN_STAMP(3);
          }
          if (leader) {
          if (CG == 1 || crank == 0) mbar_expect_tx(&full_bar[s], tx_bytes);
          if (CG == 2) {
            const uint32_t fb = lead_full + (uint32_t)(s * 8);
            if (p.mode == 0) {
              tma_load_2d_cg2(sa, &tmA, fb, kx, c.m0);
            } else {
              tma_load_4d_cg2(sa, &tmA, fb, cb * BK, c.aw0 + dx, c.ah0 + dy, c.img);
            }
          } else if (p.mode == 0) {
            tma_load_2d(sa, &tmA, &full_bar[s], kx, c.m0);
          } else {
            tma_load_4d(sa, &tmA, &full_bar[s], cb * BK, c.aw0 + dx, c.ah0 + dy, c.img);
          }
          if (MC == 1) {
            if (kProducers == 1) tma_load_2d(sa + A_BYTES, &tmB, &full_bar[s], kx, c.n0);
          } else {
            // my 1/MC slice of the B tile goes to every CTA of the cluster (same n-tile, consecutive m-tiles)
            constexpr int SLICE = BN / MC;
            tma_load_2d_mc(sa + A_BYTES + (int)crank * SLICE * 128, &tmB, &full_bar[s], kx, c.n0 + (int)crank * SLICE,
                           (uint16_t)((1u << MC) - 1));
          }
          }                                        // leader
          if (p.mode != 0 && ++cb == p.cblocks) { cb = 0; if (++dx == 3) { dx = 0; ++dy; } }
          kx += BK;
          if (++s == STAGES) { s = 0; ph ^= 1u; }
          if (i == 0 && t == 0 && leader) VN_STAMP(4);
        }
        if (t == 0 && leader) VN_STAMP(5);
        if (tma_epi && p.R) {
          // residual tile -> staging, once the previous tile's store has finished reading it
          mbar_wait(sfree_bar, (t & 1) ^ 1);
          const int nblk = min(BN / 64, (p.N - c.n0 + 63) / 64);
          if (leader) {
            mbar_expect_tx(rfull_bar, (uint32_t)(nblk * p.rows_a * 128));
            for (int j = 0; j < nblk; ++j) {
              if (p.mode == 0) tma_load_2d(staging + j * (BM * 128), &tmR, rfull_bar, c.n0 + j * 64, c.m0);
              else tma_load_4d(staging + j * (BM * 128), &tmR, rfull_bar, c.n0 + j * 64, c.w0, c.h0, c.img);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (kProducers == 2 && MC == 1 && warp == kBWarp) {
    // ================= second TMA producer: the B (weight) tiles =================
    // Issue latency of the single producer thread bounds the main loop; the B loads of a stage need nothing from the
    // A producer but the drained stage (same empty barrier) - their complete_tx may land before its expect_tx, the
    // phase cannot complete until that arrival.
    {
      const bool leader = elect_one();
      const int b_pre_w = __shfl_sync(0xffffffffu, b_pre, 0);        // lane 0 put the first b_pre weight tiles in flight above
      int s = 0;
      uint32_t ph = 1;
      const uint32_t lead_full = CG == 2 ? mapa_shared(smem_u32(full_bar), 0) : 0;
      for (int w = tile_begin; w < num_tiles; w += tile_step) {
        int n0 = w == tile_begin ? c_first.n0 : (VN_TILE_OF(w) / p.m_tiles) * BN;
        if (CG == 2) {
          // the pair's MMA takes B rows [0, n/2) from the leader and [n/2, n) from its peer (n = the tile's UMMA N)
          const int n_eff = (min(BN, p.N - n0) + 15) & ~15;
          n0 += (int)crank * (n_eff >> 1);
        }
        int kx = kb_begin * BK;
        for (int i = 0; i < nkb; ++i) {
          if (leader && p.pf_dist > 0 && i + p.pf_dist < nkb) tma_prefetch_l2_2d(&tmB, kx + p.pf_dist * BK, n0);
          if (w != tile_begin || i >= b_pre_w) {           // (the first b_pre tiles of the first tile are already in flight)
            mbar_wait(&empty_bar[s], ph);
            if (leader) {
              if (CG == 2) tma_load_2d_cg2(smem + s * STAGE_BYTES + A_BYTES, &tmB, lead_full + (uint32_t)(s * 8), kx, n0);
              else tma_load_2d(smem + s * STAGE_BYTES + A_BYTES, &tmB, &full_bar[s], kx, n0);
            }
          }
          kx += BK;
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (CG == 2: the leader CTA issues for the pair) =================
    // The whole warp runs the loop (warp-uniform control flow) and ONE elected lane issues: descriptors and tensor-memory
    // addresses then live in uniform registers and the tcgen05.mma of a k-block go out back to back; under a divergent
    // `if (lane == 0)` each one was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~9 instructions per MMA).
    if (CG == 1 || crank == 0) {
      const bool leader = elect_one();
      int s = 0, t = 0;
      uint32_t ph = 0;                             // parity to wait for on full_bar[s]
      const uint64_t desc0 = umma_desc_k_sw128(smem_u32(smem));          // stage 0, A operand
      constexpr uint64_t kStageStep = (uint64_t)(STAGE_BYTES >> 4);      // descriptor start-address units of 16 B
      constexpr uint64_t kBOffset = (uint64_t)(A_BYTES >> 4);
      for (int w = tile_begin; w < num_tiles; w += tile_step, ++t) {
        const int n0 = w == tile_begin ? c_first.n0 : (VN_TILE_OF(w) / p.m_tiles) * BN;
        int n_eff = min(BN, p.N - n0);
        n_eff = (n_eff + 15) & ~15;               // UMMA N granularity; B rows beyond N are TMA zero-fill
        const uint32_t idesc = umma_idesc_bf16(BM * CG, n_eff);
        const int as = SPLIT ? 0 : (t & 1);
        if (!SPLIT) {
          mbar_wait(&tempty_bar[as], ((t >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (i == 0 && t == 0 && leader) VN_STAMP(6);
          const uint64_t adesc = desc0 + (uint64_t)s * kStageStep;
          const uint64_t bdesc = adesc + kBOffset;
          if (leader) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
              if (CG == 2) umma_bf16_cg2(tacc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (i | k) ? 1u : 0u);
              else umma_bf16(tacc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (i | k) ? 1u : 0u);
            }
            if (CG == 2) umma_commit_cg2(&empty_bar[s], 3);   // the stage is free in BOTH CTAs of the pair
            else if (MC == 1) umma_commit(&empty_bar[s]);   // frees this smem stage when the MMAs above have read it
            else umma_commit_mc(&empty_bar[s], (uint16_t)((1u << MC) - 1));
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (leader) {
          if (CG == 2) umma_commit_cg2(&tfull_bar[as], 3);   // both halves of the pair's accumulator are complete
          else umma_commit(&tfull_bar[as]);             // accumulator complete
        }
        if (t == 0 && leader) VN_STAMP(7);
      }
    }
    __syncwarp();
  } else if (!SPLIT && warp < 2 + 4 * kEpiHalves) {
    // ================= epilogue (warps 2..9), persistent schedule =================
    // Two warps per TMEM lane quarter: warps 2..5 finish the left half of the tile's columns, warps 6..9 the right half
    // (a lone warp per scheduler cannot hide its own instruction latency, and most launches of a batch-1 step are
    // single-tile, so the epilogue is on the critical path).  TMEM loads are software-pipelined one chunk ahead.
    pdl_wait();                                  // row-bias / residual reads and the D stores depend on the previous kernel
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                 // tile row == TMEM lane
    const int half = (warp - 2) >> 2;            // column half of the tile (0 when kEpiHalves == 1)
    constexpr int NCH = BN / 32 / kEpiHalves;    // 32-column chunks per epilogue warp
    int t = 0;
    const uint32_t lead_tempty = CG == 2 ? mapa_shared(smem_u32(tempty_bar), 0) : 0;
    for (int w = tile_begin; w < num_tiles; w += tile_step, ++t) {
      const TileCoord c = w == tile_begin ? c_first : tile_coord(p, VN_TILE_OF(w), BN);
      const int as = t & 1;
      long long gm;
      int bidx;
      const bool row_ok = tile_row(p, c, r, &gm, &bidx);
      // bias (+ the per-image time-embedding row-bias in conv mode) of this tile: global loads are issued before
      // the accumulator wait and parked in shared memory, so no global latency sits on the epilogue's critical path
      const bool smem_rowbias = p.rowbias && p.mode == 1;
      const int te = threadIdx.x - 64;           // 0..kEpiThreads-1
      constexpr int NBV = (BN + kEpiThreads - 1) / kEpiThreads;
      float bv[NBV];
#pragma unroll
      for (int u = 0; u < NBV; ++u) {
        const int col = te + u * kEpiThreads;
        bv[u] = 0.f;
        if (col < BN && c.n0 + col < p.N) {
          if (p.bias) bv[u] = __ldg(p.bias + c.n0 + col);
          if (smem_rowbias) bv[u] += __ldg(p.rowbias + (long long)min(c.img, p.nbimg - 1) * p.ld_rowbias + c.n0 + col);
        }
      }
      mbar_wait(&tfull_bar[as], (t >> 1) & 1);
      tc_fence_after();
      if (t == 0 && threadIdx.x == 64) VN_STAMP(8);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      if (tma_epi) {
        uint32_t raw[2][32];
        tmem_ld32(taddr + (half * NCH) * 32, raw[0]);          // in flight across the barrier below
#pragma unroll
        for (int u = 0; u < NBV; ++u)
          if (te + u * kEpiThreads < BN) sbias[te + u * kEpiThreads] = bv[u];
        epi_bar_sync();
        if (p.R) mbar_wait(rfull_bar, t & 1);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int cc = half * NCH + i;
          const int nb0 = c.n0 + cc * 32;
          tmem_ld_wait();
          if (i + 1 < NCH) tmem_ld32(taddr + (cc + 1) * 32, raw[(i + 1) & 1]);
          if (nb0 < p.N) {
            uint8_t* blk = staging + (cc >> 1) * (BM * 128) + r * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int n = nb0 + g * 8;
              float o[8];
              const float4 b0 = *reinterpret_cast<const float4*>(sbias + cc * 32 + g * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(sbias + cc * 32 + g * 8 + 4);
              o[0] = __uint_as_float(raw[i & 1][g * 8 + 0]) + b0.x; o[1] = __uint_as_float(raw[i & 1][g * 8 + 1]) + b0.y;
              o[2] = __uint_as_float(raw[i & 1][g * 8 + 2]) + b0.z; o[3] = __uint_as_float(raw[i & 1][g * 8 + 3]) + b0.w;
              o[4] = __uint_as_float(raw[i & 1][g * 8 + 4]) + b1.x; o[5] = __uint_as_float(raw[i & 1][g * 8 + 5]) + b1.y;
              o[6] = __uint_as_float(raw[i & 1][g * 8 + 6]) + b1.z; o[7] = __uint_as_float(raw[i & 1][g * 8 + 7]) + b1.w;
              if (p.rowbias && !smem_rowbias && n < p.N && row_ok) {
                const float* rb = p.rowbias + (long long)bidx * p.ld_rowbias + n;
                const float4 c0 = __ldg(reinterpret_cast<const float4*>(rb));
                const float4 c1 = __ldg(reinterpret_cast<const float4*>(rb + 4));
                o[0] += c0.x; o[1] += c0.y; o[2] += c0.z; o[3] += c0.w;
                o[4] += c1.x; o[5] += c1.y; o[6] += c1.z; o[7] += c1.w;
              }
              uint4* slot = reinterpret_cast<uint4*>(blk + ((((cc & 1) * 4 + g) ^ (r & 7)) << 4));
              if (p.R && n < p.N) {
                const uint4 rr = *slot;
                float2 f;
                f = unpack_bf162(rr.x); o[0] += f.x; o[1] += f.y;
                f = unpack_bf162(rr.y); o[2] += f.x; o[3] += f.y;
                f = unpack_bf162(rr.z); o[4] += f.x; o[5] += f.y;
                f = unpack_bf162(rr.w); o[6] += f.x; o[7] += f.y;
              }
              uint4 w;
              w.x = pack_bf162(o[0], o[1]); w.y = pack_bf162(o[2], o[3]);
              w.z = pack_bf162(o[4], o[5]); w.w = pack_bf162(o[6], o[7]);
              *slot = w;
            }
          }
        }
        tc_fence_before();
        if (lane == 0) {                                   // TMEM stage may be overwritten by tile t+2
          if (CG == 2) mbar_arrive_cluster(lead_tempty + (uint32_t)(as * 8));
          else mbar_arrive(&tempty_bar[as]);
        }
        if (t == 0 && threadIdx.x == 64) VN_STAMP(14);
        fence_proxy_async();
        epi_bar_sync();
        if (threadIdx.x == 64) {
          const int nblk = min(BN / 64, (p.N - c.n0 + 63) / 64);
          for (int j = 0; j < nblk; ++j) {
            if (p.mode == 0) tma_store_2d(&tmD, staging + j * (BM * 128), c.n0 + j * 64, c.m0);
            else tma_store_4d(&tmD, staging + j * (BM * 128), c.n0 + j * 64, c.w0, c.h0, c.img);
          }
          tma_store_commit();
          if (t == 0) VN_STAMP(15);
          tma_store_wait_read();
          mbar_arrive(sfree_bar);
        }
        epi_bar_sync();                                      // staging reusable by all epilogue warps
      } else {
        // direct register -> global path (fp32 outputs, odd strides)
#pragma unroll 1
        for (int cc = half * NCH; cc < (half + 1) * NCH; ++cc) {
          const int nb0 = c.n0 + cc * 32;
          if (nb0 >= p.N) break;
          uint32_t raw[32];
          tmem_ld32(taddr + cc * 32, raw);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int n = nb0 + g * 8;
              if (n < p.N) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(raw[g * 8 + j]);
                store8(p, o, gm, bidx, n);
              }
            }
          }
        }
        tc_fence_before();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(lead_tempty + (uint32_t)(as * 8));
          else mbar_arrive(&tempty_bar[as]);
        }
      }
    }
    if (threadIdx.x == 64) VN_STAMP(9);
    // (the bulk stores were drained from shared memory by wait_group.read above; their global visibility is ordered by
    //  grid completion, which is what the next kernel's griddepcontrol.wait / stream order waits for)
  }

  if (SPLIT) {
    // ================= split-K: exchange partial tiles through distributed shared memory =================
    const int S = p.splits;
    const int rpo = BM / S;                       // rows reduced (and finished) by each CTA of the cluster
    const int cpt = BN / S;                       // columns finished per epilogue thread
    // [S src][BN/4 column quads][rpo row][4] fp32, aliases the pipeline stages.  The 32 lanes of a warp (32 consecutive
    // tile rows) write one contiguous 512-byte run per 16-byte remote store - whole 128-byte packets on the SM-to-SM
    // network with a quarter of the instructions of scalar stores (a row-major float4 layout, 32 scattered 16-byte
    // pieces per instruction, measured 1.7x slower on B200); the finishing threads read it back with LDS.128,
    // conflict-free for the same reason.
    float* exch = reinterpret_cast<float*>(kExchOwn ? staging : smem);
    const TileCoord c = c_first;
    // Operands of the finishing pass that do not depend on the partials - bias (+ the per-image row-bias of a conv) of the
    // tile and this thread's residual values - are requested NOW, while the main loop still runs: read after the second
    // cluster barrier they cost one exposed L2 round trip per 8-column group (4 x 0.57 us on a 128-wide tile split 4
    // ways, profiles/r1_kernel_timeline_v4.txt: "cluster sync 2" -> "epilogue done" 2.3 us).
    constexpr int PRE = 4;                        // residual groups kept in registers per chunk
    const int te = (int)threadIdx.x - 64;         // 0..127 in the epilogue warps
    const int rl = te % rpo, cg = te / rpo;
    const int ngrp = cpt / 8;
    const bool smem_rowbias = p.rowbias && p.mode == 1;
    const bool r16 = p.R && !p.r_fp32;
    long long gm = 0;
    int bidx = 0;
    bool row_ok = false;
    uint4 rpre[PRE];
    if (warp >= 2 && warp < 6) {
      pdl_wait();                                 // row-bias / residual reads and the D stores depend on the previous kernel
      for (int col = te; col < BN; col += 128) {
        float b = 0.f;
        if (c.n0 + col < p.N) {
          if (p.bias) b = __ldg(p.bias + c.n0 + col);
          if (smem_rowbias) b += __ldg(p.rowbias + (long long)min(c.img, p.nbimg - 1) * p.ld_rowbias + c.n0 + col);
        }
        sbias[col] = b;                           // visible to the finishing threads through the two cluster barriers below
      }
      row_ok = tile_row(p, c, (int)crank * rpo + rl, &gm, &bidx);
#pragma unroll
      for (int g = 0; g < PRE; ++g) {
        rpre[g] = make_uint4(0u, 0u, 0u, 0u);
        const int n = c.n0 + cg * cpt + g * 8;
        if (r16 && row_ok && g < ngrp && n < p.N) rpre[g] = *reinterpret_cast<const uint4*>(p.R + gm * p.ldr + n);
      }
      mbar_wait(&tfull_bar[0], 0);                // my accumulator is complete => my stages are no longer read
      tc_fence_after();
      if (threadIdx.x == 64) VN_STAMP(8);
    }
    __syncwarp();
    if (kExchOwn) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // completes the arrive made at set-up
    else cluster_sync_all();                      // every CTA of the cluster is done with its pipeline stages
    if (threadIdx.x == 64) VN_STAMP(12);
    if (warp >= 2 && warp < 6) {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const uint32_t owner = (uint32_t)(r / rpo);
      const uint32_t remote = mapa_shared(smem_u32(exch), owner) + (uint32_t)((((int)crank * (BN / 4)) * rpo + (r % rpo)) * 16);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      // tensor-memory loads run one 32-column chunk ahead of the remote stores.  Rows of the tile that do not exist (M = 64
      // launches fill half a tile) are not sent: the SM-to-SM network moves ~20 B/clk per SM, the exchange is bound by it,
      // and their owners never read them.
      constexpr int NCC = BN / 32;
      long long gm_x;
      int b_x;
      const bool send = tile_row(p, c, r, &gm_x, &b_x);
      const bool wsend = __any_sync(0xffffffffu, send);        // tcgen05.ld is warp-collective: loads stay warp-uniform
      uint32_t raw[2][32];
      if (wsend) tmem_ld32(taddr, raw[0]);
#pragma unroll
      for (int cc = 0; cc < NCC; ++cc) {
        if (wsend && c.n0 + cc * 32 < p.N) {
          tmem_ld_wait();
          if (cc + 1 < NCC && c.n0 + (cc + 1) * 32 < p.N) tmem_ld32(taddr + (cc + 1) * 32, raw[(cc + 1) & 1]);
          if (send)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(remote + (uint32_t)((cc * 8 + j) * rpo * 16)),
                         "r"(raw[cc & 1][4 * j]), "r"(raw[cc & 1][4 * j + 1]), "r"(raw[cc & 1][4 * j + 2]), "r"(raw[cc & 1][4 * j + 3])
                         : "memory");
        }
      }
      tmem_ld_wait();
    }
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();                           // all partials have landed (release / acquire at cluster scope)
    if (threadIdx.x == 64) VN_STAMP(13);
    if (warp >= 2 && warp < 6) {
      if (row_ok) {
        // Finishing pass with the split count as a COMPILE-TIME constant (2, 4 or 8): the partials of up to four 8-column
        // groups (<= 16 LDS.128, explicit shared-space loads) are all requested before the first add.  The generic-pointer
        // version issued one dependent load -> add -> store chain per group, and the compiler could not move a group's
        // loads above the previous group's global store: 600 cycles per group, 3.4 us for the ten groups of a BN-160 tile
        // split two ways (in-kernel timeline, "cluster sync 2" -> "epilogue done").
        const uint32_t ex_s = smem_u32(exch), sb_s = smem_u32(sbias);
        auto finish = [&](auto sc) {
          constexpr int SS = decltype(sc)::value;
          constexpr int NG = BN / (8 * SS);        // 8-column groups per finishing thread
          constexpr int RPO = BM / SS;
          constexpr int G = (8 / SS) < 1 ? 1 : ((8 / SS) < NG ? (8 / SS) : NG);     // groups per chunk
#pragma unroll
          for (int g0 = 0; g0 < NG; g0 += G) {
            float4 v[G][SS][2], bb[G][2];
            uint4 rr[G];
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
              const int g = g0 + gi;
              if (g < NG) {
                const int col = cg * (BN / SS) + g * 8;
                bb[gi][0] = lds128(sb_s + (uint32_t)(col * 4));
                bb[gi][1] = lds128(sb_s + (uint32_t)(col * 4 + 16));
#pragma unroll
                for (int src = 0; src < SS; ++src) {
                  const uint32_t a = ex_s + (uint32_t)(((src * (BN / 4) + (col >> 2)) * RPO + rl) * 16);
                  v[gi][src][0] = lds128(a);
                  v[gi][src][1] = lds128(a + RPO * 16);
                }
                rr[gi] = make_uint4(0u, 0u, 0u, 0u);
                if (g < PRE) rr[gi] = rpre[g];
                else if (r16 && c.n0 + col < p.N) rr[gi] = *reinterpret_cast<const uint4*>(p.R + gm * p.ldr + c.n0 + col);
              }
            }
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
              const int g = g0 + gi;
              if (g < NG) {
                const int col = cg * (BN / SS) + g * 8;
                const int n = c.n0 + col;
                if (n < p.N) {
                  float o[8] = {bb[gi][0].x, bb[gi][0].y, bb[gi][0].z, bb[gi][0].w, bb[gi][1].x, bb[gi][1].y, bb[gi][1].z, bb[gi][1].w};
#pragma unroll
                  for (int src = 0; src < SS; ++src) {        // fixed order: deterministic
                    o[0] += v[gi][src][0].x; o[1] += v[gi][src][0].y; o[2] += v[gi][src][0].z; o[3] += v[gi][src][0].w;
                    o[4] += v[gi][src][1].x; o[5] += v[gi][src][1].y; o[6] += v[gi][src][1].z; o[7] += v[gi][src][1].w;
                  }
                  if (r16) {
                    float2 f;
                    f = unpack_bf162(rr[gi].x); o[0] += f.x; o[1] += f.y;
                    f = unpack_bf162(rr[gi].y); o[2] += f.x; o[3] += f.y;
                    f = unpack_bf162(rr[gi].z); o[4] += f.x; o[5] += f.y;
                    f = unpack_bf162(rr[gi].w); o[6] += f.x; o[7] += f.y;
                  }
                  store8_rest(p, o, gm, bidx, n, smem_rowbias, r16);
                }
              }
            }
          }
        };
        if (S == 2) finish(std::integral_constant<int, 2>{});
        else if (S == 4) finish(std::integral_constant<int, 4>{});
        else finish(std::integral_constant<int, 8>{});
      }
      if (threadIdx.x == 64) VN_STAMP(9);
    }
  }

  tc_fence_before();
  if (MC > 1 || CG == 2) cluster_sync_all();       // no CTA leaves while a peer may still signal its barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
#undef VN_TILE_OF
#ifdef VN_TIMELINE
  if (p.dbg && threadIdx.x == 0) { VN_STAMP(10); p.dbg[(long long)blockIdx.x * 16 + 11] = gtimer_ns(); }
#endif
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box) {
  return vn_make_map(m, base, rank, dims, strides_bytes, box);
}

int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int STAGES, bool SPLIT, int CG = 1>
constexpr int smem_bytes() {
  return STAGES * (BM * BK * 2 + (BN / CG) * BK * 2) + (SPLIT ? (split_exch_own<BN, STAGES>() ? BM * BN * 4 : 0) : BM * BN * 2) + BN * 4 +
         (2 * STAGES + 6) * 8 + 16 + 1024;
}

template <int BN, int STAGES, bool SPLIT, int MC = 1, int CG = 1>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tr,
           const GemmParams& p, int grid_x, cudaStream_t st) {
  static bool configured = false;
  static int max_clusters = 0;
  constexpr int smem = smem_bytes<BN, STAGES, SPLIT, CG>();
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kern = vn_gemm_kernel<BN, STAGES, SPLIT, MC, CG>;
  if (!configured) {
    VN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (MC > 1) {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3((unsigned)(MC * 64), 1, 1);
      q.blockDim = dim3(SPLIT ? kThreadsSplit : kThreadsPers, 1, 1);
      q.dynamicSmemBytes = smem;
      cudaLaunchAttribute a[1];
      a[0].id = cudaLaunchAttributeClusterDimension;
      a[0].val.clusterDim.x = MC; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
      q.attrs = a; q.numAttrs = 1;
      VN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &q));
      VN_CHECK(max_clusters > 0, "vn_gemm: no cluster of %d CTAs fits on this device", MC);
    }
    configured = true;
  }
  if (MC > 1) {                                  // persistent clusters: never more than can be co-resident
    int clusters = grid_x / MC;
    if (clusters > max_clusters) clusters = max_clusters;
    grid_x = clusters * MC;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid_x, 1, 1);
  cfg.blockDim = dim3(SPLIT ? kThreadsSplit : kThreadsPers, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  cfg.attrs = attr;
  int na = 0;
  if (vn_pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // PDL, see vn_launch_pdl
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (SPLIT || MC > 1 || CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = SPLIT ? (unsigned)p.splits : CG == 2 ? 2u : (unsigned)MC;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.numAttrs = na;
  VN_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, td, tr, p));
  vn_count_launch();
  return 0;
}

// Pick (BN, splits) with a small cycle model: persistent tiles are processed in rounds over the SMs; a cluster
// split divides the k-loop of every tile by S but needs all tiles x S CTAs resident at once.
void choose_tiling(int m_tiles, int N, int kb_total, bool allow_split, int* bn_out, int* split_out) {
  const int sms = num_sms();
  const int cands[3] = {256, 128, 64};
  double best = 1e30;
  int best_bn = 128, best_split = 1;
  for (int ci = 0; ci < 3; ++ci) {
    const int bn = cands[ci];
    const int n_tiles = vn_cdiv(N, bn);
    const int tiles = m_tiles * n_tiles;
    // cycles per 64-deep k-block: tensor floor vs. the L2 -> SM operand stream (~40 B/cycle/SM measured)
    const double t_mma = bn == 256 ? 512.0 : bn == 128 ? 256.0 : 192.0;
    const double t_mem = (BM + bn) * 128.0 / 40.0;
    const double tk = t_mma > t_mem ? t_mma : t_mem;
    const double t_epi = 300.0 + bn * 8.0;
    {
      const int rounds = vn_cdiv(tiles, sms);
      const double cost = rounds * (kb_total * tk + 300.0) + t_epi + 3000.0;
      if (cost < best) { best = cost; best_bn = bn; best_split = 1; }
    }
    if (!allow_split) continue;
    for (int s = 2; s <= 8; s *= 2) {
      if (tiles * s > sms) break;
      const int kps = vn_cdiv(kb_total, s);
      if (kps < 2 || kb_total <= (s - 1) * kps) continue;       // every rank needs at least one k-block
      const double cost = kps * tk + 2500.0 + t_epi / s + 3500.0;
      if (cost < best) { best = cost; best_bn = bn; best_split = s; }
    }
  }
  *bn_out = best_bn;
  *split_out = best_split;
}

}  // namespace

extern "C" size_t vn_gemm_workspace_bytes(int max_M, int max_N) {
  (void)max_M; (void)max_N;
  return 256;       // split-K reduces through distributed shared memory; kept for ABI stability
}

extern "C" int vn_gemm(const vn_gemm_desc* d, vn_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VN_CHECK(d != nullptr, "vn_gemm: null descriptor");
  VN_CHECK(d->M > 0 && d->N > 0 && d->K > 0, "vn_gemm: empty problem M=%d N=%d K=%d", d->M, d->N, d->K);
  VN_CHECK(d->K % BK == 0, "vn_gemm: K=%d must be a multiple of 64", d->K);
  VN_CHECK(d->N % 8 == 0, "vn_gemm: N=%d must be a multiple of 8", d->N);
  VN_CHECK(d->lda % 8 == 0 && d->ldb % 8 == 0 && d->ldd % 8 == 0, "vn_gemm: lda/ldb/ldd must be multiples of 8");
  VN_CHECK(d->ldb >= d->K, "vn_gemm: ldb < K");
  VN_CHECK(!d->R || d->r_fp32 || d->ldr % 8 == 0, "vn_gemm: ldr must be a multiple of 8");
  VN_CHECK((reinterpret_cast<uintptr_t>(d->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->B) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(d->D) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->R) & 15) == 0,
           "vn_gemm: A/B/D/R must be 16-byte aligned");

  GemmParams p{};
  p.M = d->M; p.N = d->N;
  p.kb_total = d->K / BK;
  p.mode = d->mode ? 1 : 0;                       // the kernel knows linear (0) and conv (1); stride / pad are parameters
  p.D = d->D; p.ldd = d->ldd;
  p.bias = d->bias;
  p.rowbias = d->rowbias; p.ld_rowbias = d->ld_rowbias; p.rows_per_batch = d->rows_per_batch;
  p.R = reinterpret_cast<const bf16*>(d->R); p.ldr = d->ldr;
  p.out_fp32 = d->out_fp32;
  p.r_fp32 = d->R ? d->r_fp32 : 0;
  VN_CHECK(!p.r_fp32 || d->out_fp32, "vn_gemm: an fp32 residual needs an fp32 output (direct epilogue)");
  VN_CHECK(!p.r_fp32 || d->ldr % 4 == 0, "vn_gemm: ldr of an fp32 residual must be a multiple of 4");

  CUtensorMap ta, tb, td, tr;
  memset(&td, 0, sizeof(td));
  memset(&tr, 0, sizeof(tr));
  cuuint32_t box_a[4];
  int out_h = 0, out_w = 0;                       // conv: output grid (== input dims at stride 1)
  if (d->mode == 0) {
    VN_CHECK(d->lda >= d->K, "vn_gemm: lda < K");
    cuuint64_t dims[2] = {(cuuint64_t)d->K, (cuuint64_t)d->M};
    cuuint64_t str[1] = {(cuuint64_t)d->lda * 2};
    box_a[0] = BK; box_a[1] = BM;
    if (make_map(&ta, d->A, 2, dims, str, box_a)) return -1;
    p.m_tiles = vn_cdiv(d->M, BM);
    p.rows_a = BM;
  } else if (d->mode >= 1 && d->mode <= 3) {
    // mode 1: 3x3 / stride 1 / pad 1.  mode 2: stride 2 / pad 1 (diffusers Downsample2D).  mode 3: stride 2, no leading pad,
    // zero beyond the far edge (the VAE encoder's F.pad(x, (0,1,0,1)) + pad-0 conv).  H, W of the descriptor are the INPUT
    // dims; the tile geometry runs over the output grid and the A map traverses the input with element strides
    // (experiments/tma_stride2_probe.cu: box dims count source elements, the tw*th loaded pixels are compacted).
    const int cs = d->mode == 1 ? 1 : 2, cpad = d->mode == 3 ? 0 : 1;
    const int Ho = cs == 1 ? d->H : (cpad ? (d->H - 1) / 2 + 1 : (d->H - 2) / 2 + 1);
    const int Wo = cs == 1 ? d->W : (cpad ? (d->W - 1) / 2 + 1 : (d->W - 2) / 2 + 1);
    VN_CHECK(d->C % BK == 0 && d->K == 9 * d->C, "vn_gemm conv: need C %% 64 == 0 and K == 9*C (C=%d K=%d)", d->C, d->K);
    VN_CHECK(Ho >= 1 && Wo >= 1 && d->M == d->nb * Ho * Wo, "vn_gemm conv: M != nb*Ho*Wo (M=%d Ho=%d Wo=%d)", d->M, Ho, Wo);
    VN_CHECK(d->lda >= d->C, "vn_gemm conv: pixel stride < C");
    int tw = 1;
    while (tw < 64 && Wo % (tw * 2) == 0) tw *= 2;
    int th = BM / tw;
    while (th > 1 && th / 2 >= Ho) th /= 2;       // do not fetch far more rows than the image has
    VN_CHECK(cs * tw <= 256 && cs * th <= 256, "vn_gemm conv: tile %d x %d too large for a stride-%d box", tw, th, cs);
    p.tw = tw; p.th = th;
    p.tiles_w = vn_cdiv(Wo, tw);
    p.tiles_h = vn_cdiv(Ho, th);
    p.rows_a = tw * th;
    p.H = Ho; p.W = Wo; p.cblocks = d->C / BK;
    p.cs = cs; p.cpad = cpad;
    out_h = Ho; out_w = Wo;
    cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->nb};
    cuuint64_t str[3] = {(cuuint64_t)d->lda * 2, (cuuint64_t)d->lda * 2 * d->W, (cuuint64_t)d->lda * 2 * d->W * d->H};
    cuuint32_t es[4] = {1, (cuuint32_t)cs, (cuuint32_t)cs, 1};
    box_a[0] = BK; box_a[1] = (cuuint32_t)(cs * tw); box_a[2] = (cuuint32_t)(cs * th); box_a[3] = 1;
    if (vn_make_map(&ta, d->A, 4, dims, str, box_a, es)) return -1;
    p.m_tiles = d->nb * p.tiles_w * p.tiles_h;
  } else {
    VN_CHECK(false, "vn_gemm: unknown mode %d", d->mode);
  }

  int bn = 128, splits = 1;
  choose_tiling(p.m_tiles, d->N, p.kb_total, true, &bn, &splits);
  // force_split: low 4 bits = split-K cluster size (0 = auto), bits 4..7 = multicast cluster size MC (0 = auto),
  // bits 8..9 = CTA pairs (cta_group::2): 1 = on, 2 = off, 0 = auto
  const int force_s = d->force_split & 15, force_mc = (d->force_split >> 4) & 15, force_cg = (d->force_split >> 8) & 3;
  if (d->force_bn) bn = d->force_bn;
  if (force_s) splits = force_s;
  VN_CHECK(bn == 64 || bn == 128 || bn == 160 || bn == 192 || bn == 256, "vn_gemm: unsupported BN %d (64, 128, 160, 192, 256)", bn);
  VN_CHECK(splits == 1 || splits == 2 || splits == 4 || splits == 8, "vn_gemm: unsupported split %d (1, 2, 4, 8)", splits);
  // BN = 160 (= N/2, N/4, N/8 of the 320 / 640 / 1280-wide convolutions: no ragged last n-tile) exists in the split-K
  // schedule only (the staged epilogue needs 64-column blocks); every finishing thread owns a multiple of 8 columns
  if (bn == 160 && splits == 1) splits = 2;
  while (splits > 1 && (bn / splits) % 8 != 0) splits /= 2;
  while (splits > 1 && (vn_cdiv(p.kb_total, splits) < 1 || p.kb_total <= (splits - 1) * vn_cdiv(p.kb_total, splits)))
    splits /= 2;
  // (checked AFTER the clamp to the number of k-blocks: a forced BN 160 on a one-k-block product used to reach the split
  //  kernel with a cluster of 1 and hang on a barrier nobody arms)
  VN_CHECK(!(bn == 160 && splits == 1), "vn_gemm: BN 160 needs a split-K cluster of 2 or 4 (and at least 2 k-blocks)");
  p.kb_per_split = vn_cdiv(p.kb_total, splits);
  p.splits = splits;
  p.n_tiles = vn_cdiv(d->N, bn);
  p.nbimg = d->mode != 0 ? d->nb : 1;
  p.dbg = vn_debug_buffer();
  {
    // weight look-ahead: few m-tiles stream a big weight matrix (>= 4 MB) - keep ~256 KB per CTA requested ahead
    static int pf_env = -1;
    if (pf_env < 0) { const char* e = getenv("VN_GEMM_LOOKAHEAD"); pf_env = e ? atoi(e) : 1; }
    p.pf_dist = 0;
    if (pf_env && p.m_tiles <= 2 && (long long)d->N * d->K * 2 >= (4ll << 20)) {
      int dist = 262144 / (bn * BK * 2);
      p.pf_dist = dist < 8 ? 8 : dist > 48 ? 48 : dist;
    }
  }
  // multicast clusters along M (persistent schedule only): every CTA of a cluster works on the same n-tile
  int mc = 1;
  if (splits == 1) {
    // Measured on B200 (scripts/gemm_bench.py, FORCE_SPLIT=1+16*MC): multicast does not change the time of any layer
    // shape - the tiles are bound by the bytes LANDING in an SM per k-block (A + B either way), not by the requests
    // it issues - so it is opt-in only (force_split bits 4..7) and the default is MC = 1.
    mc = force_mc ? force_mc : 1;
    VN_CHECK(mc == 1 || mc == 2 || mc == 4, "vn_gemm: unsupported multicast cluster size %d (1, 2, 4)", mc);
    if (mc > p.m_tiles) mc = p.m_tiles >= 2 ? 2 : 1;
    p.m_tiles = vn_cdiv(p.m_tiles, mc) * mc;       // padded slots decode to out-of-range tiles (TMA zero-fill / clipping)
  }
  // CTA pairs (persistent schedule): two m-tiles of one n-tile share every B tile through cta_group::2 MMAs - each SM
  // receives A + B/2 per k-block and the pair's tensor cores work on one 256 x BN tile
  int cg = 1;
  if (splits == 1 && mc == 1 && bn >= 128 && bn != 160 && p.m_tiles >= 2 && p.m_tiles % 2 == 0 && kProducers == 2) {
    const bool auto_on = false;
    cg = force_cg == 1 ? 2 : force_cg == 2 ? 1 : (auto_on ? 2 : 1);
  }
  const int tiles = p.m_tiles * p.n_tiles;
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->K, (cuuint64_t)d->N};
    cuuint64_t str[1] = {(cuuint64_t)d->ldb * 2};
    cuuint32_t box[2] = {BK, (cuuint32_t)(bn / mc / cg)};
    if (make_map(&tb, d->B, 2, dims, str, box)) return -1;
  }
  // staged TMA-store epilogue for bf16 outputs of the persistent schedule
  p.use_tma_epilogue = (splits == 1 && !d->out_fp32) ? 1 : 0;
  if (p.use_tma_epilogue) {
    for (int which = 0; which < 2; ++which) {
      const void* base = which == 0 ? d->D : d->R;
      const long long ld = which == 0 ? d->ldd : d->ldr;
      if (!base) continue;
      CUtensorMap* m = which == 0 ? &td : &tr;
      if (d->mode == 0) {
        cuuint64_t dims[2] = {(cuuint64_t)d->N, (cuuint64_t)d->M};
        cuuint64_t str[1] = {(cuuint64_t)ld * 2};
        cuuint32_t box[2] = {64, BM};
        if (make_map(m, base, 2, dims, str, box)) return -1;
      } else {
        cuuint64_t dims[4] = {(cuuint64_t)d->N, (cuuint64_t)out_w, (cuuint64_t)out_h, (cuuint64_t)d->nb};
        cuuint64_t str[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * out_w, (cuuint64_t)ld * 2 * out_w * out_h};
        cuuint32_t box[4] = {64, (cuuint32_t)p.tw, (cuuint32_t)p.th, 1};
        if (make_map(m, base, 4, dims, str, box)) return -1;
      }
    }
  }
  if (splits == 1 && cg == 2) {
    const int pairs = tiles / 2 < num_sms() / 2 ? tiles / 2 : num_sms() / 2;
    if (bn == 256) return launch<256, 5, false, 1, 2>(ta, tb, td, tr, p, 2 * pairs, st);
    if (bn == 192) return launch<192, 6, false, 1, 2>(ta, tb, td, tr, p, 2 * pairs, st);
    return launch<128, 7, false, 1, 2>(ta, tb, td, tr, p, 2 * pairs, st);
  }
  if (splits == 1) {
    int grid = tiles < num_sms() ? tiles : num_sms();
    grid = (grid / mc) * mc;
#define VN_GEMM_CASE(BN_, ST_)                                                               \
  case BN_:                                                                                  \
    if (mc == 4) return launch<BN_, ST_, false, 4>(ta, tb, td, tr, p, grid, st);             \
    if (mc == 2) return launch<BN_, ST_, false, 2>(ta, tb, td, tr, p, grid, st);             \
    return launch<BN_, ST_, false, 1>(ta, tb, td, tr, p, grid, st);
    if (mc > 1) VN_CHECK(bn != 192, "vn_gemm: multicast clusters exist for BN 64 / 128 / 256 only");
    switch (bn) {
      VN_GEMM_CASE(64, 6)
      VN_GEMM_CASE(128, 5)
      case 192: return launch<192, 4, false, 1>(ta, tb, td, tr, p, grid, st);
      default:
        if (mc == 4) return launch<256, 3, false, 4>(ta, tb, td, tr, p, grid, st);
        if (mc == 2) return launch<256, 3, false, 2>(ta, tb, td, tr, p, grid, st);
        return launch<256, 3, false, 1>(ta, tb, td, tr, p, grid, st);
    }
#undef VN_GEMM_CASE
  }
  const int grid = tiles * splits;
  switch (bn) {
    case 64: return launch<64, 6, true>(ta, tb, td, tr, p, grid, st);
    case 128: return launch<128, 5, true>(ta, tb, td, tr, p, grid, st);
    case 160: return launch<160, 5, true>(ta, tb, td, tr, p, grid, st);
    case 192: return launch<192, 4, true>(ta, tb, td, tr, p, grid, st);
    default: return launch<256, 4, true>(ta, tb, td, tr, p, grid, st);
  }
}
