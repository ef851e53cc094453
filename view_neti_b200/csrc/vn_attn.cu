// vn_attn.cu — attention core (head_dim 64) forward and backward, logits never leave the SM.
//
// Restates, as one flash-style kernel per direction, reference models/xti_attention_processor.py:44-50:
//   head_to_batch_dim -> get_attention_scores (fp32 logits, baddbmm alpha=scale, softmax) -> bmm -> batch_to_head_dim
// and its autograd backward (training/coach.py:214).  K and V are separate tensors because XTI takes K from
// CONTEXT_TENSOR_i and V from CONTEXT_TENSOR_BYPASS_i (xti_attention_processor.py:38-42).
//
// Data layout: token-major, heads side by side: element (b, n, h, d) at base + b*bs + n*ld + h*64 + d, so the
// head split / merge copies of the reference do not exist.  bf16 operands, fp32 logits / softmax / accumulators.
//
// This file holds the BACKWARD (the forward is the tcgen05 / TMEM kernel in vn_attn_tc.cu): warp-level mma.sync
// m16n8k16 (bf16 -> fp32) with ldmatrix from XOR-swizzled shared memory and a cp.async double buffer.
//   bwd dQ: CTA = 128 queries x 8 warps, loops over 64-key tiles; recomputes P from the saved log-sum-exp.
//   bwd dK/dV: CTA = 64 keys x 4 warps, loops over 64-query tiles; optional split over the query range with fp32
//           atomics into a scratch accumulator when nk is tiny (cross-attention: nk = 77).
#include "vn_common.cuh"

namespace {

constexpr int D = 64;          // head dim
constexpr int ROWB = 128;      // bytes per tile row (64 bf16)
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t sw_addr(uint32_t base, int row, int chunk) {
  return base + row * ROWB + (((chunk ^ row) & 7) << 4);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Cooperative async load of a [ROWS x 64] bf16 tile (rows row0.. of a [*, ld] matrix) into swizzled smem;
// rows >= nrows are zero-filled.
template <int ROWS, int THREADS>
__device__ __forceinline__ void load_tile(uint32_t smem, const bf16* __restrict__ g, long long ld, int row0, int nrows) {
  static_assert((ROWS * 8) % THREADS == 0, "tile chunks must divide evenly over the CTA");
#pragma unroll
  for (int it = 0; it < ROWS * 8 / THREADS; ++it) {
    const int i = threadIdx.x + it * THREADS;
    const int r = i >> 3, c = i & 7;
    const bool ok = (row0 + r) < nrows;
    const bf16* src = g + (long long)(ok ? row0 + r : 0) * ld + c * 8;
    cp_async16(sw_addr(smem, r, c), src, ok);
  }
}

// A fragments (16 rows r0.., all 4 k-chunks) of a [rows x 64] tile
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], uint32_t tile, int r0, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldsm_x4(f[kk], sw_addr(tile, r0 + (lane & 15), kk * 2 + (lane >> 4)));
}

// acc[8][4] (16 x 64 fp32) += A(16 x 64, frags) * T^T where T is a [64 n][64 k] tile (both K-major)
__device__ __forceinline__ void mma_nt(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane) {
  const int mid = lane >> 3;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4(b, sw_addr(tile, np * 16 + (lane & 7) + 8 * (mid >> 1), kk * 2 + (mid & 1)));
      mma16816(acc[2 * np], a[kk], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[kk], b[2], b[3]);
    }
  }
}

// acc[8][4] (16 x 64) += A(16 x 64 over k, frags) * T where T is a [64 k][64 n] tile (n contiguous)
__device__ __forceinline__ void mma_nn(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane) {
  const int mid = lane >> 3;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, sw_addr(tile, kk * 16 + (lane & 7) + 8 * (mid & 1), np * 2 + (mid >> 1)));
      mma16816(acc[2 * np], a[kk], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[kk], b[2], b[3]);
    }
  }
}

// fp32 C fragments (16 x 64) -> bf16 A fragments over the 64 columns
__device__ __forceinline__ void c_to_a(uint32_t (&a)[4][4], const float (&c)[8][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = pack_bf162(c[2 * kk][0], c[2 * kk][1]);
    a[kk][1] = pack_bf162(c[2 * kk][2], c[2 * kk][3]);
    a[kk][2] = pack_bf162(c[2 * kk + 1][0], c[2 * kk + 1][1]);
    a[kk][3] = pack_bf162(c[2 * kk + 1][2], c[2 * kk + 1][3]);
  }
}

struct AttnParams {
  int nb, heads, nq, nk;
  float scale;
  const bf16* q; long long ldq, bsq;
  const bf16* k; long long ldk, bsk;
  const bf16* v; long long ldv, bsv;
  bf16* o; long long ldo, bso;
  float* lse;
  const bf16* d_o; long long lddo, bsdo;
  float* delta;
  bf16* dq; long long lddq, bsdq;
  bf16* dk; long long lddk, bsdk;
  bf16* dv; long long lddv, bsdv;
  double* dkv_acc;
  int qsplits, qtiles_per_split;
};

// tile geometry shared by the backward kernels (the forward lives in vn_attn_tc.cu)
constexpr int FWD_THREADS = 256;
constexpr int FWD_BM = 128;
constexpr int BN = 64;

// =================================================================================================
// backward: delta = rowsum(dO * O)
// =================================================================================================
__global__ void __launch_bounds__(256) attn_delta_kernel(const AttnParams p) {
  // one 16-byte vector per lane: 8 lanes cover one (row, head) = 64 elements, a warp covers 4 consecutive (row, head)
  const int C = p.heads * D;
  const long long nvec = (long long)p.nb * p.nq * p.heads * 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f;
  long long rh = 0;
  const bool ok = i < nvec;
  if (ok) {
    rh = i >> 3;                                   // (b*nq + n)*heads + h
    const int h = (int)(rh % p.heads);
    const long long bn = rh / p.heads;
    const int n = (int)(bn % p.nq), b = (int)(bn / p.nq);
    const int c = h * D + (int)(i & 7) * 8;
    const uint4 ov = *reinterpret_cast<const uint4*>(p.o + (long long)b * p.bso + (long long)n * p.ldo + c);
    const uint4 dv = *reinterpret_cast<const uint4*>(p.d_o + (long long)b * p.bsdo + (long long)n * p.lddo + c);
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
    const int h = (int)(rh % p.heads);
    const long long bn = rh / p.heads;
    const int n = (int)(bn % p.nq), b = (int)(bn / p.nq);
    p.delta[((long long)b * p.heads + h) * p.nq + n] = s;
  }
  (void)C;
}

// =================================================================================================
// backward: dQ
// =================================================================================================
constexpr int DQ_SMEM = 2 * FWD_BM * ROWB + 2 * 2 * BN * ROWB;   // Q, dO, 2 x (K, V) = 64 KB

__global__ void __launch_bounds__(FWD_THREADS) attn_bwd_dq_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sQ = smem_u32(smem_raw);
  const uint32_t sdO = sQ + FWD_BM * ROWB;
  const uint32_t sKV = sdO + FWD_BM * ROWB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * FWD_BM, h = blockIdx.y, b = blockIdx.z;
  const bf16* gq = p.q + (long long)b * p.bsq + h * D;
  const bf16* gdo = p.d_o + (long long)b * p.bsdo + h * D;
  const bf16* gk = p.k + (long long)b * p.bsk + h * D;
  const bf16* gv = p.v + (long long)b * p.bsv + h * D;
  const int ntiles = (p.nk + BN - 1) / BN;

  load_tile<FWD_BM, FWD_THREADS>(sQ, gq, p.ldq, q0, p.nq);
  load_tile<FWD_BM, FWD_THREADS>(sdO, gdo, p.lddo, q0, p.nq);
  load_tile<BN, FWD_THREADS>(sKV, gk, p.ldk, 0, p.nk);
  load_tile<BN, FWD_THREADS>(sKV + BN * ROWB, gv, p.ldv, 0, p.nk);
  cp_async_commit();

  const int row0 = q0 + warp * 16 + (lane >> 2), row1 = row0 + 8;
  const long long sidx = ((long long)b * p.heads + h) * p.nq;
  const float sl2 = p.scale * kLog2e;
  // invalid rows: lse = +inf -> P = 0
  const float lse0 = row0 < p.nq ? p.lse[sidx + row0] * kLog2e : INFINITY;
  const float lse1 = row1 < p.nq ? p.lse[sidx + row1] * kLog2e : INFINITY;
  const float dl0 = row0 < p.nq ? p.delta[sidx + row0] : 0.f;
  const float dl1 = row1 < p.nq ? p.delta[sidx + row1] : 0.f;

  uint32_t qf[4][4], dof[4][4];
  float dq[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f; }

  for (int j = 0; j < ntiles; ++j) {
    if (j + 1 < ntiles) {
      const uint32_t nb_ = sKV + ((j + 1) & 1) * 2 * BN * ROWB;
      load_tile<BN, FWD_THREADS>(nb_, gk, p.ldk, (j + 1) * BN, p.nk);
      load_tile<BN, FWD_THREADS>(nb_ + BN * ROWB, gv, p.ldv, (j + 1) * BN, p.nk);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (j == 0) { load_a_frags(qf, sQ, warp * 16, lane); load_a_frags(dof, sdO, warp * 16, lane); }
    const uint32_t sK = sKV + (j & 1) * 2 * BN * ROWB;
    const uint32_t sV = sK + BN * ROWB;

    float s[8][4], dp[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f;
      dp[t][0] = dp[t][1] = dp[t][2] = dp[t][3] = 0.f;
    }
    mma_nt(s, qf, sK, lane);
    mma_nt(dp, dof, sV, lane);
    const int kbase = j * BN + 2 * (lane & 3);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int kc = kbase + t * 8;
      const bool ok0 = kc < p.nk, ok1 = kc + 1 < p.nk;
      const float p00 = ok0 ? exp2f(s[t][0] * sl2 - lse0) : 0.f;
      const float p01 = ok1 ? exp2f(s[t][1] * sl2 - lse0) : 0.f;
      const float p10 = ok0 ? exp2f(s[t][2] * sl2 - lse1) : 0.f;
      const float p11 = ok1 ? exp2f(s[t][3] * sl2 - lse1) : 0.f;
      s[t][0] = p00 * (dp[t][0] - dl0) * p.scale;
      s[t][1] = p01 * (dp[t][1] - dl0) * p.scale;
      s[t][2] = p10 * (dp[t][2] - dl1) * p.scale;
      s[t][3] = p11 * (dp[t][3] - dl1) * p.scale;
    }
    uint32_t dsf[4][4];
    c_to_a(dsf, s);
    mma_nn(dq, dsf, sK, lane);
    __syncthreads();
  }
  cp_async_wait<0>();

  bf16* gdq = p.dq + (long long)b * p.bsdq + h * D + 2 * (lane & 3);
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    if (row0 < p.nq) *reinterpret_cast<uint32_t*>(gdq + (long long)row0 * p.lddq + t * 8) = pack_bf162(dq[t][0], dq[t][1]);
    if (row1 < p.nq) *reinterpret_cast<uint32_t*>(gdq + (long long)row1 * p.lddq + t * 8) = pack_bf162(dq[t][2], dq[t][3]);
  }
}

// =================================================================================================
// backward: dK, dV
// =================================================================================================
constexpr int DKV_THREADS = 128;
constexpr int DKV_BK = 64;       // keys per CTA
constexpr int DKV_BQ = 64;       // queries per inner tile
constexpr int DKV_STAGE = 2 * DKV_BQ * ROWB + 2 * DKV_BQ * 4;     // Q, dO, lse, delta
constexpr int DKV_SMEM = 2 * DKV_BK * ROWB + 2 * DKV_STAGE;

template <bool ATOMIC>
__global__ void __launch_bounds__(DKV_THREADS) attn_bwd_dkv_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sK = smem_u32(smem_raw);
  const uint32_t sV = sK + DKV_BK * ROWB;
  const uint32_t sStage = sV + DKV_BK * ROWB;
  uint8_t* stage_ptr = smem_raw + 2 * DKV_BK * ROWB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k0 = blockIdx.x * DKV_BK, h = blockIdx.y;
  const int b = blockIdx.z / p.qsplits, split = blockIdx.z % p.qsplits;
  const bf16* gq = p.q + (long long)b * p.bsq + h * D;
  const bf16* gdo = p.d_o + (long long)b * p.bsdo + h * D;
  const bf16* gk = p.k + (long long)b * p.bsk + h * D;
  const bf16* gv = p.v + (long long)b * p.bsv + h * D;
  const long long sidx = ((long long)b * p.heads + h) * p.nq;
  const int total_qt = (p.nq + DKV_BQ - 1) / DKV_BQ;
  const int qt_begin = split * p.qtiles_per_split;
  const int qt_end = min(total_qt, qt_begin + p.qtiles_per_split);
  const int ntiles = qt_end - qt_begin;
  if (ntiles <= 0) return;
  const float sl2 = p.scale * kLog2e;

  auto issue_stage = [&](int st, int qt) {
    const uint32_t base = sStage + st * DKV_STAGE;
    load_tile<DKV_BQ, DKV_THREADS>(base, gq, p.ldq, qt * DKV_BQ, p.nq);
    load_tile<DKV_BQ, DKV_THREADS>(base + DKV_BQ * ROWB, gdo, p.lddo, qt * DKV_BQ, p.nq);
    float* sl = reinterpret_cast<float*>(stage_ptr + st * DKV_STAGE + 2 * DKV_BQ * ROWB);
    if (threadIdx.x < DKV_BQ) {
      const int qq = qt * DKV_BQ + threadIdx.x;
      sl[threadIdx.x] = qq < p.nq ? p.lse[sidx + qq] * kLog2e : INFINITY;     // +inf -> P = 0 for padded queries
      sl[DKV_BQ + threadIdx.x] = qq < p.nq ? p.delta[sidx + qq] : 0.f;
    }
  };

  load_tile<DKV_BK, DKV_THREADS>(sK, gk, p.ldk, k0, p.nk);
  load_tile<DKV_BK, DKV_THREADS>(sV, gv, p.ldv, k0, p.nk);
  issue_stage(0, qt_begin);
  cp_async_commit();

  uint32_t kf[4][4], vf[4][4];
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }

  for (int j = 0; j < ntiles; ++j) {
    if (j + 1 < ntiles) issue_stage((j + 1) & 1, qt_begin + j + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (j == 0) { load_a_frags(kf, sK, warp * 16, lane); load_a_frags(vf, sV, warp * 16, lane); }
    const uint32_t sQ = sStage + (j & 1) * DKV_STAGE;
    const uint32_t sdO = sQ + DKV_BQ * ROWB;
    const float* sl = reinterpret_cast<const float*>(stage_ptr + (j & 1) * DKV_STAGE + 2 * DKV_BQ * ROWB);

    float st[8][4], dpt[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      st[t][0] = st[t][1] = st[t][2] = st[t][3] = 0.f;
      dpt[t][0] = dpt[t][1] = dpt[t][2] = dpt[t][3] = 0.f;
    }
    mma_nt(st, kf, sQ, lane);        // S^T = K Q^T      [16 keys x 64 queries]
    mma_nt(dpt, vf, sdO, lane);      // dP^T = V dO^T
    // P^T = exp2(S^T * sl2 - lse[q]);  columns are queries
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int qc = t * 8 + 2 * (lane & 3);
      const float ls0 = sl[qc], ls1 = sl[qc + 1];
      st[t][0] = exp2f(st[t][0] * sl2 - ls0);
      st[t][1] = exp2f(st[t][1] * sl2 - ls1);
      st[t][2] = exp2f(st[t][2] * sl2 - ls0);
      st[t][3] = exp2f(st[t][3] * sl2 - ls1);
    }
    uint32_t af[4][4];
    c_to_a(af, st);
    mma_nn(dv, af, sdO, lane);       // dV += P^T dO
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int qc = t * 8 + 2 * (lane & 3);
      const float d0 = sl[DKV_BQ + qc], d1 = sl[DKV_BQ + qc + 1];
      st[t][0] = st[t][0] * (dpt[t][0] - d0) * p.scale;
      st[t][1] = st[t][1] * (dpt[t][1] - d1) * p.scale;
      st[t][2] = st[t][2] * (dpt[t][2] - d0) * p.scale;
      st[t][3] = st[t][3] * (dpt[t][3] - d1) * p.scale;
    }
    c_to_a(af, st);
    mma_nn(dk, af, sQ, lane);        // dK += dS^T Q
    __syncthreads();
  }
  cp_async_wait<0>();

  const int row0 = k0 + warp * 16 + (lane >> 2), row1 = row0 + 8;
  const int col = h * D + 2 * (lane & 3);
  if (ATOMIC) {
    const long long C = (long long)p.heads * D;
    double* ak = p.dkv_acc + ((long long)b * p.nk) * C;
    double* av = p.dkv_acc + ((long long)p.nb * p.nk) * C + ((long long)b * p.nk) * C;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (row0 < p.nk) {
        atomicAdd(ak + row0 * C + col + t * 8, (double)dk[t][0]); atomicAdd(ak + row0 * C + col + t * 8 + 1, (double)dk[t][1]);
        atomicAdd(av + row0 * C + col + t * 8, (double)dv[t][0]); atomicAdd(av + row0 * C + col + t * 8 + 1, (double)dv[t][1]);
      }
      if (row1 < p.nk) {
        atomicAdd(ak + row1 * C + col + t * 8, (double)dk[t][2]); atomicAdd(ak + row1 * C + col + t * 8 + 1, (double)dk[t][3]);
        atomicAdd(av + row1 * C + col + t * 8, (double)dv[t][2]); atomicAdd(av + row1 * C + col + t * 8 + 1, (double)dv[t][3]);
      }
    }
  } else {
    bf16* gdk = p.dk + (long long)b * p.bsdk + col;
    bf16* gdv = p.dv + (long long)b * p.bsdv + col;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (row0 < p.nk) {
        *reinterpret_cast<uint32_t*>(gdk + (long long)row0 * p.lddk + t * 8) = pack_bf162(dk[t][0], dk[t][1]);
        *reinterpret_cast<uint32_t*>(gdv + (long long)row0 * p.lddv + t * 8) = pack_bf162(dv[t][0], dv[t][1]);
      }
      if (row1 < p.nk) {
        *reinterpret_cast<uint32_t*>(gdk + (long long)row1 * p.lddk + t * 8) = pack_bf162(dk[t][2], dk[t][3]);
        *reinterpret_cast<uint32_t*>(gdv + (long long)row1 * p.lddv + t * 8) = pack_bf162(dv[t][2], dv[t][3]);
      }
    }
  }
}

// fp32 scratch -> bf16 dk / dv, leaving the scratch zeroed for the next launch
__global__ void __launch_bounds__(256) attn_dkv_finish_kernel(const AttnParams p) {
  const int C = p.heads * D;
  const long long per = (long long)p.nb * p.nk * C;
  const long long total = per / 2;       // bf16 pairs per tensor
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

int fill_params(const vn_attn_desc* d, AttnParams* p, bool bwd) {
  VN_CHECK(d != nullptr, "attention: null descriptor");
  VN_CHECK(d->nb > 0 && d->heads > 0 && d->nq > 0 && d->nk > 0, "attention: empty problem");
  VN_CHECK(d->ldq % 8 == 0 && d->ldk % 8 == 0 && d->ldv % 8 == 0 && d->ldo % 8 == 0 && d->bsq % 8 == 0 &&
               d->bsk % 8 == 0 && d->bsv % 8 == 0 && d->bso % 8 == 0,
           "attention: strides must be multiples of 8 elements");
  VN_CHECK(((reinterpret_cast<uintptr_t>(d->q) | reinterpret_cast<uintptr_t>(d->k) | reinterpret_cast<uintptr_t>(d->v) |
             reinterpret_cast<uintptr_t>(d->o)) & 15) == 0, "attention: q/k/v/o must be 16-byte aligned");
  p->nb = d->nb; p->heads = d->heads; p->nq = d->nq; p->nk = d->nk; p->scale = d->scale;
  p->q = (const bf16*)d->q; p->ldq = d->ldq; p->bsq = d->bsq;
  p->k = (const bf16*)d->k; p->ldk = d->ldk; p->bsk = d->bsk;
  p->v = (const bf16*)d->v; p->ldv = d->ldv; p->bsv = d->bsv;
  p->o = (bf16*)d->o; p->ldo = d->ldo; p->bso = d->bso;
  p->lse = d->lse;
  p->qsplits = 1; p->qtiles_per_split = 1 << 30;
  if (bwd) {
    VN_CHECK(d->lse && d->delta && d->d_o && d->dk && d->dv, "attention bwd: lse, delta, d_o, dk, dv are required");
    VN_CHECK(d->lddo % 8 == 0 && d->bsdo % 8 == 0 && d->lddk % 8 == 0 && d->lddv % 8 == 0 && d->bsdk % 8 == 0 &&
                 d->bsdv % 8 == 0 && (!d->dq || (d->lddq % 8 == 0 && d->bsdq % 8 == 0)),
             "attention bwd: strides must be multiples of 8 elements");
    p->d_o = (const bf16*)d->d_o; p->lddo = d->lddo; p->bsdo = d->bsdo;
    p->delta = d->delta;
    p->dq = (bf16*)d->dq; p->lddq = d->lddq; p->bsdq = d->bsdq;
    p->dk = (bf16*)d->dk; p->lddk = d->lddk; p->bsdk = d->bsdk;
    p->dv = (bf16*)d->dv; p->lddv = d->lddv; p->bsdv = d->bsdv;
    p->dkv_acc = d->dkv_acc;
  }
  return 0;
}

template <typename K>
int set_smem(K kernel, int bytes) {
  VN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

}  // namespace

extern "C" int vn_attention_bwd(const vn_attn_desc* d, vn_stream_t s) {
  AttnParams p{};
  if (fill_params(d, &p, true)) return -1;
  cudaStream_t st = (cudaStream_t)s;
  static bool configured = false;
  if (!configured) {
    if (set_smem(attn_bwd_dq_kernel, DQ_SMEM)) return -2;
    if (set_smem(attn_bwd_dkv_kernel<false>, DKV_SMEM)) return -2;
    if (set_smem(attn_bwd_dkv_kernel<true>, DKV_SMEM)) return -2;
    configured = true;
  }
  attn_delta_kernel<<<(unsigned)vn_cdiv64((long long)p.nb * p.nq * p.heads * 8, 256), 256, 0, st>>>(p);
  VN_LAUNCH_OK();
  if (p.dq) {
    dim3 grid(vn_cdiv(p.nq, FWD_BM), p.heads, p.nb);
    attn_bwd_dq_kernel<<<grid, FWD_THREADS, DQ_SMEM, st>>>(p);
    VN_LAUNCH_OK();
  }
  const int ktiles = vn_cdiv(p.nk, DKV_BK);
  const int qtiles = vn_cdiv(p.nq, DKV_BQ);
  const long long base_ctas = (long long)ktiles * p.heads * p.nb;
  int splits = 1;
  if (p.dkv_acc && base_ctas < 148) {
    splits = (int)((148 * 2 + base_ctas - 1) / base_ctas);
    if (splits > qtiles) splits = qtiles;
    if (splits < 1) splits = 1;
  }
  p.qtiles_per_split = vn_cdiv(qtiles, splits);
  splits = vn_cdiv(qtiles, p.qtiles_per_split);
  p.qsplits = splits;
  dim3 grid(ktiles, p.heads, p.nb * splits);
  if (splits > 1) {
    attn_bwd_dkv_kernel<true><<<grid, DKV_THREADS, DKV_SMEM, st>>>(p);
    VN_LAUNCH_OK();
    const long long pairs = (long long)p.nb * p.nk * p.heads * D / 2;
    int blocks = (int)vn_cdiv64(pairs, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    attn_dkv_finish_kernel<<<blocks, 256, 0, st>>>(p);
    VN_LAUNCH_OK();
  } else {
    attn_bwd_dkv_kernel<false><<<grid, DKV_THREADS, DKV_SMEM, st>>>(p);
    VN_LAUNCH_OK();
  }
  return 0;
}
