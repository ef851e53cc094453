// vn_gemm.cu — tcgen05 / TMEM / TMA GEMM and implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T (+bias) (+rowbias) (+residual)      bf16 x bf16 -> fp32 (TMEM) -> bf16|fp32
//
// Replaces the cuBLAS / cuDNN calls under diffusers' Linear / Conv2d modules on the reference hot path
// (xti_attention_processor.py:30,38-42,53; ResnetBlock2D conv1/conv2/conv_shortcut; Transformer2DModel
// proj_in/out; FeedForward) and, with pre-transposed weights, their dgrads (coach.py:214).
//
// Kernel shape (one 128 x BN output tile per CTA, optional split-K over gridDim.z):
//   warp 0      : TMA producer   — one elected lane; A tile [128 rows x 64 k] + B tile [BN x 64 k] per stage,
//                                  128B-swizzled, completion on the stage's `full` mbarrier.
//                                  conv mode: A comes from a 4-D NHWC tensor map, one (tap, 64-channel) slab per
//                                  k-block at coordinates (c0, w0+dx-1, h0+dy-1, b); TMA zero-fills the halo,
//                                  so padding costs nothing and no im2col buffer exists.
//   warp 1      : MMA issuer     — allocates TMEM, one lane issues 4 x tcgen05.mma (K=16) per stage,
//                                  tcgen05.commit releases the stage / signals the epilogue.
//   warps 2..5  : epilogue       — tcgen05.ld 32 lanes x 32 columns at a time; fused bias / time-embedding
//                                  row-bias / residual; 16-byte stores.  Split-K: fp32 red.add into a
//                                  self-cleaning workspace, last-arriving CTA of a tile runs the epilogue.
#include "vn_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;

struct GemmParams {
  int M, N;
  int kb_total, kb_per_split, splits;
  int mode;
  int H, W, cblocks, tw, th, tiles_w, tiles_h, rows_a;
  void* D; long long ldd;
  const float* bias;
  const float* rowbias; long long ld_rowbias; int rows_per_batch;
  const bf16* R; long long ldr;
  int out_fp32;
  float* ws; int* counters;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Finish one 32-column chunk of one output row: v[] holds fp32 sums.
__device__ __forceinline__ void epilogue_store(const GemmParams& p, float (&v)[32], long long gm, int bidx, int n_base) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n = n_base + g * 8;
    if (n >= p.N) break;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = v[g * 8 + j];
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
    if (p.R) {
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
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads) vn_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            const GemmParams p) {
  constexpr int A_BYTES = BM * BK * 2;            // 16 KB
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // multiple of 1024 for BN % 8 == 0
  constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024-byte alignment");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  int* flag_slot = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates ----
  const int n0 = blockIdx.y * BN;
  int m0 = 0, img = 0, h0 = 0, w0 = 0;
  if (p.mode == 0) {
    m0 = blockIdx.x * BM;
  } else {
    const int tpi = p.tiles_w * p.tiles_h;
    img = blockIdx.x / tpi;
    const int r = blockIdx.x - img * tpi;
    h0 = (r / p.tiles_w) * p.th;
    w0 = (r % p.tiles_w) * p.tw;
  }
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;   // host guarantees >= 1

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)(p.rows_a * BK * 2 + B_BYTES);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const int round = i / STAGES;
        mbar_wait(&empty_bar[s], (round & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], tx_bytes);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        const int kb = kb_begin + i;
        if (p.mode == 0) {
          tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m0);
        } else {
          const int tap = kb / p.cblocks;
          const int cb = kb - tap * p.cblocks;
          const int dy = tap / 3, dx = tap - dy * 3;
          tma_load_4d(sa, &tmA, &full_bar[s], cb * BK, w0 + dx - 1, h0 + dy - 1, img);
        }
        tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const int round = i / STAGES;
        mbar_wait(&full_bar[s], round & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = umma_desc_k_sw128(sa);
        const uint64_t bdesc = umma_desc_k_sw128(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (i | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);          // frees this smem stage when the MMAs above have read it
      }
      umma_commit(accum_bar);                // accumulator complete
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..5) =================
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;             // tile row == TMEM lane
    long long gm;
    int bidx;
    bool row_ok;
    if (p.mode == 0) {
      gm = (long long)m0 + r;
      row_ok = gm < p.M;
      bidx = p.rows_per_batch > 0 ? (int)(gm / p.rows_per_batch) : 0;
    } else {
      const int ty = r / p.tw, tx = r - ty * p.tw;
      const int h = h0 + ty, w = w0 + tx;
      row_ok = (r < p.rows_a) && (h < p.H) && (w < p.W);
      gm = ((long long)img * p.H + h) * p.W + w;
      bidx = img;
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);

    if (p.splits == 1) {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t raw[32];
        tmem_ld32(taddr + c * 32, raw);
        tmem_ld_wait();
        if (row_ok) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          epilogue_store(p, v, gm, bidx, n0 + c * 32);
        }
      }
    } else {
      // ---- split-K: accumulate partial tile into the zeroed fp32 workspace ----
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t raw[32];
        tmem_ld32(taddr + c * 32, raw);
        tmem_ld_wait();
        if (row_ok) {
          float* wrow = p.ws + gm * p.N + n0 + c * 32;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (n0 + c * 32 + g * 4 < p.N)
              red_add_v4(wrow + g * 4, __uint_as_float(raw[g * 4]), __uint_as_float(raw[g * 4 + 1]),
                         __uint_as_float(raw[g * 4 + 2]), __uint_as_float(raw[g * 4 + 3]));
          }
        }
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        const int tile = blockIdx.y * gridDim.x + blockIdx.x;
        const int old = atomicAdd(&p.counters[tile], 1);
        const int last = (old == p.splits - 1);
        if (last) p.counters[tile] = 0;      // self-reset for the next launch on this stream
        *flag_slot = last;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*flag_slot) {
        __threadfence();
        if (row_ok) {
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            const int nb = n0 + c * 32;
            if (nb >= p.N) break;
            float v[32];
            float* wrow = p.ws + gm * p.N + nb;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
              if (nb + g * 4 < p.N) {
                t = __ldcg(reinterpret_cast<const float4*>(wrow + g * 4));
                __stcg(reinterpret_cast<float4*>(wrow + g * 4), make_float4(0.f, 0.f, 0.f, 0.f));   // leave it clean
              }
              v[g * 4] = t.x; v[g * 4 + 1] = t.y; v[g * 4 + 2] = t.z; v[g * 4 + 3] = t.w;
            }
            epilogue_store(p, v, gm, bidx, nb);
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  VN_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not found (no CUDA driver?)");
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                  box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
           (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return 0;
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

template <int BN, int STAGES>
constexpr int smem_bytes() {
  return STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 1) * 8 + 16 + 1024;
}

template <int BN, int STAGES>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, dim3 grid, cudaStream_t st) {
  static bool configured = false;
  constexpr int smem = smem_bytes<BN, STAGES>();
  if (!configured) {
    VN_CUDA(cudaFuncSetAttribute(vn_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  vn_gemm_kernel<BN, STAGES><<<grid, kThreads, smem, st>>>(ta, tb, p);
  VN_LAUNCH_OK();
  return 0;
}

constexpr size_t kCounterBytes = 64 * 1024;   // 16384 tile counters

// Pick (BN, splits).  Model: a CTA costs  fixed + kblocks * t_kb(BN);  the launch costs waves * that.
void choose_tiling(int m_tiles, int N, int kb_total, bool allow_split, int* bn_out, int* split_out) {
  const int sms = num_sms();
  const int cands[3] = {160, 128, 64};
  double best = 1e30;
  int best_bn = 128, best_split = 1;
  for (int ci = 0; ci < 3; ++ci) {
    const int bn = cands[ci];
    const int n_tiles = vn_cdiv(N, bn);
    const double waste = (double)(n_tiles * bn) / (double)N;
    const double t_kb = (bn >= 128 ? bn : 96 + bn / 4) * 2.0;       // cycles per 64-deep k-block (MMA floor bn/2*4), small tiles are operand-bound
    const double t_epi = 600.0 + bn * 6.0;
    const int max_split = allow_split ? 16 : 1;
    for (int s = 1; s <= max_split; ++s) {
      const int kps = vn_cdiv(kb_total, s);
      if (s > 1 && (kps < 4 || vn_cdiv(kb_total, kps) != s)) continue;
      const long long ctas = (long long)m_tiles * n_tiles * s;
      const int per_sm = bn == 64 ? 2 : 2;
      const double waves = (double)vn_cdiv64(ctas, (long long)sms * per_sm);
      // two co-resident CTAs share one tensor pipe: each wave of 2*sms CTAs takes ~2x the MMA time of one CTA
      const double t_cta = 2500.0 + kps * t_kb * per_sm + t_epi * (s > 1 ? 2.2 : 1.0);
      double cost = waves * t_cta * (0.9 + 0.1 * waste);
      if (ctas < sms) cost *= 1.0;   // under-filled single wave: cost is just t_cta
      if (cost < best) { best = cost; best_bn = bn; best_split = s; }
    }
  }
  *bn_out = best_bn;
  *split_out = best_split;
}

}  // namespace

extern "C" size_t vn_gemm_workspace_bytes(int max_M, int max_N) {
  return kCounterBytes + (size_t)max_M * (size_t)max_N * sizeof(float);
}

extern "C" int vn_gemm(const vn_gemm_desc* d, vn_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VN_CHECK(d != nullptr, "vn_gemm: null descriptor");
  VN_CHECK(d->M > 0 && d->N > 0 && d->K > 0, "vn_gemm: empty problem M=%d N=%d K=%d", d->M, d->N, d->K);
  VN_CHECK(d->K % BK == 0, "vn_gemm: K=%d must be a multiple of 64", d->K);
  VN_CHECK(d->N % 8 == 0, "vn_gemm: N=%d must be a multiple of 8", d->N);
  VN_CHECK(d->lda % 8 == 0 && d->ldb % 8 == 0 && d->ldd % 8 == 0, "vn_gemm: lda/ldb/ldd must be multiples of 8");
  VN_CHECK(d->ldb >= d->K, "vn_gemm: ldb < K");
  VN_CHECK(!d->R || d->ldr % 8 == 0, "vn_gemm: ldr must be a multiple of 8");
  VN_CHECK((reinterpret_cast<uintptr_t>(d->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->B) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(d->D) & 15) == 0,
           "vn_gemm: A/B/D must be 16-byte aligned");

  GemmParams p{};
  p.M = d->M; p.N = d->N;
  p.kb_total = d->K / BK;
  p.mode = d->mode;
  p.D = d->D; p.ldd = d->ldd;
  p.bias = d->bias;
  p.rowbias = d->rowbias; p.ld_rowbias = d->ld_rowbias; p.rows_per_batch = d->rows_per_batch;
  p.R = reinterpret_cast<const bf16*>(d->R); p.ldr = d->ldr;
  p.out_fp32 = d->out_fp32;

  CUtensorMap ta, tb;
  int m_tiles;
  if (d->mode == 0) {
    VN_CHECK(d->lda >= d->K, "vn_gemm: lda < K");
    cuuint64_t dims[2] = {(cuuint64_t)d->K, (cuuint64_t)d->M};
    cuuint64_t str[1] = {(cuuint64_t)d->lda * 2};
    cuuint32_t box[2] = {BK, BM};
    if (make_map(&ta, d->A, 2, dims, str, box)) return -1;
    m_tiles = vn_cdiv(d->M, BM);
    p.rows_a = BM;
  } else if (d->mode == 1) {
    VN_CHECK(d->C % BK == 0 && d->K == 9 * d->C, "vn_gemm conv: need C %% 64 == 0 and K == 9*C (C=%d K=%d)", d->C, d->K);
    VN_CHECK(d->M == d->nb * d->H * d->W, "vn_gemm conv: M != nb*H*W");
    VN_CHECK(d->lda >= d->C, "vn_gemm conv: pixel stride < C");
    int tw = 1;
    while (tw < 64 && d->W % (tw * 2) == 0) tw *= 2;
    int th = BM / tw;
    while (th > 1 && th / 2 >= d->H) th /= 2;     // do not fetch far more rows than the image has
    p.tw = tw; p.th = th;
    p.tiles_w = vn_cdiv(d->W, tw);
    p.tiles_h = vn_cdiv(d->H, th);
    p.rows_a = tw * th;
    p.H = d->H; p.W = d->W; p.cblocks = d->C / BK;
    cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->nb};
    cuuint64_t str[3] = {(cuuint64_t)d->lda * 2, (cuuint64_t)d->lda * 2 * d->W, (cuuint64_t)d->lda * 2 * d->W * d->H};
    cuuint32_t box[4] = {BK, (cuuint32_t)tw, (cuuint32_t)th, 1};
    if (make_map(&ta, d->A, 4, dims, str, box)) return -1;
    m_tiles = d->nb * p.tiles_w * p.tiles_h;
  } else {
    VN_CHECK(false, "vn_gemm: unknown mode %d", d->mode);
  }

  const bool have_ws = d->workspace != nullptr && d->workspace_bytes >= vn_gemm_workspace_bytes(d->M, d->N);
  int bn = 128, splits = 1;
  choose_tiling(m_tiles, d->N, p.kb_total, have_ws, &bn, &splits);
  if (d->force_bn) bn = d->force_bn;
  if (d->force_split) splits = d->force_split;
  VN_CHECK(bn == 64 || bn == 128 || bn == 160, "vn_gemm: unsupported BN %d", bn);
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = vn_cdiv(p.kb_total, splits);
  splits = vn_cdiv(p.kb_total, p.kb_per_split);
  p.splits = splits;
  const int n_tiles = vn_cdiv(d->N, bn);
  if (splits > 1) {
    VN_CHECK(have_ws, "vn_gemm: split-K needs a workspace of %zu bytes", vn_gemm_workspace_bytes(d->M, d->N));
    VN_CHECK((size_t)m_tiles * n_tiles * sizeof(int) <= kCounterBytes, "vn_gemm: too many tiles for split-K counters");
    p.counters = reinterpret_cast<int*>(d->workspace);
    p.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d->workspace) + kCounterBytes);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->K, (cuuint64_t)d->N};
    cuuint64_t str[1] = {(cuuint64_t)d->ldb * 2};
    cuuint32_t box[2] = {BK, (cuuint32_t)bn};
    if (make_map(&tb, d->B, 2, dims, str, box)) return -1;
  }
  dim3 grid(m_tiles, n_tiles, splits);
  switch (bn) {
    case 64: return launch<64, 4>(ta, tb, p, grid, st);
    case 128: return launch<128, 3>(ta, tb, p, grid, st);
    default: return launch<160, 3>(ta, tb, p, grid, st);
  }
}
