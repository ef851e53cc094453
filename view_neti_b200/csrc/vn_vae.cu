// vn_vae.cu — the one op the SD-2.1 VAE needs beyond the UNet kernels (`vae.encode` every train
// step, reference training/coach.py:165-169; `decode_latents` per image, reference sd_pipeline_call.py:115).
// Its mid-block attention has ONE head over all 512 channels, which does not fit the head_dim-64 tcgen05 attention
// kernels; at one layer per pass it runs as two vn_gemm launches (S = Q K^T in fp32, O = P V) around this row softmax.
// HBM/L2-bound: 4 B read + 2 B written per score, one CTA per row, the row held in registers between the two passes.
#include "vn_common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* s_red) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                                  // s_red may still be read from the previous reduction
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float r = s_red[0];
#pragma unroll
  for (int i = 1; i < kThreads / 32; ++i) r = is_max ? fmaxf(r, s_red[i]) : r + s_red[i];
  return r;
}

// P[r, :] = softmax(scale * S[r, :]) as bf16.  NV float4 per thread: cols <= 4 * kThreads * NV.
template <int NV>
__global__ void __launch_bounds__(kThreads) softmax_rows_kernel(const float* __restrict__ S, long long lds,
                                                                bf16* __restrict__ P, long long ldp, int rows, int cols,
                                                                float scale_log2e) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_red[kThreads / 32];
  const int vecs = cols >> 2;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(S + (long long)r * lds);
    float4 v[NV];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = threadIdx.x + j * kThreads;
      if (i < vecs) {
        v[j] = __ldg(src + i);
        m = fmaxf(m, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
      }
    }
    // scale > 0, so the max of the scaled row is the scaled max
    m = block_reduce(m, true, s_red) * scale_log2e;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = threadIdx.x + j * kThreads;
      if (i < vecs) {
        v[j].x = exp2f(fmaf(v[j].x, scale_log2e, -m)); v[j].y = exp2f(fmaf(v[j].y, scale_log2e, -m));
        v[j].z = exp2f(fmaf(v[j].z, scale_log2e, -m)); v[j].w = exp2f(fmaf(v[j].w, scale_log2e, -m));
        sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      }
    }
    const float inv = 1.f / block_reduce(sum, false, s_red);
    uint2* dst = reinterpret_cast<uint2*>(P + (long long)r * ldp);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = threadIdx.x + j * kThreads;
      if (i < vecs) dst[i] = make_uint2(pack_bf162(v[j].x * inv, v[j].y * inv), pack_bf162(v[j].z * inv, v[j].w * inv));
    }
  }
}

}  // namespace

extern "C" int vn_softmax_rows(const float* S, int64_t lds, void* P, int64_t ldp, int rows, int cols, float scale,
                               vn_stream_t s) {
  VN_CHECK(rows >= 1 && cols >= 4 && cols % 4 == 0 && lds % 4 == 0 && ldp % 4 == 0 && lds >= cols && ldp >= cols,
           "softmax_rows: cols and strides must be multiples of 4 (rows=%d cols=%d)", rows, cols);
  VN_CHECK(scale > 0.f, "softmax_rows: scale must be positive");
  VN_CHECK((reinterpret_cast<uintptr_t>(S) & 15) == 0 && (reinterpret_cast<uintptr_t>(P) & 7) == 0,
           "softmax_rows: S must be 16-byte and P 8-byte aligned");
  const int nv = vn_cdiv(cols, 4 * kThreads);
  const int grid = rows < 148 * 8 ? rows : 148 * 8;
  const float sl = scale * 1.4426950408889634f;
#define VN_SM_CASE(NV)                                                                                              \
  case NV:                                                                                                          \
    VN_LAUNCH(softmax_rows_kernel<NV>, grid, kThreads, 0, (cudaStream_t)s, S, (long long)lds, (bf16*)P,             \
              (long long)ldp, rows, cols, sl);                                                                      \
    break;
  switch (nv <= 2 ? 2 : nv <= 4 ? 4 : nv <= 8 ? 8 : 0) {
    VN_SM_CASE(2) VN_SM_CASE(4) VN_SM_CASE(8)
    default: VN_CHECK(false, "softmax_rows: cols=%d unsupported (<= %d)", cols, 4 * kThreads * 8);
  }
#undef VN_SM_CASE
  return 0;
}
