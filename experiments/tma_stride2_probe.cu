// experiments/tma_stride2_probe.cu — NOT part of the library.  A stand-alone probe for the next kernel step listed in
// stride-2 3x3 convolutions (UNet Downsample2D x3, VAE encoder x3) as an implicit GEMM whose A tiles
// come straight from a 4-D NHWC tensor map with elementStrides = {1, 2, 2, 1}, instead of the explicit im2col they use now.
// What has to be known before vn_gemm can rely on it, and what this prints:
//   1. does boxDim count SOURCE elements (box {64, 2*tw, 2*th, 1} -> tw*th pixels land) or LOADED elements?
//   2. are the loaded pixels compacted in shared memory (row r = ty*tw + tx), as the UMMA descriptor needs?
//   3. how many bytes does complete_tx report (what expect_tx must be)?
//   4. negative start coordinates (the pad) with a traversal stride: which pixels are zero-filled?
// The kernel never waits on the mbarrier (a wrong expect_tx would hang): it arms it with a huge count, spins 200 us on
// clock64, then reads shared memory and counts the bytes that changed from a sentinel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_stride2_probe tma_stride2_probe.cu && ./tma_stride2_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int C = 64, W = 16, H = 8;
constexpr int SMEM_PIX = 64;                       // room for 64 pixels of 64 channels

__global__ void probe(const __grid_constant__ CUtensorMap tm, int w0, int h0, float* out, int* landed_bytes) {
  __shared__ __align__(128) __nv_bfloat16 tile[SMEM_PIX * C];
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < SMEM_PIX * C; i += blockDim.x) tile[i] = __float2bfloat16(-1.f);      // sentinel
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)),
                 "r"(1u << 19)
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"((uint32_t)__cvta_generic_to_shared(tile)), "l"(reinterpret_cast<uint64_t>(&tm)),
        "r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(0), "r"(w0), "r"(h0), "r"(0)
        : "memory");
    const long long t0 = clock64();
    while (clock64() - t0 < 400000) {}             // ~200 us: far longer than one 8 KB TMA load
  }
  __syncthreads();
  int changed = 0;
  for (int i = threadIdx.x; i < SMEM_PIX * C; i += blockDim.x) {
    const float v = __bfloat162float(tile[i]);
    if (v != -1.f) ++changed;
    if (i % C == 0) out[i / C] = v;                // channel 0 of every smem pixel row
  }
  atomicAdd(landed_bytes, changed * 2);
}

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess || !fnp) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeFn encode = reinterpret_cast<EncodeFn>(fnp);
  std::vector<__nv_bfloat16> h(H * W * C);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < C; ++c) h[(y * W + x) * C + c] = __float2bfloat16((float)(y * W + x + 1));   // 1..128, exact in bf16
  __nv_bfloat16* d;
  cudaMalloc(&d, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  float* out;
  int* landed;
  cudaMalloc(&out, SMEM_PIX * 4);
  cudaMalloc(&landed, 4);
  const cuuint64_t dims[4] = {C, W, H, 1};
  const cuuint64_t strides[3] = {C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
  const cuuint32_t es[4] = {1, 2, 2, 1};
  const int tw = 4, th = 2;
  struct Case { const char* name; cuuint32_t box[4]; int w0, h0; } cases[] = {
      {"box = {64, 2*tw, 2*th, 1} (source extent), start (0,0)", {64, 2 * tw, 2 * th, 1}, 0, 0},
      {"box = {64, tw, th, 1} (loaded count), start (0,0)", {64, tw, th, 1}, 0, 0},
      {"box = {64, 2*tw, 2*th, 1}, start (-1,-1): the pad-1 tap (dx=0, dy=0)", {64, 2 * tw, 2 * th, 1}, -1, -1},
      {"box = {64, 2*tw-1, 2*th-1, 1} (odd extent), start (1,1)", {64, 2 * tw - 1, 2 * th - 1, 1}, 1, 1},
      {"box = {64, 2*tw, 2*th, 1}, start (12,6): runs off the far edge", {64, 2 * tw, 2 * th, 1}, 12, 6},
  };
  for (const Case& cs : cases) {
    CUtensorMap tm;
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, cs.box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("\n%s\n  encode -> CUresult %d\n", cs.name, (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaMemset(landed, 0, 4);
    probe<<<1, 128>>>(tm, cs.w0, cs.h0, out, landed);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("  kernel error: %s\n", cudaGetErrorString(e));
      return 1;
    }
    float ho[SMEM_PIX];
    int lb = 0;
    cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
    cudaMemcpy(&lb, landed, 4, cudaMemcpyDeviceToHost);
    printf("  bytes changed in smem: %d (tw*th pixels would be %d)\n  smem pixel rows (h,w) or 0 = zero fill, . = untouched:\n   ",
           lb, tw * th * C * 2);
    for (int i = 0; i < 24; ++i) {
      if (ho[i] == -1.f) printf(" .");
      else if (ho[i] == 0.f) printf(" 0");
      else printf(" (%d,%d)", ((int)ho[i] - 1) / W, ((int)ho[i] - 1) % W);
    }
    printf("\n");
  }
  return 0;
}
