// vn_core.cu — error plumbing, launch accounting and the trivial byte movers of libviewneti_sm100a.so.
#include "vn_common.cuh"

#include <atomic>
#include <stdarg.h>
#include <string.h>

namespace {
thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};
std::atomic<int> g_pdl{1};
}  // namespace

// in-kernel timelines (scripts/kernel_timeline.py): device buffer of 16 int64 slots per CTA, or NULL (off)
static long long* g_dbg = nullptr;
long long* vn_debug_buffer() { return g_dbg; }
extern "C" void vn_set_debug_buffer(void* p) { g_dbg = reinterpret_cast<long long*>(p); }

bool vn_pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }
extern "C" void vn_set_pdl(int enabled) { g_pdl.store(enabled ? 1 : 0); }

void vn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void vn_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int vn_version(void) { return VN_ABI_VERSION; }
extern "C" const char* vn_last_error(void) { return g_err; }
extern "C" int64_t vn_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" void vn_launch_count_reset(void) { g_launches.store(0); }

namespace {

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y,
                                                            long long n) {
  pdl_trigger();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      uint2 o;
      o.x = pack_bf162(v.x, v.y);
      o.y = pack_bf162(v.z, v.w);
      *reinterpret_cast<uint2*>(y + i) = o;
    } else {
      for (long long j = i; j < n; ++j) y[j] = __float2bfloat16(x[j]);
    }
  }
}

__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y,
                                                            long long n) {
  pdl_trigger();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const uint2 v = *reinterpret_cast<const uint2*>(x + i);
      const float2 a = unpack_bf162(v.x), b = unpack_bf162(v.y);
      *reinterpret_cast<float4*>(y + i) = make_float4(a.x, a.y, b.x, b.y);
    } else {
      for (long long j = i; j < n; ++j) y[j] = __bfloat162float(x[j]);
    }
  }
}

// rows x cols (cols % 8 == 0) bf16 strided copy, optional add of a second source: dst = src (+ add)
__global__ void __launch_bounds__(256) copy2d_kernel(const bf16* __restrict__ src, long long lds,
                                                     const bf16* __restrict__ add, long long lda,
                                                     bf16* __restrict__ dst, long long ldd, long long rows, int vecs) {
  pdl_trigger();
  pdl_wait();
  const long long total = rows * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vecs;
    const int c = (int)(i - r * vecs) * 8;
    uint4 v = *reinterpret_cast<const uint4*>(src + r * lds + c);
    if (add) {
      const uint4 a = *reinterpret_cast<const uint4*>(add + r * lda + c);
      float2 x, y;
      x = unpack_bf162(v.x); y = unpack_bf162(a.x); v.x = pack_bf162(x.x + y.x, x.y + y.y);
      x = unpack_bf162(v.y); y = unpack_bf162(a.y); v.y = pack_bf162(x.x + y.x, x.y + y.y);
      x = unpack_bf162(v.z); y = unpack_bf162(a.z); v.z = pack_bf162(x.x + y.x, x.y + y.y);
      x = unpack_bf162(v.w); y = unpack_bf162(a.w); v.w = pack_bf162(x.x + y.x, x.y + y.y);
    }
    *reinterpret_cast<uint4*>(dst + r * ldd + c) = v;
  }
}

int grid_for(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int vn_cast_f32_bf16(const float* x, void* y, int64_t n, vn_stream_t s) {
  VN_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0,
           "vn_cast_f32_bf16: misaligned pointers");
  if (n <= 0) return 0;
  VN_LAUNCH(cast_f32_bf16_kernel, grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)s, x, (bf16*)y, n);
  return 0;
}

extern "C" int vn_cast_bf16_f32(const void* x, float* y, int64_t n, vn_stream_t s) {
  VN_CHECK((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0,
           "vn_cast_bf16_f32: misaligned pointers");
  if (n <= 0) return 0;
  VN_LAUNCH(cast_bf16_f32_kernel, grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)s, (const bf16*)x, y, n);
  return 0;
}

extern "C" int vn_copy2d(const void* src, int64_t lds, const void* add, int64_t ldadd, void* dst, int64_t ldd,
                         int64_t rows, int cols, vn_stream_t s) {
  VN_CHECK(cols % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && (!add || ldadd % 8 == 0),
           "vn_copy2d: cols and strides must be multiples of 8");
  if (rows <= 0 || cols <= 0) return 0;
  const int vecs = cols / 8;
  VN_LAUNCH(copy2d_kernel, grid_for(rows * vecs, 256), 256, 0, (cudaStream_t)s, (const bf16*)src, lds, (const bf16*)add,
                                                                          ldadd, (bf16*)dst, ldd, rows, vecs);
  return 0;
}

extern "C" int vn_memset(void* p, int byte_value, size_t bytes, vn_stream_t s) {
  VN_CUDA(cudaMemsetAsync(p, byte_value, bytes, (cudaStream_t)s));
  return 0;
}

extern "C" int vn_memset_zero(void* p, size_t bytes, vn_stream_t s) {
  VN_CUDA(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)s));
  return 0;
}
