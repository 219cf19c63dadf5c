// Library-wide entry points: version, error string, launch accounting, the POPC micro-benchmark.
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <utility>
#include <vector>
#include "common.cuh"

namespace hpusm {

static thread_local char t_error[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

// ---- optional event bracketing of the dominant kernels -------------------------------------------------
static std::atomic<int> g_profile{0};
static std::mutex g_profile_mu;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_profile_events;
static thread_local cudaEvent_t g_profile_open = nullptr;  // begin/end are paired on the launching thread

void profile_mark(cudaStream_t st, bool begin) {
  if (!g_profile.load(std::memory_order_relaxed)) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  if (begin) {
    if (g_profile_open) cudaEventDestroy(g_profile_open);
    g_profile_open = e;
  } else if (g_profile_open) {
    std::lock_guard<std::mutex> lock(g_profile_mu);
    g_profile_events.emplace_back(g_profile_open, e);
    g_profile_open = nullptr;
  } else {
    cudaEventDestroy(e);
  }
}

int sm_count() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// FP32 pipe: every thread issues `iters` fused multiply-adds in 8 independent chains; PACKED = fma.rn.f32x2 (FFMA2, two
// FMAs per instruction), else scalar FFMA.  The measured denominator of the RANSAC scoring kernel's roofline.
template <bool PACKED>
__global__ void ffma_microbench_kernel(int64_t iters, uint32_t* sink) {
  const float s0 = 1.0f + 1e-7f * threadIdx.x, m = 0.999999f;
  if (PACKED) {
    unsigned long long a[8], mm, cc;
    asm("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(1e-9f));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(s0 + i), "f"(s0 - i));
    for (int64_t k = 0; k < iters; k += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(mm), "l"(cc));
    }
    unsigned long long r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    if (r == 0x123456789abcdefull) sink[0] = 1;
  } else {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = s0 + i;
    for (int64_t k = 0; k < iters; k += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(m), "f"(1e-9f));
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    if (r == 123.456f) sink[0] = 1;
  }
}

// ---- asynchronous error word (pinned, mapped) ------------------------------------------------------------------------
static std::mutex g_async_mu;
static volatile int* g_async_host = nullptr;

__global__ void async_report_kernel(const int* __restrict__ device_flag, int code, volatile int* host_word) {
  if (*device_flag != 0) *host_word = code;
}

int async_report(const int* device_flag, int code, cudaStream_t st) {
  {
    std::lock_guard<std::mutex> lock(g_async_mu);
    if (!g_async_host) {
      int* p = nullptr;
      HPUSM_CUDA_OK(cudaHostAlloc(&p, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
      *p = 0;
      g_async_host = p;
    }
  }
  HPUSM_LAUNCH(async_report_kernel, 1, 1, 0, st, device_flag, code, g_async_host);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

// Each thread keeps 8 independent POPC chains so the pipe, not the dependency, is the limit.
__global__ void popc_microbench_kernel(int64_t iters, uint32_t* sink) {
  uint32_t x0 = threadIdx.x * 2654435761u + blockIdx.x, x1 = x0 ^ 0x9e3779b9u, x2 = x0 + 0x7f4a7c15u,
           x3 = ~x0, x4 = x0 * 3u, x5 = x1 * 5u, x6 = x2 * 7u, x7 = x3 * 11u;
  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  for (int64_t i = 0; i < iters; i += 8) {
    a0 += __popc(x0 ^ a7); a1 += __popc(x1 ^ a0); a2 += __popc(x2 ^ a1); a3 += __popc(x3 ^ a2);
    a4 += __popc(x4 ^ a3); a5 += __popc(x5 ^ a4); a6 += __popc(x6 ^ a5); a7 += __popc(x7 ^ a6);
  }
  uint32_t r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
  if (r == 0xffffffffu) sink[0] = r;  // never true in practice; keeps the chains alive
}

// Same for LOP3 (ALU pipe): 8 independent chains of majority/xor3 operations.
__global__ void lop3_microbench_kernel(int64_t iters, uint32_t* sink) {
  uint32_t a0 = threadIdx.x * 2654435761u + blockIdx.x, a1 = a0 ^ 0x9e3779b9u, a2 = a0 + 0x7f4a7c15u, a3 = ~a0,
           a4 = a0 * 3u, a5 = a1 * 5u, a6 = a2 * 7u, a7 = a3 * 11u;
  const uint32_t k0 = a0 * 13u + 1u, k1 = a1 * 17u + 3u;
  for (int64_t i = 0; i < iters; i += 8) {
    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a0) : "r"(k0), "r"(k1));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a1) : "r"(k0), "r"(k1));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a2) : "r"(k1), "r"(k0));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a3) : "r"(k1), "r"(k0));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a4) : "r"(k0), "r"(k1));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a5) : "r"(k0), "r"(k1));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a6) : "r"(k1), "r"(k0));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a7) : "r"(k1), "r"(k0));
  }
  const uint32_t r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
  if (r == 0xffffffffu && k0 == 7u) sink[0] = r;
}

// Min/max issue rates (the fused top-2 epilogues are max-scans): 8 independent chains of one of
//   OP 0: 3-input s32 max (VIMNMX3)   1: 3-input f32 max (FMNMX3)   2: 2-input s32 max   3: 2-input f32 max
template <int OP>
__global__ void minmax_microbench_kernel(int64_t iters, uint32_t* sink) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (threadIdx.x * 2654435761u + blockIdx.x * 40503u + i * 977u) & 0x3fffffffu;
  const uint32_t k1 = a[1] * 17u & 0x3fffffffu;
  for (int64_t it = 0; it < iters; it += 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // each value is refreshed from its neighbours: nothing for the compiler to fold
      const uint32_t x = a[(i + 1) & 7], y = a[(i + 2) & 7];
      if (OP == 0) asm volatile("max.s32 %0, %0, %1;\n\tmax.s32 %0, %0, %2;" : "+r"(a[i]) : "r"(x), "r"(y));  // ptxas fuses: VIMNMX3
      if (OP == 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));
      if (OP == 2) asm volatile("max.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(x));
      if (OP == 3) asm volatile("max.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(x));
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r ^= a[i];
  if (r == 0xffffffffu && k1 == 7u) sink[0] = r;
}

}  // namespace hpusm

using namespace hpusm;

extern "C" {

int hpusm_version(void) { return HPUSM_ABI_VERSION; }

int hpusm_async_status(void) {
  std::lock_guard<std::mutex> lock(hpusm::g_async_mu);
  if (!hpusm::g_async_host) return HPUSM_OK;
  const int code = *hpusm::g_async_host;
  *hpusm::g_async_host = 0;
  if (code != 0) hpusm::set_error("an asynchronous entry point rejected its input on the device (code %d): e.g. HPUSM_L2_TC_I8 "
                                  "on descriptors that are not integers in [0,255]", code);
  return code;
}

const char* hpusm_last_error(void) { return t_error; }

int hpusm_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return HPUSM_ERR_CUDA;
  }
  return n;
}

int64_t hpusm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int hpusm_profile_enable(int on) {
  g_profile.store(on ? 1 : 0);
  return HPUSM_OK;
}

int hpusm_profile_collect(double* total_ms, int64_t* launches) {
  std::lock_guard<std::mutex> lock(g_profile_mu);
  double sum = 0.0;
  int64_t n = 0;
  for (auto& pr : g_profile_events) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(pr.second);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, pr.first, pr.second);
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
    if (e != cudaSuccess) {
      set_error("hpusm_profile_collect: %s", cudaGetErrorString(e));
      g_profile_events.clear();
      return HPUSM_ERR_CUDA;
    }
    sum += ms;
    ++n;
  }
  g_profile_events.clear();
  if (total_ms) *total_ms = sum;
  if (launches) *launches = n;
  return HPUSM_OK;
}

int hpusm_microbench_popc(int64_t blocks, int threads, int64_t iters, uint32_t* sink, void* stream) {
  HPUSM_REQUIRE(blocks > 0 && threads > 0 && threads <= 1024 && iters > 0 && sink, HPUSM_ERR_INVALID,
                "hpusm_microbench_popc: bad arguments");
  HPUSM_LAUNCH(popc_microbench_kernel, (unsigned)blocks, threads, 0, as_stream(stream), iters, sink);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_microbench_lop3(int64_t blocks, int threads, int64_t iters, uint32_t* sink, void* stream) {
  HPUSM_REQUIRE(blocks > 0 && threads > 0 && threads <= 1024 && iters > 0 && sink, HPUSM_ERR_INVALID,
                "hpusm_microbench_lop3: bad arguments");
  HPUSM_LAUNCH(lop3_microbench_kernel, (unsigned)blocks, threads, 0, as_stream(stream), iters, sink);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_microbench_ffma(int packed, int64_t blocks, int threads, int64_t iters, uint32_t* sink, void* stream) {
  HPUSM_REQUIRE(blocks > 0 && threads > 0 && threads <= 1024 && iters > 0 && sink, HPUSM_ERR_INVALID,
                "hpusm_microbench_ffma: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (packed) { HPUSM_LAUNCH(ffma_microbench_kernel<true>, (unsigned)blocks, threads, 0, st, iters, sink); }
  else { HPUSM_LAUNCH(ffma_microbench_kernel<false>, (unsigned)blocks, threads, 0, st, iters, sink); }
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_microbench_minmax(int op, int64_t blocks, int threads, int64_t iters, uint32_t* sink, void* stream) {
  HPUSM_REQUIRE(op >= 0 && op <= 3 && blocks > 0 && threads > 0 && threads <= 1024 && iters > 0 && sink, HPUSM_ERR_INVALID,
                "hpusm_microbench_minmax: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (op == 0) { HPUSM_LAUNCH(minmax_microbench_kernel<0>, (unsigned)blocks, threads, 0, st, iters, sink); }
  else if (op == 1) { HPUSM_LAUNCH(minmax_microbench_kernel<1>, (unsigned)blocks, threads, 0, st, iters, sink); }
  else if (op == 2) { HPUSM_LAUNCH(minmax_microbench_kernel<2>, (unsigned)blocks, threads, 0, st, iters, sink); }
  else { HPUSM_LAUNCH(minmax_microbench_kernel<3>, (unsigned)blocks, threads, 0, st, iters, sink); }
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

}  // extern "C"
