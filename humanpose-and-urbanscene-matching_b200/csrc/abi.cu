// Library-wide entry points: version, error string, launch accounting, the POPC micro-benchmark.
#include <stdarg.h>
#include <string.h>
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

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
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

}  // namespace hpusm

using namespace hpusm;

extern "C" {

int hpusm_version(void) { return HPUSM_ABI_VERSION; }

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

int hpusm_microbench_popc(int64_t blocks, int threads, int64_t iters, uint32_t* sink, void* stream) {
  HPUSM_REQUIRE(blocks > 0 && threads > 0 && threads <= 1024 && iters > 0 && sink, HPUSM_ERR_INVALID,
                "hpusm_microbench_popc: bad arguments");
  HPUSM_LAUNCH(popc_microbench_kernel, (unsigned)blocks, threads, 0, as_stream(stream), iters, sink);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

}  // extern "C"
