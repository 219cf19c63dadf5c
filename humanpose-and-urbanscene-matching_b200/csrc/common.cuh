// Shared plumbing for libhpusm.so: error reporting, launch accounting, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/hpusm.h"

namespace hpusm {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Every kernel launch goes through this so that hpusm_launch_count() is an honest count.
#define HPUSM_LAUNCH(kernel, grid, block, smem, stream, ...)                         \
  do {                                                                               \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                      \
    ::hpusm::g_launches.fetch_add(1, std::memory_order_relaxed);                     \
  } while (0)

// Same, bracketed by CUDA events on the launching stream when hpusm_profile_enable(1) is in force: the
// dominant kernel of each path is launched through this so bench.py can report its duration (roofline).
void profile_mark(cudaStream_t st, bool begin);
#define HPUSM_LAUNCH_TIMED(kernel, grid, block, smem, stream, ...)                   \
  do {                                                                               \
    ::hpusm::profile_mark((stream), true);                                           \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                      \
    ::hpusm::profile_mark((stream), false);                                          \
    ::hpusm::g_launches.fetch_add(1, std::memory_order_relaxed);                     \
  } while (0)

#define HPUSM_CUDA_OK(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::hpusm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                         __FILE__, __LINE__);                                        \
      return HPUSM_ERR_CUDA;                                                         \
    }                                                                                \
  } while (0)

#define HPUSM_CHECK_LAUNCH()                                                         \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      ::hpusm::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                        \
      return HPUSM_ERR_CUDA;                                                         \
    }                                                                                \
  } while (0)

#define HPUSM_REQUIRE(cond, code, ...)                                               \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      ::hpusm::set_error(__VA_ARGS__);                                               \
      return (code);                                                                 \
    }                                                                                \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline bool is_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int sm_count();  // SMs of the current device (148 on B200), cached

// Asynchronous error reporting: enqueues a one-thread kernel that, if *device_flag != 0, stores `code` into a pinned host
// word; hpusm_async_status() returns and clears it.  Lets an entry point stay asynchronous where it used to read a flag back.
int async_report(const int* device_flag, int code, cudaStream_t st);

// ---- device helpers -------------------------------------------------------------------------------

// A (distance, index) candidate pair list of length 2 kept sorted lexicographically.
template <typename D>
struct Top2 {
  D d0, d1;
  int32_t i0, i1;
};

template <typename D>
__device__ __forceinline__ bool lex_less(D da, int32_t ia, D db, int32_t ib) {
  // Empty slots carry index -1 and must lose against every real candidate.
  if (ia < 0) return false;
  if (ib < 0) return true;
  return (da < db) || (da == db && ia < ib);
}

// Insert (d, i) into a sorted top-2 under the lexicographic order.
template <typename D>
__device__ __forceinline__ void top2_insert(Top2<D>& t, D d, int32_t i) {
  if (lex_less<D>(d, i, t.d1, t.i1)) {
    if (lex_less<D>(d, i, t.d0, t.i0)) {
      t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = i;
    } else {
      t.d1 = d; t.i1 = i;
    }
  }
}

template <typename D>
__device__ __forceinline__ void top2_merge(Top2<D>& a, const Top2<D>& b) {
  top2_insert<D>(a, b.d0, b.i0);
  top2_insert<D>(a, b.d1, b.i1);
}

}  // namespace hpusm
