// Tensor-pipe issue-rate microbenchmark: one CTA per SM issues back-to-back tcgen05.mma (M = 128, N = 256, one 32-byte K
// step, both operands in shared memory) into two alternating accumulators and does nothing else.  It gives the roofline
// of the L2 engines a MEASURED denominator for the integer pipe (MEASURED_PEAKS.json only has a bf16 figure) and the
// same figure for bf16 as a cross-check against the cuBLAS number.
#include <cuda.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace hpusm {

constexpr int MB_N = 256;
constexpr int MB_MMAS_PER_ROUND = 64;                       // 64 x 128 tensor cycles between two completion waits
constexpr int MB_A_BYTES = 128 * 128, MB_B_BYTES = MB_N * 128;
constexpr int MB_SMEM = MB_A_BYTES + MB_B_BYTES + 1024;

// 0: kind::f16 with bf16 operands (K = 16 per step), 1: kind::i8 with u8 operands (K = 32 per step); N = the MMA's N;
// A_TMEM: the A operand is read from tensor memory (the form the kNN engines use) instead of shared memory
template <int KIND, int N = MB_N, bool A_TMEM = false>
__global__ void __launch_bounds__(128, 1) mma_microbench_kernel(int rounds) {
  extern __shared__ uint8_t mb_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mb_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (MB_A_BYTES + MB_B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u * (KIND == 0) + 0x01010101u * (KIND == 1);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the generic-proxy stores above are read by the tensor pipe
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = slot;
  if (warp == 0) {
    const uint64_t adesc = umma_desc_sw128(smem_u32(smem)), bdesc = umma_desc_sw128(smem_u32(smem + MB_A_BYTES));
    const uint32_t idesc = KIND == 0 ? tc_idesc_bf16(N) : ((2u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24));
    for (int r = 0; r < rounds; ++r) {
      if (elect_one()) {
#pragma unroll 8
        for (int i = 0; i < MB_MMAS_PER_ROUND; ++i) {
          const uint32_t d = tmem_base + (uint32_t)((i & 1) * N);
          const uint64_t koff = (uint64_t)(2 * (i & 3));
          if (A_TMEM) {
            tc_mma_i8_ts(d, tmem_base + (uint32_t)(2 * N + 8 * (i & 3)), bdesc + koff, idesc, 1u);
          } else if (KIND == 0) {
            tc_mma_bf16(d, adesc + koff, bdesc + koff, idesc, 1u);
          } else {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
                "}" ::"r"(d), "l"(adesc + koff), "l"(bdesc + koff), "r"(idesc), "r"(1u) : "memory");
          }
        }
        tc_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, (uint32_t)(r & 1));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

}  // namespace hpusm

using namespace hpusm;

extern "C" {

// kind 0 = bf16 (kind::f16), 1 = u8 (kind::i8).  Every SM issues rounds x 64 MMAs of 128 x 256 x (16 | 32).
// kinds 2..5 = u8 at N = 64 / 64 with A in tensor memory / 128 / 128 with A in tensor memory (the shapes of the kNN engines).
int hpusm_microbench_mma(int kind, int rounds, void* stream) {
  HPUSM_REQUIRE(kind >= 0 && kind <= 5 && rounds > 0 && rounds <= (1 << 20), HPUSM_ERR_INVALID,
                "hpusm_microbench_mma: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (kind == 0) {
    HPUSM_CUDA_OK(cudaFuncSetAttribute(mma_microbench_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
    HPUSM_LAUNCH(mma_microbench_kernel<0>, sm_count(), 128, MB_SMEM, st, rounds);
  } else if (kind == 1) {
    HPUSM_CUDA_OK(cudaFuncSetAttribute(mma_microbench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
    HPUSM_LAUNCH(mma_microbench_kernel<1>, sm_count(), 128, MB_SMEM, st, rounds);
  } else if (kind == 2) {   // u8, N = 64, both operands in shared memory
    HPUSM_CUDA_OK(cudaFuncSetAttribute(mma_microbench_kernel<1, 64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
    HPUSM_LAUNCH((mma_microbench_kernel<1, 64, false>), sm_count(), 128, MB_SMEM, st, rounds);
  } else if (kind == 3) {   // u8, N = 64, A in tensor memory
    HPUSM_CUDA_OK(cudaFuncSetAttribute(mma_microbench_kernel<1, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
    HPUSM_LAUNCH((mma_microbench_kernel<1, 64, true>), sm_count(), 128, MB_SMEM, st, rounds);
  } else if (kind == 4) {   // u8, N = 128, both operands in shared memory
    HPUSM_CUDA_OK(cudaFuncSetAttribute(mma_microbench_kernel<1, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
    HPUSM_LAUNCH((mma_microbench_kernel<1, 128, false>), sm_count(), 128, MB_SMEM, st, rounds);
  } else {                  // u8, N = 128, A in tensor memory
    HPUSM_CUDA_OK(cudaFuncSetAttribute(mma_microbench_kernel<1, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
    HPUSM_LAUNCH((mma_microbench_kernel<1, 128, true>), sm_count(), 128, MB_SMEM, st, rounds);
  }
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

}  // extern "C"
