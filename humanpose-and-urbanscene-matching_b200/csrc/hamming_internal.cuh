// Internal interface between hamming.cu (entry points, __popc engine) and hamming_tc.cu (tensor-core engine).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hpusm {

bool hamming_tc_eligible(int64_t nq, int64_t nt, int64_t nseg, int nbytes);
size_t hamming_tc_workspace_bytes(int64_t nq, int64_t nt, int64_t nseg);
int hamming_tc_segmented(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, const int64_t* seg_offsets, int64_t nseg,
                         double ratio, int32_t* score, int32_t* best_idx, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace hpusm
