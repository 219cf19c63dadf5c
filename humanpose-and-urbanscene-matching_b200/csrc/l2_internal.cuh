// Internal interfaces between the L2 engines (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace hpusm {

// l2_simt.cu
int l2_simt_splits(int64_t nq_rows, int64_t nt);
int l2_simt_candidates(const float* q, int64_t nq, const float* t, int64_t nt, int d, const int32_t* row_list,
                       int64_t n_list, int splits, int32_t* part_idx, cudaStream_t st);
int l2_rerank_finalize(const float* q, int64_t nq, const float* t, int d, const int32_t* cand, int nparts, int m,
                       const int32_t* row_list, int64_t n_list, int32_t* idx, float* dist, cudaStream_t st);

// l2_tc.cu (tcgen05 + TMA engine)
bool l2_tc_available();
size_t l2_tc_workspace_bytes(int64_t nq, int64_t nt, int d, int mode);
int l2_tc_knn2(const float* q, int64_t nq, const float* t, int64_t nt, int d, int mode, int32_t* idx, float* dist,
               void* ws, size_t ws_bytes, cudaStream_t st, int* resolved_mode, int64_t* fallback_rows);

// l2_tc_split.cu (tcgen05 engine for general float descriptors: hi/lo bf16 split, top-4 candidates, verified re-rank)
size_t l2_tc_split_workspace_bytes(int64_t nq, int64_t nt);
int l2_tc_split_knn2(const float* q, int64_t nq, const float* t, int64_t nt, int32_t* idx, float* dist, void* ws,
                     size_t ws_bytes, cudaStream_t st, int64_t* fallback_rows);

// l2_tc_i8.cu (tcgen05 kind::i8 engine for integer descriptors in [0,255]; *eligible = 0 when the data does not fit)
size_t l2_tc_i8_workspace_bytes(int64_t nq, int64_t nt);
int l2_tc_i8_knn2(const float* q, int64_t nq, const float* t, int64_t nt, int32_t* idx, float* dist, void* ws,
                  size_t ws_bytes, cudaStream_t st, bool require, int* eligible);

// l2_tc_q16.cu (tcgen05 kind::i8 engine for general float descriptors: 15-bit fixed point in two signed bytes, two
// accumulators per tile, top-M candidates, verified re-rank)
size_t l2_tc_q16_workspace_bytes(int64_t nq, int64_t nt);
int l2_tc_q16_knn2(const float* q, int64_t nq, const float* t, int64_t nt, int32_t* idx, float* dist, void* ws,
                   size_t ws_bytes, cudaStream_t st, int64_t* fallback_rows);

}  // namespace hpusm
