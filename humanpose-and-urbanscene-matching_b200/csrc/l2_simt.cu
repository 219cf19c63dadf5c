// K1 (CUDA-core engine) + the exact FP32 re-rank shared by every L2 engine.
//
// Replaces cv2.BFMatcher(cv2.NORM_L2).knnMatch(..., k=2) at reference urbanscene/features.py:59.
//
// l2_simt_knn2_kernel is the exact brute force: d2 = sum_k fma(a_k-b_k, a_k-b_k, d2) in increasing k,
// FP32, i.e. the same number the re-rank produces.  It is the cross-check for the tcgen05 engine at
// sizes no CPU oracle reaches, the engine for shapes the tensor path does not take (d not a multiple of
// 128), and the fallback for query rows whose split-engine candidate list could not be proven complete.  64x64 register-tiled, query tile resident in shared memory for the whole sweep, database
// chunks double-buffered with cp.async, per-thread running top-2 (no reduction in the hot loop) and one
// lexicographic warp-shuffle merge at the end.
//
// l2_rerank_finalize_kernel takes candidate index lists (from any engine / any number of database
// splits), recomputes the exact FP32 distance of every candidate from the original float32 rows, keeps
// the lexicographic top-2 on (d2, index) and writes sqrtf(d2) -- the value DMatch.distance carries.
#include <float.h>
#include <limits.h>
#include "common.cuh"

namespace hpusm {

constexpr int SIMT_TQ = 64;       // queries per CTA
constexpr int SIMT_TT = 64;       // database rows per stage
constexpr int SIMT_KC = 32;       // k-chunk staged per cp.async group
constexpr int SIMT_THREADS = 256; // 16 x 16 threads, 4 x 4 pairs each
constexpr int SIMT_LD = SIMT_TT + 4;  // padded leading dimension of the transposed tiles
constexpr int SIMT_KCP = SIMT_KC + 4; // padded row pitch of the row-major staging buffer (conflict-free LDS.128)

__device__ __forceinline__ void cp_async16_l2(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}

struct F32Top2 {
  float d0, d1;
  int i0, i1;
};

// grid = (query tiles, splits); dynamic smem = A tile [d][SIMT_LD] + B staging 2 x [TT][KC] + Bt [KC][LD].
__global__ void __launch_bounds__(SIMT_THREADS)
l2_simt_knn2_kernel(const float* __restrict__ q, int64_t nq, const float* __restrict__ t, int64_t nt, int d,
                    int64_t rows_per_split, const int32_t* __restrict__ row_list, int64_t n_list,
                    int32_t* __restrict__ part_idx) {
  extern __shared__ __align__(16) float smem_f[];
  float* As = smem_f;                                   // [d][SIMT_LD]   As[k][query]
  float* Braw = As + (size_t)d * SIMT_LD;               // 2 x [SIMT_TT][SIMT_KCP] row-major staging
  float* Bt = Braw + 2 * SIMT_TT * SIMT_KCP;            // [SIMT_KC][SIMT_LD]   Bt[k][row]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  // With a row list (fallback mode) the CTA's queries are row_list[q_base ..], otherwise q_base .. .
  const int64_t n_rows_q = row_list ? n_list : nq;
  const int64_t q_base = (int64_t)blockIdx.x * SIMT_TQ;
  const int split = blockIdx.y;
  const int64_t t_begin = (int64_t)split * rows_per_split;
  const int64_t t_end = min(nt, t_begin + rows_per_split);

  // Query tile -> shared, transposed (k-major) so the inner loop reads 4 queries with one LDS.128.
  const int dv = d >> 2;
  for (int e = tid; e < SIMT_TQ * dv; e += SIMT_THREADS) {
    const int kv = e / SIMT_TQ, r = e - kv * SIMT_TQ;  // consecutive threads -> consecutive rows: no bank conflicts
    const int64_t lrow = q_base + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lrow < n_rows_q) {
      const int64_t row = row_list ? (int64_t)row_list[lrow] : lrow;
      v = __ldg(reinterpret_cast<const float4*>(q + row * d) + kv);
    }
    As[(4 * kv + 0) * SIMT_LD + r] = v.x; As[(4 * kv + 1) * SIMT_LD + r] = v.y;
    As[(4 * kv + 2) * SIMT_LD + r] = v.z; As[(4 * kv + 3) * SIMT_LD + r] = v.w;
  }

  F32Top2 top[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { top[i].d0 = top[i].d1 = FLT_MAX; top[i].i0 = top[i].i1 = -1; }

  const int64_t n_rows = t_end - t_begin;
  const int n_tiles = n_rows > 0 ? (int)((n_rows + SIMT_TT - 1) / SIMT_TT) : 0;
  const int n_chunks = (d + SIMT_KC - 1) / SIMT_KC;
  const int total_stages = n_tiles * n_chunks;

  // Stage s = (tile, chunk): rows [tile*TT, +TT), columns [chunk*KC, +KC) of the database, row-major.
  auto issue = [&](int s, int buf) {
    const int tile = s / n_chunks, chunk = s - tile * n_chunks;
    const int64_t row0 = t_begin + (int64_t)tile * SIMT_TT;
    const int k0 = chunk * SIMT_KC;
    float* dst = Braw + buf * SIMT_TT * SIMT_KCP;
    for (int e = tid; e < SIMT_TT * (SIMT_KC / 4); e += SIMT_THREADS) {
      const int r = e / (SIMT_KC / 4), kv = e - r * (SIMT_KC / 4);  // 8 threads cover one 128-byte row chunk
      const int64_t row = row0 + r;
      const int k = k0 + 4 * kv;
      if (row < t_end && k < d) cp_async16_l2(dst + r * SIMT_KCP + 4 * kv, t + row * d + k);
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  float acc[4][4];
  if (total_stages > 0) issue(0, 0);
  __syncthreads();  // As visible
  for (int s = 0; s < total_stages; ++s) {
    const int buf = s & 1;
    const int tile = s / n_chunks, chunk = s - tile * n_chunks;
    if (s + 1 < total_stages) {
      issue(s + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    // transpose the staged chunk into Bt[k][row]
    {
      const float* src = Braw + buf * SIMT_TT * SIMT_KCP;
      for (int e = tid; e < SIMT_TT * (SIMT_KC / 4); e += SIMT_THREADS) {
        const int kv = e / SIMT_TT, r = e - kv * SIMT_TT;
        float4 v = *reinterpret_cast<const float4*>(src + r * SIMT_KCP + 4 * kv);
        Bt[(4 * kv + 0) * SIMT_LD + r] = v.x; Bt[(4 * kv + 1) * SIMT_LD + r] = v.y;
        Bt[(4 * kv + 2) * SIMT_LD + r] = v.z; Bt[(4 * kv + 3) * SIMT_LD + r] = v.w;
      }
    }
    __syncthreads();
    if (chunk == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    const int k0 = chunk * SIMT_KC;
    const int kn = min(SIMT_KC, d - k0);
#pragma unroll 8
    for (int k = 0; k < kn; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(As + (size_t)(k0 + k) * SIMT_LD + ty * 4);
      const float4 b = *reinterpret_cast<const float4*>(Bt + k * SIMT_LD + tx * 4);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float df = __fsub_rn(av[i], bv[j]);
          acc[i][j] = __fmaf_rn(df, df, acc[i][j]);
        }
    }
    if (chunk == n_chunks - 1) {
      const int64_t col0 = t_begin + (int64_t)tile * SIMT_TT + tx * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t col = col0 + j;
        if (col < t_end) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float v = acc[i][j];
            if (v < top[i].d1) {  // strict: columns arrive in increasing index per thread
              if (v < top[i].d0) { top[i].d1 = top[i].d0; top[i].i1 = top[i].i0; top[i].d0 = v; top[i].i0 = (int)col; }
              else { top[i].d1 = v; top[i].i1 = (int)col; }
            }
          }
        }
      }
    }
    __syncthreads();
  }

  // Lexicographic merge across the 16 threads (tx) that share the same 4 query rows.
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Top2<float> m; m.d0 = top[i].d0; m.d1 = top[i].d1; m.i0 = top[i].i0; m.i1 = top[i].i1;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      Top2<float> other;
      other.d0 = __shfl_xor_sync(0xffffffffu, m.d0, o); other.d1 = __shfl_xor_sync(0xffffffffu, m.d1, o);
      other.i0 = __shfl_xor_sync(0xffffffffu, m.i0, o); other.i1 = __shfl_xor_sync(0xffffffffu, m.i1, o);
      top2_merge<float>(m, other);
    }
    const int64_t lrow = q_base + ty * 4 + i;
    if (tx == 0 && lrow < n_rows_q) {
      int32_t* o = part_idx + ((int64_t)split * n_rows_q + lrow) * 2;
      o[0] = m.i0; o[1] = m.i1;
    }
  }
}

// One thread per (query, candidate): exact FP32 squared distance in canonical order, then the first
// n_q_per_block threads pick each query's lexicographic top-2.
//   cand [nparts][n_rows][m]  candidate train indices (-1 = empty), n_rows = n_list or nq.
//   row_list (optional): the queries handled are row_list[0..n_list); outputs are scattered to them.
constexpr int RERANK_QB = 32;
__global__ void __launch_bounds__(256)
l2_rerank_finalize_kernel(const float* __restrict__ q, int64_t nq, const float* __restrict__ t, int d,
                          const int32_t* __restrict__ cand, int nparts, int m,
                          const int32_t* __restrict__ row_list, int64_t n_list, int32_t* __restrict__ idx,
                          float* __restrict__ dist) {
  extern __shared__ float sd2[];  // [RERANK_QB][C]
  const int C = nparts * m;
  const int64_t n_rows = row_list ? n_list : nq;
  const int64_t q_base = (int64_t)blockIdx.x * RERANK_QB;
  int32_t* sidx = reinterpret_cast<int32_t*>(sd2 + RERANK_QB * C);
  for (int e = threadIdx.x; e < RERANK_QB * C; e += blockDim.x) {
    const int ql = e / C, c = e - ql * C;
    const int64_t lrow = q_base + ql;
    float d2 = FLT_MAX;
    int32_t ci = -1;
    if (lrow < n_rows) {
      const int p = c / m, mm = c - p * m;
      ci = cand[((int64_t)p * n_rows + lrow) * m + mm];
      if (ci >= 0) {
        const int64_t row = row_list ? (int64_t)row_list[lrow] : lrow;
        const float4* a = reinterpret_cast<const float4*>(q + row * d);
        const float4* b = reinterpret_cast<const float4*>(t + (int64_t)ci * d);
        float s = 0.f;
        for (int kv = 0; kv < (d >> 2); ++kv) {
          const float4 x = __ldg(a + kv), y = __ldg(b + kv);
          float df;
          df = __fsub_rn(x.x, y.x); s = __fmaf_rn(df, df, s);
          df = __fsub_rn(x.y, y.y); s = __fmaf_rn(df, df, s);
          df = __fsub_rn(x.z, y.z); s = __fmaf_rn(df, df, s);
          df = __fsub_rn(x.w, y.w); s = __fmaf_rn(df, df, s);
        }
        d2 = s;
      }
    }
    sd2[e] = d2; sidx[e] = ci;
  }
  __syncthreads();
  if (threadIdx.x < RERANK_QB) {
    const int64_t lrow = q_base + threadIdx.x;
    if (lrow < n_rows) {
      Top2<float> best; best.d0 = best.d1 = 0.f; best.i0 = best.i1 = -1;
      for (int c = 0; c < C; ++c) {
        const int32_t ci = sidx[threadIdx.x * C + c];
        // the same train row may be offered twice (tensor candidates + fallback): keep one copy
        if (ci >= 0 && ci != best.i0 && ci != best.i1) top2_insert<float>(best, sd2[threadIdx.x * C + c], ci);
      }
      const int64_t row = row_list ? (int64_t)row_list[lrow] : lrow;
      idx[row * 2] = best.i0; idx[row * 2 + 1] = best.i1;
      dist[row * 2] = best.i0 < 0 ? 0.f : __fsqrt_rn(best.d0);
      dist[row * 2 + 1] = best.i1 < 0 ? 0.f : __fsqrt_rn(best.d1);
    }
  }
}

int l2_simt_splits(int64_t nq_rows, int64_t nt) {
  const int64_t tiles = (nq_rows + SIMT_TQ - 1) / SIMT_TQ;
  const int64_t want = (int64_t)sm_count() * 4;
  int64_t s = (want + tiles - 1) / tiles;
  const int64_t max_s = (nt + 4095) / 4096;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return (int)s;
}

size_t l2_simt_smem_bytes(int d) {
  return ((size_t)d * SIMT_LD + 2 * SIMT_TT * SIMT_KCP + SIMT_KC * SIMT_LD) * sizeof(float);
}

// Brute-force top-2 candidate indices for nq rows (or for row_list) -> part_idx [splits][rows][2].
int l2_simt_candidates(const float* q, int64_t nq, const float* t, int64_t nt, int d, const int32_t* row_list,
                       int64_t n_list, int splits, int32_t* part_idx, cudaStream_t st) {
  const int64_t rows = row_list ? n_list : nq;
  if (rows == 0) return HPUSM_OK;
  const size_t smem = l2_simt_smem_bytes(d);
  static std::atomic<bool> attr_done[64];
  int dev = 0;
  HPUSM_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev].load(std::memory_order_acquire)) {
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_simt_knn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)l2_simt_smem_bytes(512)));
    attr_done[dev].store(true, std::memory_order_release);
  }
  const int64_t tiles = (rows + SIMT_TQ - 1) / SIMT_TQ;
  const int64_t rows_per_split = nt == 0 ? 1 : (nt + splits - 1) / splits;
  dim3 grid((unsigned)tiles, (unsigned)splits);
  // bracketed for the roofline only when it is the whole sweep (as a fallback of the split engine it is a side show)
  if (row_list) HPUSM_LAUNCH(l2_simt_knn2_kernel, grid, SIMT_THREADS, smem, st, q, nq, t, nt, d, rows_per_split, row_list,
                             n_list, part_idx);
  else HPUSM_LAUNCH_TIMED(l2_simt_knn2_kernel, grid, SIMT_THREADS, smem, st, q, nq, t, nt, d, rows_per_split, row_list,
                          n_list, part_idx);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int l2_rerank_finalize(const float* q, int64_t nq, const float* t, int d, const int32_t* cand, int nparts,
                       int m, const int32_t* row_list, int64_t n_list, int32_t* idx, float* dist,
                       cudaStream_t st) {
  const int64_t rows = row_list ? n_list : nq;
  if (rows == 0) return HPUSM_OK;
  const int C = nparts * m;
  const size_t smem = (size_t)RERANK_QB * C * 8;
  HPUSM_REQUIRE(smem <= 48 * 1024, HPUSM_ERR_UNSUPPORTED, "l2 re-rank: %d candidates per query is too many", C);
  HPUSM_LAUNCH(l2_rerank_finalize_kernel, (unsigned)((rows + RERANK_QB - 1) / RERANK_QB), 256, smem, st, q, nq,
               t, d, cand, nparts, m, row_list, n_list, idx, dist);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

}  // namespace hpusm
