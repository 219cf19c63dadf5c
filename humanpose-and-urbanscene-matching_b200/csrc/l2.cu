// hpusm_knn2_l2: engine selection and workspace carving for the L2 kNN (k=2) path.
// Replaces cv2.BFMatcher(cv2.NORM_L2).knnMatch(..., k=2) (reference urbanscene/features.py:42,:59).
#include <limits.h>
#include "common.cuh"
#include "l2_internal.cuh"

namespace hpusm {

static thread_local int t_last_mode = HPUSM_L2_SIMT;
static thread_local int64_t t_last_fallback = 0;

static bool tc_shape_ok(int d) { return d == 128; }  // SIFT width; other widths take the FP32 engine ...

// ... unless HPUSM_L2_AUTO sees a NARROWER descriptor (SURF: 64 floats, reference features.py:16) on a problem large enough
// to pay for it: the rows are zero-padded to 128 columns inside the workspace and take the tensor path.  Zeros change no
// distance, and the FP32 re-rank adds fma(0, 0, s) = s for them: the answer is bit for bit the d-wide one.
constexpr int64_t kPadMinPairs = (int64_t)1 << 24;
static bool pad_to_tc(int d, int mode, int64_t nq, int64_t nt) {
  return mode == HPUSM_L2_AUTO && d < 128 && l2_tc_available() && nt >= 2 && nq > 0 && nq * nt >= kPadMinPairs;
}
static size_t pad_bytes(int64_t n) { return align_up((size_t)n * 128 * sizeof(float), 1024); }

}  // namespace hpusm

using namespace hpusm;

extern "C" {

int hpusm_knn2_l2_last_mode(void) { return t_last_mode; }
int64_t hpusm_knn2_l2_last_fallback_rows(void) { return t_last_fallback; }

size_t hpusm_knn2_l2_workspace_bytes(int64_t nq, int64_t nt, int d, int mode) {
  if (nq <= 0) return 256;
  size_t simt = 256;
  {
    const int s = l2_simt_splits(nq, nt > 0 ? nt : 1);
    simt = align_up((size_t)s * nq * 2 * sizeof(int32_t), 256) + 256;
  }
  if (pad_to_tc(d, mode, nq, nt)) {
    const size_t tc = l2_tc_workspace_bytes(nq, nt, 128, mode) + pad_bytes(nq) + pad_bytes(nt);
    return tc > simt ? tc : simt;
  }
  if (mode == HPUSM_L2_SIMT || !tc_shape_ok(d)) return simt;
  const size_t tc = l2_tc_workspace_bytes(nq, nt, d, mode);
  return tc > simt ? tc : simt;
}

int hpusm_knn2_l2(const float* q, int64_t nq, const float* t, int64_t nt, int d, int mode, int32_t* idx,
                  float* dist, void* ws, size_t ws_bytes, void* stream) {
  HPUSM_REQUIRE(nq >= 0 && nt >= 0, HPUSM_ERR_INVALID, "hpusm_knn2_l2: negative size");
  HPUSM_REQUIRE(d > 0 && d % 4 == 0 && d <= 512, HPUSM_ERR_UNSUPPORTED,
                "hpusm_knn2_l2: descriptor length %d (need a multiple of 4, at most 512)", d);
  HPUSM_REQUIRE(mode >= HPUSM_L2_AUTO && mode <= HPUSM_L2_TC_Q16, HPUSM_ERR_INVALID, "hpusm_knn2_l2: bad mode %d", mode);
  if (nq == 0) return HPUSM_OK;
  HPUSM_REQUIRE(q && idx && dist && (t || nt == 0), HPUSM_ERR_INVALID, "hpusm_knn2_l2: null pointer");
  HPUSM_REQUIRE(is_aligned(q, 16) && is_aligned(t, 16), HPUSM_ERR_INVALID,
                "hpusm_knn2_l2: descriptor arrays must be 16-byte aligned");
  HPUSM_REQUIRE(nt < (int64_t)INT_MAX - 4096 && nq < (int64_t)INT_MAX - 4096, HPUSM_ERR_UNSUPPORTED,
                "hpusm_knn2_l2: size out of range");
  const size_t need = hpusm_knn2_l2_workspace_bytes(nq, nt, d, mode);
  HPUSM_REQUIRE(ws && is_aligned(ws, 256) && ws_bytes >= need, HPUSM_ERR_WORKSPACE,
                "hpusm_knn2_l2: workspace needs %zu bytes, got %zu", need, ws_bytes);
  cudaStream_t st = as_stream(stream);
  t_last_fallback = 0;

  if (pad_to_tc(d, mode, nq, nt)) {
    char* base = reinterpret_cast<char*>(ws);
    float* qp = reinterpret_cast<float*>(base);
    float* tp = reinterpret_cast<float*>(base + pad_bytes(nq));
    const size_t used = pad_bytes(nq) + pad_bytes(nt);
    HPUSM_CUDA_OK(cudaMemsetAsync(base, 0, used, st));
    HPUSM_CUDA_OK(cudaMemcpy2DAsync(qp, 128 * sizeof(float), q, (size_t)d * sizeof(float), (size_t)d * sizeof(float), (size_t)nq,
                                    cudaMemcpyDeviceToDevice, st));
    HPUSM_CUDA_OK(cudaMemcpy2DAsync(tp, 128 * sizeof(float), t, (size_t)d * sizeof(float), (size_t)d * sizeof(float), (size_t)nt,
                                    cudaMemcpyDeviceToDevice, st));
    int resolved = HPUSM_L2_AUTO;
    int64_t fallback = 0;
    int rc = l2_tc_knn2(qp, nq, tp, nt, 128, HPUSM_L2_AUTO, idx, dist, base + used, ws_bytes - used, st, &resolved, &fallback);
    t_last_mode = resolved;
    t_last_fallback = fallback;
    return rc;
  }
  int engine = mode;
  if (mode == HPUSM_L2_AUTO) engine = (l2_tc_available() && tc_shape_ok(d) && nt >= 2) ? HPUSM_L2_AUTO : HPUSM_L2_SIMT;
  if (mode == HPUSM_L2_TC_INT || mode == HPUSM_L2_TC_SPLIT || mode == HPUSM_L2_TC_I8 || mode == HPUSM_L2_TC_Q16)
    HPUSM_REQUIRE(tc_shape_ok(d), HPUSM_ERR_UNSUPPORTED,
                  "hpusm_knn2_l2: the tensor-core engine is built for d == 128, got %d", d);
  if (engine != HPUSM_L2_SIMT && nt < 2)
    HPUSM_REQUIRE(false, HPUSM_ERR_UNSUPPORTED,
                  "hpusm_knn2_l2: the tensor-core engines need at least 2 database rows (got %lld); use HPUSM_L2_AUTO or HPUSM_L2_SIMT",
                  (long long)nt);

  if (engine == HPUSM_L2_SIMT) {
    t_last_mode = HPUSM_L2_SIMT;
    if (nt == 0) {
      HPUSM_CUDA_OK(cudaMemsetAsync(idx, 0xff, (size_t)nq * 2 * sizeof(int32_t), st));
      HPUSM_CUDA_OK(cudaMemsetAsync(dist, 0, (size_t)nq * 2 * sizeof(float), st));
      return HPUSM_OK;
    }
    const int splits = l2_simt_splits(nq, nt);
    int32_t* part = reinterpret_cast<int32_t*>(ws);
    int rc = l2_simt_candidates(q, nq, t, nt, d, nullptr, 0, splits, part, st);
    if (rc != HPUSM_OK) return rc;
    return l2_rerank_finalize(q, nq, t, d, part, splits, 2, nullptr, 0, idx, dist, st);
  }
  int resolved = engine;
  int64_t fallback = 0;
  int rc = l2_tc_knn2(q, nq, t, nt, d, engine, idx, dist, ws, ws_bytes, st, &resolved, &fallback);
  t_last_mode = resolved;
  t_last_fallback = fallback;
  return rc;
}

}  // extern "C"
