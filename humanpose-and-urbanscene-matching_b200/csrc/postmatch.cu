// Post-match glue kept on the device so a retrieval query never bounces through the host:
//   - order-preserving compaction of the Lowe-ratio survivors + keypoint gather (features.py:49-53)
//   - projective point transform (cv2.perspectiveTransform at features.py:77, transformation.py:38)
#include "common.cuh"

namespace hpusm {

constexpr int GM_THREADS = 256;
constexpr int GM_ROWS = 4;                       // rows per thread
constexpr int GM_TILE = GM_THREADS * GM_ROWS;    // rows per CTA

__global__ void __launch_bounds__(GM_THREADS)
gm_count_kernel(const uint8_t* __restrict__ keep, int64_t nq, int32_t* __restrict__ block_counts) {
  __shared__ int wsum[GM_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * GM_TILE + (int64_t)threadIdx.x * GM_ROWS;
  int c = 0;
#pragma unroll
  for (int r = 0; r < GM_ROWS; ++r) c += (base + r < nq && keep[base + r]) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < GM_THREADS / 32; ++w) s += wsum[w];
    block_counts[blockIdx.x] = s;
  }
}

// Single CTA: exclusive scan of the per-CTA counts (a few thousand entries at most per pass).
__global__ void __launch_bounds__(1024)
gm_scan_kernel(int32_t* __restrict__ block_counts, int64_t nblocks, int32_t* __restrict__ total) {
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nblocks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    int v = i < nblocks ? block_counts[i] : 0;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = wsum[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      wsum[threadIdx.x] = w;
    }
    __syncthreads();
    const int warp_off = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0;
    const int incl = x + warp_off + carry;
    if (i < nblocks) block_counts[i] = incl - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(GM_THREADS)
gm_scatter_kernel(const float* __restrict__ kp_q, const float* __restrict__ kp_t, const int32_t* __restrict__ idx,
                  const uint8_t* __restrict__ keep, int64_t nq, const int32_t* __restrict__ block_offsets,
                  float* __restrict__ p_q, float* __restrict__ p_t, int32_t* __restrict__ sel) {
  __shared__ int wsum[GM_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * GM_TILE + (int64_t)threadIdx.x * GM_ROWS;
  bool k[GM_ROWS];
  int c = 0;
#pragma unroll
  for (int r = 0; r < GM_ROWS; ++r) { k[r] = base + r < nq && keep[base + r]; c += k[r] ? 1 : 0; }
  int x = c;
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
  __syncthreads();
  int warp_off = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) warp_off += wsum[w];
  int pos = block_offsets[blockIdx.x] + warp_off + x - c;
#pragma unroll
  for (int r = 0; r < GM_ROWS; ++r) {
    if (k[r]) {
      const int64_t row = base + r;
      const int32_t ti = idx[row * 2];
      reinterpret_cast<float2*>(p_q)[pos] = reinterpret_cast<const float2*>(kp_q)[row];
      reinterpret_cast<float2*>(p_t)[pos] = reinterpret_cast<const float2*>(kp_t)[ti];
      if (sel) sel[pos] = (int32_t)row;
      ++pos;
    }
  }
}

// cv2.perspectiveTransform on CV_32FC2 input with a CV_64F matrix: accumulate in double, w == 0 -> 0.
__global__ void perspective_transform_kernel(const float* __restrict__ pts, int64_t n, const double* __restrict__ H,
                                             float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 p = reinterpret_cast<const float2*>(pts)[i];
  const double x = p.x, y = p.y;
  double w = x * H[6] + y * H[7] + H[8];
  float2 o;
  if (fabs(w) > 2.220446049250313e-16) {
    w = 1.0 / w;
    o.x = (float)((x * H[0] + y * H[1] + H[2]) * w);
    o.y = (float)((x * H[3] + y * H[4] + H[5]) * w);
  } else {
    o.x = 0.f; o.y = 0.f;
  }
  reinterpret_cast<float2*>(out)[i] = o;
}

// ---- per-image match lists for batched RANSAC (query vs image database) ----------------------------------------
// pair_offsets[s] = sum over s' < s of (count[s'] >= min_count ? count[s'] : 0): single CTA, int64 running carry.
__global__ void __launch_bounds__(1024)
seg_offsets_kernel(const int32_t* __restrict__ count, int64_t nseg, int min_count, int64_t* __restrict__ offsets) {
  __shared__ long long wsum[32];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nseg; base += 1024) {
    const int64_t i = base + threadIdx.x;
    long long v = 0;
    if (i < nseg) { const int c = count[i]; v = c >= min_count ? c : 0; }
    long long x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      long long w = wsum[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      wsum[threadIdx.x] = w;
    }
    __syncthreads();
    const long long incl = x + ((threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0) + carry;
    if (i < nseg) offsets[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[nseg] = carry;
}

// One CTA per database image: order-preserving compaction of the query rows whose ratio test passed against this
// image (best_idx >= 0) into src = query keypoint, dst = database keypoint.  Images below min_count emit nothing.
__global__ void __launch_bounds__(256)
seg_gather_kernel(const int32_t* __restrict__ best_idx, int64_t nq, const float* __restrict__ kp_q,
                  const float* __restrict__ kp_t, const int64_t* __restrict__ offsets, float* __restrict__ src,
                  float* __restrict__ dst, int32_t* __restrict__ sel) {
  __shared__ int wsum[8];
  __shared__ int carry;
  const int64_t seg = blockIdx.x;
  const int64_t o0 = offsets[seg];
  if (offsets[seg + 1] == o0) return;
  const int32_t* b = best_idx + seg * nq;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nq; base += 256) {
    const int64_t i = base + threadIdx.x;
    const int32_t t = i < nq ? b[i] : -1;
    const int k = t >= 0 ? 1 : 0;
    int x = k;
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wsum[w];
    const int pos = carry + woff + x - k;
    if (k) {
      reinterpret_cast<float2*>(src)[o0 + pos] = reinterpret_cast<const float2*>(kp_q)[i];
      reinterpret_cast<float2*>(dst)[o0 + pos] = reinterpret_cast<const float2*>(kp_t)[t];
      if (sel) sel[o0 + pos] = (int32_t)i;
    }
    __syncthreads();
    if (threadIdx.x == 255) carry = pos + k;
    __syncthreads();
  }
}

// features.validate_homography (features.py:177-211) for a batch of 3x3 matrices; NaN (no model) -> invalid.
__global__ void validate_homography_kernel(const double* __restrict__ H, int64_t n, uint8_t* __restrict__ valid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* h = H + i * 9;
  bool ok = h[0] == h[0];
  // Python evaluates these in double with one rounding per operation: no FMA contraction here, or a singular
  // top-left block (h00*h11 == h01*h10 after rounding) could come out as a tiny negative determinant
  const double det = __dsub_rn(__dmul_rn(h[0], h[4]), __dmul_rn(h[1], h[3]));
  if (det < 0) ok = false;
  const double n1 = sqrt(__dadd_rn(__dmul_rn(h[0], h[0]), __dmul_rn(h[1], h[1])));
  if (n1 > 4 || n1 < 0.1) ok = false;
  const double n2 = sqrt(__dadd_rn(__dmul_rn(h[3], h[3]), __dmul_rn(h[4], h[4])));
  if (n2 > 4 || n2 < 0.1) ok = false;
  const double n3 = sqrt(__dadd_rn(__dmul_rn(h[6], h[6]), __dmul_rn(h[7], h[7])));
  if (n3 > 0.002) ok = false;
  valid[i] = ok ? 1 : 0;
}

}  // namespace hpusm

using namespace hpusm;

extern "C" {

size_t hpusm_gather_matches_workspace_bytes(int64_t nq) {
  const int64_t nblocks = nq > 0 ? (nq + GM_TILE - 1) / GM_TILE : 1;
  return align_up((size_t)nblocks * sizeof(int32_t), 256);
}

int hpusm_gather_matches(const float* kp_q, const float* kp_t, const int32_t* idx, const uint8_t* keep,
                         int64_t nq, float* p_q, float* p_t, int32_t* sel, int32_t* count, void* ws,
                         size_t ws_bytes, void* stream) {
  HPUSM_REQUIRE(nq >= 0 && count, HPUSM_ERR_INVALID, "hpusm_gather_matches: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (nq == 0) {
    HPUSM_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
    return HPUSM_OK;
  }
  HPUSM_REQUIRE(kp_q && kp_t && idx && keep && p_q && p_t, HPUSM_ERR_INVALID, "hpusm_gather_matches: null pointer");
  const int64_t nblocks = (nq + GM_TILE - 1) / GM_TILE;
  HPUSM_REQUIRE(ws && is_aligned(ws, 256) && ws_bytes >= (size_t)nblocks * sizeof(int32_t), HPUSM_ERR_WORKSPACE,
                "hpusm_gather_matches: workspace needs %zu bytes, got %zu", (size_t)nblocks * sizeof(int32_t),
                ws_bytes);
  int32_t* bc = reinterpret_cast<int32_t*>(ws);
  HPUSM_LAUNCH(gm_count_kernel, (unsigned)nblocks, GM_THREADS, 0, st, keep, nq, bc);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(gm_scan_kernel, 1, 1024, 0, st, bc, nblocks, count);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(gm_scatter_kernel, (unsigned)nblocks, GM_THREADS, 0, st, kp_q, kp_t, idx, keep, nq, bc, p_q, p_t,
               sel);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_segment_matches(const int32_t* best_idx, const int32_t* count, int64_t nseg, int64_t nq, const float* kp_q,
                          const float* kp_t, int min_count, int64_t* pair_offsets, float* src, float* dst, int32_t* sel,
                          void* stream) {
  HPUSM_REQUIRE(nseg >= 0 && nq >= 0, HPUSM_ERR_INVALID, "hpusm_segment_matches: negative size");
  HPUSM_REQUIRE(pair_offsets, HPUSM_ERR_INVALID, "hpusm_segment_matches: null pair_offsets");
  cudaStream_t st = as_stream(stream);
  if (nseg == 0) {
    HPUSM_CUDA_OK(cudaMemsetAsync(pair_offsets, 0, sizeof(int64_t), st));
    return HPUSM_OK;
  }
  HPUSM_REQUIRE(best_idx && count && kp_q && kp_t && src && dst, HPUSM_ERR_INVALID, "hpusm_segment_matches: null pointer");
  HPUSM_REQUIRE(nseg < 0x7fffffff, HPUSM_ERR_UNSUPPORTED, "hpusm_segment_matches: too many segments");
  HPUSM_LAUNCH(seg_offsets_kernel, 1, 1024, 0, st, count, nseg, min_count, pair_offsets);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(seg_gather_kernel, (unsigned)nseg, 256, 0, st, best_idx, nq, kp_q, kp_t, pair_offsets, src, dst, sel);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_validate_homography_batch(const double* H, int64_t n, uint8_t* valid, void* stream) {
  HPUSM_REQUIRE(n >= 0, HPUSM_ERR_INVALID, "hpusm_validate_homography_batch: negative size");
  if (n == 0) return HPUSM_OK;
  HPUSM_REQUIRE(H && valid, HPUSM_ERR_INVALID, "hpusm_validate_homography_batch: null pointer");
  HPUSM_LAUNCH(validate_homography_kernel, (unsigned)((n + 127) / 128), 128, 0, as_stream(stream), H, n, valid);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_perspective_transform(const float* pts, int64_t n, const double* H, float* out, void* stream) {
  HPUSM_REQUIRE(n >= 0, HPUSM_ERR_INVALID, "hpusm_perspective_transform: negative size");
  if (n == 0) return HPUSM_OK;
  HPUSM_REQUIRE(pts && H && out, HPUSM_ERR_INVALID, "hpusm_perspective_transform: null pointer");
  HPUSM_LAUNCH(perspective_transform_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(stream), pts, n, H, out);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

}  // extern "C"
