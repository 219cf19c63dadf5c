// ORB description on the GPU (first slice): cv2.ORB.compute for given keypoints, bit for bit.
//
// The reference describes both images of every pair on the CPU (urbanscene/urban_scene.py:31-32,
// detector.detectAndCompute with cv2.ORB_create(nfeatures=100000, scaleFactor=1.2, nlevels=8, edgeThreshold=31,
// firstLevel=0, WTA_K=2, scoreType=ORB_FAST_SCORE, patchSize=31), urbanscene/features.py:19-20); with matching at ~2 ms
// that step is the scene pipeline.  hpusm_orb_compute describes keypoints OpenCV detected; hpusm_orb_detect_and_compute
// also detects them (FAST-9/16 + non-maximum suppression + border filter + intensity-centroid orientation, the
// reference's ORB_FAST_SCORE configuration): pyramid, detection, smoothing and the rotated 256-pair rBRIEF tests run
// here and leave keypoints and descriptors in HBM, where hpusm_knn2_hamming consumes them without a host round trip.  OpenCV is not vendored under /root/reference; the
// arithmetic is restated in oracle/orb_oracle.py (pinned bit for bit on cv2's own descriptors) and mirrored here:
//   * level L = resize(level L-1, INTER_LINEAR_EXACT): separable bilinear, Q8.8 weights (fraction rounded to 1/256),
//     Q16.16 accumulation, one rounding; sizes (cvRound(cols / s_L), cvRound(rows / s_L)), s_L = (float)pow(1.2, L);
//   * GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) as ORB runs it (generic separable filter on a view of its pyramid
//     atlas, NOT OpenCV's fixed-point Gaussian): uint8 -> float32 rows by a left-to-right FMA chain -> columns in the
//     symmetric form k3*c + k2*(u1+d1) + k1*(u2+d2) + k0*(u3+d3) with FMAs -> cvRound;
//   * descriptor bit k = level[cy + cvRound(x0*b + y0*a)][cx + cvRound(x0*a - y0*b)] < the same for (x1, y1), float32
//     products and sums without contraction, (a, b) = (cos, sin) of the keypoint angle as computed on the host.
// The 256 test pairs are OpenCV's learned pattern, read off the binary by oracle/record_golden.py (record_orb_pattern).
#include <math.h>
#include "common.cuh"

namespace hpusm {

constexpr int ORB_MAX_LEVELS = 16;

struct OrbLevels {
  int64_t off[ORB_MAX_LEVELS];  // byte offset of the level inside the pyramid buffer
  int w[ORB_MAX_LEVELS], h[ORB_MAX_LEVELS];
  float scale[ORB_MAX_LEVELS];  // (float)pow(scale_factor, level)
  int n;
};

// (x0, y0, x1, y1) of the 256 tests (tests/golden/orb_pattern.npy)
__constant__ signed char c_orb_pattern[256 * 4] = {
#include "orb_pattern.inc"
};

// ---- INTER_LINEAR_EXACT, uint8, one channel ----------------------------------------------------------------
__device__ __forceinline__ void linear_exact_coeff(int d, double scale, int src_n, int& i, int& q) {
  const double f = __dadd_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), -0.5);
  const double fl = floor(f);
  i = (int)fl;
  double a = __dsub_rn(f, fl);
  if (i < 0) { i = 0; a = 0.0; }
  if (i >= src_n - 1) { i = src_n - 1; a = 0.0; }
  q = (int)floor(__dadd_rn(__dmul_rn(a, 256.0), 0.5));
}

__global__ void orb_resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, int64_t spitch, uint8_t* __restrict__ dst,
                                  int dw, int dh, double scale_x, double scale_y) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dw || y >= dh) return;
  int ix, ax, iy, ay;
  linear_exact_coeff(x, scale_x, sw, ix, ax);
  linear_exact_coeff(y, scale_y, sh, iy, ay);
  const int x1 = min(ix + 1, sw - 1), y1 = min(iy + 1, sh - 1);
  const uint8_t* r0 = src + (int64_t)iy * spitch;
  const uint8_t* r1 = src + (int64_t)y1 * spitch;
  const int top = (256 - ax) * (int)r0[ix] + ax * (int)r0[x1];  // Q8.8
  const int bot = (256 - ax) * (int)r1[ix] + ax * (int)r1[x1];
  const int v = (256 - ay) * top + ay * bot;                     // Q16.16
  dst[(int64_t)y * dw + x] = (uint8_t)((v + 32768) >> 16);
}

// ---- 7x7 Gaussian, sigma 2, BORDER_REFLECT_101, OpenCV's float path ----------------------------------------------
constexpr int BL_TX = 32, BL_TY = 16;

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

__global__ void __launch_bounds__(BL_TX * BL_TY)
orb_blur_kernel(const uint8_t* __restrict__ src, int w, int h, int64_t spitch, uint8_t* __restrict__ dst) {
  __shared__ float rows[BL_TY + 6][BL_TX];
  // cv2.getGaussianKernel(7, 2, CV_32F)
  const float k0 = __uint_as_float(0x3D8FAFB1u), k1 = __uint_as_float(0x3E06387Eu), k2 = __uint_as_float(0x3E434A39u),
              k3 = __uint_as_float(0x3E5D4AE0u);
  const int x = blockIdx.x * BL_TX + threadIdx.x, y0 = blockIdx.y * BL_TY;
  // row pass for the BL_TY + 6 source rows this tile needs: left-to-right FMA chain
  for (int r = threadIdx.y; r < BL_TY + 6; r += BL_TY) {
    const int sy = reflect101(y0 + r - 3, h);
    const uint8_t* row = src + (int64_t)sy * spitch;
    float t = 0.f;
    if (x < w) {
      const float taps[7] = {k0, k1, k2, k3, k2, k1, k0};
      t = __fmul_rn((float)row[reflect101(x - 3, w)], taps[0]);
#pragma unroll
      for (int i = 1; i < 7; ++i) t = __fmaf_rn((float)row[reflect101(x - 3 + i, w)], taps[i], t);
    }
    rows[r][threadIdx.x] = t;
  }
  __syncthreads();
  const int y = y0 + threadIdx.y;
  if (x >= w || y >= h) return;
  const int r = threadIdx.y + 3;
  float c = __fmul_rn(rows[r][threadIdx.x], k3);
  c = __fmaf_rn(__fadd_rn(rows[r - 1][threadIdx.x], rows[r + 1][threadIdx.x]), k2, c);
  c = __fmaf_rn(__fadd_rn(rows[r - 2][threadIdx.x], rows[r + 2][threadIdx.x]), k1, c);
  c = __fmaf_rn(__fadd_rn(rows[r - 3][threadIdx.x], rows[r + 3][threadIdx.x]), k0, c);
  const int v = __float2int_rn(c);
  dst[(int64_t)y * w + x] = (uint8_t)min(max(v, 0), 255);
}

// ---- rBRIEF: one warp per keypoint, lane = descriptor byte ----------------------------------------------------------
__global__ void __launch_bounds__(256)
orb_describe_kernel(const uint8_t* __restrict__ pyr, OrbLevels lv, const int32_t* __restrict__ kp_lxy,
                    const float* __restrict__ kp_cs, int64_t n, uint8_t* __restrict__ desc) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= n) return;
  int level = kp_lxy[3 * k];
  level = min(max(level, 0), lv.n - 1);
  const int cx = kp_lxy[3 * k + 1], cy = kp_lxy[3 * k + 2];
  const float a = kp_cs[2 * k], b = kp_cs[2 * k + 1];
  const int w = lv.w[level], h = lv.h[level];
  const uint8_t* img = pyr + lv.off[level];
  auto sample = [&](int px, int py) -> int {
    const float fx = (float)px, fy = (float)py;
    const int ix = __float2int_rn(__fsub_rn(__fmul_rn(fx, a), __fmul_rn(fy, b)));
    const int iy = __float2int_rn(__fadd_rn(__fmul_rn(fx, b), __fmul_rn(fy, a)));
    const int x = min(max(cx + ix, 0), w - 1), y = min(max(cy + iy, 0), h - 1);  // OpenCV keeps keypoints 31 px inside
    return (int)__ldg(img + (int64_t)y * w + x);
  };
  unsigned byte = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const signed char* p = c_orb_pattern + 4 * (lane * 8 + t);
    const int t0 = sample(p[0], p[1]), t1 = sample(p[2], p[3]);
    byte |= (t0 < t1 ? 1u : 0u) << t;
  }
  desc[k * 32 + lane] = (uint8_t)byte;
}

// ---- detection: FAST-9/16 + non-maximum suppression + border filter + intensity-centroid angle ----------------------------
// cv2.ORB.detect restated (oracle/orb_oracle.py: detect, pinned on cv2's own keypoint lists): per pyramid level (the
// UNSMOOTHED one) the FAST score of every pixel, survivors of the 3x3 non-maximum suppression that lie edge_threshold
// inside the level, in row-major order; orientation = fastAtan2(m01, m10) over the 31-pixel circular patch.

// cornerScore<16> of OpenCV where the pixel is a FAST-9 corner for `threshold`, else 0.
__device__ __forceinline__ int fast_score(const uint8_t* __restrict__ p, int64_t pitch, int threshold) {
  const int v = p[0];
  int d[16];
  d[0] = v - p[3 * pitch];      d[1] = v - p[3 * pitch + 1];  d[2] = v - p[2 * pitch + 2];  d[3] = v - p[pitch + 3];
  d[4] = v - p[3];              d[5] = v - p[-pitch + 3];     d[6] = v - p[-2 * pitch + 2]; d[7] = v - p[-3 * pitch + 1];
  d[8] = v - p[-3 * pitch];     d[9] = v - p[-3 * pitch - 1]; d[10] = v - p[-2 * pitch - 2]; d[11] = v - p[-pitch - 3];
  d[12] = v - p[-3];            d[13] = v - p[pitch - 3];     d[14] = v - p[2 * pitch - 2]; d[15] = v - p[3 * pitch - 1];
  // quick reject: 9 contiguous of 16 need at least 9 of one polarity
  unsigned dark = 0, bright = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { dark |= (d[i] > threshold ? 1u : 0u) << i; bright |= (d[i] < -threshold ? 1u : 0u) << i; }
  auto has_arc = [](unsigned m) {
    m |= m << 16;
    unsigned r = m & (m >> 1);   // runs of 2
    r &= r >> 2;                 // runs of 4
    r &= r >> 4;                 // runs of 8
    r &= m >> 8;                 // runs of 9
    return (r & 0xFFFFu) != 0u;
  };
  if (!has_arc(dark) && !has_arc(bright)) return 0;
  // largest t' for which it is still a corner: max over the 16 arcs of min |v - c| (one polarity), by doubling
  int best = -1000;
#pragma unroll
  for (int pol = 0; pol < 2; ++pol) {
    int w2[16], w4[16], w8[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { const int a = pol ? -d[i] : d[i], b = pol ? -d[(i + 1) & 15] : d[(i + 1) & 15]; w2[i] = min(a, b); }
#pragma unroll
    for (int i = 0; i < 16; ++i) w4[i] = min(w2[i], w2[(i + 2) & 15]);
#pragma unroll
    for (int i = 0; i < 16; ++i) w8[i] = min(w4[i], w4[(i + 4) & 15]);
#pragma unroll
    for (int i = 0; i < 16; ++i) best = max(best, min(w8[i], pol ? -d[(i + 8) & 15] : d[(i + 8) & 15]));
  }
  return best > threshold ? best - 1 : 0;
}

__global__ void orb_fast_score_kernel(const uint8_t* __restrict__ img, int w, int h, int64_t pitch, int threshold,
                                      uint8_t* __restrict__ score) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  int s = 0;
  if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) s = fast_score(img + (int64_t)y * pitch + x, pitch, threshold);
  score[(int64_t)y * w + x] = (uint8_t)s;  // scores are at most 254
}

// One warp per row of one level: survivors of the non-maximum suppression inside the border, in increasing x.
// EMIT == false: counts them into row_count; EMIT == true: writes them behind row_offset (an exclusive scan of the counts).
template <bool EMIT>
__global__ void __launch_bounds__(128)
orb_nms_kernel(const uint8_t* __restrict__ score, int w, int h, int edge, int row0, int32_t* __restrict__ row_count,
               const int32_t* __restrict__ row_offset, int level, int64_t capacity, int32_t* __restrict__ kp_lxy,
               float* __restrict__ kp_response) {
  const int lane = threadIdx.x & 31;
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (y >= h) return;
  int n = 0;
  int64_t out = EMIT ? row_offset[row0 + y] : 0;
  if (y >= edge && y < h - edge) {
    const uint8_t* r0 = score + (int64_t)(y - 1) * w;
    const uint8_t* r1 = r0 + w;
    const uint8_t* r2 = r1 + w;
    for (int x0 = edge; x0 < w - edge; x0 += 32) {
      const int x = x0 + lane;
      bool keep = false;
      int s = 0;
      if (x < w - edge) {
        s = r1[x];
        keep = s > 0 && s > r0[x - 1] && s > r0[x] && s > r0[x + 1] && s > r1[x - 1] && s > r1[x + 1] && s > r2[x - 1] &&
               s > r2[x] && s > r2[x + 1];
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (EMIT && keep) {
        const int64_t k = out + n + __popc(bal & ((1u << lane) - 1u));
        if (k < capacity) {
          kp_lxy[3 * k] = level; kp_lxy[3 * k + 1] = x; kp_lxy[3 * k + 2] = y;
          kp_response[k] = (float)s;
        }
      }
      n += __popc(bal);
    }
  }
  if (!EMIT && lane == 0) row_count[row0 + y] = n;
}

// Exclusive scan of the row counts of all levels (a few thousand rows: one block), per-level totals, grand total.
__global__ void __launch_bounds__(1024)
orb_scan_kernel(const int32_t* __restrict__ row_count, int nrows, int32_t* __restrict__ row_offset, OrbLevels lv,
                int32_t* __restrict__ level_counts) {
  __shared__ int part[1024];
  const int tid = threadIdx.x;
  const int per = (nrows + 1023) / 1024;
  const int r0 = tid * per, r1 = min(nrows, r0 + per);
  int sum = 0;
  for (int r = r0; r < r1; ++r) sum += row_count[r];
  part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < 1024; ++t) { const int v = part[t]; part[t] = run; run += v; }
  }
  __syncthreads();
  int run = part[tid];
  for (int r = r0; r < r1; ++r) { row_offset[r] = run; run += row_count[r]; }
  __syncthreads();
  if (tid == 0) {
    int row = 0, total = 0;
    for (int l = 0; l < lv.n; ++l) {
      int c = 0;
      for (int y = 0; y < lv.h[l]; ++y) c += row_count[row + y];
      level_counts[l] = c;
      total += c;
      row += lv.h[l];
    }
    level_counts[lv.n] = total;
  }
}

// cv::fastAtan2 (degrees): float32, one rounding per operation
__device__ __forceinline__ float cv_fast_atan2(float y, float x) {
  const float k = (float)(180.0 / 3.14159265358979323846);
  const float p1 = __fmul_rn(0.9997878412794807f, k), p3 = __fmul_rn(-0.3258083974640975f, k),
              p5 = __fmul_rn(0.1555786518463281f, k), p7 = __fmul_rn(-0.04432655554792128f, k);
  const float eps = (float)2.2204460492503131e-16;
  const float ax = fabsf(x), ay = fabsf(y);
  float a;
  if (ax >= ay) {
    const float c = __fdiv_rn(ay, __fadd_rn(ax, eps)), c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    const float c = __fdiv_rn(ax, __fadd_rn(ay, eps)), c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0.f) a = __fsub_rn(180.f, a);
  if (y < 0.f) a = __fsub_rn(360.f, a);
  return a;
}

// Half-widths of the rows of ORB's circular patch of radius 15 (computeKeyPoints' umax; oracle umax_table()).
__constant__ int c_orb_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// One warp per keypoint: lane = patch row v = lane - 15; the keypoint's final fields.
__global__ void __launch_bounds__(256)
orb_angle_kernel(const uint8_t* __restrict__ raw, OrbLevels lv, const int32_t* __restrict__ total, int64_t capacity,
                 const int32_t* __restrict__ kp_lxy, float* __restrict__ kp_xy, int32_t* __restrict__ kp_octave,
                 float* __restrict__ kp_angle, float* __restrict__ kp_size, float* __restrict__ kp_cs, int patch_size) {
  const int lane = threadIdx.x & 31;
  const int64_t n = min((int64_t)*total, capacity);
  for (int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < n; k += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int level = kp_lxy[3 * k], x = kp_lxy[3 * k + 1], y = kp_lxy[3 * k + 2];
    const int w = lv.w[level];
    const uint8_t* img = raw + lv.off[level];
    int m01 = 0, m10 = 0;
    if (lane < 31) {
      const int v = lane - 15;
      const int d = c_orb_umax[v < 0 ? -v : v];
      const uint8_t* row = img + (int64_t)(y + v) * w + x;
      int s = 0, us = 0;
      for (int u = -d; u <= d; ++u) { const int val = row[u]; s += val; us += u * val; }
      m01 = v * s;
      m10 = us;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m01 += __shfl_xor_sync(0xffffffffu, m01, o); m10 += __shfl_xor_sync(0xffffffffu, m10, o); }
    if (lane == 0) {
      const float ang = cv_fast_atan2((float)m01, (float)m10);
      const float sf = lv.scale[level];
      kp_xy[2 * k] = __fmul_rn((float)x, sf);
      kp_xy[2 * k + 1] = __fmul_rn((float)y, sf);
      kp_octave[k] = level;
      kp_angle[k] = ang;
      kp_size[k] = __fmul_rn((float)patch_size, sf);
      const double ar = (double)__fmul_rn(ang, (float)(3.14159265358979323846 / 180.0));
      kp_cs[2 * k] = (float)cos(ar);
      kp_cs[2 * k + 1] = (float)sin(ar);
    }
  }
}

// rBRIEF over the keypoints a detection pass left on the device (their number is only known there).
__global__ void __launch_bounds__(256)
orb_describe_counted_kernel(const uint8_t* __restrict__ pyr, OrbLevels lv, const int32_t* __restrict__ kp_lxy,
                            const float* __restrict__ kp_cs, const int32_t* __restrict__ total, int64_t capacity,
                            uint8_t* __restrict__ desc);

static int orb_levels(int width, int height, int nlevels, double scale_factor, OrbLevels& lv) {
  if (nlevels < 1 || nlevels > ORB_MAX_LEVELS || width < 1 || height < 1 || !(scale_factor > 1.0)) return 0;
  lv.n = nlevels;
  int64_t off = 0;
  for (int l = 0; l < nlevels; ++l) {
    const float s = (float)pow(scale_factor, (double)l);
    lv.scale[l] = s;
    lv.w[l] = l == 0 ? width : (int)lrintf((float)width / s);
    lv.h[l] = l == 0 ? height : (int)lrintf((float)height / s);
    if (lv.w[l] < 1 || lv.h[l] < 1) return 0;
    lv.off[l] = off;
    off += (int64_t)align_up((size_t)lv.w[l] * lv.h[l], 256);
  }
  return 1;
}


__global__ void __launch_bounds__(256)
orb_describe_counted_kernel(const uint8_t* __restrict__ pyr, OrbLevels lv, const int32_t* __restrict__ kp_lxy,
                            const float* __restrict__ kp_cs, const int32_t* __restrict__ total, int64_t capacity,
                            uint8_t* __restrict__ desc) {
  const int lane = threadIdx.x & 31;
  const int64_t n = min((int64_t)*total, capacity);
  for (int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < n; k += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int level = kp_lxy[3 * k], cx = kp_lxy[3 * k + 1], cy = kp_lxy[3 * k + 2];
    const float a = kp_cs[2 * k], b = kp_cs[2 * k + 1];
    const int w = lv.w[level], h = lv.h[level];
    const uint8_t* img = pyr + lv.off[level];
    unsigned byte = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const signed char* p = c_orb_pattern + 4 * (lane * 8 + t);
      int val[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float fx = (float)p[2 * e], fy = (float)p[2 * e + 1];
        const int ix = __float2int_rn(__fsub_rn(__fmul_rn(fx, a), __fmul_rn(fy, b)));
        const int iy = __float2int_rn(__fadd_rn(__fmul_rn(fx, b), __fmul_rn(fy, a)));
        const int x = min(max(cx + ix, 0), w - 1), y = min(max(cy + iy, 0), h - 1);
        val[e] = (int)__ldg(img + (int64_t)y * w + x);
      }
      byte |= (val[0] < val[1] ? 1u : 0u) << t;
    }
    desc[k * 32 + lane] = (uint8_t)byte;
  }
}

// raw levels (each the source of the next) and smoothed levels into the two halves of the workspace
static int orb_build_pyramid(const uint8_t* img, int64_t pitch, const OrbLevels& lv, uint8_t* raw, uint8_t* smooth,
                             cudaStream_t st) {
  const dim3 rb(32, 8);
  for (int l = 0; l < lv.n; ++l) {
    const uint8_t* src = l == 0 ? img : raw + lv.off[l];
    const int64_t sp = l == 0 ? pitch : lv.w[l];
    if (l > 0) {
      const uint8_t* prev = l == 1 ? img : raw + lv.off[l - 1];
      const int64_t pp = l == 1 ? pitch : lv.w[l - 1];
      const dim3 rg((lv.w[l] + 31) / 32, (lv.h[l] + 7) / 8);
      HPUSM_LAUNCH(orb_resize_kernel, rg, rb, 0, st, prev, lv.w[l - 1], lv.h[l - 1], pp, raw + lv.off[l], lv.w[l], lv.h[l],
                   (double)lv.w[l - 1] / (double)lv.w[l], (double)lv.h[l - 1] / (double)lv.h[l]);
    }
    const dim3 bg((lv.w[l] + BL_TX - 1) / BL_TX, (lv.h[l] + BL_TY - 1) / BL_TY);
    HPUSM_LAUNCH(orb_blur_kernel, bg, dim3(BL_TX, BL_TY), 0, st, src, lv.w[l], lv.h[l], sp, smooth + lv.off[l]);
  }
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

}  // namespace hpusm

using namespace hpusm;

extern "C" {

size_t hpusm_orb_workspace_bytes(int width, int height, int nlevels, double scale_factor) {
  OrbLevels lv;
  if (!orb_levels(width, height, nlevels, scale_factor, lv)) return 256;
  const size_t one = (size_t)lv.off[nlevels - 1] + align_up((size_t)lv.w[nlevels - 1] * lv.h[nlevels - 1], 256);
  return 2 * one;  // the raw levels (each is the source of the next) and the smoothed ones
}

int hpusm_orb_compute(const uint8_t* img, int width, int height, int64_t pitch, int nlevels, double scale_factor,
                      const int32_t* kp_lxy, const float* kp_cs, int64_t n, uint8_t* desc, void* ws, size_t ws_bytes,
                      void* stream) {
  HPUSM_REQUIRE(n >= 0 && pitch >= width, HPUSM_ERR_INVALID, "hpusm_orb_compute: bad sizes");
  OrbLevels lv;
  HPUSM_REQUIRE(orb_levels(width, height, nlevels, scale_factor, lv), HPUSM_ERR_INVALID,
                "hpusm_orb_compute: %d x %d image, %d levels, scale factor %g", width, height, nlevels, scale_factor);
  HPUSM_REQUIRE(img && (n == 0 || (kp_lxy && kp_cs && desc)), HPUSM_ERR_INVALID, "hpusm_orb_compute: null pointer");
  const size_t need = hpusm_orb_workspace_bytes(width, height, nlevels, scale_factor);
  HPUSM_REQUIRE(ws && is_aligned(ws, 256) && ws_bytes >= need, HPUSM_ERR_WORKSPACE,
                "hpusm_orb_compute: workspace needs %zu bytes, got %zu", need, ws_bytes);
  cudaStream_t st = as_stream(stream);
  uint8_t* raw = reinterpret_cast<uint8_t*>(ws);
  uint8_t* smooth = raw + need / 2;
  const int rc = orb_build_pyramid(img, pitch, lv, raw, smooth, st);
  if (rc != HPUSM_OK) return rc;
  if (n > 0) {
    HPUSM_LAUNCH_TIMED(orb_describe_kernel, (unsigned)((n + 7) / 8), 256, 0, st, smooth, lv, kp_lxy, kp_cs, n, desc);
    HPUSM_CHECK_LAUNCH();
  }
  return HPUSM_OK;
}


static size_t orb_pyramid_bytes(const OrbLevels& lv) {
  return (size_t)lv.off[lv.n - 1] + align_up((size_t)lv.w[lv.n - 1] * lv.h[lv.n - 1], 256);
}

size_t hpusm_orb_detect_workspace_bytes(int width, int height, int nlevels, double scale_factor) {
  OrbLevels lv;
  if (!orb_levels(width, height, nlevels, scale_factor, lv)) return 256;
  int rows = 0;
  for (int l = 0; l < nlevels; ++l) rows += lv.h[l];
  // raw levels, smoothed levels, FAST score maps, per-row counts and offsets
  return 3 * orb_pyramid_bytes(lv) + 2 * align_up((size_t)(rows + 1) * sizeof(int32_t), 256);
}

int hpusm_orb_detect_and_compute(const uint8_t* img, int width, int height, int64_t pitch, int nlevels, double scale_factor,
                                 int fast_threshold, int edge_threshold, int patch_size, int64_t capacity, float* kp_xy,
                                 int32_t* kp_octave, float* kp_angle, float* kp_response, float* kp_size, int32_t* kp_lxy,
                                 float* kp_cs, uint8_t* desc, int32_t* level_counts, void* ws, size_t ws_bytes, void* stream) {
  HPUSM_REQUIRE(capacity > 0 && pitch >= width && fast_threshold >= 0 && fast_threshold < 255 && edge_threshold >= 19 &&
                patch_size == 31, HPUSM_ERR_INVALID,
                "hpusm_orb_detect_and_compute: bad arguments (the kernels are built for patchSize 31 and an edge threshold "
                "that keeps the patch and the rotated tests inside the level, >= 19)");
  OrbLevels lv;
  HPUSM_REQUIRE(orb_levels(width, height, nlevels, scale_factor, lv), HPUSM_ERR_INVALID,
                "hpusm_orb_detect_and_compute: %d x %d image, %d levels, scale factor %g", width, height, nlevels, scale_factor);
  HPUSM_REQUIRE(img && kp_xy && kp_octave && kp_angle && kp_response && kp_size && kp_lxy && kp_cs && level_counts,
                HPUSM_ERR_INVALID, "hpusm_orb_detect_and_compute: null pointer");
  const size_t need = hpusm_orb_detect_workspace_bytes(width, height, nlevels, scale_factor);
  HPUSM_REQUIRE(ws && is_aligned(ws, 256) && ws_bytes >= need, HPUSM_ERR_WORKSPACE,
                "hpusm_orb_detect_and_compute: workspace needs %zu bytes, got %zu", need, ws_bytes);
  cudaStream_t st = as_stream(stream);
  const size_t pyr = orb_pyramid_bytes(lv);
  int rows = 0;
  for (int l = 0; l < nlevels; ++l) rows += lv.h[l];
  uint8_t* raw = reinterpret_cast<uint8_t*>(ws);
  uint8_t* smooth = raw + pyr;
  uint8_t* score = smooth + pyr;
  int32_t* row_count = reinterpret_cast<int32_t*>(score + pyr);
  int32_t* row_offset = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(row_count) + align_up((size_t)(rows + 1) * sizeof(int32_t), 256));
  int rc = orb_build_pyramid(img, pitch, lv, raw, smooth, st);
  if (rc != HPUSM_OK) return rc;
  const dim3 fb(32, 8);
  for (int pass = 0; pass < 2; ++pass) {
    int row0 = 0;
    for (int l = 0; l < nlevels; ++l) {
      const uint8_t* src = l == 0 ? img : raw + lv.off[l];
      const int64_t sp = l == 0 ? pitch : lv.w[l];
      if (pass == 0) {
        const dim3 fg((lv.w[l] + 31) / 32, (lv.h[l] + 7) / 8);
        HPUSM_LAUNCH(orb_fast_score_kernel, fg, fb, 0, st, src, lv.w[l], lv.h[l], sp, fast_threshold, score + lv.off[l]);
        HPUSM_LAUNCH(orb_nms_kernel<false>, (unsigned)((lv.h[l] + 3) / 4), 128, 0, st, score + lv.off[l], lv.w[l], lv.h[l],
                     edge_threshold, row0, row_count, nullptr, l, capacity, nullptr, nullptr);
      } else {
        HPUSM_LAUNCH(orb_nms_kernel<true>, (unsigned)((lv.h[l] + 3) / 4), 128, 0, st, score + lv.off[l], lv.w[l], lv.h[l],
                     edge_threshold, row0, nullptr, row_offset, l, capacity, kp_lxy, kp_response);
      }
      row0 += lv.h[l];
    }
    if (pass == 0) HPUSM_LAUNCH(orb_scan_kernel, 1, 1024, 0, st, row_count, rows, row_offset, lv, level_counts);
    HPUSM_CHECK_LAUNCH();
  }
  // level 0 of `raw` is the caller's image: the angle kernel needs it at lv.off[0] with the level's own pitch
  HPUSM_CUDA_OK(cudaMemcpy2DAsync(raw, (size_t)width, img, (size_t)pitch, (size_t)width, (size_t)height, cudaMemcpyDeviceToDevice, st));
  const unsigned grid = (unsigned)sm_count() * 8;
  HPUSM_LAUNCH(orb_angle_kernel, grid, 256, 0, st, raw, lv, level_counts + nlevels, capacity, kp_lxy, kp_xy, kp_octave,
               kp_angle, kp_size, kp_cs, patch_size);
  if (desc) {
    HPUSM_LAUNCH_TIMED(orb_describe_counted_kernel, grid, 256, 0, st, smooth, lv, kp_lxy, kp_cs, level_counts + nlevels,
                       capacity, desc);
  }
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

}  // extern "C"
