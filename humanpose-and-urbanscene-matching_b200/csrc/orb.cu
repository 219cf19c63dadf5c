// ORB description on the GPU (first slice): cv2.ORB.compute for given keypoints, bit for bit.
//
// The reference describes both images of every pair on the CPU (urbanscene/urban_scene.py:31-32,
// detector.detectAndCompute with cv2.ORB_create(nfeatures=100000, scaleFactor=1.2, nlevels=8, edgeThreshold=31,
// firstLevel=0, WTA_K=2, scoreType=ORB_FAST_SCORE, patchSize=31), urbanscene/features.py:19-20); with matching at ~2 ms
// that step is the scene pipeline.  Detection (FAST + Harris/FAST score + orientation) stays with OpenCV; everything
// after it -- pyramid, smoothing, the rotated 256-pair rBRIEF tests -- runs here and leaves the descriptors in HBM, where
// hpusm_knn2_hamming consumes them without a host round trip.  OpenCV is not vendored under /root/reference; the
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

static int orb_levels(int width, int height, int nlevels, double scale_factor, OrbLevels& lv) {
  if (nlevels < 1 || nlevels > ORB_MAX_LEVELS || width < 1 || height < 1 || !(scale_factor > 1.0)) return 0;
  lv.n = nlevels;
  int64_t off = 0;
  for (int l = 0; l < nlevels; ++l) {
    const float s = (float)pow(scale_factor, (double)l);
    lv.w[l] = l == 0 ? width : (int)lrintf((float)width / s);
    lv.h[l] = l == 0 ? height : (int)lrintf((float)height / s);
    if (lv.w[l] < 1 || lv.h[l] < 1) return 0;
    lv.off[l] = off;
    off += (int64_t)align_up((size_t)lv.w[l] * lv.h[l], 256);
  }
  return 1;
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
  const dim3 rb(32, 8);
  for (int l = 0; l < nlevels; ++l) {
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
  if (n > 0) {
    HPUSM_LAUNCH_TIMED(orb_describe_kernel, (unsigned)((n + 7) / 8), 256, 0, st, smooth, lv, kp_lxy, kp_cs, n, desc);
    HPUSM_CHECK_LAUNCH();
  }
  return HPUSM_OK;
}

}  // extern "C"
