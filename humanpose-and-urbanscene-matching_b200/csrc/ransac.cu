// K3: batched RANSAC homography.
//
// Replaces cv2.findHomography(src, dst, cv2.RANSAC, 5.0) at reference urbanscene/features.py:66,:69.  OpenCV is not
// vendored under /root/reference; its algorithm is restated function by function in oracle/ransac_oracle.c (pinned on
// cv2's own outputs, see the header there) and this file computes the same thing on the GPU:
//   * sampler.  HPUSM_RANSAC_SAMPLER_CV2 replays OpenCV's fixed-seed generator (RNG((uint64)-1), multiply-with-carry,
//     uniform(0, count), duplicate redraw, checkSubset, up to 10000 attempts) in a per-pair sequential pre-pass, so the
//     hypotheses -- and with them mask and H -- are cv2.findHomography's own.  HPUSM_RANSAC_SAMPLER_COUNTER draws from a
//     counter-based table keyed on (seed, pair, hypothesis, attempt): every hypothesis is a pure function of its key and
//     the lanes draw them in parallel.  Both redraw until OpenCV's checkSubset passes, so each of the nhyp iterations
//     scores a model (round 1 drew once and left 79 % of C4's hypotheses unscored).
//   * minimal solver: closed form in FP64 (unit square -> quad on both sides, H = S_dst * adj(S_src)), op for op the
//     oracle's, instead of OpenCV's 9x9 eigen decomposition (same unique homography, rounding apart).
//   * scoring.  One CTA per image pair, matches staged in shared memory as float4 pairs; a warp takes 32 hypotheses
//     (one per LANE, model in registers) and streams the matches as broadcast LDS.128 through sm_100's packed FP32
//     pipe.  Under the counter sampler the test is division free, (u - x'w)^2 + (v - y'w)^2 <= thr^2 w^2; under the cv2
//     sampler it is OpenCV's computeError, float32 with a correctly rounded division, one rounding per operation.
//   * the sequential early exit of RANSAC (RANSACUpdateNumIters) is replayed over the counts by one thread, so the
//     parallel evaluation returns what the sequential loop would.
//   * refinement (findHomography's post-processing): one WARP per pair.  runKernel's normalised DLT on the inliers of
//     the winning model (smallest eigenvector of the 9x9 LtL by inverse iteration on a Cholesky factor), then
//     LMSolver's <= 10 Levenberg-Marquardt iterations with its lambda schedule on HomographyRefineCallback's residuals
//     (FP64, pixel coordinates), then the mask of the REFINED model under computeError -- what cv2 4.x returns.
#include <float.h>
#include <math.h>
#include "common.cuh"

namespace hpusm {

constexpr int RS_THREADS = 256;
#ifndef RS_MIN_BLOCKS
#define RS_MIN_BLOCKS 3  // 80 registers, 73 KB of shared memory per CTA: 126.4 ms on C4 vs 129.7 ms at 2 CTAs per SM
#endif
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_CHUNK = 4096;  // matches staged in shared memory at a time (64 KB)
constexpr int COUNTER_MAX_ATTEMPTS = 256;
constexpr int CV_MAX_ATTEMPTS = 10000;

// ---- counter-based sample table -----------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Four distinct indices in [0, K), K >= 4, as a function of (seed, pair, hyp, attempt) only: two splitmix64 outputs of the
// key give four 32-bit uniforms; each becomes a rank among the indices not yet used.  sample_key() is the part of the key
// that does not depend on the attempt (hoisted out of the redraw loop).
__host__ __device__ __forceinline__ uint64_t sample_key(uint64_t seed, uint64_t pair, uint64_t hyp) {
  return seed ^ (0xD1B54A32D192ED03ull * (pair + 1)) ^ (0x8CB92BA72F3D8DD7ull * (hyp + 1));
}
__host__ __device__ __forceinline__ void sample4_keyed(uint64_t key, uint32_t attempt, uint32_t K, uint32_t out[4]) {
  uint64_t s = key ^ (0xA0761D6478BD642Full * (uint64_t)attempt);
  const uint64_t r0 = splitmix64(s), r1 = splitmix64(s);
  const uint32_t u[4] = {(uint32_t)(r0 >> 32), (uint32_t)r0, (uint32_t)(r1 >> 32), (uint32_t)r1};
  uint32_t sorted[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t v = (uint32_t)(((uint64_t)u[j] * (uint64_t)(K - j)) >> 32);  // rank among the unused indices
    int pos = 0;
#pragma unroll
    for (int c = 0; c < j; ++c) {
      if (v >= sorted[c]) { ++v; pos = c + 1; }
    }
#pragma unroll
    for (int c = j; c > pos; --c) sorted[c] = sorted[c - 1];
    sorted[pos] = v;
    out[j] = v;
  }
}
__host__ __device__ __forceinline__ void sample4(uint64_t seed, uint64_t pair, uint64_t hyp, uint32_t attempt, uint32_t K,
                                                 uint32_t out[4]) {
  sample4_keyed(sample_key(seed, pair, hyp), attempt, K, out);
}

// OpenCV's RNG::next(): multiply-with-carry.
__device__ __forceinline__ uint32_t cv_rng_next(uint64_t& state) {
  state = (uint64_t)(uint32_t)state * 4164903690ull + (uint32_t)(state >> 32);
  return (uint32_t)state;
}

// ---- minimal solver (FP64, explicit rounding so that the C oracle reproduces every bit) ----------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// 2x2 determinant a*d - b*c with two roundings of the products and one of the difference.
__device__ __forceinline__ double det2(double a, double b, double c, double d) {
  return dsub(dmul(a, d), dmul(b, c));
}

// Projective map of the unit square (0,0),(1,0),(1,1),(0,1) onto the quad p0..p3, row-major 3x3.
// Returns false when the quad has no such map (vanishing denominator).
__device__ __forceinline__ bool square_to_quad(const double* x, const double* y, double* M) {
  const double dx1 = dsub(x[1], x[2]), dx2 = dsub(x[3], x[2]);
  const double dy1 = dsub(y[1], y[2]), dy2 = dsub(y[3], y[2]);
  const double px = dadd(dsub(x[0], x[1]), dsub(x[2], x[3]));
  const double py = dadd(dsub(y[0], y[1]), dsub(y[2], y[3]));
  const double den = det2(dx1, dx2, dy1, dy2);
  if (den == 0.0) return false;
  const double g = ddiv(det2(px, dx2, py, dy2), den);
  const double h = ddiv(det2(dx1, px, dy1, py), den);
  M[0] = dadd(dsub(x[1], x[0]), dmul(g, x[1])); M[1] = dadd(dsub(x[3], x[0]), dmul(h, x[3])); M[2] = x[0];
  M[3] = dadd(dsub(y[1], y[0]), dmul(g, y[1])); M[4] = dadd(dsub(y[3], y[0]), dmul(h, y[3])); M[5] = y[0];
  M[6] = g; M[7] = h; M[8] = 1.0;
  return true;
}

// H = D * adj(S), normalised to H[8] == 1.  Returns false if H[8] vanishes or anything is non-finite.
__device__ __forceinline__ bool homography_from_maps(const double* S, const double* D, double* H) {
  double A[9];
  A[0] = det2(S[4], S[5], S[7], S[8]); A[1] = det2(S[2], S[1], S[8], S[7]); A[2] = det2(S[1], S[2], S[4], S[5]);
  A[3] = det2(S[5], S[3], S[8], S[6]); A[4] = det2(S[0], S[2], S[6], S[8]); A[5] = det2(S[2], S[0], S[5], S[3]);
  A[6] = det2(S[3], S[4], S[6], S[7]); A[7] = det2(S[1], S[0], S[7], S[6]); A[8] = det2(S[0], S[1], S[3], S[4]);
  double R[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      R[3 * i + j] = dadd(dadd(dmul(D[3 * i], A[j]), dmul(D[3 * i + 1], A[3 + j])), dmul(D[3 * i + 2], A[6 + j]));
  if (R[8] == 0.0) return false;
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    H[i] = ddiv(R[i], R[8]);
    ok = ok && (fabs(H[i]) <= DBL_MAX);  // rejects inf and NaN
  }
  return ok;
}

// Exact 4-point homography of a non-degenerate sample. Points are translated to the first sample point before the
// closed-form solve (exact in FP64 for pixel-range FP32 input) and the translation is folded back in afterwards.
__device__ bool hypothesis_solve(double* sx, double* sy, double* dx, double* dy, double* H) {
  const double osx = sx[0], osy = sy[0], odx = dx[0], ody = dy[0];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sx[j] = dsub(sx[j], osx); sy[j] = dsub(sy[j], osy); dx[j] = dsub(dx[j], odx); dy[j] = dsub(dy[j], ody);
  }
  double S[9], D[9], Hc[9];
  if (!square_to_quad(sx, sy, S)) return false;
  if (!square_to_quad(dx, dy, D)) return false;
  if (!homography_from_maps(S, D, Hc)) return false;
  // H = T(+od) * Hc * T(-os)
  double G[9];
  G[0] = dadd(Hc[0], dmul(odx, Hc[6])); G[1] = dadd(Hc[1], dmul(odx, Hc[7])); G[2] = dadd(Hc[2], dmul(odx, Hc[8]));
  G[3] = dadd(Hc[3], dmul(ody, Hc[6])); G[4] = dadd(Hc[4], dmul(ody, Hc[7])); G[5] = dadd(Hc[5], dmul(ody, Hc[8]));
  G[6] = Hc[6]; G[7] = Hc[7]; G[8] = Hc[8];
  double R[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    R[3 * i] = G[3 * i]; R[3 * i + 1] = G[3 * i + 1];
    R[3 * i + 2] = dsub(dsub(G[3 * i + 2], dmul(G[3 * i], osx)), dmul(G[3 * i + 1], osy));
  }
  if (R[8] == 0.0) return false;
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    H[i] = ddiv(R[i], R[8]);
    ok = ok && (fabs(H[i]) <= DBL_MAX);
  }
  return ok;
}


// ---- HomographyEstimatorCallback::checkSubset for a 4-point sample (FP64, one rounding per operation) ----------------
__device__ __forceinline__ bool cv_collinear_with_last(const double* x, const double* y) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double dx1 = dsub(x[j], x[3]), dy1 = dsub(y[j], y[3]);
#pragma unroll
    for (int k = 0; k < j; ++k) {
      const double dx2 = dsub(x[k], x[3]), dy2 = dsub(y[k], y[3]);
      const double lhs = fabs(dsub(dmul(dx2, dy1), dmul(dy2, dx1)));
      const double rhs = dmul(1.1920928955078125e-07, dadd(dadd(dadd(fabs(dx1), fabs(dy1)), fabs(dx2)), fabs(dy2)));
      if (lhs <= rhs) return true;
    }
  }
  return false;
}

// determinant of [[x0,y0,1],[x1,y1,1],[x2,y2,1]] in the operation order of cv::Matx33d's determinant
__device__ __forceinline__ double cv_det3(double x0, double y0, double x1, double y1, double x2, double y2) {
  const double m0 = dsub(y1, y2), m1 = dsub(x1, x2);
  const double m2 = dsub(dmul(x1, y2), dmul(x2, y1));
  return dadd(dsub(dmul(x0, m0), dmul(y0, m1)), m2);
}

__device__ __forceinline__ bool cv_check_subset(const double* sx, const double* sy, const double* dx, const double* dy) {
  if (cv_collinear_with_last(sx, sy) || cv_collinear_with_last(dx, dy)) return false;
  const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
  int negative = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = tt[i][0], b = tt[i][1], c = tt[i][2];
    const double dA = cv_det3(sx[a], sy[a], sx[b], sy[b], sx[c], sy[c]);
    const double dB = cv_det3(dx[a], dy[a], dx[b], dy[b], dx[c], dy[c]);
    negative += (dmul(dA, dB) < 0.0) ? 1 : 0;
  }
  return negative == 0 || negative == 4;
}

// The four correspondences id[0..3] as doubles, from global memory ...
__device__ __forceinline__ void sample_points(const float* __restrict__ src, const float* __restrict__ dst, const uint32_t* id,
                                              double* sx, double* sy, double* dx, double* dy) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 s = __ldg(reinterpret_cast<const float2*>(src) + id[j]);
    const float2 d = __ldg(reinterpret_cast<const float2*>(dst) + id[j]);
    sx[j] = (double)s.x; sy[j] = (double)s.y; dx[j] = (double)d.x; dy[j] = (double)d.y;
  }
}

// ... or from the pair-packed shared-memory staging of the scoring kernel:
// pts[2j] = (x0, x1, y0, y1), pts[2j+1] = (-x'0, -x'1, -y'0, -y'1) for matches 2j, 2j+1 (negation is exact).
__device__ __forceinline__ void sample_points_smem(const float4* pts, const uint32_t* id, double* sx, double* sy, double* dx,
                                                   double* dy) {
  const float* f = reinterpret_cast<const float*>(pts);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t base = (id[j] >> 1) * 8 + (id[j] & 1);
    sx[j] = (double)f[base]; sy[j] = (double)f[base + 2];
    dx[j] = (double)(-f[base + 4]); dy[j] = (double)(-f[base + 6]);
  }
}

// ---- inlier tests -----------------------------------------------------------------------------------------------------
// Division-free test in FP32 with explicit FMAs (mirrored by the oracle with fmaf()).
__device__ __forceinline__ bool fast_is_inlier(const float* h, float x, float y, float xp, float yp, float thr2) {
  const float w = __fmaf_rn(h[6], x, __fmaf_rn(h[7], y, h[8]));
  const float u = __fmaf_rn(h[0], x, __fmaf_rn(h[1], y, h[2]));
  const float v = __fmaf_rn(h[3], x, __fmaf_rn(h[4], y, h[5]));
  const float du = __fmaf_rn(-xp, w, u);
  const float dv = __fmaf_rn(-yp, w, v);
  const float e = __fmaf_rn(du, du, __fmul_rn(dv, dv));
  const float lim = __fmul_rn(__fmul_rn(thr2, w), w);
  return e <= lim;
}

// OpenCV's HomographyEstimatorCallback::computeError: float32, a correctly rounded division, no contraction.
__device__ __forceinline__ bool cv_is_inlier(const float* h, float x, float y, float xp, float yp, float thr2) {
  const float ww = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(h[6], x), __fmul_rn(h[7], y)), 1.f));
  const float px = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(h[0], x), __fmul_rn(h[1], y)), h[2]), ww);
  const float py = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(h[3], x), __fmul_rn(h[4], y)), h[5]), ww);
  const float ex = __fsub_rn(px, xp), ey = __fsub_rn(py, yp);
  return __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)) <= thr2;
}

// ---- packed FP32 (sm_100 FFMA2 / FMUL2 / FADD2): two matches per instruction, each half rounded exactly like the scalar op
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// fast_is_inlier() on TWO matches at once.  xy = (x0, x1, y0, y1), nq = (-x'0, -x'1, -y'0, -y'1); h2[i] holds model
// coefficient i in both halves.  Same operations in the same order as the scalar test: counts stay bit-identical to the oracle.
__device__ __forceinline__ void fast_count2(const f32x2* h2, f32x2 thr2, float4 xy, float4 nq, int& cnt) {
  const f32x2 X = pack2(xy.x, xy.y), Y = pack2(xy.z, xy.w), NXP = pack2(nq.x, nq.y), NYP = pack2(nq.z, nq.w);
  const f32x2 w = fma2(h2[6], X, fma2(h2[7], Y, h2[8]));
  const f32x2 u = fma2(h2[0], X, fma2(h2[1], Y, h2[2]));
  const f32x2 v = fma2(h2[3], X, fma2(h2[4], Y, h2[5]));
  const f32x2 du = fma2(NXP, w, u);
  const f32x2 dv = fma2(NYP, w, v);
  const f32x2 e = fma2(du, du, mul2(dv, dv));
  const f32x2 lim = mul2(mul2(thr2, w), w);
  float e0, e1, l0, l1;
  unpack2(e, e0, e1);
  unpack2(lim, l0, l1);
  asm("{\n\t.reg .pred p, q;\n\tsetp.le.f32 p, %1, %2;\n\tsetp.le.f32 q, %3, %4;\n\t@p add.s32 %0, %0, 1;\n\t@q add.s32 %0, %0, 1;\n\t}"
      : "+r"(cnt) : "f"(e0), "f"(l0), "f"(e1), "f"(l1));
}

// cv_is_inlier() on two matches at once (h2[8] holds 1.0f in both halves).
__device__ __forceinline__ void cv_count2(const f32x2* h2, f32x2 thr2, float4 xy, float4 nq, int& cnt) {
  const f32x2 X = pack2(xy.x, xy.y), Y = pack2(xy.z, xy.w), NXP = pack2(nq.x, nq.y), NYP = pack2(nq.z, nq.w);
  const f32x2 den = add2(add2(mul2(h2[6], X), mul2(h2[7], Y)), h2[8]);
  float d0, d1;
  unpack2(den, d0, d1);
  const f32x2 ww = pack2(__fdiv_rn(1.f, d0), __fdiv_rn(1.f, d1));
  const f32x2 px = mul2(add2(add2(mul2(h2[0], X), mul2(h2[1], Y)), h2[2]), ww);
  const f32x2 py = mul2(add2(add2(mul2(h2[3], X), mul2(h2[4], Y)), h2[5]), ww);
  const f32x2 ex = add2(px, NXP), ey = add2(py, NYP);  // px - x' == px + (-x') in IEEE arithmetic
  const f32x2 e = add2(mul2(ex, ex), mul2(ey, ey));
  float e0, e1, l0, l1;
  unpack2(e, e0, e1);
  unpack2(thr2, l0, l1);
  asm("{\n\t.reg .pred p, q;\n\tsetp.le.f32 p, %1, %2;\n\tsetp.le.f32 q, %3, %4;\n\t@p add.s32 %0, %0, 1;\n\t@q add.s32 %0, %0, 1;\n\t}"
      : "+r"(cnt) : "f"(e0), "f"(l0), "f"(e1), "f"(l1));
}

// cv2's adaptive bound on the number of iterations (RANSACUpdateNumIters restated, modelPoints = 4).
__device__ int update_num_iters(double conf, double outlier_ratio, int max_iters) {
  double p = fmin(fmax(conf, 0.0), 1.0);
  double ep = fmin(fmax(outlier_ratio, 0.0), 1.0);
  double num = fmax(dsub(1.0, p), DBL_MIN);
  const double g = dsub(1.0, ep);
  double denom = dsub(1.0, dmul(dmul(g, g), dmul(g, g)));  // 1 - (1-ep)^4 without contraction
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  if (denom >= 0.0 || -num >= (double)max_iters * (-denom)) return max_iters;
  return (int)llrint(num / denom);
}

// ---- cv2 sampler pre-pass: one THREAD per pair walks OpenCV's sequence; samples [npairs][nhyp] int4, x = -1: none -------
__global__ void __launch_bounds__(64)
ransac_sample_cv2_kernel(const float* __restrict__ src, const float* __restrict__ dst, const int64_t* __restrict__ pair_offsets,
                         int64_t npairs, int nhyp, int4* __restrict__ samples) {
  const int64_t pair = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= npairs) return;
  const int64_t m0 = pair_offsets[pair];
  const uint32_t K = (uint32_t)(pair_offsets[pair + 1] - m0);
  const float* s = src + m0 * 2;
  const float* d = dst + m0 * 2;
  int4* out = samples + pair * nhyp;
  uint64_t state = 0xFFFFFFFFFFFFFFFFull;
  int h = 0;
  if (K >= 4) {
    for (; h < nhyp; ++h) {
      bool found = false;
      uint32_t id[4];
      for (int it = 0; it < CV_MAX_ATTEMPTS && !found; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          for (;;) {
            const uint32_t v = cv_rng_next(state) % K;
            bool dup = false;
#pragma unroll
            for (int c = 0; c < i; ++c) dup |= (id[c] == v);
            if (!dup) { id[i] = v; break; }
          }
        }
        double sx[4], sy[4], dx[4], dy[4];
        sample_points(s, d, id, sx, sy, dx, dy);
        found = cv_check_subset(sx, sy, dx, dy);
      }
      if (!found) break;  // OpenCV ends the loop when an iteration finds no subset
      out[h] = make_int4((int)id[0], (int)id[1], (int)id[2], (int)id[3]);
    }
  }
  for (; h < nhyp; ++h) out[h] = make_int4(-1, -1, -1, -1);
}

// The same sequence with one WARP per pair, for small batches (the drop-in's two calls per image pair), where the
// thread-per-pair walk is a chain of ~10^4 dependent attempts.  OpenCV's generator itself is cheap; what an attempt costs
// is fetching its four matches and the subset check.  So every lane walks the generator redundantly through the next 128
// draws, keeps the four that belong to ITS attempt (lane l = attempt l of the batch, assuming no duplicate redraws
// before it), and checks that attempt; the lanes below the first one that saw a duplicate are exactly OpenCV's next
// attempts, and the ones among them whose subset passes are its next iterations, in order.  The attempt with the
// duplicate is then replayed the sequential way (uniformly by all lanes) and the next batch starts behind it.
__global__ void __launch_bounds__(128)
ransac_sample_cv2_warp_kernel(const float* __restrict__ src, const float* __restrict__ dst,
                              const int64_t* __restrict__ pair_offsets, int64_t npairs, int nhyp, int4* __restrict__ samples) {
  const int lane = threadIdx.x & 31;
  const int64_t pair = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= npairs) return;
  const int64_t m0 = pair_offsets[pair];
  const uint32_t K = (uint32_t)(pair_offsets[pair + 1] - m0);
  const float* s = src + m0 * 2;
  const float* d = dst + m0 * 2;
  int4* out = samples + pair * nhyp;
  uint64_t state = 0xFFFFFFFFFFFFFFFFull;  // uniform over the warp
  int h = 0, fails = 0;                     // iterations written; attempts since the last accepted subset
  bool dead = K < 4;
  while (h < nhyp && !dead) {
    // 128 draws; lane l keeps draws 4l .. 4l+3 and the generator state in front of them
    uint64_t st = state, my_state = state;
    uint32_t id[4] = {0, 0, 0, 0};
    for (int a = 0; a < 32; ++a) {
      if (a == lane) my_state = st;
      const uint32_t v0 = cv_rng_next(st), v1 = cv_rng_next(st), v2 = cv_rng_next(st), v3 = cv_rng_next(st);
      if (a == lane) { id[0] = v0; id[1] = v1; id[2] = v2; id[3] = v3; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) id[i] %= K;
    const bool dup = id[0] == id[1] || id[0] == id[2] || id[0] == id[3] || id[1] == id[2] || id[1] == id[3] || id[2] == id[3];
    bool ok = false;
    if (!dup) {
      double sx[4], sy[4], dx[4], dy[4];
      sample_points(s, d, id, sx, sy, dx, dy);
      ok = cv_check_subset(sx, sy, dx, dy);
    }
    const unsigned dupmask = __ballot_sync(0xffffffffu, dup);
    const int first_dup = dupmask ? __ffs(dupmask) - 1 : 32;
    const unsigned valid = first_dup == 32 ? 0xffffffffu : ((1u << first_dup) - 1u);
    const unsigned acc = __ballot_sync(0xffffffffu, ok) & valid;
    // OpenCV gives up on an iteration after CV_MAX_ATTEMPTS failed attempts: find where, if at all, inside this batch
    // (attempt a of the batch is the (fails + a - last accepted before a)-th of its iteration)
    int limit = 32;  // first attempt of the batch that would be attempt number CV_MAX_ATTEMPTS + 1 of its iteration
    {
      const unsigned below = acc & ((1u << lane) - 1u);
      const int since = below ? lane - (32 - __clz(below)) : fails + lane;  // failed attempts of this iteration before mine
      const unsigned over = __ballot_sync(0xffffffffu, since >= CV_MAX_ATTEMPTS) & valid;
      if (over) limit = __ffs(over) - 1;
    }
    const unsigned live = limit == 32 ? 0xffffffffu : ((1u << limit) - 1u);
    const unsigned acc_live = acc & live;
    if ((acc_live >> lane) & 1u) {
      const int hh = h + __popc(acc_live & ((1u << lane) - 1u));
      if (hh < nhyp) out[hh] = make_int4((int)id[0], (int)id[1], (int)id[2], (int)id[3]);
    }
    h += __popc(acc_live);
    if (limit < 32 && limit <= first_dup) { dead = true; break; }
    const int consumed = first_dup;  // whole attempts of the batch that were OpenCV's attempts
    {
      const unsigned upto = acc_live & valid;
      fails = upto ? consumed - (32 - __clz(upto)) : fails + consumed;
    }
    if (first_dup == 32) { state = st; continue; }
    // the attempt with the duplicate, replayed the sequential way from the state in front of it
    state = __shfl_sync(0xffffffffu, my_state, first_dup);
    if (h >= nhyp) break;
    if (fails >= CV_MAX_ATTEMPTS) { dead = true; break; }
    uint32_t q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      for (;;) {
        const uint32_t v = cv_rng_next(state) % K;
        bool dd = false;
#pragma unroll
        for (int c = 0; c < i; ++c) dd |= (q[c] == v);
        if (!dd) { q[i] = v; break; }
      }
    }
    double sx[4], sy[4], dx[4], dy[4];
    sample_points(s, d, q, sx, sy, dx, dy);
    if (cv_check_subset(sx, sy, dx, dy)) {
      if (lane == 0) out[h] = make_int4((int)q[0], (int)q[1], (int)q[2], (int)q[3]);
      ++h;
      fails = 0;
    } else {
      ++fails;
    }
  }
  for (int i = h + lane; i < nhyp; i += 32) out[i] = make_int4(-1, -1, -1, -1);
}

// grid.x = pairs. dynamic smem: float4 pts[chunk]; float models[nhyp][8] (compacted; h8 == 1 exactly after the
// normalisation of the solver); int counts[nhyp]; int hyp_of[nhyp].
//   A   one LANE per hypothesis: sample (redraw until the subset check passes | read the cv2 table) + FP64 closed-form
//       solve; the models that exist are compacted (ballot + one smem atomic per warp), FP32 model parked in shared memory
//   B   lane = hypothesis (model in 9 registers), the pair's matches stream through as broadcast LDS.128; work items
//       are (32 hypotheses) x (RS_SEG matches) so all warps stay busy; the count lives in a register
//   C   sequential replay of RANSAC's early exit over the counts; the winner's FP64 model goes to best_H
constexpr int RS_SEG = 256;

template <bool CV>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_BLOCKS)
ransac_score_kernel(const float* __restrict__ src, const float* __restrict__ dst,
                    const int64_t* __restrict__ pair_offsets, float thresh, int nhyp, double confidence,
                    uint64_t seed, int64_t pair0, int chunk, const int4* __restrict__ samples,
                    int32_t* __restrict__ counts_out, int32_t* __restrict__ best_hyp, double* __restrict__ best_H) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  float4* pts = reinterpret_cast<float4*>(rs_smem);
  float* models = reinterpret_cast<float*>(rs_smem + (size_t)chunk * sizeof(float4));
  int* counts = reinterpret_cast<int*>(models + (size_t)nhyp * 8);
  int* hyp_of = counts + nhyp;
  __shared__ int n_valid;
  const int64_t pair = blockIdx.x;
  const int64_t m0 = pair_offsets[pair];
  const int64_t K64 = pair_offsets[pair + 1] - m0;
  const uint32_t K = (uint32_t)K64;
  const float* s = src + m0 * 2;
  const float* d = dst + m0 * 2;
  const int4* smp = CV ? samples + pair * nhyp : nullptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float thr2 = __fmul_rn(thresh, thresh);
  if (tid == 0) n_valid = 0;

  // Stage `cn` matches starting at c0, two per entry pair; an odd tail is padded with a match no model can explain.
  auto stage = [&](int64_t c0, int cn) {
    const int npair = (cn + 1) >> 1;
    for (int j = tid; j < npair; j += RS_THREADS) {
      const int i0 = 2 * j, i1 = 2 * j + 1;
      const float2 a0 = reinterpret_cast<const float2*>(s)[c0 + i0];
      const float2 b0 = reinterpret_cast<const float2*>(d)[c0 + i0];
      float2 a1 = make_float2(0.f, 0.f), b1 = make_float2(1e30f, 1e30f);
      if (i1 < cn) {
        a1 = reinterpret_cast<const float2*>(s)[c0 + i1];
        b1 = reinterpret_cast<const float2*>(d)[c0 + i1];
      }
      pts[2 * j] = make_float4(a0.x, a1.x, a0.y, a1.y);
      pts[2 * j + 1] = make_float4(-b0.x, -b1.x, -b0.y, -b1.y);
    }
  };
  // When the whole pair fits (the common case) it is staged ONCE, before hypothesis generation, so the scattered
  // sample reads of every draw hit shared memory instead of L2.
  const bool resident = K64 <= chunk;
  if (resident) stage(0, (int)K64);
  __syncthreads();

  // The points of hypothesis h (after redraws / from the cv2 table); false when the iteration has no sample.
  // `a`: counter sampler only -- the attempt whose draw is wanted (phase A1 leaves the first passing one in counts[h])
  auto hypothesis_points = [&](int h, int a, double* sx, double* sy, double* dx, double* dy) -> bool {
    if (K < 4) return false;
    if (CV) {
      const int4 q = __ldg(smp + h);
      if (q.x < 0) return false;
      const uint32_t id[4] = {(uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w};
      if (resident) sample_points_smem(pts, id, sx, sy, dx, dy);
      else sample_points(s, d, id, sx, sy, dx, dy);
      return true;
    }
    if (a < 0) return false;
    uint32_t id[4];
    sample4(seed, (uint64_t)(pair + pair0), (uint64_t)h, (uint32_t)a, K, id);
    if (resident) sample_points_smem(pts, id, sx, sy, dx, dy);
    else sample_points(s, d, id, sx, sy, dx, dy);
    return true;
  };

  // phase A1 (counter sampler): find, for every hypothesis, the FIRST attempt whose sample passes OpenCV's subset check.
  // A lane per hypothesis looping until its own draw passes leaves the warp running for the unluckiest of its 32 lanes
  // (one draw in five passes on C4's data: 17 iterations per warp where the average lane needs 5; the sampling was 39 %
  // of the kernel).  Here a warp works on 8 hypotheses at a time, 4 consecutive attempts of each in parallel (lane =
  // (slot, attempt)); the lowest passing attempt wins, which is what the sequential loop returns; a resolved slot pulls
  // the warp's next hypothesis, so all lanes stay busy until the hypotheses run out.
  if (!CV) {
    const int slot = lane >> 2, sub = lane & 3;
    // every warp owns a contiguous run of hypotheses and hands them to its slots from a register counter (a block-wide
    // shared-memory counter put an atomic's latency into every round of the dependent chain)
    const int per_warp = (nhyp + RS_WARPS - 1) / RS_WARPS;
    int next = warp * per_warp;
    const int h_end = min(nhyp, next + per_warp);
    int h = -1, r = 0;
    uint64_t key = 0;
    for (;;) {
      const unsigned needm = __ballot_sync(0xffffffffu, h < 0 && sub == 0);
      if (needm) {
        if (h < 0) {
          h = next + __popc(needm & ((1u << (slot * 4)) - 1u));
          r = 0;
          key = sample_key(seed, (uint64_t)(pair + pair0), (uint64_t)h);
        }
        next += __popc(needm);
      }
      const bool active = h < h_end;
      if (!__any_sync(0xffffffffu, active)) break;
      bool pass = false;
      const int a = r * 4 + sub;
      if (active && K >= 4 && a < COUNTER_MAX_ATTEMPTS) {
        uint32_t id[4];
        double sx[4], sy[4], dx[4], dy[4];
        sample4_keyed(key, (uint32_t)a, K, id);
        if (resident) sample_points_smem(pts, id, sx, sy, dx, dy);
        else sample_points(s, d, id, sx, sy, dx, dy);
        pass = cv_check_subset(sx, sy, dx, dy);
      }
      const unsigned mine = (__ballot_sync(0xffffffffu, pass) >> (slot * 4)) & 0xFu;
      if (active) {
        if (mine) {
          if (sub == 0) counts[h] = r * 4 + __ffs(mine) - 1;
          h = -1;
        } else if (K < 4 || (r + 1) * 4 >= COUNTER_MAX_ATTEMPTS) {
          if (sub == 0) counts[h] = -1;
          h = -1;
        } else {
          ++r;
        }
      }
    }
    __syncthreads();
  }

  // phase A2: one LANE per hypothesis: the sample's points + the FP64 closed-form solve, compaction of the models that exist
  const int rounds = (nhyp + RS_THREADS - 1) / RS_THREADS;
  for (int r = 0; r < rounds; ++r) {
    const int h = r * RS_THREADS + tid;
    bool ok = false;
    double H[9];
    if (h < nhyp) {
      double sx[4], sy[4], dx[4], dy[4];
      ok = hypothesis_points(h, CV ? 0 : counts[h], sx, sy, dx, dy) && hypothesis_solve(sx, sy, dx, dy, H);
    }
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    int base = 0;
    if (lane == 0 && b) base = atomicAdd(&n_valid, __popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (ok) {
      const int j = base + __popc(b & ((1u << lane) - 1u));
      hyp_of[j] = h;
#pragma unroll
      for (int i = 0; i < 8; ++i) models[j * 8 + i] = (float)H[i];
    }
    if (h < nhyp) counts[h] = ok ? 0 : -1;
  }
  __syncthreads();
  const int nv = n_valid;

  // phase B
  const int groups = (nv + 31) >> 5;
  const f32x2 thr2x2 = pack2(thr2, thr2);
  for (int64_t c0 = 0; c0 < K64; c0 += chunk) {
    const int cn = (int)min((int64_t)chunk, K64 - c0);
    if (!resident) {
      __syncthreads();
      stage(c0, cn);
      __syncthreads();
    }
    const int npair = (cn + 1) >> 1;
    const int segs = (npair + RS_SEG / 2 - 1) / (RS_SEG / 2);
    for (int item = warp; item < groups * segs; item += RS_WARPS) {
      const int g = item / segs, sg = item - g * segs;
      const int j = g * 32 + lane;
      const int h = j < nv ? hyp_of[j] : -1;
      f32x2 h2[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const float c = h >= 0 ? (i < 8 ? models[j * 8 + i] : 1.f) : 0.f;
        h2[i] = pack2(c, c);
      }
      const int p0 = sg * (RS_SEG / 2), p1 = min(npair, p0 + RS_SEG / 2);
      int cnt = 0;
#pragma unroll 4
      for (int i = p0; i < p1; ++i) {
        if (CV) cv_count2(h2, thr2x2, pts[2 * i], pts[2 * i + 1], cnt);
        else fast_count2(h2, thr2x2, pts[2 * i], pts[2 * i + 1], cnt);
      }
      if (h >= 0 && cnt) atomicAdd(&counts[h], cnt);
    }
  }
  __syncthreads();

  if (counts_out) {
    for (int h = tid; h < nhyp; h += RS_THREADS) counts_out[pair * nhyp + h] = counts[h];
  }

  // phase C: replay the sequential loop (exact early-exit semantics), lowest hypothesis wins ties.  The sequential loop
  // only ever changes state at a RECORD (a count above every earlier count and above 3), so all threads first mark the
  // records (slice maxima -> exclusive prefix maximum -> one pass over the slice) in a bit mask, and one thread then walks
  // the dozen or so set bits instead of all nhyp counts.
  if (best_hyp) {
    __shared__ int slice_max[RS_THREADS];
    unsigned* rec = reinterpret_cast<unsigned*>(hyp_of);  // the compaction list is no longer needed: nhyp ints >= nhyp / 32 words
    const int per = (nhyp + RS_THREADS - 1) / RS_THREADS;
    const int h0 = tid * per, h1 = min(nhyp, h0 + per);
    int mx = 3;
    for (int h = h0; h < h1; ++h) mx = max(mx, counts[h]);
    __syncthreads();  // every thread is done reading hyp_of (phase B) before it is reused
    slice_max[tid] = mx;
    for (int wd = tid; wd < (nhyp + 31) / 32; wd += RS_THREADS) rec[wd] = 0u;
    __syncthreads();
    int run = 3;
    for (int t = 0; t < tid; ++t) run = max(run, slice_max[t]);  // exclusive prefix maximum (256 broadcast reads)
    for (int h = h0; h < h1; ++h) {
      const int c = counts[h];
      if (c > run) { run = c; atomicOr(&rec[h >> 5], 1u << (h & 31)); }
    }
    __syncthreads();
    if (tid == 0) {
      int best = -1, niters = nhyp;
      for (int wd = 0; wd < (nhyp + 31) / 32 && wd * 32 < niters; ++wd) {
        unsigned bits = rec[wd];
        while (bits) {
          const int h = wd * 32 + __ffs(bits) - 1;
          bits &= bits - 1;
          if (h >= niters) { bits = 0; break; }
          best = h;
          if (confidence > 0.0) niters = update_num_iters(confidence, (double)(K64 - counts[h]) / (double)K64, niters);
        }
      }
      best_hyp[pair] = best;
      if (best >= 0) {
        double sx[4], sy[4], dx[4], dy[4], H[9];
        // counts[] now holds inlier counts: the winner's draw is found again by walking its attempts (it has one: it scored)
        for (int a = 0; a < (CV ? 1 : COUNTER_MAX_ATTEMPTS); ++a)
          if (hypothesis_points(best, a, sx, sy, dx, dy) && (CV || cv_check_subset(sx, sy, dx, dy))) break;
        (void)hypothesis_solve(sx, sy, dx, dy, H);
#pragma unroll
        for (int i = 0; i < 9; ++i) best_H[pair * 9 + i] = H[i];
      }
    }
  }
}

// ---- refinement: findHomography's post-processing, one warp per pair ---------------------------------------------------
constexpr int RF_WARPS = 4;

template <int N>
__device__ __forceinline__ void warp_sum(double (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
}

// Cholesky factor (lower, in place, row-major NxN) of a symmetric positive definite matrix.  A pivot that is not safely
// positive marks a direction the data do not determine: its reciprocal is set to 0, which removes the direction from the
// solve (what the eigenvalue threshold of cv::solve(DECOMP_EIG) does to it).  rdiag receives 1 / L[j][j].
template <int N>
__device__ __forceinline__ void cholesky(double (&A)[N * N], double (&rdiag)[N], double tiny) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double sum = A[j * N + j];
#pragma unroll
    for (int k = 0; k < j; ++k) sum -= A[j * N + k] * A[j * N + k];
    const bool ok = sum > tiny;
    const double ljj = ok ? sqrt(sum) : 1.0;
    const double r = ok ? 1.0 / ljj : 0.0;
    A[j * N + j] = ljj;
    rdiag[j] = r;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double s2 = A[i * N + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s2 -= A[i * N + k] * A[j * N + k];
      A[i * N + j] = s2 * r;
    }
  }
}

// x = (L L^T)^-1 b
template <int N>
__device__ __forceinline__ void cholesky_solve(const double (&L)[N * N], const double (&rdiag)[N], const double (&b)[N],
                                               double (&x)[N]) {
  double y[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s2 = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s2 -= L[i * N + k] * y[k];
    y[i] = s2 * rdiag[i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double s2 = y[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s2 -= L[k * N + i] * x[k];
    x[i] = s2 * rdiag[i];
  }
}

// d = (A + lambda * diag(D))^-1 v for the symmetric 8x8 A, through the Cholesky factor of the diagonally scaled matrix.
__device__ __noinline__ void lm_solve(const double* A /*64*/, const double* D, const double* v, double lambda, double* d) {
  double C[64], sc[8], b[8], z[8], rd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double aii = A[9 * i] + lambda * D[i];
    sc[i] = aii > 0.0 ? rsqrt(aii) : 1.0;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      const double a = A[8 * i + j] + (i == j ? lambda * D[i] : 0.0);
      C[8 * i + j] = a * sc[i] * sc[j];
    }
    b[i] = v[i] * sc[i];
  }
  cholesky<8>(C, rd, 1e-14);
  cholesky_solve<8>(C, rd, b, z);
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = z[i] * sc[i];
}

// largest |diagonal entry| of A^-1 (LMSolver re-seeds lambda with its reciprocal)
__device__ __noinline__ double lm_inverse_max_diag(const double* A /*64*/) {
  double C[64], sc[8], rd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sc[i] = A[9 * i] > 0.0 ? rsqrt(A[9 * i]) : 1.0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) C[8 * i + j] = A[8 * i + j] * sc[i] * sc[j];
  cholesky<8>(C, rd, 1e-14);
  double maxval = DBL_EPSILON;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // column j of L^-1 by forward substitution; (C^-1)[j][j] = sum_i (L^-1)[i][j]^2
    double col[8];
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < j) { col[i] = 0.0; continue; }
      double s2 = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) s2 -= C[8 * i + k] * col[k];
      col[i] = s2 * rd[i];
      acc += col[i] * col[i];
    }
    maxval = fmax(maxval, fabs(acc * sc[j] * sc[j]));
  }
  return maxval;
}

// Eigenvector of the smallest eigenvalue of the symmetric positive semi-definite 9x9 matrix M (what cv::eigen hands
// runKernel as V[8]): inverse iteration on the Cholesky factor of M + mu I.
__device__ __noinline__ void smallest_eigenvector9(const double* M /*81, lower triangle used*/, double* vec) {
  double C[81], rd[9], x[9], y[9];
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) tr += M[10 * i];
  const double mu = 1e-12 * tr;
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) C[9 * i + j] = M[9 * i + j] + (i == j ? mu : 0.0);
  cholesky<9>(C, rd, 0.0);
#pragma unroll
  for (int i = 0; i < 9; ++i) x[i] = 1.0 - 0.07 * i;
  for (int it = 0; it < 64; ++it) {
    cholesky_solve<9>(C, rd, x, y);
    double nn = 0.0, big = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) { nn += y[i] * y[i]; if (fabs(y[i]) > fabs(big)) big = y[i]; }
    const double inv = (big < 0.0 ? -1.0 : 1.0) / sqrt(nn);
    double change = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double t = y[i] * inv;
      change = fmax(change, fabs(t - x[i]));
      x[i] = t;
    }
    if (!(change > 4e-16)) break;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) vec[i] = x[i];
}

// grid.x = ceil(npairs / RF_WARPS); warp = pair.  Every lane holds the same reduced sums and walks the (uniform) solver code.
template <bool CV>
__global__ void __launch_bounds__(RF_WARPS * 32)
ransac_finalize_kernel(const float* __restrict__ src, const float* __restrict__ dst, const int64_t* __restrict__ pair_offsets,
                       int64_t npairs, float thresh, int refine, const int32_t* __restrict__ best_hyp,
                       const double* __restrict__ best_H, int32_t* __restrict__ sel_ws, double* __restrict__ Hout,
                       uint8_t* __restrict__ mask, int32_t* __restrict__ ninl) {
  const int lane = threadIdx.x & 31;
  const int64_t pair = (int64_t)blockIdx.x * RF_WARPS + (threadIdx.x >> 5);
  if (pair >= npairs) return;
  const int64_t m0 = pair_offsets[pair];
  const int K = (int)(pair_offsets[pair + 1] - m0);
  const float2* s = reinterpret_cast<const float2*>(src) + m0;
  const float2* d = reinterpret_cast<const float2*>(dst) + m0;
  uint8_t* mk = mask + m0;
  int32_t* sel = sel_ws + m0;
  const float thr2 = __fmul_rn(thresh, thresh);
  const int bh = best_hyp[pair];
  if (bh < 0) {
    for (int i = lane; i < K; i += 32) mk[i] = 0;
    if (lane < 9) Hout[pair * 9 + lane] = nan("");
    if (lane == 0) ninl[pair] = 0;
    return;
  }
  double Hm[9];
  float hf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { Hm[i] = best_H[pair * 9 + i]; hf[i] = (float)Hm[i]; }

  // inliers of the winning minimal model, compacted in increasing order
  int n = 0;
  for (int base = 0; base < K; base += 32) {
    const int i = base + lane;
    bool in = false;
    if (i < K) {
      const float2 a = s[i], b = d[i];
      in = CV ? cv_is_inlier(hf, a.x, a.y, b.x, b.y, thr2) : fast_is_inlier(hf, a.x, a.y, b.x, b.y, thr2);
      mk[i] = in ? 1 : 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if (in) sel[n + __popc(bal & ((1u << lane) - 1u))] = i;
    n += __popc(bal);
  }
  __syncwarp();
  bool refined = false;
  double H[9];
  if (refine && n >= 4) {
    // ---- runKernel: normalised DLT (M = src, m = dst) ----
    double c4[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = lane; i < n; i += 32) {
      const int r = sel[i];
      const float2 a = s[r], b = d[r];
      c4[0] += b.x; c4[1] += b.y; c4[2] += a.x; c4[3] += a.y;
    }
    warp_sum<4>(c4);
    const double cmx = c4[0] / n, cmy = c4[1] / n, cMx = c4[2] / n, cMy = c4[3] / n;
    double s4[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = lane; i < n; i += 32) {
      const int r = sel[i];
      const float2 a = s[r], b = d[r];
      s4[0] += fabs(b.x - cmx); s4[1] += fabs(b.y - cmy); s4[2] += fabs(a.x - cMx); s4[3] += fabs(a.y - cMy);
    }
    warp_sum<4>(s4);
    if (!(fabs(s4[0]) < DBL_EPSILON || fabs(s4[1]) < DBL_EPSILON || fabs(s4[2]) < DBL_EPSILON || fabs(s4[3]) < DBL_EPSILON)) {
      const double smx = n / s4[0], smy = n / s4[1], sMx = n / s4[2], sMy = n / s4[3];
      // LtL = [[PP, 0, -xPP], [0, PP, -yPP], [-xPP, -yPP, rPP]] with p = (X, Y, 1), r = x^2 + y^2
      double q[24];
#pragma unroll
      for (int i = 0; i < 24; ++i) q[i] = 0.0;
      for (int i = lane; i < n; i += 32) {
        const int r = sel[i];
        const float2 a = s[r], b = d[r];
        const double x = (b.x - cmx) * smx, y = (b.y - cmy) * smy;
        const double X = (a.x - cMx) * sMx, Y = (a.y - cMy) * sMy;
        const double pp[6] = {X * X, X * Y, X, Y * Y, Y, 1.0};
        const double rr = x * x + y * y;
#pragma unroll
        for (int k = 0; k < 6; ++k) { q[k] += pp[k]; q[6 + k] += x * pp[k]; q[12 + k] += y * pp[k]; q[18 + k] += rr * pp[k]; }
      }
      warp_sum<24>(q);
      double L[81];
      // symmetric 3x3 block from its 6 unique sums (XX XY X YY Y 1)
      const int blk[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
#pragma unroll
      for (int i = 0; i < 81; ++i) L[i] = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int u = blk[3 * i + j];
          L[9 * i + j] = q[u]; L[9 * (3 + i) + (3 + j)] = q[u];
          L[9 * (6 + i) + j] = -q[6 + u]; L[9 * i + (6 + j)] = -q[6 + u];
          L[9 * (6 + i) + (3 + j)] = -q[12 + u]; L[9 * (3 + i) + (6 + j)] = -q[12 + u];
          L[9 * (6 + i) + (6 + j)] = q[18 + u];
        }
      double H0[9];
      smallest_eigenvector9(L, H0);
      // H = invHnorm * H0 * Hnorm2, invHnorm = [[1/smx,0,cmx],[0,1/smy,cmy],[0,0,1]], Hnorm2 = [[sMx,0,-cMx sMx],[0,sMy,-cMy sMy],[0,0,1]]
      double T[9], R[9];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T[j] = H0[j] / smx + cmx * H0[6 + j];
        T[3 + j] = H0[3 + j] / smy + cmy * H0[6 + j];
        T[6 + j] = H0[6 + j];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        R[3 * i] = T[3 * i] * sMx;
        R[3 * i + 1] = T[3 * i + 1] * sMy;
        R[3 * i + 2] = T[3 * i + 2] - T[3 * i] * (cMx * sMx) - T[3 * i + 1] * (cMy * sMy);
      }
      double x8[8];
      const double sc0 = 1.0 / R[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x8[i] = R[i] * sc0;

      // ---- LMSolver(10 iterations, eps = FLT_EPSILON) on HomographyRefineCallback ----
      // One pass over the inliers at parameters h: S = sum of squared residuals, max |residual|, and the unique sums of
      // J^T J and J^T r:  acc[0..5] pp^T, [6..11] -xi p_a p_b, [12..17] -yi p_a p_b, [18..20] (xi^2+yi^2) p_a p_b (a,b<2),
      // [21..28] J^T r, [29] S.
      double A[64], v[8], D[8], S = 0.0, rinf = 0.0;
      auto lm_pass = [&](const double* h, double* Ao, double* vo, double& So, double& rinfo) {
        double acc[30];
#pragma unroll
        for (int i = 0; i < 30; ++i) acc[i] = 0.0;
        double rmax = 0.0;
        for (int i = lane; i < n; i += 32) {
          const int r = sel[i];
          const float2 a = s[r], b = d[r];
          const double Mx = a.x, My = a.y;
          double ww = h[6] * Mx + h[7] * My + 1.0;
          ww = fabs(ww) > DBL_EPSILON ? 1.0 / ww : 0.0;
          const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
          const double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
          const double e0 = xi - b.x, e1 = yi - b.y;
          const double p0 = Mx * ww, p1 = My * ww, p2 = ww;
          const double pp[6] = {p0 * p0, p0 * p1, p0 * p2, p1 * p1, p1 * p2, p2 * p2};
          const double pq[6] = {p0 * p0, p0 * p1, p1 * p0, p1 * p1, p2 * p0, p2 * p1};  // p_a p_b, a = 0..2, b = 0..1
          const double rr = xi * xi + yi * yi;
#pragma unroll
          for (int k = 0; k < 6; ++k) { acc[k] += pp[k]; acc[6 + k] -= xi * pq[k]; acc[12 + k] -= yi * pq[k]; }
          acc[18] += rr * pq[0]; acc[19] += rr * pq[1]; acc[20] += rr * pq[3];
          acc[21] += p0 * e0; acc[22] += p1 * e0; acc[23] += p2 * e0;
          acc[24] += p0 * e1; acc[25] += p1 * e1; acc[26] += p2 * e1;
          const double g = xi * e0 + yi * e1;
          acc[27] -= p0 * g; acc[28] -= p1 * g;
          acc[29] += e0 * e0 + e1 * e1;
          rmax = fmax(rmax, fmax(fabs(e0), fabs(e1)));
        }
        warp_sum<30>(acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
#pragma unroll
        for (int i = 0; i < 64; ++i) Ao[i] = 0.0;
        const int blk2[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) { Ao[8 * i + j] = acc[blk2[3 * i + j]]; Ao[8 * (3 + i) + (3 + j)] = acc[blk2[3 * i + j]]; }
        // pq index: (a, b) -> 0:(0,0) 1:(0,1) 2:(1,0) 3:(1,1) 4:(2,0) 5:(2,1)
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            Ao[8 * a + (6 + b)] = acc[6 + 2 * a + b]; Ao[8 * (6 + b) + a] = acc[6 + 2 * a + b];
            Ao[8 * (3 + a) + (6 + b)] = acc[12 + 2 * a + b]; Ao[8 * (6 + b) + (3 + a)] = acc[12 + 2 * a + b];
          }
        Ao[8 * 6 + 6] = acc[18]; Ao[8 * 6 + 7] = acc[19]; Ao[8 * 7 + 6] = acc[19]; Ao[8 * 7 + 7] = acc[20];
#pragma unroll
        for (int i = 0; i < 8; ++i) vo[i] = acc[21 + i];
        So = acc[29];
        rinfo = rmax;
      };
      lm_pass(x8, A, v, S, rinf);
#pragma unroll
      for (int i = 0; i < 8; ++i) D[i] = A[9 * i];
      double lambda = 1.0, lc = 0.75;
      for (int iter = 0;;) {
        double dd[8], xd[8], Ad[64], vd[8], Sd, rinfd;
        lm_solve(A, D, v, lambda, dd);
#pragma unroll
        for (int i = 0; i < 8; ++i) xd[i] = x8[i] - dd[i];
        lm_pass(xd, Ad, vd, Sd, rinfd);
        double dS = 0.0, dinf = 0.0, tdv = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          double t = 0.0;
#pragma unroll
          for (int j = 0; j < 8; ++j) t += A[8 * i + j] * dd[j];
          dS += dd[i] * (-t + 2.0 * v[i]);
          tdv += dd[i] * v[i];
          dinf = fmax(dinf, fabs(dd[i]));
        }
        const double Rr = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1.0);
        if (Rr > 0.75) {
          lambda *= 0.5;
          if (lambda < lc) lambda = 0.0;
        } else if (Rr < 0.25) {
          double nu = (Sd - S) / (fabs(tdv) > DBL_EPSILON ? tdv : 1.0) + 2.0;
          nu = fmin(fmax(nu, 2.0), 10.0);
          if (lambda == 0.0) {
            lambda = lc = 1.0 / lm_inverse_max_diag(A);
            nu *= 0.5;
          }
          lambda *= nu;
        }
        if (Sd < S) {
          S = Sd; rinf = rinfd;
#pragma unroll
          for (int i = 0; i < 8; ++i) { x8[i] = xd[i]; v[i] = vd[i]; }
#pragma unroll
          for (int i = 0; i < 64; ++i) A[i] = Ad[i];
        }
        ++iter;
        if (!(iter < 10 && dinf >= (double)FLT_EPSILON && rinf >= (double)FLT_EPSILON)) break;
      }
      bool fin = true;
#pragma unroll
      for (int i = 0; i < 8; ++i) { H[i] = x8[i]; fin = fin && (fabs(x8[i]) <= DBL_MAX); }
      H[8] = 1.0;
      refined = fin;
    }
  }
  if (!refined) {
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = Hm[i];
  }
  if (lane < 9) {
    double hv = H[0];
#pragma unroll
    for (int i = 1; i < 9; ++i) hv = lane == i ? H[i] : hv;
    Hout[pair * 9 + lane] = hv;
  }
  if (!refined) {
    if (lane == 0) ninl[pair] = n;
    return;
  }
  // OpenCV returns the inliers of the REFINED homography under computeError
#pragma unroll
  for (int i = 0; i < 9; ++i) hf[i] = (float)H[i];
  int cnt = 0;
  for (int i = lane; i < K; i += 32) {
    const float2 a = s[i], b = d[i];
    const bool in = cv_is_inlier(hf, a.x, a.y, b.x, b.y, thr2);
    mk[i] = in ? 1 : 0;
    cnt += in ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) ninl[pair] = cnt;
}

static size_t ransac_smem_bytes(int nhyp, int chunk) {
  return (size_t)chunk * sizeof(float4) + (size_t)nhyp * (8 * sizeof(float) + 2 * sizeof(int));
}

static int ransac_pick_chunk(int nhyp) {
  // keep RS_MIN_BLOCKS CTAs per SM resident when the hypothesis table allows it
  const size_t table = (size_t)nhyp * 40;
  size_t budget = (RS_MIN_BLOCKS >= 3 ? 73 : 100) * 1024;
  if (table + 2048 * sizeof(float4) > budget) budget = 100 * 1024;
  if (table + 1024 * sizeof(float4) > budget) budget = 220 * 1024;
  if (table + 256 * sizeof(float4) > budget) return -1;
  size_t c = (budget - table) / sizeof(float4);
  if (c > RS_CHUNK) c = RS_CHUNK;
  c = c / 32 * 32;
  return (int)c;
}

struct RansacWs {
  int32_t* best_hyp;
  double* best_H;
  int32_t* sel;
  int4* samples;
  size_t bytes;
};

static RansacWs ransac_ws_layout(void* ws, int64_t total, int64_t npairs, int nhyp, int sampler, bool full) {
  char* p = reinterpret_cast<char*>(ws);
  size_t off = 0;
  RansacWs w{};
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes > 0 ? bytes : 1, 256); return r; };
  if (full) {
    w.best_hyp = reinterpret_cast<int32_t*>(take((size_t)npairs * sizeof(int32_t)));
    w.best_H = reinterpret_cast<double*>(take((size_t)npairs * 9 * sizeof(double)));
    w.sel = reinterpret_cast<int32_t*>(take((size_t)total * sizeof(int32_t)));
  }
  if (sampler == HPUSM_RANSAC_SAMPLER_CV2) w.samples = reinterpret_cast<int4*>(take((size_t)npairs * nhyp * sizeof(int4)));
  w.bytes = off > 0 ? off : 256;
  return w;
}

static int launch_sampler(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs, int nhyp,
                          int4* samples, cudaStream_t st) {
  if (npairs < 4 * (int64_t)sm_count()) {
    // few pairs: one warp per pair tests 32 attempts at a time (a thread per pair would walk ~10^4 dependent attempts)
    HPUSM_LAUNCH(ransac_sample_cv2_warp_kernel, (unsigned)((npairs + 3) / 4), 128, 0, st, src, dst, pair_offsets, npairs, nhyp,
                 samples);
  } else {
    const int threads = npairs >= 64 * 148 ? 64 : 32;
    const unsigned grid = (unsigned)((npairs + threads - 1) / threads);
    HPUSM_LAUNCH(ransac_sample_cv2_kernel, grid, threads, 0, st, src, dst, pair_offsets, npairs, nhyp, samples);
  }
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

static int launch_score(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs,
                        float thresh, int nhyp, double confidence, int sampler, uint64_t seed, int64_t pair0,
                        const int4* samples, int32_t* counts, int32_t* best_hyp, double* best_H, cudaStream_t st) {
  const int chunk = ransac_pick_chunk(nhyp);
  HPUSM_REQUIRE(chunk > 0, HPUSM_ERR_UNSUPPORTED, "ransac: %d hypotheses do not fit in shared memory", nhyp);
  const size_t smem = ransac_smem_bytes(nhyp, chunk);
  if (sampler == HPUSM_RANSAC_SAMPLER_CV2) {
    HPUSM_CUDA_OK(cudaFuncSetAttribute(ransac_score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HPUSM_LAUNCH_TIMED(ransac_score_kernel<true>, (unsigned)npairs, RS_THREADS, smem, st, src, dst, pair_offsets, thresh, nhyp,
                       confidence, seed, pair0, chunk, samples, counts, best_hyp, best_H);
  } else {
    HPUSM_CUDA_OK(cudaFuncSetAttribute(ransac_score_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HPUSM_LAUNCH_TIMED(ransac_score_kernel<false>, (unsigned)npairs, RS_THREADS, smem, st, src, dst, pair_offsets, thresh, nhyp,
                       confidence, seed, pair0, chunk, samples, counts, best_hyp, best_H);
  }
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

}  // namespace hpusm

using namespace hpusm;

extern "C" {

size_t hpusm_ransac_homography_workspace_bytes(int64_t total_matches, int64_t npairs, int nhyp, int sampler) {
  if (total_matches < 0 || npairs < 0 || nhyp <= 0) return 256;
  return ransac_ws_layout(nullptr, total_matches, npairs, nhyp, sampler, true).bytes;
}

size_t hpusm_ransac_score_workspace_bytes(int64_t total_matches, int64_t npairs, int nhyp, int sampler) {
  if (total_matches < 0 || npairs < 0 || nhyp <= 0) return 256;
  return ransac_ws_layout(nullptr, total_matches, npairs, nhyp, sampler, false).bytes;
}

int hpusm_ransac_score_batch(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs,
                             int64_t total_matches, float thresh, int nhyp, int sampler, uint64_t seed, int64_t pair_index0,
                             int32_t* counts, void* ws, size_t ws_bytes, void* stream) {
  HPUSM_REQUIRE(npairs >= 0 && total_matches >= 0 && nhyp > 0 && pair_index0 >= 0, HPUSM_ERR_INVALID,
                "hpusm_ransac_score_batch: bad sizes");
  HPUSM_REQUIRE(sampler == HPUSM_RANSAC_SAMPLER_COUNTER || sampler == HPUSM_RANSAC_SAMPLER_CV2, HPUSM_ERR_INVALID,
                "hpusm_ransac_score_batch: unknown sampler %d", sampler);
  if (npairs == 0) return HPUSM_OK;
  HPUSM_REQUIRE(pair_offsets && counts && ((src && dst) || total_matches == 0), HPUSM_ERR_INVALID,
                "hpusm_ransac_score_batch: null pointer");
  HPUSM_REQUIRE(npairs < 0x40000000, HPUSM_ERR_UNSUPPORTED, "hpusm_ransac_score_batch: too many pairs for one call");
  cudaStream_t st = as_stream(stream);
  RansacWs w{};
  if (sampler == HPUSM_RANSAC_SAMPLER_CV2) {
    w = ransac_ws_layout(ws, total_matches, npairs, nhyp, sampler, false);
    HPUSM_REQUIRE(ws && is_aligned(ws, 256) && ws_bytes >= w.bytes, HPUSM_ERR_WORKSPACE,
                  "hpusm_ransac_score_batch: workspace needs %zu bytes, got %zu", w.bytes, ws_bytes);
    const int rc = launch_sampler(src, dst, pair_offsets, npairs, nhyp, w.samples, st);
    if (rc != HPUSM_OK) return rc;
  }
  return launch_score(src, dst, pair_offsets, npairs, thresh, nhyp, 0.0, sampler, seed, pair_index0, w.samples, counts, nullptr,
                      nullptr, st);
}

int hpusm_ransac_homography_batch(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs,
                                  int64_t total_matches, float thresh, int nhyp, double confidence, int sampler,
                                  uint64_t seed, int64_t pair_index0, int refine, double* H, uint8_t* mask, int32_t* ninl,
                                  int32_t* best_hyp, void* ws, size_t ws_bytes, void* stream) {
  HPUSM_REQUIRE(npairs >= 0 && total_matches >= 0 && nhyp > 0 && pair_index0 >= 0, HPUSM_ERR_INVALID,
                "hpusm_ransac_homography_batch: bad sizes");
  HPUSM_REQUIRE(sampler == HPUSM_RANSAC_SAMPLER_COUNTER || sampler == HPUSM_RANSAC_SAMPLER_CV2, HPUSM_ERR_INVALID,
                "hpusm_ransac_homography_batch: unknown sampler %d", sampler);
  if (npairs == 0) return HPUSM_OK;
  HPUSM_REQUIRE(pair_offsets && H && ninl && (mask || total_matches == 0) && ((src && dst) || total_matches == 0),
                HPUSM_ERR_INVALID, "hpusm_ransac_homography_batch: null pointer");
  HPUSM_REQUIRE(npairs < 0x40000000 && total_matches < 0x7fffffff, HPUSM_ERR_UNSUPPORTED,
                "hpusm_ransac_homography_batch: too many pairs / matches for one call");
  const RansacWs w = ransac_ws_layout(ws, total_matches, npairs, nhyp, sampler, true);
  HPUSM_REQUIRE(ws && is_aligned(ws, 256) && ws_bytes >= w.bytes, HPUSM_ERR_WORKSPACE,
                "hpusm_ransac_homography_batch: workspace needs %zu bytes, got %zu", w.bytes, ws_bytes);
  cudaStream_t st = as_stream(stream);
  int rc;
  if (sampler == HPUSM_RANSAC_SAMPLER_CV2) {
    rc = launch_sampler(src, dst, pair_offsets, npairs, nhyp, w.samples, st);
    if (rc != HPUSM_OK) return rc;
  }
  rc = launch_score(src, dst, pair_offsets, npairs, thresh, nhyp, confidence, sampler, seed, pair_index0, w.samples, nullptr,
                    w.best_hyp, w.best_H, st);
  if (rc != HPUSM_OK) return rc;
  const unsigned grid = (unsigned)((npairs + RF_WARPS - 1) / RF_WARPS);
  if (sampler == HPUSM_RANSAC_SAMPLER_CV2)
    HPUSM_LAUNCH(ransac_finalize_kernel<true>, grid, RF_WARPS * 32, 0, st, src, dst, pair_offsets, npairs, thresh, refine,
                 w.best_hyp, w.best_H, w.sel, H, mask, ninl);
  else
    HPUSM_LAUNCH(ransac_finalize_kernel<false>, grid, RF_WARPS * 32, 0, st, src, dst, pair_offsets, npairs, thresh, refine,
                 w.best_hyp, w.best_H, w.sel, H, mask, ninl);
  HPUSM_CHECK_LAUNCH();
  if (best_hyp) HPUSM_CUDA_OK(cudaMemcpyAsync(best_hyp, w.best_hyp, (size_t)npairs * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return HPUSM_OK;
}

}  // extern "C"
