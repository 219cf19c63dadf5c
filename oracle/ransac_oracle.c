/*
 * ransac_oracle.c -- CPU restatement of the batched RANSAC homography path.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product never does.
 *
 * What it restates.  The reference calls cv2.findHomography(src, dst, cv2.RANSAC, 5.0) at
 * urbanscene/features.py:66 and :69.  The arithmetic lives in OpenCV, which is NOT vendored under /root/reference
 * (installed binary here: opencv-python-headless 4.13.0.92), so its published algorithm is restated function by function:
 *   RANSACPointSetRegistrator::run / getSubset      (modules/calib3d/src/ptsetreg.cpp)  -> run_pair(), sampler_cv2()
 *   HomographyEstimatorCallback::checkSubset,
 *     haveCollinearPoints                            (ptsetreg.cpp / fundam.cpp)         -> cv_check_subset()
 *   HomographyEstimatorCallback::computeError        (fundam.cpp)                        -> cv_error_mask()
 *   HomographyEstimatorCallback::runKernel           (fundam.cpp, normalised DLT)        -> cv_dlt()
 *   RANSACUpdateNumIters                             (ptsetreg.cpp)                      -> adaptive_iters()
 *   cv::RNG (multiply-with-carry, RNG((uint64)-1))   (core/include/opencv2/core.hpp)     -> cv_rng_next()
 *   findHomography's post-processing: runKernel on the inliers + LMSolver(10 iterations) on HomographyRefineCallback,
 *     then the mask of the refined model           (fundam.cpp, levmarq.cpp)            -> cv_refine(), cv_lm()
 *   cv::eigen / cv::solve(DECOMP_EIG) / cv::invert(DECOMP_EIG) = JacobiImpl_ + SVBkSb  (core/src/lapack.cpp)  -> jacobi_eig()
 *
 * Parity pinning (tests/test_oracle_golden.py, tests/golden/ransac_real_pairs.npz, ransac_synth_cv2.npz):
 * with sampler == SAMPLER_CV2 the sample sequence is OpenCV's own (its generator is seeded with a constant and is
 * independent of cv2.setRNGSeed), so this file reproduces cv2.findHomography itself: identical masks and H to ~1e-8 on
 * every recorded real image pair whose homography is well posed and on 1000 noisy synthetic pairs.
 * With sampler == SAMPLER_COUNTER the 4-point samples come from a counter-based table keyed on
 * (seed, pair, hypothesis, attempt) -- every hypothesis is an independent function of its key, which is what lets the
 * CUDA kernel draw them in parallel -- and everything else is unchanged (OpenCV's subset check, redraw until it passes).
 *
 * The one piece that is not OpenCV's: the 4-point minimal solver is a closed form (unit square -> quad on both sides)
 * instead of runKernel's 9x9 eigen decomposition.  Both return the unique homography through 4 points in general
 * position; they differ by rounding only (checked: the RANSAC stage picks the same winner as cv2 on every golden).
 * The scoring formula is OpenCV's float32 computeError under SAMPLER_CV2; under SAMPLER_COUNTER the CUDA kernel's
 * division-free FMA form (u - x'w)^2 + (v - y'w)^2 <= thr^2 w^2 is used, op for op (fmaf() where the kernel uses
 * __fmaf_rn), so that counts are bit-identical between this file and the kernel in both modes.
 * Build with -ffp-contract=off.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SAMPLER_COUNTER 0
#define SAMPLER_CV2 1
#define COUNTER_MAX_ATTEMPTS 256 /* redraws per hypothesis before it is given up (counts -1) */
#define CV_MAX_ATTEMPTS 10000    /* getSubset(..., 10000) in RANSACPointSetRegistrator::run */

/* ---- counter-based sample table: 4 distinct indices from (seed, pair, hypothesis, attempt) ------------------------ */
static uint64_t mix64(uint64_t* s) {
  *s += 0x9E3779B97F4A7C15ull;
  uint64_t z = *s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void ransac_oracle_sample4(uint64_t seed, uint64_t pair, uint64_t hyp, uint32_t attempt, uint32_t K, uint32_t out[4]) {
  uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (pair + 1)) ^ (0x8CB92BA72F3D8DD7ull * (hyp + 1)) ^
               (0xA0761D6478BD642Full * (uint64_t)attempt);
  /* two 64-bit outputs = four 32-bit uniforms */
  const uint64_t r0 = mix64(&s), r1 = mix64(&s);
  const uint32_t u[4] = {(uint32_t)(r0 >> 32), (uint32_t)r0, (uint32_t)(r1 >> 32), (uint32_t)r1};
  uint32_t taken[4]; /* ascending list of the indices already drawn */
  for (int j = 0; j < 4; ++j) {
    /* rank of the new index among the K-j still unused ones, then skip over the used ones */
    uint32_t v = (uint32_t)(((uint64_t)u[j] * (uint64_t)(K - (uint32_t)j)) >> 32);
    int pos = 0;
    for (int c = 0; c < j; ++c)
      if (v >= taken[c]) { ++v; pos = c + 1; }
    for (int c = j; c > pos; --c) taken[c] = taken[c - 1];
    taken[pos] = v;
    out[j] = v;
  }
}

/* ---- OpenCV's RNG: multiply-with-carry, state' = (uint32)state * 4164903690 + (state >> 32) ----------------------- */
static uint32_t cv_rng_next(uint64_t* state) {
  *state = (uint64_t)(uint32_t)(*state) * 4164903690ull + (uint32_t)(*state >> 32);
  return (uint32_t)(*state);
}

/* ---- HomographyEstimatorCallback::checkSubset for a 4-point sample ---------------------------------------------- */
static int cv_collinear_with_last(const double* x, const double* y) {
  /* haveCollinearPoints: only triples that contain the LAST sample point are tested */
  const int i = 3;
  for (int j = 0; j < i; ++j) {
    const double dx1 = x[j] - x[i], dy1 = y[j] - y[i];
    for (int k = 0; k < j; ++k) {
      const double dx2 = x[k] - x[i], dy2 = y[k] - y[i];
      const double a = dx2 * dy1, b = dy2 * dx1;
      const double lhs = fabs(a - b);
      const double rhs = (double)FLT_EPSILON * (((fabs(dx1) + fabs(dy1)) + fabs(dx2)) + fabs(dy2));
      if (lhs <= rhs) return 1;
    }
  }
  return 0;
}

/* determinant of [[x0,y0,1],[x1,y1,1],[x2,y2,1]] in the operation order of cv::Matx33d's determinant */
static double cv_det3(double x0, double y0, double x1, double y1, double x2, double y2) {
  const double m0 = y1 * 1.0 - y2 * 1.0;
  const double m1 = x1 * 1.0 - x2 * 1.0;
  const double p12 = x1 * y2, p21 = x2 * y1;
  const double m2 = p12 - p21;
  const double t0 = x0 * m0, t1 = y0 * m1, t2 = 1.0 * m2;
  return (t0 - t1) + t2;
}

static int cv_check_subset(const double* sx, const double* sy, const double* dx, const double* dy) {
  static const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
  if (cv_collinear_with_last(sx, sy) || cv_collinear_with_last(dx, dy)) return 0;
  int negative = 0;
  for (int i = 0; i < 4; ++i) {
    const int* t = tt[i];
    const double dA = cv_det3(sx[t[0]], sy[t[0]], sx[t[1]], sy[t[1]], sx[t[2]], sy[t[2]]);
    const double dB = cv_det3(dx[t[0]], dy[t[0]], dx[t[1]], dy[t[1]], dx[t[2]], dy[t[2]]);
    negative += (dA * dB < 0.0) ? 1 : 0;
  }
  return negative == 0 || negative == 4;
}

static void load4(const float* src, const float* dst, const uint32_t id[4], double* sx, double* sy, double* dx, double* dy) {
  for (int j = 0; j < 4; ++j) {
    sx[j] = (double)src[2 * id[j]]; sy[j] = (double)src[2 * id[j] + 1];
    dx[j] = (double)dst[2 * id[j]]; dy[j] = (double)dst[2 * id[j] + 1];
  }
}

/* Sample table of one pair: samples[nhyp][4], -1 where no sample exists.  Returns the number of iterations that have one. */
static int sampler_counter(const float* src, const float* dst, uint32_t K, uint64_t seed, uint64_t pair, int nhyp,
                           int32_t* samples) {
  int n = 0;
  for (int h = 0; h < nhyp; ++h) {
    int ok = 0;
    uint32_t id[4];
    if (K >= 4) {
      for (uint32_t a = 0; a < COUNTER_MAX_ATTEMPTS && !ok; ++a) {
        double sx[4], sy[4], dx[4], dy[4];
        ransac_oracle_sample4(seed, pair, (uint64_t)h, a, K, id);
        load4(src, dst, id, sx, sy, dx, dy);
        ok = cv_check_subset(sx, sy, dx, dy);
      }
    }
    for (int j = 0; j < 4; ++j) samples[4 * h + j] = ok ? (int32_t)id[j] : -1;
    n += ok;
  }
  return n;
}

/* OpenCV's own sequence: RNG((uint64)-1), rng.uniform(0, count) with duplicate redraw, checkSubset, up to 10000 attempts
 * per iteration; when an iteration finds no subset the loop ends (all later iterations have no sample). */
static int sampler_cv2(const float* src, const float* dst, uint32_t K, int nhyp, int32_t* samples) {
  uint64_t state = 0xFFFFFFFFFFFFFFFFull;
  int h = 0;
  if (K >= 4) {
    for (; h < nhyp; ++h) {
      int found = 0;
      uint32_t id[4];
      for (int it = 0; it < CV_MAX_ATTEMPTS && !found; ++it) {
        for (int i = 0; i < 4; ++i) {
          for (;;) {
            const uint32_t v = cv_rng_next(&state) % K;
            int dup = 0;
            for (int c = 0; c < i; ++c) dup |= (id[c] == v);
            if (!dup) { id[i] = v; break; }
          }
        }
        double sx[4], sy[4], dx[4], dy[4];
        load4(src, dst, id, sx, sy, dx, dy);
        found = cv_check_subset(sx, sy, dx, dy);
      }
      if (!found) break;
      for (int j = 0; j < 4; ++j) samples[4 * h + j] = (int32_t)id[j];
    }
  }
  const int n = h;
  for (; h < nhyp; ++h)
    for (int j = 0; j < 4; ++j) samples[4 * h + j] = -1;
  return n;
}

int ransac_oracle_samples(const float* src, const float* dst, uint32_t K, int sampler, uint64_t seed, uint64_t pair, int nhyp,
                          int32_t* samples) {
  return sampler == SAMPLER_CV2 ? sampler_cv2(src, dst, K, nhyp, samples) : sampler_counter(src, dst, K, seed, pair, nhyp, samples);
}

/* ---- minimal 4-point solver (closed form, shared op for op with csrc/ransac.cu) ----------------------------------- */
static double det2(double a, double b, double c, double d) {
  const double ad = a * d, bc = b * c;
  return ad - bc;
}

/* projective map of the unit square onto the quad (x,y)[0..3]; 0 when it does not exist */
static int unit_square_to_quad(const double* x, const double* y, double* M) {
  const double dx1 = x[1] - x[2], dx2 = x[3] - x[2];
  const double dy1 = y[1] - y[2], dy2 = y[3] - y[2];
  const double px = (x[0] - x[1]) + (x[2] - x[3]);
  const double py = (y[0] - y[1]) + (y[2] - y[3]);
  const double den = det2(dx1, dx2, dy1, dy2);
  if (den == 0.0) return 0;
  const double g = det2(px, dx2, py, dy2) / den;
  const double h = det2(dx1, px, dy1, py) / den;
  M[0] = (x[1] - x[0]) + g * x[1]; M[1] = (x[3] - x[0]) + h * x[3]; M[2] = x[0];
  M[3] = (y[1] - y[0]) + g * y[1]; M[4] = (y[3] - y[0]) + h * y[3]; M[5] = y[0];
  M[6] = g; M[7] = h; M[8] = 1.0;
  return 1;
}

static int finite_all(const double* H) {
  for (int i = 0; i < 9; ++i)
    if (!(fabs(H[i]) <= DBL_MAX)) return 0;
  return 1;
}

/* H = D * adj(S) / (..)[8] */
static int compose_maps(const double* S, const double* D, double* H) {
  double A[9], R[9];
  A[0] = det2(S[4], S[5], S[7], S[8]); A[1] = det2(S[2], S[1], S[8], S[7]); A[2] = det2(S[1], S[2], S[4], S[5]);
  A[3] = det2(S[5], S[3], S[8], S[6]); A[4] = det2(S[0], S[2], S[6], S[8]); A[5] = det2(S[2], S[0], S[5], S[3]);
  A[6] = det2(S[3], S[4], S[6], S[7]); A[7] = det2(S[1], S[0], S[7], S[6]); A[8] = det2(S[0], S[1], S[3], S[4]);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double t0 = D[3 * i] * A[j], t1 = D[3 * i + 1] * A[3 + j], t2 = D[3 * i + 2] * A[6 + j];
      R[3 * i + j] = (t0 + t1) + t2;
    }
  if (R[8] == 0.0) return 0;
  for (int i = 0; i < 9; ++i) H[i] = R[i] / R[8];
  return finite_all(H);
}

/* exact homography through the four correspondences id[0..3]; 0 when it does not exist */
static int minimal_model(const float* src, const float* dst, const int32_t* id, double* H) {
  double sx[4], sy[4], dx[4], dy[4];
  if (id[0] < 0) return 0;
  const uint32_t u[4] = {(uint32_t)id[0], (uint32_t)id[1], (uint32_t)id[2], (uint32_t)id[3]};
  load4(src, dst, u, sx, sy, dx, dy);
  const double osx = sx[0], osy = sy[0], odx = dx[0], ody = dy[0];
  for (int j = 0; j < 4; ++j) { sx[j] -= osx; sy[j] -= osy; dx[j] -= odx; dy[j] -= ody; }
  double S[9], D[9], Hc[9], G[9], R[9];
  if (!unit_square_to_quad(sx, sy, S)) return 0;
  if (!unit_square_to_quad(dx, dy, D)) return 0;
  if (!compose_maps(S, D, Hc)) return 0;
  /* fold the two translations back in: H = T(+od) * Hc * T(-os) */
  for (int j = 0; j < 3; ++j) {
    const double a = odx * Hc[6 + j], b = ody * Hc[6 + j];
    G[j] = Hc[j] + a;
    G[3 + j] = Hc[3 + j] + b;
    G[6 + j] = Hc[6 + j];
  }
  for (int i = 0; i < 3; ++i) {
    R[3 * i] = G[3 * i]; R[3 * i + 1] = G[3 * i + 1];
    const double a = G[3 * i] * osx, b = G[3 * i + 1] * osy;
    R[3 * i + 2] = (G[3 * i + 2] - a) - b;
  }
  if (R[8] == 0.0) return 0;
  for (int i = 0; i < 9; ++i) H[i] = R[i] / R[8];
  return finite_all(H);
}

int ransac_oracle_hypothesis(const float* src, const float* dst, uint32_t K, int sampler, uint64_t seed, uint64_t pair,
                             int hyp, double* H) {
  int32_t* samples = (int32_t*)malloc((size_t)(hyp + 1) * 4 * sizeof(int32_t));
  ransac_oracle_samples(src, dst, K, sampler, seed, pair, hyp + 1, samples);
  const int ok = minimal_model(src, dst, samples + 4 * hyp, H);
  free(samples);
  return ok;
}

/* ---- inlier tests ------------------------------------------------------------------------------------------------ */
/* division-free test in float, the CUDA kernel's form under SAMPLER_COUNTER */
static int fast_is_inlier(const float* h, float x, float y, float xp, float yp, float thr2) {
  const float w = fmaf(h[6], x, fmaf(h[7], y, h[8]));
  const float u = fmaf(h[0], x, fmaf(h[1], y, h[2]));
  const float v = fmaf(h[3], x, fmaf(h[4], y, h[5]));
  const float du = fmaf(-xp, w, u);
  const float dv = fmaf(-yp, w, v);
  const float dv2 = dv * dv;
  const float e = fmaf(du, du, dv2);
  const float tw = thr2 * w;
  const float lim = tw * w;
  return e <= lim;
}

/* HomographyEstimatorCallback::computeError, float32 with a division, one rounding per operation; H[8] is taken as 1 */
static int cv_is_inlier(const float* h, float x, float y, float xp, float yp, float thr2) {
  const float a6 = h[6] * x, a7 = h[7] * y;
  const float ww = 1.f / ((a6 + a7) + 1.f);
  const float a0 = h[0] * x, a1 = h[1] * y, a3 = h[3] * x, a4 = h[4] * y;
  const float px = ((a0 + a1) + h[2]) * ww, py = ((a3 + a4) + h[5]) * ww;
  const float ex = px - xp, ey = py - yp;
  const float ex2 = ex * ex, ey2 = ey * ey;
  return (ex2 + ey2) <= thr2;
}

static int count_inliers(const double* H, const float* src, const float* dst, int64_t K, float thr2, int cv_formula,
                         uint8_t* mask) {
  float h[9];
  int c = 0;
  for (int i = 0; i < 9; ++i) h[i] = (float)H[i];
  for (int64_t i = 0; i < K; ++i) {
    const int in = cv_formula ? cv_is_inlier(h, src[2 * i], src[2 * i + 1], dst[2 * i], dst[2 * i + 1], thr2)
                              : fast_is_inlier(h, src[2 * i], src[2 * i + 1], dst[2 * i], dst[2 * i + 1], thr2);
    if (mask) mask[i] = (uint8_t)in;
    c += in;
  }
  return c;
}

/* RANSACUpdateNumIters with modelPoints = 4 */
static int adaptive_iters(double conf, double outlier_ratio, int max_iters) {
  double p = fmin(fmax(conf, 0.0), 1.0);
  double ep = fmin(fmax(outlier_ratio, 0.0), 1.0);
  double num = fmax(1.0 - p, DBL_MIN);
  const double g = 1.0 - ep;
  const double g2 = g * g;
  double denom = 1.0 - g2 * g2;
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  if (denom >= 0.0 || -num >= (double)max_iters * (-denom)) return max_iters;
  return (int)llrint(num / denom);
}

/* ---- cv::eigen for symmetric matrices: JacobiImpl_ (classical Jacobi with tracked row/column pivots) ------------ */
#define JN 9
static void jacobi_eig(const double* Ain, int n, double* W, double* V /* rows = eigenvectors, descending W */) {
  double A[JN * JN];
  int indR[JN], indC[JN];
  const double eps = DBL_EPSILON;
  memcpy(A, Ain, sizeof(double) * (size_t)n * (size_t)n);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) V[i * n + j] = 0.0;
    V[i * n + i] = 1.0;
  }
  double mv = 0.0;
  int m;
  for (int k = 0; k < n; ++k) {
    W[k] = A[(n + 1) * k];
    if (k < n - 1) {
      m = k + 1; mv = fabs(A[n * k + m]);
      for (int i = k + 2; i < n; ++i) { const double val = fabs(A[n * k + i]); if (mv < val) { mv = val; m = i; } }
      indR[k] = m;
    }
    if (k > 0) {
      m = 0; mv = fabs(A[k]);
      for (int i = 1; i < k; ++i) { const double val = fabs(A[n * i + k]); if (mv < val) { mv = val; m = i; } }
      indC[k] = m;
    }
  }
  if (n > 1) {
    const int max_iters = n * n * 30;
    for (int iters = 0; iters < max_iters; ++iters) {
      int k = 0, l;
      mv = fabs(A[indR[0]]);
      for (int i = 1; i < n - 1; ++i) { const double val = fabs(A[n * i + indR[i]]); if (mv < val) { mv = val; k = i; } }
      l = indR[k];
      for (int i = 1; i < n; ++i) { const double val = fabs(A[n * indC[i] + i]); if (mv < val) { mv = val; k = indC[i]; l = i; } }
      const double p = A[n * k + l];
      if (fabs(p) <= eps) break;
      const double y = (W[l] - W[k]) * 0.5;
      double t = fabs(y) + hypot(p, y);
      double s = hypot(p, t);
      const double c = t / s;
      s = p / s; t = (p / t) * p;
      if (y < 0) { s = -s; t = -t; }
      A[n * k + l] = 0;
      W[k] -= t;
      W[l] += t;
#define ROTATE(v0, v1) do { const double a0 = (v0), b0 = (v1); const double q0 = a0 * c, q1 = b0 * s, q2 = a0 * s, q3 = b0 * c; \
                            (v0) = q0 - q1; (v1) = q2 + q3; } while (0)
      for (int i = 0; i < k; ++i) ROTATE(A[n * i + k], A[n * i + l]);
      for (int i = k + 1; i < l; ++i) ROTATE(A[n * k + i], A[n * i + l]);
      for (int i = l + 1; i < n; ++i) ROTATE(A[n * k + i], A[n * l + i]);
      for (int i = 0; i < n; ++i) ROTATE(V[n * k + i], V[n * l + i]);
#undef ROTATE
      for (int j = 0; j < 2; ++j) {
        const int idx = j == 0 ? k : l;
        if (idx < n - 1) {
          m = idx + 1; mv = fabs(A[n * idx + m]);
          for (int i = idx + 2; i < n; ++i) { const double val = fabs(A[n * idx + i]); if (mv < val) { mv = val; m = i; } }
          indR[idx] = m;
        }
        if (idx > 0) {
          m = 0; mv = fabs(A[idx]);
          for (int i = 1; i < idx; ++i) { const double val = fabs(A[n * i + idx]); if (mv < val) { mv = val; m = i; } }
          indC[idx] = m;
        }
      }
    }
  }
  for (int k = 0; k < n - 1; ++k) {
    m = k;
    for (int i = k + 1; i < n; ++i)
      if (W[m] < W[i]) m = i;
    if (k != m) {
      const double tw = W[m]; W[m] = W[k]; W[k] = tw;
      for (int i = 0; i < n; ++i) { const double tv = V[n * m + i]; V[n * m + i] = V[n * k + i]; V[n * k + i] = tv; }
    }
  }
}

/* cv::solve(A, b, x, DECOMP_EIG) for symmetric A: eigen decomposition + SVBkSb (eigenvalues <= 2*DBL_EPSILON*sum dropped) */
static void eig_solve(const double* A, const double* b, int n, double* x) {
  double W[JN], V[JN * JN];
  jacobi_eig(A, n, W, V);
  double threshold = 0.0;
  for (int i = 0; i < n; ++i) threshold += W[i];
  threshold *= DBL_EPSILON * 2;
  for (int j = 0; j < n; ++j) x[j] = 0.0;
  for (int i = 0; i < n; ++i) {
    double wi = W[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += V[n * i + j] * b[j];
    s *= wi;
    for (int j = 0; j < n; ++j) x[j] = x[j] + s * V[n * i + j];
  }
}

/* largest |diagonal entry| of cv::invert(A, DECOMP_EIG) */
static double eig_inverse_max_diag(const double* A, int n) {
  double W[JN], V[JN * JN];
  jacobi_eig(A, n, W, V);
  double threshold = 0.0;
  for (int i = 0; i < n; ++i) threshold += W[i];
  threshold *= DBL_EPSILON * 2;
  double maxval = DBL_EPSILON;
  for (int j = 0; j < n; ++j) {
    double d = 0.0;
    for (int i = 0; i < n; ++i) {
      if (fabs(W[i]) <= threshold) continue;
      d += V[n * i + j] * V[n * i + j] / W[i];
    }
    maxval = fmax(maxval, fabs(d));
  }
  return maxval;
}

/* ---- HomographyEstimatorCallback::runKernel: normalised DLT, smallest eigenvector of LtL -------------------------- */
static int cv_dlt(const float* M /* src */, const float* m /* dst */, const int32_t* sel, int count, double* H) {
  double cMx = 0, cMy = 0, cmx = 0, cmy = 0, sMx = 0, sMy = 0, smx = 0, smy = 0;
  for (int i = 0; i < count; ++i) {
    const int64_t r = sel[i];
    cmx += m[2 * r]; cmy += m[2 * r + 1]; cMx += M[2 * r]; cMy += M[2 * r + 1];
  }
  cmx /= count; cmy /= count; cMx /= count; cMy /= count;
  for (int i = 0; i < count; ++i) {
    const int64_t r = sel[i];
    smx += fabs(m[2 * r] - cmx); smy += fabs(m[2 * r + 1] - cmy);
    sMx += fabs(M[2 * r] - cMx); sMy += fabs(M[2 * r + 1] - cMy);
  }
  if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON) return 0;
  smx = count / smx; smy = count / smy; sMx = count / sMx; sMy = count / sMy;
  const double invHnorm[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
  const double Hnorm2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
  double LtL[81];
  memset(LtL, 0, sizeof(LtL));
  for (int i = 0; i < count; ++i) {
    const int64_t r = sel[i];
    const double x = (m[2 * r] - cmx) * smx, y = (m[2 * r + 1] - cmy) * smy;
    const double X = (M[2 * r] - cMx) * sMx, Y = (M[2 * r + 1] - cMy) * sMy;
    const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
    const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
    for (int j = 0; j < 9; ++j)
      for (int k = j; k < 9; ++k) LtL[9 * j + k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
  }
  for (int j = 0; j < 9; ++j)
    for (int k = 0; k < j; ++k) LtL[9 * j + k] = LtL[9 * k + j];
  double W[9], V[81];
  jacobi_eig(LtL, 9, W, V);
  const double* H0 = V + 72; /* eigenvector of the smallest eigenvalue */
  double T[9], R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += invHnorm[3 * i + k] * H0[3 * k + j];
      T[3 * i + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += T[3 * i + k] * Hnorm2[3 * k + j];
      R[3 * i + j] = s;
    }
  const double scale = 1. / R[8];
  for (int i = 0; i < 9; ++i) H[i] = R[i] * scale;
  return 1;
}

/* ---- HomographyRefineCallback::compute: residuals (and JtJ, Jtr) in pixel coordinates, h[8] == 1 ----------------- */
static double lm_compute(const double* h, const float* M, const float* m, const int32_t* sel, int count, double* JtJ /*64*/,
                         double* Jtr /*8*/, double* rinf) {
  double S = 0.0, rmax = 0.0;
  if (JtJ) { memset(JtJ, 0, 64 * sizeof(double)); memset(Jtr, 0, 8 * sizeof(double)); }
  for (int i = 0; i < count; ++i) {
    const int64_t r = sel[i];
    const double Mx = M[2 * r], My = M[2 * r + 1];
    double ww = h[6] * Mx + h[7] * My + 1.;
    ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
    const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
    const double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
    const double e0 = xi - m[2 * r], e1 = yi - m[2 * r + 1];
    S += e0 * e0; S += e1 * e1;
    rmax = fmax(rmax, fmax(fabs(e0), fabs(e1)));
    if (JtJ) {
      const double J0[8] = {Mx * ww, My * ww, ww, 0, 0, 0, -Mx * ww * xi, -My * ww * xi};
      const double J1[8] = {0, 0, 0, Mx * ww, My * ww, ww, -Mx * ww * yi, -My * ww * yi};
      for (int a = 0; a < 8; ++a) {
        for (int b = 0; b < 8; ++b) JtJ[8 * a + b] += J0[a] * J0[b] + J1[a] * J1[b];
        Jtr[a] += J0[a] * e0 + J1[a] * e1;
      }
    }
  }
  if (rinf) *rinf = rmax;
  return S;
}

/* LMSolverImpl::run (levmarq.cpp) with maxIters iterations and eps = FLT_EPSILON; x holds 8 parameters */
static int cv_lm(double* x, const float* M, const float* m, const int32_t* sel, int count, int max_iters) {
  const int lx = 8;
  double A[64], v[8], D[8], Ap[64], d[8], xd[8], r_inf = 0.0;
  double S = lm_compute(x, M, m, sel, count, A, v, &r_inf);
  for (int i = 0; i < lx; ++i) D[i] = A[9 * i];
  const double Rlo = 0.25, Rhi = 0.75;
  double lambda = 1, lc = 0.75;
  int iter = 0;
  for (;;) {
    memcpy(Ap, A, sizeof(A));
    for (int i = 0; i < lx; ++i) Ap[9 * i] += lambda * D[i];
    eig_solve(Ap, v, lx, d);
    for (int i = 0; i < lx; ++i) xd[i] = x[i] - d[i];
    const double Sd = lm_compute(xd, M, m, sel, count, NULL, NULL, NULL);
    double dS = 0.0, dinf = 0.0;
    for (int i = 0; i < lx; ++i) {
      double t = 0.0;
      for (int j = 0; j < lx; ++j) t += A[8 * i + j] * d[j];
      dS += d[i] * (-t + 2 * v[i]);
      dinf = fmax(dinf, fabs(d[i]));
    }
    const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
    if (R > Rhi) {
      lambda *= 0.5;
      if (lambda < lc) lambda = 0;
    } else if (R < Rlo) {
      double t = 0.0;
      for (int i = 0; i < lx; ++i) t += d[i] * v[i];
      double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
      nu = fmin(fmax(nu, 2.), 10.);
      if (lambda == 0) {
        const double maxval = eig_inverse_max_diag(A, lx);
        lambda = lc = 1. / maxval;
        nu *= 0.5;
      }
      lambda *= nu;
    }
    if (Sd < S) {
      S = Sd;
      memcpy(x, xd, sizeof(xd));
      (void)lm_compute(x, M, m, sel, count, A, v, &r_inf);
    }
    ++iter;
    const int proceed = iter < max_iters && dinf >= (double)FLT_EPSILON && r_inf >= (double)FLT_EPSILON;
    if (!proceed) break;
  }
  return iter;
}

/* findHomography's post-processing on the rows listed in sel: runKernel + LM(10); 0 when runKernel refuses */
static int cv_refine(const float* src, const float* dst, const int32_t* sel, int count, double* H) {
  double Hk[9];
  if (!cv_dlt(src, dst, sel, count, Hk)) return 0;
  (void)cv_lm(Hk, src, dst, sel, count, 10);
  const double scale = 1. / Hk[8]; /* H.convertTo(H, type, 1/H(2,2)); H(2,2) is 1 here already */
  for (int i = 0; i < 9; ++i) H[i] = Hk[i] * scale;
  return 1;
}

/* Refinement alone on the rows flagged in `mask` (what cv2 does after its RANSAC loop). */
int ransac_oracle_refine(const float* src, const float* dst, const uint8_t* mask, int64_t K, double* Hout) {
  int32_t* sel = (int32_t*)malloc((size_t)(K > 0 ? K : 1) * sizeof(int32_t));
  int n = 0;
  for (int64_t i = 0; i < K; ++i)
    if (mask[i]) sel[n++] = (int32_t)i;
  const int ok = n >= 4 && cv_refine(src, dst, sel, n, Hout);
  free(sel);
  return ok ? 0 : -1;
}

/* counts [npairs][nhyp]: inlier count of every hypothesis, -1 where the iteration has no sample / no model */
int ransac_oracle_score(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs, float thresh,
                        int nhyp, int sampler, uint64_t seed, int64_t pair_index0, int32_t* counts) {
  const float thr2 = thresh * thresh;
  int32_t* samples = (int32_t*)malloc((size_t)nhyp * 4 * sizeof(int32_t));
  for (int64_t p = 0; p < npairs; ++p) {
    const int64_t m0 = pair_offsets[p], K = pair_offsets[p + 1] - m0;
    const float* s = src + 2 * m0;
    const float* d = dst + 2 * m0;
    ransac_oracle_samples(s, d, (uint32_t)K, sampler, seed, (uint64_t)(p + pair_index0), nhyp, samples);
    for (int h = 0; h < nhyp; ++h) {
      double H[9];
      const int ok = minimal_model(s, d, samples + 4 * h, H);
      counts[p * nhyp + h] = ok ? count_inliers(H, s, d, K, thr2, sampler == SAMPLER_CV2, NULL) : -1;
    }
  }
  free(samples);
  return 0;
}

/* Full per-pair result, same contract as hpusm_ransac_homography_batch (include/hpusm.h).
 * min_mask (optional) receives the inlier flags of the winning minimal model (the set the refinement is fitted to). */
int ransac_oracle_homography(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs,
                             float thresh, int nhyp, double confidence, int sampler, uint64_t seed, int64_t pair_index0,
                             int refine, double* H, uint8_t* mask, int32_t* ninl, int32_t* best_hyp, uint8_t* min_mask) {
  const float thr2 = thresh * thresh;
  const int cvf = sampler == SAMPLER_CV2;
  int32_t* samples = (int32_t*)malloc((size_t)nhyp * 4 * sizeof(int32_t));
  for (int64_t p = 0; p < npairs; ++p) {
    const int64_t m0 = pair_offsets[p], K = pair_offsets[p + 1] - m0;
    const float* s = src + 2 * m0;
    const float* d = dst + 2 * m0;
    ransac_oracle_samples(s, d, (uint32_t)K, sampler, seed, (uint64_t)(p + pair_index0), nhyp, samples);
    int best = -1, best_c = 3, niters = nhyp;
    double Hbest[9];
    for (int h = 0; h < nhyp && h < niters; ++h) {
      double Hh[9];
      if (!minimal_model(s, d, samples + 4 * h, Hh)) continue;
      const int c = count_inliers(Hh, s, d, K, thr2, cvf, NULL);
      if (c > best_c) {
        best_c = c; best = h;
        memcpy(Hbest, Hh, sizeof(Hbest));
        if (confidence > 0.0) niters = adaptive_iters(confidence, (double)(K - c) / (double)K, niters);
      }
    }
    if (best_hyp) best_hyp[p] = best;
    if (best < 0) {
      for (int i = 0; i < 9; ++i) H[9 * p + i] = NAN;
      memset(mask + m0, 0, (size_t)K);
      if (min_mask) memset(min_mask + m0, 0, (size_t)K);
      ninl[p] = 0;
      continue;
    }
    const int n_in = count_inliers(Hbest, s, d, K, thr2, cvf, mask + m0);
    if (min_mask) memcpy(min_mask + m0, mask + m0, (size_t)K);
    ninl[p] = n_in;
    memcpy(H + 9 * p, Hbest, sizeof(Hbest));
    if (refine) {
      int32_t* sel = (int32_t*)malloc((size_t)K * sizeof(int32_t));
      int n = 0;
      for (int64_t i = 0; i < K; ++i)
        if (mask[m0 + i]) sel[n++] = (int32_t)i;
      if (cv_refine(s, d, sel, n, H + 9 * p))
        ninl[p] = count_inliers(H + 9 * p, s, d, K, thr2, 1, mask + m0); /* mask of the refined model, computeError formula */
      free(sel);
    }
  }
  free(samples);
  return 0;
}
