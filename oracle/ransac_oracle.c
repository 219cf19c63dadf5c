/*
 * ransac_oracle.c -- CPU restatement of the batched RANSAC homography path.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product never does.
 *
 * What it restates.  The reference calls cv2.findHomography(src, dst, cv2.RANSAC, 5.0) at
 * urbanscene/features.py:66 and :69.  The arithmetic lives in OpenCV (modules/calib3d fundam.cpp /
 * ptsetreg.cpp, NOT vendored under /root/reference; installed binary here: opencv-python-headless
 * 4.13.0.92).  Its published algorithm is restated: random 4-point samples -> degenerate-sample
 * rejection (collinear triples / orientation flips) -> exact 4-point homography -> squared
 * reprojection error in the destination image, inlier iff err <= thresh^2 -> keep the hypothesis with
 * the most inliers -> adaptive iteration bound from the confidence -> least-squares DLT on the inliers
 * + Levenberg-Marquardt on the reprojection error, H normalised to H[2][2] == 1.
 *
 * Parity pinning.  cv2's sampler is an internal fixed-seed RNG that cannot be injected, so
 * "identical inlier sets under the shared hypothesis seed" is defined between THIS file and the CUDA
 * kernels (csrc/ransac.cu), which share the counter-based sample table below and the operation order
 * of the minimal solver and of the inlier test (no FMA contraction here: build with
 * -ffp-contract=off; the inlier test uses fmaf() where the kernel uses __fmaf_rn).  cv2 itself pins
 * the final H / mask on inputs with a unique consensus set (tests/golden/ransac_cv2_*.npz).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- counter-based sample table: 4 distinct indices from (seed, pair, hypothesis) ---------------- */
static uint64_t mix64(uint64_t* s) {
  *s += 0x9E3779B97F4A7C15ull;
  uint64_t z = *s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void ransac_oracle_sample4(uint64_t seed, uint64_t pair, uint64_t hyp, uint32_t K, uint32_t out[4]) {
  uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (pair + 1)) ^ (0x8CB92BA72F3D8DD7ull * (hyp + 1));
  (void)mix64(&s);
  uint32_t taken[4]; /* ascending list of the indices already drawn */
  for (int j = 0; j < 4; ++j) {
    const uint64_t r = mix64(&s);
    /* rank of the new index among the K-j still unused ones, then skip over the used ones */
    uint32_t v = (uint32_t)(((r >> 32) * (uint64_t)(K - (uint32_t)j)) >> 32);
    int pos = 0;
    for (int c = 0; c < j; ++c)
      if (v >= taken[c]) { ++v; pos = c + 1; }
    for (int c = j; c > pos; --c) taken[c] = taken[c - 1];
    taken[pos] = v;
    out[j] = v;
  }
}

/* ---- minimal 4-point solver ----------------------------------------------------------------------- */
static double det2(double a, double b, double c, double d) {
  const double ad = a * d, bc = b * c;
  return ad - bc;
}

static double tri_orient(const double* x, const double* y, int p, int q, int r) {
  return det2(x[q] - x[p], y[q] - y[p], x[r] - x[p], y[r] - y[p]);
}

static double tri_extent(const double* x, const double* y, int p, int q, int r) {
  return (fabs(x[q] - x[p]) + fabs(y[q] - y[p])) + (fabs(x[r] - x[p]) + fabs(y[r] - y[p]));
}

static int sample_is_degenerate(const double* sx, const double* sy, const double* dx, const double* dy) {
  static const int tri[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
  const double eps = 1.1920928955078125e-07; /* 2^-23 */
  int flipped = 0;
  for (int k = 0; k < 4; ++k) {
    const int p = tri[k][0], q = tri[k][1], r = tri[k][2];
    const double os = tri_orient(sx, sy, p, q, r), od = tri_orient(dx, dy, p, q, r);
    const double ms = tri_extent(sx, sy, p, q, r), md = tri_extent(dx, dy, p, q, r);
    if (fabs(os) <= eps * ms) return 1;
    if (fabs(od) <= eps * md) return 1;
    if ((os < 0.0) != (od < 0.0)) ++flipped;
  }
  return flipped != 0 && flipped != 4;
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

int ransac_oracle_hypothesis(const float* src, const float* dst, uint32_t K, uint64_t seed, uint64_t pair,
                             uint64_t hyp, double* H) {
  uint32_t id[4];
  double sx[4], sy[4], dx[4], dy[4];
  if (K < 4) return 0;
  ransac_oracle_sample4(seed, pair, hyp, K, id);
  for (int j = 0; j < 4; ++j) {
    sx[j] = (double)src[2 * id[j]]; sy[j] = (double)src[2 * id[j] + 1];
    dx[j] = (double)dst[2 * id[j]]; dy[j] = (double)dst[2 * id[j] + 1];
  }
  if (sample_is_degenerate(sx, sy, dx, dy)) return 0;
  const double osx = sx[0], osy = sy[0], odx = dx[0], ody = dy[0];
  for (int j = 0; j < 4; ++j) { sx[j] -= osx; sy[j] -= osy; dx[j] -= odx; dy[j] -= ody; }
  double S[9], D[9], Hc[9], G[9], R[9];
  if (!unit_square_to_quad(sx, sy, S)) return 0;
  if (!unit_square_to_quad(dx, dy, D)) return 0;
  if (!compose_maps(S, D, Hc)) return 0;
  /* fold the two translations back in: H = T(+od) * Hc * T(-os) */
  for (int j = 0; j < 3; ++j) {
    G[j] = Hc[j] + odx * Hc[6 + j];
    G[3 + j] = Hc[3 + j] + ody * Hc[6 + j];
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

/* division-free inlier test in float: (u - x'w)^2 + (v - y'w)^2 <= thr^2 w^2 */
static int point_is_inlier(const float* h, float x, float y, float xp, float yp, float thr2) {
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

static int count_inliers(const double* H, const float* src, const float* dst, int64_t K, float thr2, uint8_t* mask) {
  float h[9];
  int c = 0;
  for (int i = 0; i < 9; ++i) h[i] = (float)H[i];
  for (int64_t i = 0; i < K; ++i) {
    const int in = point_is_inlier(h, src[2 * i], src[2 * i + 1], dst[2 * i], dst[2 * i + 1], thr2);
    if (mask) mask[i] = (uint8_t)in;
    c += in;
  }
  return c;
}

/* Final mask: OpenCV (4.13) returns the inliers of the REFINED homography, not of the winning minimal model
 * (probe: with maxIters=5 the returned mask holds 221 points while H is the fit of a much smaller set).  Error
 * formula of its HomographyEstimatorCallback::computeError, float32, with division. */
static int final_mask(const double* H, const float* src, const float* dst, int64_t K, float thr2, uint8_t* mask) {
  float h[9];
  int c = 0;
  for (int i = 0; i < 9; ++i) h[i] = (float)H[i];
  for (int64_t i = 0; i < K; ++i) {
    const float x = src[2 * i], y = src[2 * i + 1];
    const float a6 = h[6] * x, a7 = h[7] * y;
    const float ww = 1.f / ((a6 + a7) + 1.f);
    const float a0 = h[0] * x, a1 = h[1] * y, a3 = h[3] * x, a4 = h[4] * y;
    const float px = ((a0 + a1) + h[2]) * ww, py = ((a3 + a4) + h[5]) * ww;
    const float dx = px - dst[2 * i], dy = py - dst[2 * i + 1];
    const float dx2 = dx * dx, dy2 = dy * dy;
    const int in = (dx2 + dy2) <= thr2;
    mask[i] = (uint8_t)in;
    c += in;
  }
  return c;
}

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

/* counts [npairs][nhyp]: inlier count of every hypothesis, -1 for a rejected sample */
int ransac_oracle_score(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs, float thresh,
                        int nhyp, uint64_t seed, int32_t* counts) {
  const float thr2 = thresh * thresh;
  for (int64_t p = 0; p < npairs; ++p) {
    const int64_t m0 = pair_offsets[p], K = pair_offsets[p + 1] - m0;
    for (int h = 0; h < nhyp; ++h) {
      double H[9];
      const int ok = ransac_oracle_hypothesis(src + 2 * m0, dst + 2 * m0, (uint32_t)K, seed, (uint64_t)p, (uint64_t)h, H);
      counts[p * nhyp + h] = ok ? count_inliers(H, src + 2 * m0, dst + 2 * m0, K, thr2, NULL) : -1;
    }
  }
  return 0;
}

/* ---- refinement: normalised linear least squares (h8 = 1) + Levenberg-Marquardt ------------------- */
static int chol_solve8(const double* Ain /*8x8 sym*/, const double* b, double lambda, double* x) {
  double A[8][8], L[8][8], y[8];
  memset(L, 0, sizeof(L));
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) A[i][j] = Ain[8 * i + j];
  for (int i = 0; i < 8; ++i) A[i][i] *= (1.0 + lambda);
  for (int j = 0; j < 8; ++j) {
    double s = A[j][j];
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    if (!(s > 0.0)) return 0;
    L[j][j] = sqrt(s);
    for (int i = j + 1; i < 8; ++i) {
      double t = A[i][j];
      for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
      L[i][j] = t / L[j][j];
    }
  }
  for (int i = 0; i < 8; ++i) {
    double t = b[i];
    for (int k = 0; k < i; ++k) t -= L[i][k] * y[k];
    y[i] = t / L[i][i];
  }
  for (int i = 7; i >= 0; --i) {
    double t = y[i];
    for (int k = i + 1; k < 8; ++k) t -= L[k][i] * x[k];
    x[i] = t / L[i][i];
  }
  return 1;
}

typedef struct { double csx, csy, cdx, cdy, ss, sd; } Norm;

static void rank1(double* JtJ, double* Jtr, const double* r, double b) {
  for (int i = 0; i < 8; ++i) {
    if (r[i] == 0.0) continue;
    for (int j = 0; j < 8; ++j) JtJ[8 * i + j] += r[i] * r[j];
    Jtr[i] += r[i] * b;
  }
}

static double reproj_cost(const double* h, const float* src, const float* dst, const uint8_t* mask, int64_t K,
                          const Norm* nm, double* JtJ, double* Jtr) {
  double cost = 0.0;
  if (JtJ) { memset(JtJ, 0, 64 * sizeof(double)); memset(Jtr, 0, 8 * sizeof(double)); }
  for (int64_t i = 0; i < K; ++i) {
    if (!mask[i]) continue;
    const double x = ((double)src[2 * i] - nm->csx) * nm->ss, y = ((double)src[2 * i + 1] - nm->csy) * nm->ss;
    const double u = ((double)dst[2 * i] - nm->cdx) * nm->sd, v = ((double)dst[2 * i + 1] - nm->cdy) * nm->sd;
    const double w = h[6] * x + h[7] * y + 1.0;
    const double iw = fabs(w) > DBL_EPSILON ? 1.0 / w : 0.0;
    const double pu = (h[0] * x + h[1] * y + h[2]) * iw, pv = (h[3] * x + h[4] * y + h[5]) * iw;
    const double ru = pu - u, rv = pv - v;
    cost += ru * ru + rv * rv;
    if (JtJ) {
      const double r1[8] = {x * iw, y * iw, iw, 0, 0, 0, -x * iw * pu, -y * iw * pu};
      const double r2[8] = {0, 0, 0, x * iw, y * iw, iw, -x * iw * pv, -y * iw * pv};
      rank1(JtJ, Jtr, r1, -ru);
      rank1(JtJ, Jtr, r2, -rv);
    }
  }
  return cost;
}

static void refine_on_inliers(const double* Hmin, const float* src, const float* dst, const uint8_t* mask, int64_t K,
                              int n_in, double* Hout) {
  Norm nm;
  double s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0;
  for (int64_t i = 0; i < K; ++i) {
    if (!mask[i]) continue;
    const double ax = src[2 * i], ay = src[2 * i + 1], bx = dst[2 * i], by = dst[2 * i + 1];
    s1 += ax; s2 += ay; s3 += ax * ax + ay * ay; s4 += bx; s5 += by; s6 += bx * bx + by * by;
  }
  const double n = (double)n_in;
  nm.csx = s1 / n; nm.csy = s2 / n; nm.cdx = s4 / n; nm.cdy = s5 / n;
  const double vs = fmax(s3 / n - (nm.csx * nm.csx + nm.csy * nm.csy), 1e-300);
  const double vd = fmax(s6 / n - (nm.cdx * nm.cdx + nm.cdy * nm.cdy), 1e-300);
  nm.ss = sqrt(2.0 / vs); nm.sd = sqrt(2.0 / vd);

  double JtJ[64], Jtr[8], h[8];
  memset(JtJ, 0, sizeof(JtJ)); memset(Jtr, 0, sizeof(Jtr));
  for (int64_t i = 0; i < K; ++i) {
    if (!mask[i]) continue;
    const double x = ((double)src[2 * i] - nm.csx) * nm.ss, y = ((double)src[2 * i + 1] - nm.csy) * nm.ss;
    const double u = ((double)dst[2 * i] - nm.cdx) * nm.sd, v = ((double)dst[2 * i + 1] - nm.cdy) * nm.sd;
    const double r1[8] = {x, y, 1, 0, 0, 0, -x * u, -y * u};
    const double r2[8] = {0, 0, 0, x, y, 1, -x * v, -y * v};
    rank1(JtJ, Jtr, r1, u);
    rank1(JtJ, Jtr, r2, v);
  }
  if (!chol_solve8(JtJ, Jtr, 0.0, h)) {
    /* minimal model moved into the normalised frames: Hn = Td * H * Ts^-1 */
    double G[9], R[9];
    for (int j = 0; j < 3; ++j) {
      G[j] = nm.sd * (Hmin[j] - nm.cdx * Hmin[6 + j]);
      G[3 + j] = nm.sd * (Hmin[3 + j] - nm.cdy * Hmin[6 + j]);
      G[6 + j] = Hmin[6 + j];
    }
    for (int i = 0; i < 3; ++i) {
      R[3 * i] = G[3 * i] / nm.ss; R[3 * i + 1] = G[3 * i + 1] / nm.ss;
      R[3 * i + 2] = G[3 * i] * nm.csx + G[3 * i + 1] * nm.csy + G[3 * i + 2];
    }
    for (int i = 0; i < 8; ++i) h[i] = R[i] / R[8];
  }
  double lambda = 1e-3;
  for (int iter = 0; iter < 30; ++iter) {
    const double cost = reproj_cost(h, src, dst, mask, K, &nm, JtJ, Jtr);
    int improved = 0;
    double trial[8];
    for (int attempt = 0; attempt < 12 && !improved; ++attempt) {
      double dx[8];
      const int ok = chol_solve8(JtJ, Jtr, lambda, dx);
      if (ok) {
        for (int i = 0; i < 8; ++i) trial[i] = h[i] + dx[i];
        if (reproj_cost(trial, src, dst, mask, K, &nm, NULL, NULL) <= cost) improved = 1;
      }
      if (improved) lambda = fmax(lambda * 0.1, 1e-15);
      else lambda *= 10.0;
    }
    if (!improved) break;
    double step = 0.0, scale = 0.0;
    for (int i = 0; i < 8; ++i) { step = fmax(step, fabs(trial[i] - h[i])); scale = fmax(scale, fabs(h[i])); }
    for (int i = 0; i < 8; ++i) h[i] = trial[i];
    if (step <= 1e-13 * (1.0 + scale)) break;
  }
  double Hn[9], G[9], R[9];
  for (int i = 0; i < 8; ++i) Hn[i] = h[i];
  Hn[8] = 1.0;
  for (int i = 0; i < 3; ++i) {
    G[3 * i] = Hn[3 * i] * nm.ss; G[3 * i + 1] = Hn[3 * i + 1] * nm.ss;
    G[3 * i + 2] = Hn[3 * i + 2] - nm.ss * (Hn[3 * i] * nm.csx + Hn[3 * i + 1] * nm.csy);
  }
  for (int j = 0; j < 3; ++j) {
    R[j] = G[j] / nm.sd + nm.cdx * G[6 + j];
    R[3 + j] = G[3 + j] / nm.sd + nm.cdy * G[6 + j];
    R[6 + j] = G[6 + j];
  }
  for (int i = 0; i < 9; ++i) Hout[i] = R[i] / R[8];
}

/* Refinement alone: least-squares + LM fit on the rows flagged in `mask` (what cv2 does after its RANSAC loop);
 * Hmin is only used when the linear system is singular. */
int ransac_oracle_refine(const float* src, const float* dst, const uint8_t* mask, int64_t K, const double* Hmin,
                         double* Hout) {
  int n_in = 0;
  for (int64_t i = 0; i < K; ++i) n_in += mask[i] ? 1 : 0;
  if (n_in < 4) return -1;
  refine_on_inliers(Hmin, src, dst, mask, K, n_in, Hout);
  return 0;
}

/* Full per-pair result, same contract as hpusm_ransac_homography_batch (include/hpusm.h). */
int ransac_oracle_homography(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs,
                             float thresh, int nhyp, double confidence, uint64_t seed, int refine, double* H,
                             uint8_t* mask, int32_t* ninl, int32_t* best_hyp) {
  const float thr2 = thresh * thresh;
  for (int64_t p = 0; p < npairs; ++p) {
    const int64_t m0 = pair_offsets[p], K = pair_offsets[p + 1] - m0;
    const float* s = src + 2 * m0;
    const float* d = dst + 2 * m0;
    int best = -1, best_c = 3, niters = nhyp;
    double Hbest[9];
    for (int h = 0; h < nhyp && h < niters; ++h) {
      double Hh[9];
      if (!ransac_oracle_hypothesis(s, d, (uint32_t)K, seed, (uint64_t)p, (uint64_t)h, Hh)) continue;
      const int c = count_inliers(Hh, s, d, K, thr2, NULL);
      if (c > best_c) {
        best_c = c; best = h;
        memcpy(Hbest, Hh, sizeof(Hbest));
        if (confidence > 0.0) niters = adaptive_iters(confidence, (double)(K - c) / (double)K, nhyp);
      }
    }
    if (best_hyp) best_hyp[p] = best;
    if (best < 0) {
      for (int i = 0; i < 9; ++i) H[9 * p + i] = NAN;
      memset(mask + m0, 0, (size_t)K);
      ninl[p] = 0;
      continue;
    }
    const int n_in = count_inliers(Hbest, s, d, K, thr2, mask + m0);
    ninl[p] = n_in;
    if (!refine || n_in < 5) {
      memcpy(H + 9 * p, Hbest, sizeof(Hbest));
    } else {
      refine_on_inliers(Hbest, s, d, mask + m0, K, n_in, H + 9 * p);
      ninl[p] = final_mask(H + 9 * p, s, d, K, thr2, mask + m0);
    }
  }
  return 0;
}
