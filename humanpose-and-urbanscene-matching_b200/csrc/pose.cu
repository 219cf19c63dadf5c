// K4: pose scoring -- batched affine least squares, Procrustes, and the single-person decision.
//
// Replaces, in one launch over a whole pose database, the per-pair Python of the reference:
//   common.find_transformation (common.py:37-101, np.linalg.lstsq :84)
//   common.feature_scaling (common.py:688-708), handle_undetected_points (common.py:281-348),
//   split_in_face_legs_torso_v2 / unsplit (common.py:269-279)
//   pose_comparison.* (posematching/pose_comparison.py:8-81)
//   single_person.match_single (posematching/single_person.py:80-189)
//   procrustes.procrustes (posematching/procrustes.py:150-259, np.linalg.svd :218)
//
// The reference works in float64 and its thresholds sit at 3 significant digits, so the arithmetic here
// is FP64 as well (B200 keeps a full-rate FP64 pipe; the path is a few hundred FMAs per 144-byte pose).
// One thread owns one pose pair: the query pose sits in shared memory (broadcast reads), the database
// pose in registers as the raw float32 it was stored as; masking and min/max scaling are recomputed on
// the fly per body part instead of materialising 72 doubles per thread.
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include "common.cuh"
#include "pose_fast.cuh"

namespace hpusm {

constexpr int NJ = HPUSM_POSE_JOINTS;  // 18 COCO joints

// ---- 3x3 symmetric eigen (cyclic Jacobi), only used on the rank-deficient path -------------------
__device__ __noinline__ void jacobi3(double M[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(M[0][1]) + fabs(M[0][2]) + fabs(M[1][2]);
    if (off <= 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(M[p][q]) <= 1e-300) continue;
        const double theta = (M[q][q] - M[p][p]) / (2.0 * M[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double mkp = M[k][p], mkq = M[k][q];
          M[k][p] = c * mkp - s * mkq; M[k][q] = s * mkp + c * mkq;
        }
        for (int k = 0; k < 3; ++k) {
          const double mpk = M[p][k], mqk = M[q][k];
          M[p][k] = c * mpk - s * mqk; M[q][k] = s * mpk + c * mqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) w[i] = M[i][i];
}

// Raw moments of one least-squares problem  [x y 1] * A = [X Y 1]  over the detected rows.
struct LsqSums {
  double n, sx, sy, sxx, sxy, syy;  // moments of the input points
  double sX, sY;                    // sums of the model points
  double sxX, syX, sxY, syY;        // cross moments
};

__device__ __forceinline__ void lsq_add(LsqSums& m, double x, double y, double X, double Y) {
  m.n += 1.0; m.sx += x; m.sy += y; m.sxx += x * x; m.sxy += x * y; m.syy += y * y;
  m.sX += X; m.sY += Y; m.sxX += x * X; m.syX += y * X; m.sxY += x * Y; m.syY += y * Y;
}

// Minimum-norm solution via the pseudo-inverse of the 3x3 normal matrix (what an SVD-based lstsq
// returns when the design matrix has rank < 3: fewer than three detected joints, or collinear joints).
// Takes the sums BY VALUE and fills a caller-local scratch array: nothing of the hot path has its address taken, so
// the moments and the 3x3 result stay in registers for the (overwhelmingly common) full-rank problems.
__device__ __noinline__ void lsq_solve_deficient(LsqSums m, double* A) {
  double M[3][3] = {{m.sxx, m.sxy, m.sx}, {m.sxy, m.syy, m.sy}, {m.sx, m.sy, m.n}};
  const double R[3][3] = {{m.sxX, m.sxY, m.sx}, {m.syX, m.syY, m.sy}, {m.sX, m.sY, m.n}};  // X^T [X Y 1]
  double V[3][3], w[3];
  jacobi3(M, V, w);
  const double wmax = fmax(fmax(fabs(w[0]), fabs(w[1])), fabs(w[2]));
  const double tol = wmax * 1e-12;
  for (int c = 0; c < 3; ++c) {
    double coef[3];
    for (int e = 0; e < 3; ++e) {
      double dot = V[0][e] * R[0][c] + V[1][e] * R[1][c] + V[2][e] * R[2][c];
      coef[e] = (w[e] > tol) ? dot / w[e] : 0.0;
    }
    for (int r = 0; r < 3; ++r) A[3 * r + c] = V[r][0] * coef[0] + V[r][1] * coef[1] + V[r][2] * coef[2];
  }
}

// Solve for A (row-major 3x3, row-vector convention). Full-rank problems go through the centred 2x2
// system (better conditioned than the raw normal equations); the rest through the pseudo-inverse.
__device__ __forceinline__ void lsq_solve(const LsqSums& m, double A[9]) {
  const double inv_n = 1.0 / m.n;
  const double mx = m.sx * inv_n, my = m.sy * inv_n, mX = m.sX * inv_n, mY = m.sY * inv_n;
  const double cxx = m.sxx - m.sx * mx, cxy = m.sxy - m.sx * my, cyy = m.syy - m.sy * my;
  const double det = cxx * cyy - cxy * cxy;
  const double tr = cxx + cyy;
  if (m.n >= 2.5 && det > 1e-12 * tr * tr) {
    const double cxX = m.sxX - m.sx * mX, cyX = m.syX - m.sy * mX;
    const double cxY = m.sxY - m.sx * mY, cyY = m.syY - m.sy * mY;
    const double id = 1.0 / det;
    const double a00 = (cyy * cxX - cxy * cyX) * id, a10 = (cxx * cyX - cxy * cxX) * id;
    const double a01 = (cyy * cxY - cxy * cyY) * id, a11 = (cxx * cyY - cxy * cxY) * id;
    A[0] = a00; A[1] = a01; A[2] = 0.0;
    A[3] = a10; A[4] = a11; A[5] = 0.0;
    A[6] = mX - mx * a00 - my * a10; A[7] = mY - mx * a01 - my * a11; A[8] = 1.0;
  } else {
    double Ad[9];
    lsq_solve_deficient(m, Ad);
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = Ad[i];
  }
}

__device__ __forceinline__ double zero_small(double a) { return fabs(a) < 1e-10 ? 0.0 : a; }

// ---- find_transformation, batched (FP64 in/out, arbitrary point count) ----------------------------
__global__ void affine_lsq_batch_kernel(const double* __restrict__ model, const double* __restrict__ input, int64_t n,
                                        int npts, double* __restrict__ out, double* __restrict__ Aout,
                                        uint8_t* __restrict__ has_A) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const double* M = model + b * npts * 2;
  const double* I = input + b * npts * 2;
  double* O = out + b * npts * 2;
  LsqSums m = {};
  for (int j = 0; j < npts; ++j) {
    const double x = I[2 * j], y = I[2 * j + 1];
    if (!(x == 0.0 && y == 0.0)) lsq_add(m, x, y, M[2 * j], M[2 * j + 1]);
  }
  if (m.n < 0.5) {  // common.py:70-72: nothing to fit, hand the input back
    for (int j = 0; j < 2 * npts; ++j) O[j] = I[j];
    for (int j = 0; j < 9; ++j) Aout[b * 9 + j] = 0.0;
    has_A[b] = 0;
    return;
  }
  double A[9];
  lsq_solve(m, A);
  for (int j = 0; j < npts; ++j) {
    const double x = I[2 * j], y = I[2 * j + 1];
    if (x == 0.0 && y == 0.0) { O[2 * j] = 0.0; O[2 * j + 1] = 0.0; }
    else { O[2 * j] = x * A[0] + y * A[3] + A[6]; O[2 * j + 1] = x * A[1] + y * A[4] + A[7]; }
  }
  for (int j = 0; j < 9; ++j) Aout[b * 9 + j] = zero_small(A[j]);
  has_A[b] = 1;
}

// ---- Procrustes, batched ---------------------------------------------------------------------------
__global__ void procrustes_batch_kernel(const double* __restrict__ X, const double* __restrict__ Y, int64_t n, int npts,
                                        int scaling, int reflection, double* __restrict__ dout, double* __restrict__ Z,
                                        double* __restrict__ tform) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const double* x = X + b * npts * 2;
  const double* y = Y + b * npts * 2;
  double mux = 0, muy = 0, mvx = 0, mvy = 0;
  for (int j = 0; j < npts; ++j) { mux += x[2 * j]; muy += x[2 * j + 1]; mvx += y[2 * j]; mvy += y[2 * j + 1]; }
  mux /= npts; muy /= npts; mvx /= npts; mvy /= npts;
  double ssX = 0, ssY = 0;
  for (int j = 0; j < npts; ++j) {
    const double ax = x[2 * j] - mux, ay = x[2 * j + 1] - muy, bx = y[2 * j] - mvx, by = y[2 * j + 1] - mvy;
    ssX += ax * ax + ay * ay; ssY += bx * bx + by * by;
  }
  const double normX = sqrt(ssX), normY = sqrt(ssY);
  // A = X0^T Y0 on the unit-norm point sets
  double a = 0, bb = 0, c = 0, d = 0;
  for (int j = 0; j < npts; ++j) {
    const double ax = (x[2 * j] - mux) / normX, ay = (x[2 * j + 1] - muy) / normX;
    const double bx = (y[2 * j] - mvx) / normY, by = (y[2 * j + 1] - mvy) / normY;
    a += ax * bx; bb += ax * by; c += ay * bx; d += ay * by;
  }
  // closed-form 2x2 SVD: A = rotation part + reflection part
  const double E = 0.5 * (a + d), F = 0.5 * (a - d), G = 0.5 * (c + bb), Hh = 0.5 * (c - bb);
  const double Q1 = sqrt(E * E + Hh * Hh), R1 = sqrt(F * F + G * G);
  bool use_reflection;
  if (reflection == 0) use_reflection = R1 > Q1;          // 'best': whichever fits better (det(A) < 0)
  else use_reflection = (reflection == 1);
  // T = V U^T = (orthogonal polar factor of A)^T
  double T00, T01, T10, T11, trace;
  if (!use_reflection) {
    if (Q1 > 0) { T00 = E / Q1; T01 = Hh / Q1; T10 = -Hh / Q1; T11 = E / Q1; }
    else { T00 = 1; T01 = 0; T10 = 0; T11 = 1; }
    trace = 2.0 * Q1;
  } else {
    if (R1 > 0) { T00 = F / R1; T01 = G / R1; T10 = G / R1; T11 = -F / R1; }
    else { T00 = 1; T01 = 0; T10 = 0; T11 = -1; }
    trace = 2.0 * R1;
  }
  if (reflection != 0) {
    // forcing the non-optimal branch flips the sign of the smallest singular value
    const bool best_is_reflection = R1 > Q1;
    if (use_reflection != best_is_reflection) trace = 2.0 * (use_reflection ? R1 : Q1);
  }
  double bscale, dval, zs;
  if (scaling) {
    bscale = trace * normX / normY;
    dval = 1.0 - trace * trace;
    zs = normX * trace;
  } else {
    bscale = 1.0;
    dval = 1.0 + ssY / ssX - 2.0 * trace * normY / normX;
    zs = normY;
  }
  double* z = Z + b * npts * 2;
  for (int j = 0; j < npts; ++j) {
    const double bx = (y[2 * j] - mvx) / normY, by = (y[2 * j + 1] - mvy) / normY;
    z[2 * j] = zs * (bx * T00 + by * T10) + mux;
    z[2 * j + 1] = zs * (bx * T01 + by * T11) + muy;
  }
  dout[b] = dval;
  double* tf = tform + b * 7;
  tf[0] = T00; tf[1] = T01; tf[2] = T10; tf[3] = T11; tf[4] = bscale;
  tf[5] = mux - bscale * (mvx * T00 + mvy * T10);
  tf[6] = muy - bscale * (mvx * T01 + mvy * T11);
}

// ---- match_single, batched ---------------------------------------------------------------------------
// Body parts (common.py:269-274); constexpr so that every joint index is a compile-time constant after unrolling
// (the pose lives in registers: a run-time index would push it to local memory).
struct PartTorso {
  static constexpr int N = 10;
  __host__ __device__ static constexpr int at(int k) { return k < 9 ? k : 11; }
};
struct PartLegs {
  static constexpr int N = 7;
  __host__ __device__ static constexpr int at(int k) { return k == 0 ? 1 : 7 + k; }
};
struct PartFace {
  static constexpr int N = 5;
  __host__ __device__ static constexpr int at(int k) { return k == 0 ? 0 : 13 + k; }
};

struct PoseParams {
  double d_torso, tan_torso, rot_torso, d_legs, tan_legs, rot_legs, d_shoulder;
  int torso_use_tan, legs_use_tan;
};

// min/max scaling of one masked pose (common.feature_scaling, including its quirk that each minimum runs over BOTH
// coordinates of the rows whose x (resp. y) is non-zero).  The extrema are found in the storage type T (comparisons
// only, exact); a scaled coordinate is  ((double)v - lo) * inv  so that v == lo gives an exact 0, as in the reference
// (the part gates count exact zeros).  A literal 0 coordinate (undetected joint) scales to max(0, -lo*inv): 0 unless
// the pose has negative coordinates (`general`).
__device__ __forceinline__ float tmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float tmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double tmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ double tmin(double a, double b) { return fmin(a, b); }

template <typename T>
struct Scale {
  T xlo, ylo;
  double ix, iy, zx, zy;
  bool valid, general;
};

template <typename T, typename Get>
__device__ __forceinline__ Scale<T> make_scale(Get get, unsigned undetected) {
  T xmax = -INFINITY, ymax = -INFINITY, xmin = INFINITY, ymin = INFINITY;
  bool anyx = false, anyy = false;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const bool u = (undetected >> j) & 1u;
    const T x = u ? T(0) : get(2 * j), y = u ? T(0) : get(2 * j + 1);
    xmax = tmax(xmax, x); ymax = tmax(ymax, y);
    const T lo = tmin(x, y);
    // rows with x == 0 (resp. y == 0) do not take part in the minimum: feed +inf instead of branching
    xmin = tmin(xmin, x != T(0) ? lo : T(INFINITY));
    ymin = tmin(ymin, y != T(0) ? lo : T(INFINITY));
    anyx |= x != T(0); anyy |= y != T(0);
  }
  Scale<T> s;
  // The reference raises here instead of returning: ValueError (np.min of an empty array) when no joint is
  // detected, LinAlgError (NaN/inf into LAPACK) when xmax == xmin or ymax == ymin.  Such poses never match.
  s.valid = anyx && anyy && xmax != xmin && ymax != ymin;
  s.xlo = xmin; s.ylo = ymin;
  s.ix = 1.0 / ((double)xmax - (double)xmin); s.iy = 1.0 / ((double)ymax - (double)ymin);
  const double cx = (0.0 - (double)xmin) * s.ix, cy = (0.0 - (double)ymin) * s.iy;
  s.zx = cx < 0.0 ? 0.0 : cx; s.zy = cy < 0.0 ? 0.0 : cy;
  s.general = !(s.zx == 0.0 && s.zy == 0.0);
  return s;
}

template <typename T, bool GENERAL>
__device__ __forceinline__ double scale_coord(T raw, bool undetected, T lo, double inv, double z) {
  const bool is0 = undetected || raw == T(0);
  const T m = is0 ? lo : raw;                       // (lo - lo) * inv == exact 0
  double r = ((double)m - (double)lo) * inv;
  if (GENERAL) r = is0 ? z : (r < 0.0 ? 0.0 : r);   // poses with negative coordinates: the reference's clamp
  return r;
}

struct PartResult {
  double max_d2, shoulder_d2, A[9];  // squared distances: sqrt is monotone, one sqrt per part at the end
  int nz_model, nz_input;
  bool has_A;
};

// One body part: scale, fit, transform, distances.  model(j, x, y) / input(j, x, y) return SCALED coordinates of
// joint j; `out` (optional) receives the transformed input rows [NP][2].
template <typename Rows, typename Model, typename Input>
__device__ __forceinline__ PartResult eval_part(Model model, Input input, double* out) {
  constexpr int NP = Rows::N;
  PartResult r;
  double xm[NP], ym[NP], xi[NP], yi[NP];
  LsqSums m = {};
  r.nz_model = 0; r.nz_input = 0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    model(Rows::at(k), xm[k], ym[k]);
    input(Rows::at(k), xi[k], yi[k]);
    r.nz_model += (xm[k] != 0.0) + (ym[k] != 0.0);
    r.nz_input += (xi[k] != 0.0) + (yi[k] != 0.0);
    if (!(xi[k] == 0.0 && yi[k] == 0.0)) lsq_add(m, xi[k], yi[k], xm[k], ym[k]);
  }
  r.has_A = m.n > 0.5;
  if (r.has_A) lsq_solve(m, r.A);
  else {
#pragma unroll
    for (int i = 0; i < 9; ++i) r.A[i] = 0.0;
  }
  r.max_d2 = -1.0; r.shoulder_d2 = -1.0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    double tx, ty;
    if (!r.has_A) { tx = xi[k]; ty = yi[k]; }
    else if (xi[k] == 0.0 && yi[k] == 0.0) { tx = 0.0; ty = 0.0; }
    else { tx = xi[k] * r.A[0] + yi[k] * r.A[3] + r.A[6]; ty = xi[k] * r.A[1] + yi[k] * r.A[4] + r.A[7]; }
    if (out) { out[2 * k] = tx; out[2 * k + 1] = ty; }
    const double dx = xm[k] - tx, dy = ym[k] - ty;
    const double d2 = dx * dx + dy * dy;
    // Python's max() keeps the first element unless a later one compares greater: a NaN in first position poisons
    // the result, later NaNs are skipped.  The same holds for these comparisons on the squared values.
    if (k == 0) r.max_d2 = d2;
    else if (d2 > r.max_d2) r.max_d2 = d2;
    if (k == 1) r.shoulder_d2 = d2;
    if (k == 4 && d2 > r.shoulder_d2) r.shoulder_d2 = d2;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) r.A[i] = zero_small(r.A[i]);
  return r;
}

// rot = max(|atan2(-A01, A00)|, |atan2(A10, A11)|) * 57.3 <= thr, without the atan2 when thr < ~86 deg.
__device__ __forceinline__ bool rotation_ok(const PartResult& r, double tan_thr, double thr, int use_tan) {
  if (!r.has_A) return true;  // rot_max = 0 for an empty matrix; thresholds are non-negative
  if (use_tan) {
    const bool ok1 = r.A[0] >= 0.0 && fabs(r.A[1]) <= tan_thr * r.A[0];
    const bool ok2 = r.A[4] >= 0.0 && fabs(r.A[3]) <= tan_thr * r.A[4];
    return ok1 && ok2;
  }
  const double r1 = fabs(atan2(-r.A[1], r.A[0]) * 57.3), r2 = fabs(atan2(r.A[3], r.A[4]) * 57.3);
  return fmax(r2, r1) <= thr;
}

__device__ __noinline__ double rotation_value(const PartResult& r) {
  if (!r.has_A) return 0.0;
  const double r1 = fabs(atan2(-r.A[1], r.A[0]) * 57.3), r2 = fabs(atan2(r.A[3], r.A[4]) * 57.3);
  return (r1 > r2) ? r1 : r2;  // Python max(rotation_2, rotation_1): the first argument wins unless the second is greater
}

// common.feature_scaling on n independent point sets (FP64, arbitrary point count), one thread each.
__global__ void feature_scaling_batch_kernel(const double* __restrict__ pts, int64_t n, int npts, double* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const double* P = pts + b * npts * 2;
  double* O = out + b * npts * 2;
  double xmax = -DBL_MAX, ymax = -DBL_MAX, xmin = DBL_MAX, ymin = DBL_MAX;
  for (int j = 0; j < npts; ++j) {
    const double x = P[2 * j], y = P[2 * j + 1];
    xmax = fmax(xmax, x); ymax = fmax(ymax, y);
    const double lo = fmin(x, y);
    if (x != 0.0) xmin = fmin(xmin, lo);
    if (y != 0.0) ymin = fmin(ymin, lo);
  }
  for (int j = 0; j < npts; ++j) {
    const double sx = (P[2 * j] - xmin) / (xmax - xmin), sy = (P[2 * j + 1] - ymin) / (ymax - ymin);
    O[2 * j] = sx < 0.0 ? 0.0 : sx;
    O[2 * j + 1] = sy < 0.0 ? 0.0 : sy;
  }
}

// Decision for one pose pair (single_person.py:100-189) given the raw poses, the joint mask and both scalings.
template <typename T, bool GENERAL, bool DETAIL>
__device__ __forceinline__ void pose_decide(const T* sq /*shared*/, const T (&pr)[NJ * 2], unsigned und, const Scale<T>& s_q,
                                            const Scale<T>& s_p, int query_is_model, const PoseParams& prm,
                                            uint8_t* match, float* score, double* det) {
  auto q_scaled = [&](int j, double& x, double& y) {
    const bool u = (und >> j) & 1u;
    x = scale_coord<T, GENERAL>(sq[2 * j], u, s_q.xlo, s_q.ix, s_q.zx);
    y = scale_coord<T, GENERAL>(sq[2 * j + 1], u, s_q.ylo, s_q.iy, s_q.zy);
  };
  auto p_scaled = [&](int j, double& x, double& y) {
    const bool u = (und >> j) & 1u;
    x = scale_coord<T, GENERAL>(pr[2 * j], u, s_p.xlo, s_p.ix, s_p.zx);
    y = scale_coord<T, GENERAL>(pr[2 * j + 1], u, s_p.ylo, s_p.iy, s_p.zy);
  };
  // detail layout: [0..7] scalars, [8..34] A face/torso/legs, [35..78] unsplit (common.py:276) transformed input:
  // face[0] | torso (10) | legs (7) | face[1:5]
  PartResult torso, legs;
  if (query_is_model) {
    torso = eval_part<PartTorso>(q_scaled, p_scaled, DETAIL ? det + 35 + 2 : nullptr);
    legs = eval_part<PartLegs>(q_scaled, p_scaled, DETAIL ? det + 35 + 22 : nullptr);
  } else {
    torso = eval_part<PartTorso>(p_scaled, q_scaled, DETAIL ? det + 35 + 2 : nullptr);
    legs = eval_part<PartLegs>(p_scaled, q_scaled, DETAIL ? det + 35 + 22 : nullptr);
  }
  const double torso_dist = sqrt(torso.max_d2), legs_dist = sqrt(legs.max_d2), shoulder_dist = sqrt(torso.shoulder_d2);

  // single_person.py:150-178
  bool ok_torso, ok_legs;
  if (torso.nz_model > 4) {
    if (torso.nz_model - torso.nz_input < 2)
      ok_torso = torso_dist <= prm.d_torso && rotation_ok(torso, prm.tan_torso, prm.rot_torso, prm.torso_use_tan) &&
                 shoulder_dist <= prm.d_shoulder;
    else ok_torso = false;
  } else ok_torso = true;
  if (legs.nz_model > 8) {
    if (legs.nz_model - legs.nz_input < 2)
      ok_legs = legs_dist <= prm.d_legs && rotation_ok(legs, prm.tan_legs, prm.rot_legs, prm.legs_use_tan);
    else ok_legs = false;
  } else ok_legs = true;
  const bool valid = s_q.valid && s_p.valid;
  const double err = valid ? (torso_dist + legs_dist) * 0.5 : nan("");
  *match = (valid && ok_torso && ok_legs) ? 1 : 0;
  *score = (float)err;

  if (DETAIL) {
    double tf_face[10];
    PartResult face = query_is_model ? eval_part<PartFace>(q_scaled, p_scaled, tf_face)
                                     : eval_part<PartFace>(p_scaled, q_scaled, tf_face);
    det[0] = torso_dist; det[1] = legs_dist; det[2] = sqrt(face.max_d2); det[3] = shoulder_dist;
    det[4] = rotation_value(torso); det[5] = rotation_value(legs);
    det[6] = (double)(torso.has_A ? 1 : 0) + 2.0 * (legs.has_A ? 1 : 0) + 4.0 * (face.has_A ? 1 : 0);
    det[7] = err;
    for (int i = 0; i < 9; ++i) { det[8 + i] = face.A[i]; det[17 + i] = torso.A[i]; det[26 + i] = legs.A[i]; }
    double* o = det + 35;
    o[0] = tf_face[0]; o[1] = tf_face[1];
    for (int i = 0; i < 8; ++i) o[36 + i] = tf_face[2 + i];
  }
}

// Load one pose row into registers (float: 9 x LDG.128; double: 18 x LDG.128).
template <typename T>
__device__ __forceinline__ void load_pose(const T* __restrict__ row, T (&pr)[NJ * 2]) {
  if constexpr (sizeof(T) == 4) {
    const float4* src = reinterpret_cast<const float4*>(row);
#pragma unroll
    for (int v = 0; v < NJ * 2 / 4; ++v) {
      const float4 x = __ldg(src + v);
      pr[4 * v] = x.x; pr[4 * v + 1] = x.y; pr[4 * v + 2] = x.z; pr[4 * v + 3] = x.w;
    }
  } else {
    const double2* src = reinterpret_cast<const double2*>(row);
#pragma unroll
    for (int v = 0; v < NJ; ++v) {
      const double2 x = __ldg(src + v);
      pr[2 * v] = x.x; pr[2 * v + 1] = x.y;
    }
  }
}

// Out-of-line variant for the rare poses with negative coordinates and for callers that want the per-part detail.
// It re-reads the pose itself so that nothing of the fast path has its address taken (which would push the pose
// registers of EVERY thread into local memory).
template <typename T>
__device__ __noinline__ void pose_decide_slow(const T* sq, const T* __restrict__ row, unsigned und, int query_is_model,
                                              const PoseParams* prm, uint8_t* match, float* score, double* det) {
  T pr[NJ * 2];
  load_pose<T>(row, pr);
  const Scale<T> s_q = make_scale<T>([&](int i) { return sq[i]; }, und);
  const Scale<T> s_p = make_scale<T>([&](int i) { return pr[i]; }, und);
  double scratch[HPUSM_POSE_DETAIL];
  pose_decide<T, true, true>(sq, pr, und, s_q, s_p, query_is_model, *prm, match, score, det ? det : scratch);
}

constexpr int PM_THREADS = 128;

#ifndef HPUSM_POSE_MIN_BLOCKS
#define HPUSM_POSE_MIN_BLOCKS 4  // 128 registers: measured 1.39 ms per 10 M poses vs 1.59 ms at 2 blocks (233 registers)
#endif
// One database pose through the exact FP64 decision.
template <typename T>
__device__ __forceinline__ void pose_exact_one(const T* sq, unsigned sq_undetected, const T* __restrict__ db, int64_t p,
                                               int query_is_model, const PoseParams& prm, uint8_t* __restrict__ match,
                                               float* __restrict__ score, double* __restrict__ detail) {
  T pr[NJ * 2];
  load_pose<T>(db + p * (NJ * 2), pr);
  // common.handle_undetected_points: a joint missing on either side is (0,0) on both
  unsigned und = sq_undetected;
#pragma unroll
  for (int j = 0; j < NJ; ++j)
    if (pr[2 * j] == T(0) && pr[2 * j + 1] == T(0)) und |= 1u << j;

  const Scale<T> s_q = make_scale<T>([&](int i) { return sq[i]; }, und);
  const Scale<T> s_p = make_scale<T>([&](int i) { return pr[i]; }, und);
  double* det = detail ? detail + p * HPUSM_POSE_DETAIL : nullptr;
  if (s_q.general || s_p.general || det)  // negative coordinates / per-part detail requested: the roomy variant
    pose_decide_slow<T>(sq, db + p * (NJ * 2), und, query_is_model, &prm, match + p, score + p, det);
  else
    pose_decide<T, false, false>(sq, pr, und, s_q, s_p, query_is_model, prm, match + p, score + p, nullptr);
}

// Exact FP64 decision for every database pose (list == nullptr: thread = pose) or for the poses the FP32 fast path
// could not certify (list / list_count: grid-stride over the list).
template <typename T>
__global__ void __launch_bounds__(PM_THREADS, HPUSM_POSE_MIN_BLOCKS)
pose_match_batch_kernel(const T* __restrict__ query, const T* __restrict__ db, int64_t n, int query_is_model,
                        PoseParams prm, uint8_t* __restrict__ match, float* __restrict__ score,
                        double* __restrict__ detail, const int32_t* __restrict__ list, const int32_t* __restrict__ list_count) {
  __shared__ T sq[NJ * 2];
  __shared__ unsigned sq_undetected;
  if (threadIdx.x < NJ * 2) sq[threadIdx.x] = query[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned u = 0;
    for (int j = 0; j < NJ; ++j) if (sq[2 * j] == T(0) && sq[2 * j + 1] == T(0)) u |= 1u << j;
    sq_undetected = u;
  }
  __syncthreads();
  if (list) {
    const int64_t cnt = *list_count;
    for (int64_t i = (int64_t)blockIdx.x * PM_THREADS + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * PM_THREADS)
      pose_exact_one<T>(sq, sq_undetected, db, (int64_t)list[i], query_is_model, prm, match, score, detail);
    return;
  }
  const int64_t p = (int64_t)blockIdx.x * PM_THREADS + threadIdx.x;
  if (p >= n) return;
  pose_exact_one<T>(sq, sq_undetected, db, p, query_is_model, prm, match, score, detail);
}

// ---- FP32 fast path (pose_fast.cuh) ----------------------------------------------------------------------------------
__global__ void pose_query_prep_kernel(const float* __restrict__ query, PoseQueryPrep* __restrict__ prep) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    PoseQueryPrep pq;
    pose_query_prep(query, pq);
    *prep = pq;
  }
}

constexpr int PF_THREADS = 128;
#ifndef PF_MIN_BLOCKS
#define PF_MIN_BLOCKS 5  // 96 registers: measured 0.49 ms per 10 M poses vs 0.55 ms at 4 blocks (120 registers) and 0.57 ms at 6 (80, spills)
#endif

// thread = database pose.  A pose the FP32 evaluation cannot certify goes to `list` (one atomic per warp) for the exact kernel.
template <bool QIM>
__global__ void __launch_bounds__(PF_THREADS, PF_MIN_BLOCKS)
pose_match_fast_kernel(const PoseQueryPrep* __restrict__ prep, const float* __restrict__ db, int64_t n, PoseFastParams prm,
                       uint8_t* __restrict__ match, float* __restrict__ score, int32_t* __restrict__ list,
                       int32_t* __restrict__ list_count) {
  __shared__ PoseQueryPrep pq;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(prep);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&pq);
    for (int i = threadIdx.x; i < (int)(sizeof(PoseQueryPrep) / 4); i += PF_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * PF_THREADS + threadIdx.x;
  bool defer = false;
  if (p < n) {
    float pr[NJ * 2], px[NJ], py[NJ];
    load_pose<float>(db + p * (NJ * 2), pr);
#pragma unroll
    for (int j = 0; j < NJ; ++j) { px[j] = pr[2 * j]; py[j] = pr[2 * j + 1]; }
    uint8_t m = 0;
    float sc = 0.f;
    const bool certain = pq.usable && pose_fast_decide<float, QIM>(pq, px, py, prm, PF_DIST_BAND, PF_ROT_BAND, m, sc);
    if (certain) { match[p] = m; score[p] = sc; }
    defer = !certain;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, defer);
  if (bal) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(list_count, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (defer) list[base + __popc(bal & ((1u << lane) - 1u))] = (int32_t)p;
  }
}

// Arbitrary (model, input) pose pairs in one launch: the candidate stage of multi_person.match
// (multi_person.py:169-207 calls match_single for every model person x input person).
template <typename T>
__global__ void __launch_bounds__(PM_THREADS)
pose_match_pairs_kernel(const T* __restrict__ model, const T* __restrict__ input, int64_t n, PoseParams prm,
                        uint8_t* __restrict__ match, float* __restrict__ score, double* __restrict__ detail) {
  const int64_t p = (int64_t)blockIdx.x * PM_THREADS + threadIdx.x;
  if (p >= n) return;
  T mo[NJ * 2], pr[NJ * 2];
  load_pose<T>(model + p * (NJ * 2), mo);
  load_pose<T>(input + p * (NJ * 2), pr);
  unsigned und = 0;
#pragma unroll
  for (int j = 0; j < NJ; ++j)
    if ((mo[2 * j] == T(0) && mo[2 * j + 1] == T(0)) || (pr[2 * j] == T(0) && pr[2 * j + 1] == T(0))) und |= 1u << j;
  const Scale<T> s_q = make_scale<T>([&](int i) { return mo[i]; }, und);
  const Scale<T> s_p = make_scale<T>([&](int i) { return pr[i]; }, und);
  double* det = detail ? detail + p * HPUSM_POSE_DETAIL : nullptr;
  if (s_q.general || s_p.general) {
    if (det) pose_decide<T, true, true>(mo, pr, und, s_q, s_p, 1, prm, match + p, score + p, det);
    else pose_decide<T, true, false>(mo, pr, und, s_q, s_p, 1, prm, match + p, score + p, nullptr);
  } else {
    if (det) pose_decide<T, false, true>(mo, pr, und, s_q, s_p, 1, prm, match + p, score + p, det);
    else pose_decide<T, false, false>(mo, pr, und, s_q, s_p, 1, prm, match + p, score + p, nullptr);
  }
}

// ---- scene score: perspective correction + affine fit of pose and random background matches -------------------
// Batched urban_scene.match_scene_multi steps 3.1-5 (urban_scene.py:80-145): transformation.perspective_correction
// (transformation.py:22-53: cv2.perspectiveTransform of the inlier matches and the input pose with H2) followed by
// transformation.affine_multi (transformation.py:101-173,:413: affine least squares over the pose and
// AMOUNT_BACKGROUND_FEATURES = 5 randomly chosen inlier matches, normalisation by the model image size, maximum
// distance).  The reference draws the 5 matches with Python's unseeded random.sample; here they come from a
// counter-based table keyed on (seed, pair) -- or from the caller (`picks`).
__device__ __forceinline__ uint64_t scene_mix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// cv2.perspectiveTransform on a CV_32FC2 point with a CV_64F matrix: double accumulate, float32 result
__device__ __forceinline__ void persp32(const double* H, double x, double y, double& ox, double& oy) {
  double w = x * H[6] + y * H[7] + H[8];
  if (fabs(w) > 2.220446049250313e-16) {
    w = 1.0 / w;
    ox = (double)(float)((x * H[0] + y * H[1] + H[2]) * w);
    oy = (double)(float)((x * H[3] + y * H[4] + H[5]) * w);
  } else { ox = 0.0; oy = 0.0; }
}

constexpr int SCENE_BG = 5;  // thresholds.AMOUNT_BACKGROUND_FEATURES

// One thread per image pair.
__global__ void scene_score_batch_kernel(const float* __restrict__ src, const float* __restrict__ dst,
                                         const int64_t* __restrict__ pair_offsets, const uint8_t* __restrict__ mask,
                                         const double* __restrict__ H2, const double* __restrict__ model_pose,
                                         const double* __restrict__ input_pose, int64_t npairs, int model_pose_stride,
                                         int njoints, double model_w, double model_h, uint64_t seed,
                                         const int32_t* __restrict__ picks, double* __restrict__ score,
                                         int32_t* __restrict__ picks_out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const int64_t m0 = pair_offsets[p], K = pair_offsets[p + 1] - m0;
  const double* H = H2 + p * 9;
  int n_in = 0;
  for (int64_t i = 0; i < K; ++i) n_in += mask[m0 + i] ? 1 : 0;
  if (!(H[0] == H[0])) { score[p] = INFINITY; return; }             // no homography (urban_scene.py:45-47)
  if (n_in < SCENE_BG) { score[p] = 1.0; return; }                  // transformation.py:107-108
  // the five background matches: ordinals among the inliers
  int pick[SCENE_BG];
  if (picks) {
#pragma unroll
    for (int j = 0; j < SCENE_BG; ++j) pick[j] = picks[p * SCENE_BG + j];
  } else {
    uint64_t st = seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(p + 1));
    (void)scene_mix64(st);
    int taken[SCENE_BG];
    for (int j = 0; j < SCENE_BG; ++j) {
      const uint64_t r = scene_mix64(st);
      int v = (int)(((r >> 32) * (uint64_t)(n_in - j)) >> 32);
      int pos = 0;
      for (int c = 0; c < j; ++c)
        if (v >= taken[c]) { ++v; pos = c + 1; }
      for (int c = j; c > pos; --c) taken[c] = taken[c - 1];
      taken[pos] = v;
      pick[j] = v;
    }
  }
  if (picks_out)
    for (int j = 0; j < SCENE_BG; ++j) picks_out[p * SCENE_BG + j] = pick[j];
  // coordinates of the picked matches (model side as is, input side perspective corrected)
  double bmx[SCENE_BG], bmy[SCENE_BG], bix[SCENE_BG], biy[SCENE_BG];
  {
    int ord = 0;
    for (int64_t i = 0; i < K; ++i) {
      if (!mask[m0 + i]) continue;
#pragma unroll
      for (int j = 0; j < SCENE_BG; ++j)
        if (pick[j] == ord) {
          bmx[j] = (double)src[2 * (m0 + i)]; bmy[j] = (double)src[2 * (m0 + i) + 1];
          persp32(H, (double)dst[2 * (m0 + i)], (double)dst[2 * (m0 + i) + 1], bix[j], biy[j]);
        }
      ++ord;
    }
  }
  const double* mp = model_pose + p * model_pose_stride;
  const double* ip = input_pose + p * njoints * 2;
  // find_transformation(model = pose + picks, input = corrected pose + corrected picks) (transformation.py:129)
  LsqSums m = {};
  for (int j = 0; j < njoints; ++j) {
    double x, y;
    persp32(H, (double)(float)ip[2 * j], (double)(float)ip[2 * j + 1], x, y);
    if (!(x == 0.0 && y == 0.0)) lsq_add(m, x, y, mp[2 * j], mp[2 * j + 1]);
  }
#pragma unroll
  for (int j = 0; j < SCENE_BG; ++j)
    if (!(bix[j] == 0.0 && biy[j] == 0.0)) lsq_add(m, bix[j], biy[j], bmx[j], bmy[j]);
  if (m.n < 0.5) { score[p] = NAN; return; }
  double A[9];
  lsq_solve(m, A);
#pragma unroll
  for (int i = 0; i < 9; ++i) A[i] = zero_small(A[i]);  // find_transformation returns M with |a| < 1e-10 zeroed (common.py:99)
  // np.dot(pad(input_pose), M) for every row (zero rows included), scaled by the model image size, max distance
  double best = -1.0;
  for (int j = 0; j < njoints + SCENE_BG; ++j) {
    double x, y, X, Y;
    if (j < njoints) { persp32(H, (double)(float)ip[2 * j], (double)(float)ip[2 * j + 1], x, y); X = mp[2 * j]; Y = mp[2 * j + 1]; }
    else { x = bix[j - njoints]; y = biy[j - njoints]; X = bmx[j - njoints]; Y = bmy[j - njoints]; }
    const double tx = x * A[0] + y * A[3] + A[6], ty = x * A[1] + y * A[4] + A[7];
    const double dx = fabs(X / model_w - tx / model_w), dy = fabs(Y / model_h - ty / model_h);
    const double dist = sqrt(dx * dx + dy * dy);
    if (j == 0) best = dist;
    else if (dist > best) best = dist;
  }
  score[p] = best;
}

// ---- per-rank top-k of the matching poses ---------------------------------------------------------------
constexpr int TK_THREADS = 256;
constexpr int TK_MAXK = 64;

// Block-wide extraction of the k lexicographically smallest (score, index) among candidates produced by
// `get(i)` for i in [0, count): k rounds of arg-min with the previous winner as a strict lower bound.
template <typename Get>
__device__ void block_topk(Get get, int64_t count, int k, float* out_s, int64_t* out_i, float* red_s, int64_t* red_i) {
  float last_s = -INFINITY;
  int64_t last_i = -1;
  for (int r = 0; r < k; ++r) {
    float bs = INFINITY;
    int64_t bi = INT64_MAX;
    for (int64_t i = threadIdx.x; i < count; i += TK_THREADS) {
      float s; int64_t id;
      if (!get(i, s, id)) continue;
      const bool after_last = (s > last_s) || (s == last_s && id > last_i);
      if (after_last && ((s < bs) || (s == bs && id < bi))) { bs = s; bi = id; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if ((os < bs) || (os == bs && oi < bi)) { bs = os; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { red_s[threadIdx.x >> 5] = bs; red_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < TK_THREADS / 32; ++w)
        if ((red_s[w] < bs) || (red_s[w] == bs && red_i[w] < bi)) { bs = red_s[w]; bi = red_i[w]; }
      red_s[0] = bs; red_i[0] = bi;
      out_s[r] = bs; out_i[r] = bi == INT64_MAX ? -1 : bi;
    }
    __syncthreads();
    last_s = red_s[0]; last_i = red_i[0];
    __syncthreads();
    if (last_i == INT64_MAX) {  // ran out of candidates: pad the rest
      for (int rr = r + 1 + threadIdx.x; rr < k; rr += TK_THREADS) { out_s[rr] = INFINITY; out_i[rr] = -1; }
      break;
    }
  }
}

constexpr int TK_TILE = 4096;   // candidates one reduction CTA folds into k; smallest tile of the selection passes

// One reduction level over candidate lists: CTA b takes candidates [b * TK_TILE, (b + 1) * TK_TILE) and writes its k
// best to slot b of the output.  Levels repeat until one CTA is left (a single CTA over the 39 k candidates of a
// 10 M pose database took 0.5 ms -- a fifth of the whole retrieval step; two levels take a few microseconds).
__global__ void __launch_bounds__(TK_THREADS)
pose_topk_reduce_kernel(const float* __restrict__ part_s, const int64_t* __restrict__ part_i, int64_t count,
                        const int32_t* __restrict__ count_ptr /* optional: the count lives on the device, `count` is its cap */,
                        int k, float* __restrict__ top_s, int64_t* __restrict__ top_i) {
  __shared__ float red_s[TK_THREADS / 32];
  __shared__ int64_t red_i[TK_THREADS / 32];
  if (count_ptr) count = min(count, (int64_t)*count_ptr);
  const int64_t lo = (int64_t)blockIdx.x * TK_TILE;
  const int64_t cnt = count - lo < TK_TILE ? (count - lo > 0 ? count - lo : 0) : TK_TILE;
  auto get = [&](int64_t i, float& s, int64_t& id) {
    id = part_i[lo + i];
    if (id < 0) return false;
    s = part_s[lo + i];
    return true;
  };
  block_topk(get, cnt, k, top_s + (int64_t)blockIdx.x * k, top_i + (int64_t)blockIdx.x * k, red_s, red_i);
}

// The k best of a candidate list whose length lives on the device (all entries valid).  Short lists (the usual case: a few
// dozen candidates) are ranked by counting in shared memory and go straight to the final output from CTA 0; long lists are
// folded TK_TILE candidates per CTA into k-lists for pose_topk_final_kernel.
constexpr int TK_RANK_MAX = 1024;
__global__ void __launch_bounds__(TK_THREADS)
pose_topk_select_kernel(const float* __restrict__ cand_s, const int64_t* __restrict__ cand_i, int64_t cap,
                        const int32_t* __restrict__ count_ptr, int k, float* __restrict__ lvl_s, int64_t* __restrict__ lvl_i,
                        float* __restrict__ top_s, int64_t* __restrict__ top_i) {
  __shared__ float red_s[TK_THREADS / 32];
  __shared__ int64_t red_i[TK_THREADS / 32];
  __shared__ float rs[TK_RANK_MAX];
  __shared__ int64_t ri[TK_RANK_MAX];
  const int64_t count = min(cap, (int64_t)*count_ptr);
  const bool direct = count <= TK_TILE;
  if (direct && blockIdx.x != 0) return;
  float* out_s = direct ? top_s : lvl_s + (int64_t)blockIdx.x * k;
  int64_t* out_i = direct ? top_i : lvl_i + (int64_t)blockIdx.x * k;
  const int64_t lo = (int64_t)blockIdx.x * TK_TILE;
  const int64_t cnt = count - lo < TK_TILE ? (count - lo > 0 ? count - lo : 0) : TK_TILE;
  if (cnt <= TK_RANK_MAX) {
    for (int c = threadIdx.x; c < cnt; c += TK_THREADS) { rs[c] = cand_s[lo + c]; ri[c] = cand_i[lo + c]; }
    __syncthreads();
    for (int c = threadIdx.x; c < cnt; c += TK_THREADS) {
      const float s = rs[c];
      const int64_t id = ri[c];
      int rank = 0;
      for (int o = 0; o < cnt; ++o) rank += (rs[o] < s || (rs[o] == s && ri[o] < id)) ? 1 : 0;
      if (rank < k) { out_s[rank] = s; out_i[rank] = id; }
    }
    for (int r = (int)cnt + threadIdx.x; r < k; r += TK_THREADS) { out_s[r] = INFINITY; out_i[r] = -1; }
    return;
  }
  auto get = [&](int64_t i, float& s, int64_t& id) {
    id = cand_i[lo + i];
    s = cand_s[lo + i];
    return true;
  };
  block_topk(get, cnt, k, out_s, out_i, red_s, red_i);
}

// Last level after pose_topk_select_kernel: nothing to do when the list was short enough for the direct path.
__global__ void __launch_bounds__(TK_THREADS)
pose_topk_final_kernel(const float* __restrict__ lvl_s, const int64_t* __restrict__ lvl_i, int64_t lvl_count, int64_t cap,
                       const int32_t* __restrict__ count_ptr, int k, float* __restrict__ top_s, int64_t* __restrict__ top_i) {
  __shared__ float red_s[TK_THREADS / 32];
  __shared__ int64_t red_i[TK_THREADS / 32];
  if (min(cap, (int64_t)*count_ptr) <= TK_TILE) return;
  auto get = [&](int64_t i, float& s, int64_t& id) {
    id = lvl_i[i];
    if (id < 0) return false;
    s = lvl_s[i];
    return true;
  };
  block_topk(get, lvl_count, k, top_s, top_i, red_s, red_i);
}

// ---- threshold selection: the k best of n scored poses in two streaming passes ------------------------------------------
// Pass 1 (pose_topk_min_kernel): the lexicographic minimum (score, index) of every block's tile of entries.  The k-th
// smallest of those minima, T (pose_topk_kth_kernel, rank counting), is an upper bound of the k-th smallest entry (k entries
// are <= T: the minima themselves), and the entries <= T all lie in the k tiles whose minima are <= T: at most k * tile of
// them, usually a few dozen.  With fewer than k tiles every matching entry is a candidate.
// Pass 2 (pose_topk_filter_kernel) appends them to a candidate list; pose_topk_reduce_kernel picks the k best.  Both
// passes stream match + score (5 B per pose; the second one mostly from L2) instead of k block-wide arg-min rounds per tile.
constexpr int TK_MAX_TILES = 512;     // tiles of the selection passes: the k-th smallest of their minima is found by rank counting

struct TkThreshold {
  float s;
  int pad;
  int64_t i;
};

__device__ __forceinline__ bool tk_less(float sa, int64_t ia, float sb, int64_t ib) { return sa < sb || (sa == sb && ia < ib); }

template <typename F>
__device__ __forceinline__ void tk_scan_tile(const uint8_t* __restrict__ match, const float* __restrict__ score, int64_t lo,
                                             int64_t hi, F f) {
  // 16 consecutive entries per thread and pass: one 16-byte match load + four 16-byte score loads, all independent
  // (lo is a multiple of 16)
  const bool vec = ((reinterpret_cast<uintptr_t>(match) | reinterpret_cast<uintptr_t>(score)) & 15) == 0;
  for (int64_t g = lo + 16 * (int64_t)threadIdx.x; g < hi; g += 16 * TK_THREADS) {
    if (vec && g + 16 <= hi) {
      const uint4 m = *reinterpret_cast<const uint4*>(match + g);
      if ((m.x | m.y | m.z | m.w) == 0u) continue;
      const float4* sp = reinterpret_cast<const float4*>(score + g);
      const float4 v0 = sp[0], v1 = sp[1], v2 = sp[2], v3 = sp[3];
      const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
      const float4 vv[4] = {v0, v1, v2, v3};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (mw[q] & 0x000000ffu) f(vv[q].x, g + 4 * q);
        if (mw[q] & 0x0000ff00u) f(vv[q].y, g + 4 * q + 1);
        if (mw[q] & 0x00ff0000u) f(vv[q].z, g + 4 * q + 2);
        if (mw[q] & 0xff000000u) f(vv[q].w, g + 4 * q + 3);
      }
    } else {
      for (int64_t e = g; e < min(hi, g + 16); ++e)
        if (match[e]) f(score[e], e);
    }
  }
}

__global__ void __launch_bounds__(TK_THREADS)
pose_topk_min_kernel(const uint8_t* __restrict__ match, const float* __restrict__ score, int64_t n, int64_t tile,
                     float* __restrict__ min_s, int64_t* __restrict__ min_i, int32_t* __restrict__ match_count) {
  __shared__ float red_s[TK_THREADS / 32];
  __shared__ int64_t red_i[TK_THREADS / 32];
  const int64_t lo = (int64_t)blockIdx.x * tile, hi = min(n, lo + tile);
  float bs = INFINITY;
  int64_t bi = ((int64_t)1 << 62) + blockIdx.x;   // an empty tile ranks after every entry, tiles among themselves by number
  int cnt = 0;
  tk_scan_tile(match, score, lo, hi, [&](float x, int64_t g) {
    ++cnt;
    if (x == x && tk_less(x, g, bs, bi)) { bs = x; bi = g; }   // NaN scores never rank
  });
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    const float os = __shfl_xor_sync(0xffffffffu, bs, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (tk_less(os, oi, bs, bi)) { bs = os; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    red_s[warp] = bs; red_i[warp] = bi;
    if (cnt) atomicAdd(match_count, cnt);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < TK_THREADS / 32; ++w)
      if (tk_less(red_s[w], red_i[w], bs, bi)) { bs = red_s[w]; bi = red_i[w]; }
    min_s[blockIdx.x] = bs; min_i[blockIdx.x] = bi;
  }
}

// T = the k-th smallest tile minimum (nb >= k): every block stages all minima in shared memory, a thread per tile counts the
// minima before its own.  The minimum of tile a precedes that of tile b on equal scores iff a < b (tiles partition the index
// range in order); an empty tile (score +inf, index >= 2^62) follows every entry.
constexpr int TK_KTH_THREADS = 256;   // 8 tiles per block, a warp per tile: the lanes share the counting
__global__ void __launch_bounds__(TK_KTH_THREADS)
pose_topk_kth_kernel(const float* __restrict__ min_s, const int64_t* __restrict__ min_i, int nb, int k, TkThreshold* __restrict__ thr) {
  __shared__ float ss[TK_MAX_TILES];
  __shared__ int key[TK_MAX_TILES];
#pragma unroll
  for (int o = threadIdx.x; o < TK_MAX_TILES; o += TK_KTH_THREADS) {
    const bool in = o < nb;
    const float s = in ? __ldg(min_s + o) : INFINITY;
    const int64_t i = in ? __ldg(min_i + o) : 0;
    ss[o] = s;
    key[o] = o + (i >= ((int64_t)1 << 62) ? nb : 0);
  }
  __syncthreads();
  // fewer tiles than k: the minima bound nothing, every matching entry is a candidate (at most nb * tile <= cap of them)
  if (nb < k) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { thr->s = INFINITY; thr->i = INT64_MAX; }
    return;
  }
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (TK_KTH_THREADS / 32) + (threadIdx.x >> 5);
  if (j >= nb) return;
  const float s = ss[j];
  const int kj = key[j];
  int rank = 0;
  for (int o = lane; o < nb; o += 32) rank += (ss[o] < s || (ss[o] == s && key[o] < kj)) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
  if (lane == 0 && rank == k - 1) { thr->s = s; thr->i = min_i[j]; }
}

__global__ void __launch_bounds__(TK_THREADS)
pose_topk_filter_kernel(const uint8_t* __restrict__ match, const float* __restrict__ score, int64_t n, int64_t tile,
                        int64_t index_offset, const TkThreshold* __restrict__ thr, int64_t cap, float* __restrict__ cand_s,
                        int64_t* __restrict__ cand_i, int32_t* __restrict__ cand_count) {
  const float ts = thr->s;
  const int64_t ti = thr->i;
  const int64_t lo = (int64_t)blockIdx.x * tile, hi = min(n, lo + tile);
  tk_scan_tile(match, score, lo, hi, [&](float x, int64_t g) {
    if (x == x && !tk_less(ts, ti, x, g)) {
      const int slot = atomicAdd(cand_count, 1);
      if (slot < cap) { cand_s[slot] = x; cand_i[slot] = g + index_offset; }
    }
  });
}

static int64_t topk_tile(int64_t n) {
  int64_t tile = (n + TK_MAX_TILES - 1) / TK_MAX_TILES;
  if (tile < TK_TILE) tile = TK_TILE;
  return (tile + TK_TILE - 1) / TK_TILE * TK_TILE;
}

}  // namespace hpusm

using namespace hpusm;

extern "C" {

int hpusm_affine_lsq_batch(const double* model, const double* input, int64_t n, int npts, double* out_transform,
                           double* A, uint8_t* has_A, void* stream) {
  HPUSM_REQUIRE(n >= 0 && npts > 0, HPUSM_ERR_INVALID, "hpusm_affine_lsq_batch: bad sizes");
  if (n == 0) return HPUSM_OK;
  HPUSM_REQUIRE(model && input && out_transform && A && has_A, HPUSM_ERR_INVALID,
                "hpusm_affine_lsq_batch: null pointer");
  HPUSM_LAUNCH(affine_lsq_batch_kernel, (unsigned)((n + 127) / 128), 128, 0, as_stream(stream), model, input, n, npts,
               out_transform, A, has_A);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_procrustes_batch(const double* X, const double* Y, int64_t n, int npts, int scaling, int reflection,
                           double* d, double* Z, double* tform, void* stream) {
  HPUSM_REQUIRE(n >= 0 && npts > 0 && reflection >= 0 && reflection <= 2, HPUSM_ERR_INVALID,
                "hpusm_procrustes_batch: bad arguments");
  if (n == 0) return HPUSM_OK;
  HPUSM_REQUIRE(X && Y && d && Z && tform, HPUSM_ERR_INVALID, "hpusm_procrustes_batch: null pointer");
  HPUSM_LAUNCH(procrustes_batch_kernel, (unsigned)((n + 127) / 128), 128, 0, as_stream(stream), X, Y, n, npts,
               scaling, reflection, d, Z, tform);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

static void pose_params(const double* thresholds, PoseParams& prm) {
  prm.d_torso = thresholds[0]; prm.rot_torso = thresholds[1];
  prm.d_legs = thresholds[2]; prm.rot_legs = thresholds[3]; prm.d_shoulder = thresholds[4];
  // rot*57.3 <= thr  <=>  angle <= thr/57.3; usable as a tangent bound while that stays below ~86 degrees
  const double a_t = prm.rot_torso / 57.3, a_l = prm.rot_legs / 57.3;
  prm.torso_use_tan = (a_t >= 0.0 && a_t < 1.5) ? 1 : 0;
  prm.legs_use_tan = (a_l >= 0.0 && a_l < 1.5) ? 1 : 0;
  prm.tan_torso = prm.torso_use_tan ? tan(a_t) : 0.0;
  prm.tan_legs = prm.legs_use_tan ? tan(a_l) : 0.0;
}

struct PoseWs {
  PoseQueryPrep* prep;
  int32_t* count;
  int32_t* list;
  size_t bytes;
};

static PoseWs pose_ws_layout(void* ws, int64_t n) {
  char* p = reinterpret_cast<char*>(ws);
  PoseWs w{};
  size_t off = 0;
  w.prep = reinterpret_cast<PoseQueryPrep*>(p); off += align_up(sizeof(PoseQueryPrep), 256);
  w.count = reinterpret_cast<int32_t*>(p ? p + off : nullptr); off += 256;
  w.list = reinterpret_cast<int32_t*>(p ? p + off : nullptr); off += align_up((size_t)(n > 0 ? n : 1) * sizeof(int32_t), 256);
  w.bytes = off;
  return w;
}

size_t hpusm_pose_match_workspace_bytes(int64_t n) { return pose_ws_layout(nullptr, n).bytes; }

int hpusm_pose_match_batch(const float* query, const float* db, int64_t n, int query_is_model,
                           const double* thresholds, uint8_t* match, float* score, double* detail, void* ws,
                           size_t ws_bytes, void* stream) {
  HPUSM_REQUIRE(n >= 0, HPUSM_ERR_INVALID, "hpusm_pose_match_batch: negative size");
  if (n == 0) return HPUSM_OK;
  HPUSM_REQUIRE(query && db && thresholds && match && score, HPUSM_ERR_INVALID,
                "hpusm_pose_match_batch: null pointer");
  HPUSM_REQUIRE(is_aligned(db, 16), HPUSM_ERR_INVALID, "hpusm_pose_match_batch: db must be 16-byte aligned");
  PoseParams prm;
  pose_params(thresholds, prm);
  cudaStream_t st = as_stream(stream);
  // The FP32 fast path needs the tangent form of both rotation tests, no per-part detail, and 32-bit list indices;
  // everything else (and whatever the fast path cannot certify) is the exact FP64 kernel's.
  const bool fast = !detail && prm.torso_use_tan && prm.legs_use_tan && n < 0x7fffffff && ws != nullptr;
  if (!fast) {
    HPUSM_LAUNCH_TIMED(pose_match_batch_kernel<float>, (unsigned)((n + PM_THREADS - 1) / PM_THREADS), PM_THREADS, 0, st, query,
                       db, n, query_is_model, prm, match, score, detail, nullptr, nullptr);
    HPUSM_CHECK_LAUNCH();
    return HPUSM_OK;
  }
  const PoseWs w = pose_ws_layout(ws, n);
  HPUSM_REQUIRE(is_aligned(ws, 256) && ws_bytes >= w.bytes, HPUSM_ERR_WORKSPACE,
                "hpusm_pose_match_batch: workspace needs %zu bytes, got %zu", w.bytes, ws_bytes);
  PoseFastParams fp;
  fp.d_torso = prm.d_torso; fp.tan_torso = prm.tan_torso; fp.d_legs = prm.d_legs; fp.tan_legs = prm.tan_legs;
  fp.d_shoulder = prm.d_shoulder;
  HPUSM_CUDA_OK(cudaMemsetAsync(w.count, 0, sizeof(int32_t), st));
  HPUSM_LAUNCH(pose_query_prep_kernel, 1, 32, 0, st, query, w.prep);
  const unsigned grid = (unsigned)((n + PF_THREADS - 1) / PF_THREADS);
  if (query_is_model)
    HPUSM_LAUNCH_TIMED(pose_match_fast_kernel<true>, grid, PF_THREADS, 0, st, w.prep, db, n, fp, match, score, w.list, w.count);
  else
    HPUSM_LAUNCH_TIMED(pose_match_fast_kernel<false>, grid, PF_THREADS, 0, st, w.prep, db, n, fp, match, score, w.list, w.count);
  // measured on C5 (95 k of 10 M poses deferred): 8 CTAs per SM 0.515 ms per call, 4 CTAs 0.526 ms, 16 and 32 no better
  const unsigned lgrid = (unsigned)min((int64_t)sm_count() * 8, (n + PM_THREADS - 1) / PM_THREADS);
  HPUSM_LAUNCH(pose_match_batch_kernel<float>, lgrid, PM_THREADS, 0, st, query, db, n, query_is_model, prm, match, score,
               nullptr, w.list, w.count);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_pose_match_batch_f64(const double* query, const double* db, int64_t n, int query_is_model,
                               const double* thresholds, uint8_t* match, float* score, double* detail, void* stream) {
  HPUSM_REQUIRE(n >= 0, HPUSM_ERR_INVALID, "hpusm_pose_match_batch_f64: negative size");
  if (n == 0) return HPUSM_OK;
  HPUSM_REQUIRE(query && db && thresholds && match && score, HPUSM_ERR_INVALID,
                "hpusm_pose_match_batch_f64: null pointer");
  HPUSM_REQUIRE(is_aligned(db, 16), HPUSM_ERR_INVALID, "hpusm_pose_match_batch_f64: db must be 16-byte aligned");
  PoseParams prm;
  pose_params(thresholds, prm);
  HPUSM_LAUNCH_TIMED(pose_match_batch_kernel<double>, (unsigned)((n + PM_THREADS - 1) / PM_THREADS), PM_THREADS, 0,
                     as_stream(stream), query, db, n, query_is_model, prm, match, score, detail, nullptr, nullptr);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_pose_match_pairs(const void* model, const void* input, int64_t n, int is_f64, const double* thresholds,
                           uint8_t* match, float* score, double* detail, void* stream) {
  HPUSM_REQUIRE(n >= 0, HPUSM_ERR_INVALID, "hpusm_pose_match_pairs: negative size");
  if (n == 0) return HPUSM_OK;
  HPUSM_REQUIRE(model && input && thresholds && match && score, HPUSM_ERR_INVALID, "hpusm_pose_match_pairs: null pointer");
  HPUSM_REQUIRE(is_aligned(model, 16) && is_aligned(input, 16), HPUSM_ERR_INVALID,
                "hpusm_pose_match_pairs: pose arrays must be 16-byte aligned");
  PoseParams prm;
  pose_params(thresholds, prm);
  const unsigned grid = (unsigned)((n + PM_THREADS - 1) / PM_THREADS);
  if (is_f64)
    HPUSM_LAUNCH(pose_match_pairs_kernel<double>, grid, PM_THREADS, 0, as_stream(stream), reinterpret_cast<const double*>(model),
                 reinterpret_cast<const double*>(input), n, prm, match, score, detail);
  else
    HPUSM_LAUNCH(pose_match_pairs_kernel<float>, grid, PM_THREADS, 0, as_stream(stream), reinterpret_cast<const float*>(model),
                 reinterpret_cast<const float*>(input), n, prm, match, score, detail);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_scene_score_batch(const float* src, const float* dst, const int64_t* pair_offsets, const uint8_t* mask,
                            const double* H2, const double* model_pose, int model_pose_shared, const double* input_pose,
                            int64_t npairs, int njoints, double model_w, double model_h, uint64_t seed,
                            const int32_t* picks, double* score, int32_t* picks_out, void* stream) {
  HPUSM_REQUIRE(npairs >= 0 && njoints > 0 && model_w > 0 && model_h > 0, HPUSM_ERR_INVALID, "hpusm_scene_score_batch: bad sizes");
  if (npairs == 0) return HPUSM_OK;
  HPUSM_REQUIRE(src && dst && pair_offsets && mask && H2 && model_pose && input_pose && score, HPUSM_ERR_INVALID,
                "hpusm_scene_score_batch: null pointer");
  HPUSM_LAUNCH(scene_score_batch_kernel, (unsigned)((npairs + 63) / 64), 64, 0, as_stream(stream), src, dst, pair_offsets, mask,
               H2, model_pose, input_pose, npairs, model_pose_shared ? 0 : njoints * 2, njoints, model_w, model_h, seed, picks,
               score, picks_out);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_feature_scaling_batch(const double* pts, int64_t n, int npts, double* out, void* stream) {
  HPUSM_REQUIRE(n >= 0 && npts > 0, HPUSM_ERR_INVALID, "hpusm_feature_scaling_batch: bad sizes");
  if (n == 0) return HPUSM_OK;
  HPUSM_REQUIRE(pts && out, HPUSM_ERR_INVALID, "hpusm_feature_scaling_batch: null pointer");
  HPUSM_LAUNCH(feature_scaling_batch_kernel, (unsigned)((n + 127) / 128), 128, 0, as_stream(stream), pts, n, npts, out);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

// workspace: tile minima, the threshold, the candidate list (at most k tiles of entries) and one level of k-lists
struct TkLayout {
  size_t min_s, min_i, thr, count, cand_s, cand_i, lvl_s, lvl_i, total;
  int64_t tile, cap;
  int nb, lvl_blocks;
};

static TkLayout tk_layout(int64_t n, int k) {
  TkLayout L;
  L.tile = topk_tile(n > 0 ? n : 1);
  L.nb = (int)(((n > 0 ? n : 1) + L.tile - 1) / L.tile);
  L.cap = (int64_t)k * L.tile;
  L.lvl_blocks = (int)((L.cap + TK_TILE - 1) / TK_TILE);
  size_t off = 0;
  L.min_s = off; off += align_up((size_t)L.nb * sizeof(float), 256);
  L.min_i = off; off += align_up((size_t)L.nb * sizeof(int64_t), 256);
  L.thr = off; off += 256;
  L.count = off; off += 256;
  L.cand_s = off; off += align_up((size_t)L.cap * sizeof(float), 256);
  L.cand_i = off; off += align_up((size_t)L.cap * sizeof(int64_t), 256);
  L.lvl_s = off; off += align_up((size_t)L.lvl_blocks * k * sizeof(float), 256);
  L.lvl_i = off; off += align_up((size_t)L.lvl_blocks * k * sizeof(int64_t), 256);
  L.total = off;
  return L;
}

size_t hpusm_pose_topk_workspace_bytes(int64_t n, int k) {
  if (k <= 0 || k > TK_MAXK) return 256;
  return tk_layout(n, k).total;
}

int hpusm_pose_topk(const uint8_t* match, const float* score, int64_t n, int64_t index_offset, int k, float* top_score,
                    int64_t* top_index, int32_t* match_count, void* ws, size_t ws_bytes, void* stream) {
  HPUSM_REQUIRE(n >= 0 && k > 0 && k <= TK_MAXK, HPUSM_ERR_INVALID, "hpusm_pose_topk: k must be in 1..%d", TK_MAXK);
  HPUSM_REQUIRE(n < ((int64_t)1 << 33), HPUSM_ERR_UNSUPPORTED, "hpusm_pose_topk: too many poses for one call");
  HPUSM_REQUIRE(top_score && top_index && match_count && (n == 0 || (match && score)), HPUSM_ERR_INVALID,
                "hpusm_pose_topk: null pointer");
  cudaStream_t st = as_stream(stream);
  const TkLayout L = tk_layout(n, k);
  HPUSM_REQUIRE(ws && is_aligned(ws, 256) && ws_bytes >= L.total, HPUSM_ERR_WORKSPACE,
                "hpusm_pose_topk: workspace needs %zu bytes, got %zu", L.total, ws_bytes);
  char* w = reinterpret_cast<char*>(ws);
  float* min_s = reinterpret_cast<float*>(w + L.min_s);
  int64_t* min_i = reinterpret_cast<int64_t*>(w + L.min_i);
  TkThreshold* thr = reinterpret_cast<TkThreshold*>(w + L.thr);
  int32_t* cand_count = reinterpret_cast<int32_t*>(w + L.count);
  float* cand_s = reinterpret_cast<float*>(w + L.cand_s);
  int64_t* cand_i = reinterpret_cast<int64_t*>(w + L.cand_i);
  float* lvl_s = reinterpret_cast<float*>(w + L.lvl_s);
  int64_t* lvl_i = reinterpret_cast<int64_t*>(w + L.lvl_i);
  HPUSM_CUDA_OK(cudaMemsetAsync(match_count, 0, sizeof(int32_t), st));
  HPUSM_CUDA_OK(cudaMemsetAsync(cand_count, 0, sizeof(int32_t), st));
  HPUSM_LAUNCH(pose_topk_min_kernel, L.nb, TK_THREADS, 0, st, match, score, n, L.tile, min_s, min_i, match_count);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(pose_topk_kth_kernel, (L.nb + TK_KTH_THREADS / 32 - 1) / (TK_KTH_THREADS / 32), TK_KTH_THREADS, 0, st, min_s, min_i, L.nb, k,
               thr);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(pose_topk_filter_kernel, L.nb, TK_THREADS, 0, st, match, score, n, L.tile, index_offset, thr, L.cap, cand_s, cand_i,
               cand_count);
  HPUSM_CHECK_LAUNCH();
  // a short candidate list (the usual case) is finished by CTA 0 of the select kernel; a long one is folded TK_TILE
  // candidates per CTA into k-lists, which the final kernel reduces
  HPUSM_LAUNCH(pose_topk_select_kernel, L.lvl_blocks, TK_THREADS, 0, st, cand_s, cand_i, L.cap, cand_count, k, lvl_s, lvl_i, top_score,
               top_index);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(pose_topk_final_kernel, 1, TK_THREADS, 0, st, lvl_s, lvl_i, (int64_t)L.lvl_blocks * k, L.cap, cand_count, k, top_score,
               top_index);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_pose_topk_merge(const float* cand_score, const int64_t* cand_index, int64_t count, int k, float* top_score,
                          int64_t* top_index, void* stream) {
  HPUSM_REQUIRE(count >= 0 && count <= TK_TILE && k > 0 && k <= TK_MAXK, HPUSM_ERR_INVALID,
                "hpusm_pose_topk_merge: count must be in 0..%d and k in 1..%d", TK_TILE, TK_MAXK);
  HPUSM_REQUIRE(top_score && top_index && (count == 0 || (cand_score && cand_index)), HPUSM_ERR_INVALID,
                "hpusm_pose_topk_merge: null pointer");
  HPUSM_LAUNCH(pose_topk_reduce_kernel, 1, TK_THREADS, 0, as_stream(stream), cand_score, cand_index, count, nullptr, k, top_score,
               top_index);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

}  // extern "C"
