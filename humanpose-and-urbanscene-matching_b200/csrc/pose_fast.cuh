// K4 fast path: single_person.match_single (posematching/single_person.py:80-189) in FP32 for one (query, database pose)
// pair, with a certificate.  The reference works in float64 and its decisions are threshold comparisons; this path
// evaluates the same quantities in FP32 (no FP64 instructions, no conversions) and reports `certain = false` whenever
//   * a threshold comparison falls inside a band that bounds the FP32 error (distances: 3e-5 absolute on values <= ~1.4;
//     rotation: 1e-3 relative on the tangent),
//   * the least-squares problem of a body part is not comfortably well conditioned (fewer than 3 joints, or
//     det(C) <= 0.01 tr(C)^2, i.e. condition number above ~100), where FP32 round-off in the fit could exceed the bands,
//   * the pose needs the reference's corner cases (negative coordinates, invalid scaling, empty parts, masked-out
//     extremal joints beyond the prepared candidates).
// Uncertain poses are appended to a list and redone by the exact FP64 kernel (pose_match_list_kernel), so the decisions
// returned are the FP64 path's; the FP32 score of a certain pose differs from the FP64 one by < 2e-6.
// Everything here is __host__ __device__ so that tests/pose_fast_host.cpp can run the very same code on the CPU against
// the oracle (oracle/pose_oracle.py) over millions of poses -- that test is what pins the bands above.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define HPUSM_HD __host__ __device__ __forceinline__
#else
#define HPUSM_HD inline
#endif

namespace hpusm {

// The arithmetic type is a template parameter: float is the product path; double gives the test harness a high-precision
// evaluation of the SAME formulas to hold the float decisions against.
HPUSM_HD float pf_min(float a, float b) { return fminf(a, b); }
HPUSM_HD float pf_max(float a, float b) { return fmaxf(a, b); }
HPUSM_HD float pf_abs(float a) { return fabsf(a); }
HPUSM_HD float pf_sqrt(float a) { return sqrtf(a); }
HPUSM_HD double pf_min(double a, double b) { return fmin(a, b); }
HPUSM_HD double pf_max(double a, double b) { return fmax(a, b); }
HPUSM_HD double pf_abs(double a) { return fabs(a); }
HPUSM_HD double pf_sqrt(double a) { return sqrt(a); }
HPUSM_HD double pf_rcp(double a) { return 1.0 / a; }
HPUSM_HD int pf_popc(unsigned v) {
#ifdef __CUDA_ARCH__
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
HPUSM_HD float pf_rcp(float a) {
#ifdef __CUDA_ARCH__
  return __fdividef(1.f, a);  // MUFU.RCP: 2 ulp, far inside the certification bands
#else
  return 1.f / a;
#endif
}

constexpr int PF_NJ = 18;
constexpr int PF_CAND = 4;  // prepared candidates per query extremum

// Body parts (common.py:269-274).
struct FastTorso {
  static constexpr int N = 10;
  HPUSM_HD static constexpr int at(int k) { return k < 9 ? k : 11; }
};
struct FastLegs {
  static constexpr int N = 7;
  HPUSM_HD static constexpr int at(int k) { return k == 0 ? 1 : 7 + k; }
};

// What depends on the query only (constant over the database): built once per launch.
struct PoseQueryPrep {
  float x[PF_NJ], y[PF_NJ];     // raw coordinates, the query's own undetected joints as (0,0)
  unsigned und;                 // the query's own undetected joints
  unsigned zx, zy;              // joints whose raw x (resp. y) is exactly 0
  int xmax_j[PF_CAND], ymax_j[PF_CAND], xmin_j[PF_CAND], ymin_j[PF_CAND];  // joint indices, best first; -1 = none
  float xmin_key[PF_NJ], ymin_key[PF_NJ];  // feature_scaling's quirky minima: x != 0 ? min(x, y) : +inf (resp. y != 0)
  int usable;                   // 0: the fast path does not apply to this query (negative coordinates)
};

struct PoseFastParams {
  double d_torso, tan_torso, d_legs, tan_legs, d_shoulder;
};

HPUSM_HD void pose_query_prep(const float* q /*[18][2]*/, PoseQueryPrep& pq) {
  pq.und = 0; pq.zx = 0; pq.zy = 0; pq.usable = 1;
  for (int j = 0; j < PF_NJ; ++j) {
    const float x = q[2 * j], y = q[2 * j + 1];
    pq.x[j] = x; pq.y[j] = y;
    if (x == 0.f && y == 0.f) pq.und |= 1u << j;
    if (x == 0.f) pq.zx |= 1u << j;
    if (y == 0.f) pq.zy |= 1u << j;
    if (!(x >= 0.f) || !(y >= 0.f)) pq.usable = 0;
    const float lo = fminf(x, y);
    pq.xmin_key[j] = x != 0.f ? lo : INFINITY;
    pq.ymin_key[j] = y != 0.f ? lo : INFINITY;
  }
  // the PF_CAND best joints of each extremum (selection sort over 18 values; runs once per launch)
  for (int which = 0; which < 4; ++which) {
    unsigned taken = pq.und;
    int* out = which == 0 ? pq.xmax_j : which == 1 ? pq.ymax_j : which == 2 ? pq.xmin_j : pq.ymin_j;
    for (int c = 0; c < PF_CAND; ++c) {
      int best = -1;
      float bv = 0.f;
      for (int j = 0; j < PF_NJ; ++j) {
        if ((taken >> j) & 1u) continue;
        const float v = which == 0 ? pq.x[j] : which == 1 ? pq.y[j] : which == 2 ? -pq.xmin_key[j] : -pq.ymin_key[j];
        if (best < 0 || v > bv) { best = j; bv = v; }
      }
      out[c] = best;
      if (best >= 0) taken |= 1u << best;
    }
  }
}

// First prepared candidate that the mask leaves alive; false when all of them are masked.
HPUSM_HD bool pf_pick(const int* cand, unsigned und, int& j) {
  j = -1;
#pragma unroll
  for (int c = PF_CAND - 1; c >= 0; --c) {
    const int cj = cand[c];
    if (cj >= 0 && !((und >> cj) & 1u)) j = cj;
  }
  return j >= 0;
}

template <typename F>
struct PfPart {
  F max_d2, shoulder_d2, a00, a01, a10, a11;
  int nz_model, nz_input;
  bool ok;  // false: not certifiable in FP32 (ill conditioned, too few joints)
};

constexpr int PF_NP = 14;  // joints 0..13 are the ones torso and legs use (face joints 14..17 only enter the scaling)

// One body part.  (xm, ym) scaled model coordinates, (xi, yi) scaled input coordinates of joints 0..13; nzm_x/nzm_y and
// nzi_x/nzi_y: bit j set when the scaled model / input coordinate of joint j is non-zero.
// Branch free: a joint that takes no part in the fit (scaled input (0,0)) carries weight 0 in every sum.
template <typename F, typename Rows>
HPUSM_HD PfPart<F> pf_part(const F (&xm)[PF_NP], const F (&ym)[PF_NP], const F (&xi)[PF_NP], const F (&yi)[PF_NP],
                           unsigned nzm_x, unsigned nzm_y, unsigned nzi_x, unsigned nzi_y) {
  constexpr int NP = Rows::N;
  unsigned rows = 0;
#pragma unroll
  for (int k = 0; k < NP; ++k) rows |= 1u << Rows::at(k);
  PfPart<F> r;
  r.nz_model = pf_popc(nzm_x & rows) + pf_popc(nzm_y & rows);
  r.nz_input = pf_popc(nzi_x & rows) + pf_popc(nzi_y & rows);
  const unsigned inc = (nzi_x | nzi_y) & rows;  // joints whose scaled input is not (0,0) (common.py:62-68)
  const F n = F(pf_popc(inc));
  F sx = F(0), sy = F(0), sX = F(0), sY = F(0);
  F w[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int j = Rows::at(k);
    w[k] = ((inc >> j) & 1u) ? F(1) : F(0);
    // a left-out joint has xi = yi = 0 already; its model coordinates must not enter the model mean
    sx += xi[j]; sy += yi[j]; sX += w[k] * xm[j]; sY += w[k] * ym[j];
  }
  r.ok = n >= F(2.5);
  const F inv_n = pf_rcp(pf_max(n, F(1)));
  const F mx = sx * inv_n, my = sy * inv_n, mX = sX * inv_n, mY = sY * inv_n;
  F cxx = F(0), cxy = F(0), cyy = F(0), cxX = F(0), cyX = F(0), cxY = F(0), cyY = F(0);
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int j = Rows::at(k);
    const F dx = xi[j] - mx, dy = yi[j] - my, dX = xm[j] - mX, dY = ym[j] - mY;
    const F wx = w[k] * dx, wy = w[k] * dy;
    cxx += wx * dx; cxy += wx * dy; cyy += wy * dy;
    cxX += wx * dX; cyX += wy * dX; cxY += wx * dY; cyY += wy * dY;
  }
  const F det = cxx * cyy - cxy * cxy, tr = cxx + cyy;
  if (!(det > F(0.01) * tr * tr)) r.ok = false;
  const F id = pf_rcp(r.ok ? det : F(1));
  r.a00 = (cyy * cxX - cxy * cyX) * id; r.a10 = (cxx * cyX - cxy * cxX) * id;
  r.a01 = (cyy * cxY - cxy * cyY) * id; r.a11 = (cxx * cyY - cxy * cxY) * id;
  r.max_d2 = F(0); r.shoulder_d2 = F(0);
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int j = Rows::at(k);
    // fitted joint: model - (input * A + t), both sides centred on the means of the fitted joints;
    // a joint left out of the fit comes back as (0,0) (common.py:93-96): the residual is the model joint itself
    const F dx = xi[j] - mx, dy = yi[j] - my;
    const F fx = (xm[j] - mX) - (dx * r.a00 + dy * r.a10);
    const F fy = (ym[j] - mY) - (dx * r.a01 + dy * r.a11);
    const F rx = w[k] != F(0) ? fx : xm[j], ry = w[k] != F(0) ? fy : ym[j];
    const F d2 = rx * rx + ry * ry;
    r.max_d2 = pf_max(r.max_d2, d2);
    if (k == 1 || k == 4) r.shoulder_d2 = pf_max(r.shoulder_d2, d2);  // torso rows 1 and 4 (pose_comparison.py:62)
  }
  return r;
}

// val <= thr, flagged uncertain when the two are closer than `band`.
template <typename F>
HPUSM_HD bool pf_le(F val, F thr, F band, bool& certain) {
  if (pf_abs(val - thr) <= band) certain = false;
  return val <= thr;
}

// rot = max(|atan2(-A01, A00)|, |atan2(A10, A11)|) * 57.3 <= thr  <=>  A00 >= 0, |A01| <= tan * A00, same for (A10, A11)
template <typename F>
HPUSM_HD bool pf_rotation_ok(const PfPart<F>& p, F tan_thr, F band, bool& certain) {
  const F l1 = pf_abs(p.a01), r1 = tan_thr * p.a00, l2 = pf_abs(p.a10), r2 = tan_thr * p.a11;
  if (pf_abs(p.a00) < band * F(0.1) || pf_abs(p.a11) < band * F(0.1)) certain = false;
  if (pf_abs(l1 - r1) <= band * (l1 + pf_abs(r1)) + band * F(1e-4)) certain = false;
  if (pf_abs(l2 - r2) <= band * (l2 + pf_abs(r2)) + band * F(1e-4)) certain = false;
  return p.a00 >= F(0) && l1 <= r1 && p.a11 >= F(0) && l2 <= r2;
}

constexpr float PF_DIST_BAND = 3e-5f;  // absolute, on distances between points of the unit square
constexpr float PF_ROT_BAND = 1e-3f;   // relative, on the tangent comparison

// match_single for one pair.  px/py: the database pose (raw float32, undetected joints (0,0)).  QIM: the query is the
// reference's `model_features` and the database pose its `input_features`; otherwise the roles are swapped.
// Returns `certain`; match/score are only meaningful when it is true.
template <typename F, bool QIM>
HPUSM_HD bool pose_fast_decide(const PoseQueryPrep& pq, const float (&px_in)[PF_NJ], const float (&py_in)[PF_NJ],
                               const PoseFastParams& prm, F dist_band, F rot_band, uint8_t& match, F& score) {
  bool certain = true;
  // handle_undetected_points: a joint missing on either side is (0,0) on both
  unsigned undP = 0;
#pragma unroll
  for (int j = 0; j < PF_NJ; ++j)
    if (px_in[j] == 0.f && py_in[j] == 0.f) undP |= 1u << j;
  const unsigned und = pq.und | undP;

  // feature_scaling of the database pose (common.py:688-708 incl. the minimum over BOTH columns)
  float xmax = -INFINITY, ymax = -INFINITY, xmin = INFINITY, ymin = INFINITY;
  bool neg = false;
#pragma unroll
  for (int j = 0; j < PF_NJ; ++j) {
    const bool u = (pq.und >> j) & 1u;  // its own (0,0) joints are zero already
    const float x = u ? 0.f : px_in[j], y = u ? 0.f : py_in[j];
    xmax = fmaxf(xmax, x); ymax = fmaxf(ymax, y);
    const float lo = fminf(x, y);
    xmin = fminf(xmin, x != 0.f ? lo : INFINITY);
    ymin = fminf(ymin, y != 0.f ? lo : INFINITY);
    neg = neg || x < 0.f || y < 0.f;
  }
  if (neg || !(xmin < INFINITY) || !(ymin < INFINITY) || !(xmax > xmin) || !(ymax > ymin)) return false;
  const F ixP = pf_rcp(F(xmax) - F(xmin)), iyP = pf_rcp(F(ymax) - F(ymin));
  const float xloP = xmin, yloP = ymin;

  // feature_scaling of the masked query through the prepared candidate lists
  int jx, jy, jlx, jly;
  if (!pf_pick(pq.xmax_j, und, jx) || !pf_pick(pq.ymax_j, und, jy) || !pf_pick(pq.xmin_j, und, jlx) ||
      !pf_pick(pq.ymin_j, und, jly))
    return false;
  const float qxmax = pq.x[jx], qymax = pq.y[jy], xloQ = pq.xmin_key[jlx], yloQ = pq.ymin_key[jly];
  if (!(xloQ < INFINITY) || !(yloQ < INFINITY) || !(qxmax > xloQ) || !(qymax > yloQ)) return false;
  const F ixQ = pf_rcp(F(qxmax) - F(xloQ)), iyQ = pf_rcp(F(qymax) - F(yloQ));

  // scaled coordinates of joints 0..13 and the bit masks of the non-zero ones.  A masked or literally zero coordinate is
  // replaced by the minimum before scaling, so it comes out as an exact 0, like a coordinate that equals the minimum.
  F sxP[PF_NP], syP[PF_NP], sxQ[PF_NP], syQ[PF_NP];
  unsigned nzP_x = 0, nzP_y = 0, nzQ_x = 0, nzQ_y = 0;
#pragma unroll
  for (int j = 0; j < PF_NP; ++j) {
    const bool u = (und >> j) & 1u;
    const float mpx = (u || px_in[j] == 0.f) ? xloP : px_in[j], mpy = (u || py_in[j] == 0.f) ? yloP : py_in[j];
    const float mqx = (u || ((pq.zx >> j) & 1u)) ? xloQ : pq.x[j], mqy = (u || ((pq.zy >> j) & 1u)) ? yloQ : pq.y[j];
    sxP[j] = (F(mpx) - F(xloP)) * ixP; syP[j] = (F(mpy) - F(yloP)) * iyP;
    sxQ[j] = (F(mqx) - F(xloQ)) * ixQ; syQ[j] = (F(mqy) - F(yloQ)) * iyQ;
    nzP_x |= (sxP[j] != F(0) ? 1u : 0u) << j; nzP_y |= (syP[j] != F(0) ? 1u : 0u) << j;
    nzQ_x |= (sxQ[j] != F(0) ? 1u : 0u) << j; nzQ_y |= (syQ[j] != F(0) ? 1u : 0u) << j;
  }
  const PfPart<F> torso = QIM ? pf_part<F, FastTorso>(sxQ, syQ, sxP, syP, nzQ_x, nzQ_y, nzP_x, nzP_y)
                              : pf_part<F, FastTorso>(sxP, syP, sxQ, syQ, nzP_x, nzP_y, nzQ_x, nzQ_y);
  const PfPart<F> legs = QIM ? pf_part<F, FastLegs>(sxQ, syQ, sxP, syP, nzQ_x, nzQ_y, nzP_x, nzP_y)
                             : pf_part<F, FastLegs>(sxP, syP, sxQ, syQ, nzP_x, nzP_y, nzQ_x, nzQ_y);
  if (!torso.ok || !legs.ok) return false;
  const F torso_dist = pf_sqrt(torso.max_d2), legs_dist = pf_sqrt(legs.max_d2), shoulder_dist = pf_sqrt(torso.shoulder_d2);

  // single_person.py:150-178
  bool ok_torso = true, ok_legs = true;
  if (torso.nz_model > 4) {
    if (torso.nz_model - torso.nz_input < 2) {
      const bool c1 = pf_le<F>(torso_dist, F(prm.d_torso), dist_band, certain);
      const bool c2 = pf_rotation_ok<F>(torso, F(prm.tan_torso), rot_band, certain);
      const bool c3 = pf_le<F>(shoulder_dist, F(prm.d_shoulder), dist_band, certain);
      ok_torso = c1 && c2 && c3;
    } else ok_torso = false;
  }
  if (legs.nz_model > 8) {
    if (legs.nz_model - legs.nz_input < 2) {
      const bool c1 = pf_le<F>(legs_dist, F(prm.d_legs), dist_band, certain);
      const bool c2 = pf_rotation_ok<F>(legs, F(prm.tan_legs), rot_band, certain);
      ok_legs = c1 && c2;
    } else ok_legs = false;
  }
  const F err = (torso_dist + legs_dist) * F(0.5);
  if (!(err == err)) return false;
  match = (ok_torso && ok_legs) ? 1 : 0;
  score = err;
  return certain;
}

}  // namespace hpusm
