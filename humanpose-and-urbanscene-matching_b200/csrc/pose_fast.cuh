// K4 fast path: single_person.match_single (posematching/single_person.py:80-189) in FP32 for one (query, database pose)
// pair, with a certificate.  The reference works in float64 and its decisions are threshold comparisons; this path
// evaluates the same quantities in FP32 (no FP64 instructions, no conversions) and reports `certain = false` whenever
//   * a threshold comparison falls inside a band that bounds the FP32 error (distances: 3e-5 absolute on values <= ~1.4;
//     rotation: 1e-3 relative on the tangent),
//   * the least-squares problem of a body part is not comfortably well conditioned (fewer than 3 joints, or
//     det(C) <= 0.01 tr(C)^2, i.e. condition number above ~100), where FP32 round-off in the fit could exceed the bands,
//   * the pose needs one of the reference's corner cases: negative coordinates, a joint with exactly ONE zero coordinate
//     (feature_scaling's minima then run over different row sets), an empty or zero-extent scaling, two coordinates of
//     one body part equal to the minimum (a detected joint could then scale to (0,0) and drop out of the fit).
// Uncertain poses are appended to a list and redone by the exact FP64 kernel, so the decisions returned are the FP64
// path's; the FP32 score of a certain pose differs from the FP64 one by < 1e-6.
// With those corner cases out of the way a joint is either (0,0) on both sides or detected on both, and the quirky
// minima of feature_scaling (common.py:688-708: min over BOTH columns of the rows whose x, resp. y, is non-zero)
// coincide: one minimum `lo` per pose.  The code is written for the machine it runs on: ncu showed the first version
// bound by the ALU pipe (compares, selects, bit tests: 65 % busy at half the FMA pipe's width, FMA pipe 29 %), so what
// does not need a branch is arithmetic on the FMA pipe instead: a joint's weight is w = sat(x * 1e30) (1 for a detected
// joint, 0 for a missing one: one FMUL.SAT), masked extrema are max(w * q) and min(key + (1 - w) * 1e30), zero counts
// are sums of sat(s * 1e30), masked joints carry weight 0 in every sum.
// Everything here is __host__ __device__ so that tests/pose_fast_host.cpp can run the very same code on the CPU against
// the oracle (oracle/pose_oracle.py) over millions of poses -- that test is what pins the bands above.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define HPUSM_HD __host__ __device__ __forceinline__
#define HPUSM_UNROLL _Pragma("unroll")
#else
#define HPUSM_HD inline
#define HPUSM_UNROLL
#endif

namespace hpusm {

// The arithmetic type is a template parameter: float is the product path; double gives the test harness a high-precision
// evaluation of the SAME formulas to hold the float decisions against.
HPUSM_HD float pf_max(float a, float b) { return fmaxf(a, b); }
HPUSM_HD float pf_min(float a, float b) { return fminf(a, b); }
HPUSM_HD float pf_abs(float a) { return fabsf(a); }
HPUSM_HD float pf_sqrt(float a) { return sqrtf(a); }
HPUSM_HD double pf_max(double a, double b) { return fmax(a, b); }
HPUSM_HD double pf_min(double a, double b) { return fmin(a, b); }
HPUSM_HD double pf_abs(double a) { return fabs(a); }
HPUSM_HD double pf_sqrt(double a) { return sqrt(a); }
HPUSM_HD double pf_rcp(double a) { return 1.0 / a; }
HPUSM_HD float pf_rcp(float a) {
#ifdef __CUDA_ARCH__
  return __fdividef(1.f, a);  // MUFU.RCP: 2 ulp, far inside the certification bands
#else
  return 1.f / a;
#endif
}
HPUSM_HD unsigned pf_bits(float v) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(v);
#else
  unsigned u;
  memcpy(&u, &v, 4);
  return u;
#endif
}
// 0 for v <= 0, 1 for v >= 1e-30 (saturating multiply: one instruction on the FMA pipe).  Exact 0/1 for every value this
// file feeds it: raw coordinates of poses whose smallest non-zero coordinate is above 1e-20 (checked), and scaled
// coordinates (a non-zero one is at least one ulp of a pixel coordinate times 1/extent).
HPUSM_HD float pf_flag(float v) {
#ifdef __CUDA_ARCH__
  return __saturatef(v * 1e30f);
#else
  return fminf(fmaxf(v * 1e30f, 0.f), 1.f);
#endif
}
HPUSM_HD double pf_flag(double v) { return v > 0.0 ? 1.0 : 0.0; }

// ---- pairs: the x and y of a joint go through the same arithmetic, so they travel as one operand.  On the device a
// float pair is an aligned register pair and every operation is ONE packed sm_100 instruction (FADD2 / FMUL2 / FFMA2;
// ptxas folds the broadcasts p2s(v) into the instruction's operand modifiers): half the issue slots of the scalar form.
// Each half is rounded exactly like the scalar operation.
struct P2d { double x, y; };
HPUSM_HD P2d p2(double x, double y) { P2d r; r.x = x; r.y = y; return r; }
HPUSM_HD P2d p2s(double v) { return p2(v, v); }
HPUSM_HD double p2x(P2d a) { return a.x; }
HPUSM_HD double p2y(P2d a) { return a.y; }
HPUSM_HD P2d p2_add(P2d a, P2d b) { return p2(a.x + b.x, a.y + b.y); }
HPUSM_HD P2d p2_sub(P2d a, P2d b) { return p2(a.x - b.x, a.y - b.y); }
HPUSM_HD P2d p2_mul(P2d a, P2d b) { return p2(a.x * b.x, a.y * b.y); }
HPUSM_HD P2d p2_fma(P2d a, P2d b, P2d c) { return p2(a.x * b.x + c.x, a.y * b.y + c.y); }
#ifdef __CUDA_ARCH__
typedef unsigned long long P2f;
HPUSM_HD P2f p2(float x, float y) { P2f r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
HPUSM_HD P2f p2s(float v) { P2f r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v)); return r; }
HPUSM_HD float p2x(P2f a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); return lo; }
HPUSM_HD float p2y(P2f a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); return hi; }
HPUSM_HD P2f p2_add(P2f a, P2f b) { P2f r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
HPUSM_HD P2f p2_sub(P2f a, P2f b) { P2f r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
HPUSM_HD P2f p2_mul(P2f a, P2f b) { P2f r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
HPUSM_HD P2f p2_fma(P2f a, P2f b, P2f c) { P2f r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#else
struct P2f { float x, y; };
HPUSM_HD P2f p2(float x, float y) { P2f r; r.x = x; r.y = y; return r; }
HPUSM_HD P2f p2s(float v) { return p2(v, v); }
HPUSM_HD float p2x(P2f a) { return a.x; }
HPUSM_HD float p2y(P2f a) { return a.y; }
HPUSM_HD P2f p2_add(P2f a, P2f b) { return p2(a.x + b.x, a.y + b.y); }
HPUSM_HD P2f p2_sub(P2f a, P2f b) { return p2(a.x - b.x, a.y - b.y); }
HPUSM_HD P2f p2_mul(P2f a, P2f b) { return p2(a.x * b.x, a.y * b.y); }
HPUSM_HD P2f p2_fma(P2f a, P2f b, P2f c) { return p2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
template <typename F> struct PairT;
template <> struct PairT<float> { typedef P2f T; };
template <> struct PairT<double> { typedef P2d T; };

constexpr int PF_NJ = 18;
constexpr int PF_NP = 14;   // joints 0..13 are the ones torso and legs use (face joints 14..17 only enter the scaling)

// Body parts (common.py:269-274).
struct FastTorso {
  static constexpr int N = 10;
  HPUSM_HD static constexpr int at(int k) { return k < 9 ? k : 11; }
  static constexpr unsigned ROWS = 0x1FFu | (1u << 11);
};
struct FastLegs {
  static constexpr int N = 7;
  HPUSM_HD static constexpr int at(int k) { return k == 0 ? 1 : 7 + k; }
  static constexpr unsigned ROWS = (1u << 1) | (0x3Fu << 8);
};

// What depends on the query only (constant over the database): built once per launch.
struct PoseQueryPrep {
  float x[PF_NJ], y[PF_NJ];     // raw coordinates, the query's own undetected joints as (0,0)
  float key[PF_NJ];             // min(x, y)
  unsigned und;                 // the query's own undetected joints
  int usable;                   // 0: the fast path does not apply to this query (negative / single zero / tiny coordinates)
};

struct PoseFastParams {
  double d_torso, tan_torso, d_legs, tan_legs, d_shoulder;
};

HPUSM_HD void pose_query_prep(const float* q /*[18][2]*/, PoseQueryPrep& pq) {
  pq.und = 0; pq.usable = 1;
  for (int j = 0; j < PF_NJ; ++j) {
    const float x = q[2 * j], y = q[2 * j + 1];
    pq.x[j] = x; pq.y[j] = y;
    const bool zx = x == 0.f, zy = y == 0.f;
    if (zx && zy) pq.und |= 1u << j;
    else if (zx != zy || !(x > 1e-20f) || !(y > 1e-20f)) pq.usable = 0;
    pq.key[j] = fminf(x, y);
  }
}

template <typename F>
struct PfPart {
  F max_d2, shoulder_d2, a00, a01, a10, a11;
  bool ok;  // false: not certifiable in FP32 (ill conditioned, too few joints)
};

// One body part.  M / I: scaled (x, y) of the model / input side of joints 0..13 (exact 0 for a masked joint on both
// sides), w[j] = 1 for a joint that takes part, 0 for a masked one, n = number of joints that do.
template <typename F, typename Rows>
HPUSM_HD PfPart<F> pf_part(const typename PairT<F>::T (&M)[PF_NP], const typename PairT<F>::T (&I)[PF_NP],
                           const F (&w)[PF_NP], F n_joints) {
  typedef typename PairT<F>::T P2;
  constexpr int NP = Rows::N;
  PfPart<F> r;
  P2 sI = p2s(F(0)), sM = p2s(F(0));
  HPUSM_UNROLL
  for (int k = 0; k < NP; ++k) {
    const int j = Rows::at(k);
    sI = p2_add(sI, I[j]); sM = p2_add(sM, M[j]);  // masked joints are exact zeros on both sides
  }
  r.ok = n_joints >= F(3);
  const P2 inv_n = p2s(pf_rcp(pf_max(n_joints, F(1))));
  const P2 mI = p2_mul(sI, inv_n), mM = p2_mul(sM, inv_n);
  P2 dI[NP], dM[NP];
  P2 cd = p2s(F(0)), cX = p2s(F(0)), cY = p2s(F(0));  // (cxx, cyy), (cxX, cyX), (cxY, cyY)
  F cxy = F(0);
  HPUSM_UNROLL
  for (int k = 0; k < NP; ++k) {
    const int j = Rows::at(k);
    dI[k] = p2_sub(I[j], mI); dM[k] = p2_sub(M[j], mM);
    const P2 wd = p2_mul(p2s(w[j]), dI[k]);
    cd = p2_fma(wd, dI[k], cd);
    cxy += p2x(wd) * p2y(dI[k]);
    cX = p2_fma(wd, p2s(p2x(dM[k])), cX);
    cY = p2_fma(wd, p2s(p2y(dM[k])), cY);
  }
  const F cxx = p2x(cd), cyy = p2y(cd), cxX = p2x(cX), cyX = p2y(cX), cxY = p2x(cY), cyY = p2y(cY);
  const F det = cxx * cyy - cxy * cxy, tr = cxx + cyy;
  if (!(det > F(0.01) * tr * tr)) r.ok = false;
  const F id = pf_rcp(r.ok ? det : F(1));
  r.a00 = (cyy * cxX - cxy * cyX) * id; r.a10 = (cxx * cyX - cxy * cxX) * id;
  r.a01 = (cyy * cxY - cxy * cyY) * id; r.a11 = (cxx * cyY - cxy * cxY) * id;
  const P2 A0 = p2(r.a00, r.a01), A1 = p2(r.a10, r.a11);
  r.max_d2 = F(0); r.shoulder_d2 = F(0);
  HPUSM_UNROLL
  for (int k = 0; k < NP; ++k) {
    const int j = Rows::at(k);
    // model - (input * A + t), both sides centred on the means of the fitted joints; a masked joint is (0,0) on both
    // sides and comes back as (0,0) (common.py:93-96): residual 0
    const P2 t = p2_fma(p2s(p2x(dI[k])), A0, p2_mul(p2s(p2y(dI[k])), A1));
    const P2 f = p2_sub(dM[k], t);
    const P2 sq = p2_mul(f, f);
    const F d2 = w[j] * (p2x(sq) + p2y(sq));
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

// match_single for one pair.  px/py: the database pose (raw float32, undetected joints (0,0)); taken by reference and
// modified: the query's undetected joints are zeroed in them.  QIM: the query is the reference's `model_features` and
// the database pose its `input_features`; otherwise the roles are swapped.
// Returns `certain`; match/score are only meaningful when it is true.
template <typename F, bool QIM>
HPUSM_HD bool pose_fast_decide(const PoseQueryPrep& pq, float (&px)[PF_NJ], float (&py)[PF_NJ], const PoseFastParams& prm,
                               F dist_band, F rot_band, uint8_t& match, F& score) {
  bool certain = true;
  const F BIG = F(1e30);
  // handle_undetected_points: a joint missing on either side is (0,0) on both.  The query's missing joints (the same
  // for every pose of the launch: a uniform branch, usually not taken) are zeroed in the database pose first.
  if (pq.und) {
    HPUSM_UNROLL
    for (int j = 0; j < PF_NJ; ++j)
      if ((pq.und >> j) & 1u) { px[j] = 0.f; py[j] = 0.f; }
  }
  // which joints are detected: w = 1 / 0 per joint, from x (and from y: they must agree, or the pose is deferred)
  F w[PF_NJ];
  unsigned signs = 0;
  F disagree = F(0);
  HPUSM_UNROLL
  for (int j = 0; j < PF_NJ; ++j) {
    signs |= pf_bits(px[j]) | pf_bits(py[j]);
    w[j] = F(pf_flag(px[j]));
    const F d = w[j] - F(pf_flag(py[j]));
    disagree += d * d;
  }
  if ((signs >> 31) || disagree != F(0)) return false;  // negative coordinates / a joint with one zero coordinate

  // feature_scaling of the database pose (common.py:688-708): maxima over all rows, ONE minimum over both columns of
  // the detected joints (a missing joint is pushed out of the minimum by (1 - w) * 1e30)
  float xmax = 0.f, ymax = 0.f;
  F loPf = BIG;
  HPUSM_UNROLL
  for (int j = 0; j < PF_NJ; ++j) {
    xmax = fmaxf(xmax, px[j]); ymax = fmaxf(ymax, py[j]);
    loPf = pf_min(loPf, (F(1) - w[j]) * BIG + F(fminf(px[j], py[j])));
  }
  // exact: the winning term is a raw float32 coordinate (or 1e30 when nothing is detected)
  if (!(loPf > F(1e-20)) || !(F(xmax) > loPf) || !(F(ymax) > loPf)) return false;
  const F ixP = pf_rcp(F(xmax) - loPf), iyP = pf_rcp(F(ymax) - loPf);

  // feature_scaling of the query with the same joints masked out
  F qxmax = F(0), qymax = F(0), loQ = BIG;
  HPUSM_UNROLL
  for (int j = 0; j < PF_NJ; ++j) {
    qxmax = pf_max(qxmax, w[j] * F(pq.x[j]));
    qymax = pf_max(qymax, w[j] * F(pq.y[j]));
    loQ = pf_min(loQ, (F(1) - w[j]) * BIG + F(pq.key[j]));
  }
  if (!(qxmax > loQ) || !(qymax > loQ)) return false;
  const F ixQ = pf_rcp(qxmax - loQ), iyQ = pf_rcp(qymax - loQ);

  // scaled coordinates of joints 0..13: (raw - lo) * (w * inv): a masked joint comes out as an exact +0 on both sides,
  // a coordinate that equals the minimum as an exact 0 (the part gates count zeros).  Counts of non-zero scaled
  // coordinates per part as sums of 0/1 flags, (database pose, query) side by side.
  typedef typename PairT<F>::T P2;
  P2 SP[PF_NP], SQ[PF_NP];
  F wp[PF_NP];
  const P2 invP = p2(ixP, iyP), invQ = p2(ixQ, iyQ), loP2 = p2s(loPf), loQ2 = p2s(loQ), zero2 = p2s(F(0));
  P2 nz_t = zero2, nz_l = zero2;  // (database pose, query)
  F n_t = F(0), n_l = F(0);
  HPUSM_UNROLL
  for (int j = 0; j < PF_NP; ++j) {
    wp[j] = w[j];
    const P2 w2 = p2s(w[j]);
    SP[j] = p2_fma(p2_sub(p2(F(px[j]), F(py[j])), loP2), p2_mul(w2, invP), zero2);
    SQ[j] = p2_fma(p2_sub(p2(F(pq.x[j]), F(pq.y[j])), loQ2), p2_mul(w2, invQ), zero2);
    const P2 c = p2(F(pf_flag(p2x(SP[j]))) + F(pf_flag(p2y(SP[j]))), F(pf_flag(p2x(SQ[j]))) + F(pf_flag(p2y(SQ[j]))));
    if ((FastTorso::ROWS >> j) & 1u) { nz_t = p2_add(nz_t, c); n_t += w[j]; }
    if ((FastLegs::ROWS >> j) & 1u) { nz_l = p2_add(nz_l, c); n_l += w[j]; }
  }
  const F nzP_t = p2x(nz_t), nzQ_t = p2y(nz_t), nzP_l = p2x(nz_l), nzQ_l = p2y(nz_l);
  // A detected joint whose scaled coordinates are BOTH zero is left out of the reference's fit while its partner is
  // not: that takes two coordinates equal to the minimum inside one part -- such poses go to the exact kernel.
  if (F(2) * n_t - nzP_t >= F(2) || F(2) * n_t - nzQ_t >= F(2) || F(2) * n_l - nzP_l >= F(2) || F(2) * n_l - nzQ_l >= F(2))
    return false;
  const PfPart<F> torso = QIM ? pf_part<F, FastTorso>(SQ, SP, wp, n_t) : pf_part<F, FastTorso>(SP, SQ, wp, n_t);
  const PfPart<F> legs = QIM ? pf_part<F, FastLegs>(SQ, SP, wp, n_l) : pf_part<F, FastLegs>(SP, SQ, wp, n_l);
  if (!torso.ok || !legs.ok) return false;
  const F torso_dist = pf_sqrt(torso.max_d2), legs_dist = pf_sqrt(legs.max_d2), shoulder_dist = pf_sqrt(torso.shoulder_d2);
  const F nzm_t = QIM ? nzQ_t : nzP_t, nzi_t = QIM ? nzP_t : nzQ_t;  // small integers, held exactly
  const F nzm_l = QIM ? nzQ_l : nzP_l, nzi_l = QIM ? nzP_l : nzQ_l;

  // single_person.py:150-178
  bool ok_torso = true, ok_legs = true;
  if (nzm_t > F(4)) {
    if (nzm_t - nzi_t < F(2)) {
      const bool c1 = pf_le<F>(torso_dist, F(prm.d_torso), dist_band, certain);
      const bool c2 = pf_rotation_ok<F>(torso, F(prm.tan_torso), rot_band, certain);
      const bool c3 = pf_le<F>(shoulder_dist, F(prm.d_shoulder), dist_band, certain);
      ok_torso = c1 && c2 && c3;
    } else ok_torso = false;
  }
  if (nzm_l > F(8)) {
    if (nzm_l - nzi_l < F(2)) {
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
