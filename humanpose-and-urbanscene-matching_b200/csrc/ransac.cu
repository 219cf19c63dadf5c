// K3: batched RANSAC homography.
//
// Replaces cv2.findHomography(src, dst, cv2.RANSAC, 5.0) at reference urbanscene/features.py:66,:69.
// OpenCV's sampler cannot be seeded from Python, so parity is defined against oracle/ransac_oracle.c,
// which restates exactly the arithmetic below (same counter-based sample table, same operation order,
// explicit round-to-nearest intrinsics / no FMA contraction), giving identical hypotheses, identical
// inlier sets and an identical winner on CPU and GPU.
//
// One CTA per image pair:
//   phase A  each LANE derives one hypothesis: sample 4 matches from the (seed, pair, hypothesis) key,
//            reject degenerate samples (collinear triples, inconsistent orientation), and solve the
//            exact 4-point homography in FP64 in closed form (unit square -> quad on both sides,
//            H = S_dst * adj(S_src)); the FP32 model is parked in shared memory.
//   phase B  the surviving hypotheses are compacted; a warp takes 32 of them (one per LANE, model in registers) and
//            streams a segment of the pair's matches, staged in shared memory as float4 (x, y, x', y') and read as
//            broadcast LDS.128; the inlier test is division free, (u - x'w)^2 + (v - y'w)^2 <= thr^2 w^2, the count
//            lives in a register.  (Round-1 first version: one warp per hypothesis + popc(ballot), as north_star
//            sketches; ncu showed 2.6x the instructions per reprojection -- ballot, POPC on the 16-lane XU pipe, and
//            half-empty hypothesis pairs -- see profiles/r1_ransac_score_ncu.txt.)
//   phase C  the sequential early-exit of RANSAC (cv2's adaptive iteration bound) is replayed over the
//            counts by one thread, so the parallel evaluation returns what the sequential loop would.
// A second kernel re-derives the winner, writes its inlier mask, and refines: normalised least-squares
// DLT on the inliers, then Levenberg-Marquardt on the reprojection error (FP64, block reductions).
#include <float.h>
#include <math.h>
#include "common.cuh"

namespace hpusm {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_CHUNK = 4096;  // matches staged in shared memory at a time (64 KB)

// ---- counter-based sample table -----------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Four distinct indices in [0, K), K >= 4, as a function of (seed, pair, hyp) only.
__host__ __device__ __forceinline__ void sample4(uint64_t seed, uint64_t pair, uint64_t hyp, uint32_t K,
                                                 uint32_t out[4]) {
  uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (pair + 1)) ^ (0x8CB92BA72F3D8DD7ull * (hyp + 1));
  (void)splitmix64(s);
  uint32_t sorted[4];
  for (int j = 0; j < 4; ++j) {
    const uint64_t r = splitmix64(s);
    uint32_t v = (uint32_t)(((r >> 32) * (uint64_t)(K - j)) >> 32);  // rank among the unused indices
    int pos = 0;
    for (int c = 0; c < j; ++c) {
      if (v >= sorted[c]) { ++v; pos = c + 1; }
    }
    for (int c = j; c > pos; --c) sorted[c] = sorted[c - 1];
    sorted[pos] = v;
    out[j] = v;
  }
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

// orientation of the triangle (p, q, r): (q-p) x (r-p)
__device__ __forceinline__ double orient(const double* x, const double* y, int p, int q, int r) {
  return det2(dsub(x[q], x[p]), dsub(y[q], y[p]), dsub(x[r], x[p]), dsub(y[r], y[p]));
}

// True when the 4-point sample is unusable: a (near-)collinear triple on either side, or the four
// triangles do not keep a common orientation relation between the two images.
__device__ __forceinline__ bool degenerate4(const double* sx, const double* sy, const double* dx, const double* dy) {
  const int tri[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
  int flipped = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = tri[k][0], q = tri[k][1], r = tri[k][2];
    const double os = orient(sx, sy, p, q, r);
    const double od = orient(dx, dy, p, q, r);
    const double ms = dadd(dadd(fabs(dsub(sx[q], sx[p])), fabs(dsub(sy[q], sy[p]))),
                           dadd(fabs(dsub(sx[r], sx[p])), fabs(dsub(sy[r], sy[p]))));
    const double md = dadd(dadd(fabs(dsub(dx[q], dx[p])), fabs(dsub(dy[q], dy[p]))),
                           dadd(fabs(dsub(dx[r], dx[p])), fabs(dsub(dy[r], dy[p]))));
    if (fabs(os) <= dmul(1.1920928955078125e-07, ms)) return true;
    if (fabs(od) <= dmul(1.1920928955078125e-07, md)) return true;
    if ((os < 0.0) != (od < 0.0)) ++flipped;
  }
  return flipped != 0 && flipped != 4;
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

// The four sampled correspondences of hypothesis (pair, hyp) as doubles.
__device__ __forceinline__ void hypothesis_points(const float* __restrict__ src, const float* __restrict__ dst, uint32_t K,
                                                  uint64_t seed, uint64_t pair, uint64_t hyp, double* sx, double* sy,
                                                  double* dx, double* dy) {
  uint32_t id[4];
  sample4(seed, pair, hyp, K, id);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 s = reinterpret_cast<const float2*>(src)[id[j]];
    const float2 d = reinterpret_cast<const float2*>(dst)[id[j]];
    sx[j] = (double)s.x; sy[j] = (double)s.y; dx[j] = (double)d.x; dy[j] = (double)d.y;
  }
}

// Same, reading the matches from the pair-packed shared-memory staging of the scoring kernel:
// pts[2j] = (x0, x1, y0, y1), pts[2j+1] = (-x'0, -x'1, -y'0, -y'1) for matches 2j, 2j+1 (negation is exact).
__device__ __forceinline__ void hypothesis_points_smem(const float4* pts, uint32_t K, uint64_t seed, uint64_t pair,
                                                       uint64_t hyp, double* sx, double* sy, double* dx, double* dy) {
  uint32_t id[4];
  sample4(seed, pair, hyp, K, id);
  const float* f = reinterpret_cast<const float*>(pts);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t base = (id[j] >> 1) * 8 + (id[j] & 1);
    sx[j] = (double)f[base]; sy[j] = (double)f[base + 2];
    dx[j] = (double)(-f[base + 4]); dy[j] = (double)(-f[base + 6]);
  }
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

// Whole minimal-sample pipeline for one hypothesis.
__device__ bool hypothesis_model(const float* __restrict__ src, const float* __restrict__ dst, uint32_t K,
                                 uint64_t seed, uint64_t pair, uint64_t hyp, double* H) {
  double sx[4], sy[4], dx[4], dy[4];
  hypothesis_points(src, dst, K, seed, pair, hyp, sx, sy, dx, dy);
  if (degenerate4(sx, sy, dx, dy)) return false;
  return hypothesis_solve(sx, sy, dx, dy, H);
}

// Division-free inlier test in FP32 with explicit FMAs (mirrored by the oracle with fmaf()).
__device__ __forceinline__ bool is_inlier(const float* h, float x, float y, float xp, float yp, float thr2) {
  const float w = __fmaf_rn(h[6], x, __fmaf_rn(h[7], y, h[8]));
  const float u = __fmaf_rn(h[0], x, __fmaf_rn(h[1], y, h[2]));
  const float v = __fmaf_rn(h[3], x, __fmaf_rn(h[4], y, h[5]));
  const float du = __fmaf_rn(-xp, w, u);
  const float dv = __fmaf_rn(-yp, w, v);
  const float e = __fmaf_rn(du, du, __fmul_rn(dv, dv));
  const float lim = __fmul_rn(__fmul_rn(thr2, w), w);
  return e <= lim;
}

// Same test, accumulating into `cnt` with one predicated add (the compiler's select + add costs an extra issue slot).
__device__ __forceinline__ void inlier_count(const float* h, float x, float y, float xp, float yp, float thr2, int& cnt) {
  const float w = __fmaf_rn(h[6], x, __fmaf_rn(h[7], y, h[8]));
  const float u = __fmaf_rn(h[0], x, __fmaf_rn(h[1], y, h[2]));
  const float v = __fmaf_rn(h[3], x, __fmaf_rn(h[4], y, h[5]));
  const float du = __fmaf_rn(-xp, w, u);
  const float dv = __fmaf_rn(-yp, w, v);
  const float e = __fmaf_rn(du, du, __fmul_rn(dv, dv));
  const float lim = __fmul_rn(__fmul_rn(thr2, w), w);
  asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt) : "f"(e), "f"(lim));
}

// ---- packed FP32 (sm_100 FFMA2 / FMUL2): two matches per instruction, each half rounded exactly like the scalar op ----
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

// The inlier test of is_inlier() on TWO matches at once.  xy = (x0, x1, y0, y1), nq = (-x'0, -x'1, -y'0, -y'1);
// h2[i] holds model coefficient i in both halves.  Same operations in the same order as the scalar test, so the
// counts stay bit-identical to oracle/ransac_oracle.c.
__device__ __forceinline__ void inlier_count2(const f32x2* h2, f32x2 thr2, float4 xy, float4 nq, int& cnt) {
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

// cv2's adaptive bound on the number of iterations (RANSACUpdateNumIters restated).
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

// grid.x = pairs. dynamic smem: float4 pts[chunk]; float models[nhyp][9] (compacted); int counts[nhyp]; int hyp_of[nhyp].
//   A1  one LANE per hypothesis: sample + degeneracy test; survivors are compacted (ballot + one smem atomic per warp)
//   A2  one LANE per surviving hypothesis: FP64 closed-form solve, FP32 model parked in shared memory
//   B   lane = hypothesis (model in 9 registers), the pair's matches stream through as broadcast LDS.128; work items
//       are (32 hypotheses) x (RS_SEG matches) so all warps stay busy; the count lives in a register
//   C   sequential replay of RANSAC's early exit over the counts
constexpr int RS_SEG = 256;

__global__ void __launch_bounds__(RS_THREADS)
ransac_score_kernel(const float* __restrict__ src, const float* __restrict__ dst,
                    const int64_t* __restrict__ pair_offsets, float thresh, int nhyp, double confidence,
                    uint64_t seed, int64_t pair0, int chunk, int32_t* __restrict__ counts_out,
                    int32_t* __restrict__ best_hyp, int32_t* __restrict__ best_count) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  float4* pts = reinterpret_cast<float4*>(rs_smem);
  float* models = reinterpret_cast<float*>(rs_smem + (size_t)chunk * sizeof(float4));
  int* counts = reinterpret_cast<int*>(models + (size_t)nhyp * 9);
  int* hyp_of = counts + nhyp;
  __shared__ int n_valid;
  const int64_t pair = blockIdx.x;
  const int64_t m0 = pair_offsets[pair];
  const int64_t K64 = pair_offsets[pair + 1] - m0;
  const uint32_t K = (uint32_t)K64;
  const float* s = src + m0 * 2;
  const float* d = dst + m0 * 2;
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
  // When the whole pair fits (the common case) it is staged ONCE, before hypothesis generation, so the 8 scattered
  // sample reads per hypothesis hit shared memory instead of L2.
  const bool resident = K64 <= chunk;
  if (resident) stage(0, (int)K64);
  __syncthreads();

  // phase A1
  const int rounds = (nhyp + RS_THREADS - 1) / RS_THREADS;
  for (int r = 0; r < rounds; ++r) {
    const int h = r * RS_THREADS + tid;
    bool ok = false;
    if (h < nhyp && K >= 4) {
      double sx[4], sy[4], dx[4], dy[4];
      if (resident) hypothesis_points_smem(pts, K, seed, (uint64_t)(pair + pair0), (uint64_t)h, sx, sy, dx, dy);
      else hypothesis_points(s, d, K, seed, (uint64_t)(pair + pair0), (uint64_t)h, sx, sy, dx, dy);
      ok = !degenerate4(sx, sy, dx, dy);
    }
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    int base = 0;
    if (lane == 0 && b) base = atomicAdd(&n_valid, __popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (ok) hyp_of[base + __popc(b & ((1u << lane) - 1u))] = h;
    if (h < nhyp) counts[h] = ok ? 0 : -1;
  }
  __syncthreads();
  const int nv = n_valid;

  // phase A2
  for (int j = tid; j < nv; j += RS_THREADS) {
    const int h = hyp_of[j];
    double sx[4], sy[4], dx[4], dy[4], H[9];
    if (resident) hypothesis_points_smem(pts, K, seed, (uint64_t)(pair + pair0), (uint64_t)h, sx, sy, dx, dy);
    else hypothesis_points(s, d, K, seed, (uint64_t)(pair + pair0), (uint64_t)h, sx, sy, dx, dy);
    const bool ok = hypothesis_solve(sx, sy, dx, dy, H);
#pragma unroll
    for (int i = 0; i < 9; ++i) models[j * 9 + i] = ok ? (float)H[i] : 0.f;
    if (!ok) { counts[h] = -1; hyp_of[j] = -1; }
  }
  __syncthreads();

  // phase B
  const int groups = (nv + 31) >> 5;
  const f32x2 thr2x2 = pack2(thr2, thr2);
  for (int64_t c0 = 0; c0 < K64; c0 += chunk) {
    const int cn = (int)min((int64_t)chunk, K64 - c0);
    if (!resident) {
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
        const float c = h >= 0 ? models[j * 9 + i] : 0.f;
        h2[i] = pack2(c, c);
      }
      const int p0 = sg * (RS_SEG / 2), p1 = min(npair, p0 + RS_SEG / 2);
      int cnt = 0;
#pragma unroll 4
      for (int i = p0; i < p1; ++i) inlier_count2(h2, thr2x2, pts[2 * i], pts[2 * i + 1], cnt);
      if (h >= 0 && cnt) atomicAdd(&counts[h], cnt);
    }
    __syncthreads();
  }

  if (counts_out) {
    for (int h = tid; h < nhyp; h += RS_THREADS) counts_out[pair * nhyp + h] = counts[h];
  }

  // phase C: replay the sequential loop (exact early-exit semantics), lowest hypothesis wins ties
  if (best_hyp && tid == 0) {
    int best = -1, best_c = 3, niters = nhyp;
    for (int h = 0; h < nhyp && h < niters; ++h) {
      const int c = counts[h];
      if (c > best_c) {
        best_c = c; best = h;
        if (confidence > 0.0) niters = update_num_iters(confidence, (double)(K64 - c) / (double)K64, nhyp);
      }
    }
    best_hyp[pair] = best;
    best_count[pair] = best >= 0 ? best_c : 0;
  }
}

// ---- refinement ------------------------------------------------------------------------------------
constexpr int RF_THREADS = 256;
constexpr int RF_NSUM = 45;  // 36 (upper triangle of 8x8) + 8 (rhs) + 1 (cost)

// Block-wide sum of NS doubles per thread; result valid in `out` (shared) for all threads after return.
template <int NS>
__device__ void block_sum(double (&v)[NS], double* scratch /* [RF_THREADS/32][NS] */, double* out /* [NS] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    double x = v[i];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) scratch[warp * NS + i] = x;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NS; i += blockDim.x) {
    double x = 0.0;
    for (int w = 0; w < RF_THREADS / 32; ++w) x += scratch[w * NS + i];
    out[i] = x;
  }
  __syncthreads();
}

// Solve (A + lambda*diag(A)) x = b for the symmetric 8x8 A packed as the upper triangle (row-major) by
// Cholesky. Returns false if the matrix is not positive definite.
__device__ bool solve8_damped(const double* packed, const double* b, double lambda, double* x) {
  double L[8][8];
  int p = 0;
  double A[8][8];
  for (int i = 0; i < 8; ++i)
    for (int j = i; j < 8; ++j) { A[i][j] = packed[p]; A[j][i] = packed[p]; ++p; }
  for (int i = 0; i < 8; ++i) A[i][i] *= (1.0 + lambda);
  for (int j = 0; j < 8; ++j) {
    double sum = A[j][j];
    for (int k = 0; k < j; ++k) sum -= L[j][k] * L[j][k];
    if (!(sum > 0.0)) return false;
    const double ljj = sqrt(sum);
    L[j][j] = ljj;
    for (int i = j + 1; i < 8; ++i) {
      double s2 = A[i][j];
      for (int k = 0; k < j; ++k) s2 -= L[i][k] * L[j][k];
      L[i][j] = s2 / ljj;
    }
  }
  double y[8];
  for (int i = 0; i < 8; ++i) {
    double s2 = b[i];
    for (int k = 0; k < i; ++k) s2 -= L[i][k] * y[k];
    y[i] = s2 / L[i][i];
  }
  for (int i = 7; i >= 0; --i) {
    double s2 = y[i];
    for (int k = i + 1; k < 8; ++k) s2 -= L[k][i] * x[k];
    x[i] = s2 / L[i][i];
  }
  return true;
}

__device__ __forceinline__ void accumulate_rows(double (&acc)[RF_NSUM], const double* r1, double b1, const double* r2,
                                                double b2) {
  int p = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = i; j < 8; ++j) { acc[p] += r1[i] * r1[j] + r2[i] * r2[j]; ++p; }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[36 + i] += r1[i] * b1 + r2[i] * b2;
}

// grid.x = pairs.
__global__ void __launch_bounds__(RF_THREADS)
ransac_finalize_kernel(const float* __restrict__ src, const float* __restrict__ dst,
                       const int64_t* __restrict__ pair_offsets, float thresh, uint64_t seed, int64_t pair0, int refine,
                       const int32_t* __restrict__ best_hyp, double* __restrict__ Hout, uint8_t* __restrict__ mask,
                       int32_t* __restrict__ ninl) {
  __shared__ double sh_H[9];
  __shared__ float sh_Hf[9];
  __shared__ int sh_ok;
  __shared__ double scratch[(RF_THREADS / 32) * RF_NSUM];
  __shared__ double sums[RF_NSUM];
  __shared__ double sh_h[8], sh_trial[8];
  __shared__ int sh_flag;
  __shared__ double sh_c1[1];
  const int64_t pair = blockIdx.x;
  const int64_t m0 = pair_offsets[pair];
  const int64_t K = pair_offsets[pair + 1] - m0;
  const float2* s = reinterpret_cast<const float2*>(src) + m0;
  const float2* d = reinterpret_cast<const float2*>(dst) + m0;
  uint8_t* mk = mask + m0;
  const int tid = threadIdx.x;
  const int bh = best_hyp[pair];
  const float thr2 = __fmul_rn(thresh, thresh);

  if (tid == 0) {
    double H[9];
    bool ok = bh >= 0 && hypothesis_model(src + m0 * 2, dst + m0 * 2, (uint32_t)K, seed, (uint64_t)(pair + pair0), (uint64_t)bh, H);
    sh_ok = ok ? 1 : 0;
    for (int i = 0; i < 9; ++i) { sh_H[i] = ok ? H[i] : 0.0; sh_Hf[i] = ok ? (float)H[i] : 0.f; }
  }
  __syncthreads();
  if (!sh_ok) {
    for (int64_t i = tid; i < K; i += RF_THREADS) mk[i] = 0;
    if (tid < 9) Hout[pair * 9 + tid] = nan("");
    if (tid == 0) ninl[pair] = 0;
    return;
  }
  float hf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) hf[i] = sh_Hf[i];

  // inlier mask of the winning minimal model (what cv2 returns as `mask`) + normalisation statistics
  double st[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) st[i] = 0.0;
  for (int64_t i = tid; i < K; i += RF_THREADS) {
    const float2 a = s[i], b = d[i];
    const bool in = is_inlier(hf, a.x, a.y, b.x, b.y, thr2);
    mk[i] = in ? 1 : 0;
    if (in) {
      st[0] += 1.0;
      st[1] += a.x; st[2] += a.y; st[3] += (double)a.x * a.x + (double)a.y * a.y;
      st[4] += b.x; st[5] += b.y; st[6] += (double)b.x * b.x + (double)b.y * b.y;
    }
  }
  block_sum<9>(st, scratch, sums);
  const double n_in = sums[0];
  if (tid == 0) ninl[pair] = (int)n_in;
  if (!refine || n_in < 4.5) {
    // with exactly 4 inliers the minimal model is already the least-squares optimum
    if (tid < 9) Hout[pair * 9 + tid] = sh_H[tid];
    return;
  }
  // isotropic Hartley normalisation of both point sets (RMS distance sqrt(2))
  const double csx = sums[1] / n_in, csy = sums[2] / n_in, cdx = sums[4] / n_in, cdy = sums[5] / n_in;
  const double vs = fmax(sums[3] / n_in - (csx * csx + csy * csy), 1e-300);
  const double vd = fmax(sums[6] / n_in - (cdx * cdx + cdy * cdy), 1e-300);
  const double ss = sqrt(2.0 / vs), sd = sqrt(2.0 / vd);

  // (1) linear least squares in the normalised frame with h8 = 1
  double acc[RF_NSUM];
#pragma unroll
  for (int i = 0; i < RF_NSUM; ++i) acc[i] = 0.0;
  for (int64_t i = tid; i < K; i += RF_THREADS) {
    if (!mk[i]) continue;
    const float2 a = s[i], b = d[i];
    const double x = ((double)a.x - csx) * ss, y = ((double)a.y - csy) * ss;
    const double u = ((double)b.x - cdx) * sd, v = ((double)b.y - cdy) * sd;
    const double r1[8] = {x, y, 1.0, 0.0, 0.0, 0.0, -x * u, -y * u};
    const double r2[8] = {0.0, 0.0, 0.0, x, y, 1.0, -x * v, -y * v};
    accumulate_rows(acc, r1, u, r2, v);
  }
  block_sum<RF_NSUM>(acc, scratch, sums);
  if (tid == 0) {
    double x[8];
    bool ok = solve8_damped(sums, sums + 36, 0.0, x);
    if (!ok) {
      // fall back to the minimal model expressed in the normalised frame: Hn = Td * H * Ts^-1
      const double* H = sh_H;
      double G[9];
      // Td * H : rows scaled/shifted
      for (int j = 0; j < 3; ++j) {
        G[j] = sd * (H[j] - cdx * H[6 + j]);
        G[3 + j] = sd * (H[3 + j] - cdy * H[6 + j]);
        G[6 + j] = H[6 + j];
      }
      // * Ts^-1 : Ts^-1 = [[1/ss,0,csx],[0,1/ss,csy],[0,0,1]]
      double R[9];
      for (int i = 0; i < 3; ++i) {
        R[3 * i] = G[3 * i] / ss; R[3 * i + 1] = G[3 * i + 1] / ss;
        R[3 * i + 2] = G[3 * i] * csx + G[3 * i + 1] * csy + G[3 * i + 2];
      }
      for (int i = 0; i < 8; ++i) x[i] = R[i] / R[8];
    }
    for (int i = 0; i < 8; ++i) sh_h[i] = x[i];
  }
  __syncthreads();

  // (2) Levenberg-Marquardt on the reprojection error in the normalised destination frame
  double lambda = 1e-3;
  double cost = -1.0;
  for (int iter = 0; iter < 30; ++iter) {
    double h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = sh_h[i];
#pragma unroll
    for (int i = 0; i < RF_NSUM; ++i) acc[i] = 0.0;
    for (int64_t i = tid; i < K; i += RF_THREADS) {
      if (!mk[i]) continue;
      const float2 a = s[i], b = d[i];
      const double x = ((double)a.x - csx) * ss, y = ((double)a.y - csy) * ss;
      const double u = ((double)b.x - cdx) * sd, v = ((double)b.y - cdy) * sd;
      const double w = h[6] * x + h[7] * y + 1.0;
      const double iw = fabs(w) > DBL_EPSILON ? 1.0 / w : 0.0;
      const double pu = (h[0] * x + h[1] * y + h[2]) * iw;
      const double pv = (h[3] * x + h[4] * y + h[5]) * iw;
      const double ru = pu - u, rv = pv - v;
      const double r1[8] = {x * iw, y * iw, iw, 0.0, 0.0, 0.0, -x * iw * pu, -y * iw * pu};
      const double r2[8] = {0.0, 0.0, 0.0, x * iw, y * iw, iw, -x * iw * pv, -y * iw * pv};
      accumulate_rows(acc, r1, -ru, r2, -rv);
      acc[44] += ru * ru + rv * rv;
    }
    block_sum<RF_NSUM>(acc, scratch, sums);
    cost = sums[44];
    // inner loop: damped solves until the cost goes down
    bool improved = false;
    for (int attempt = 0; attempt < 12 && !improved; ++attempt) {
      if (tid == 0) {
        double dx[8];
        const bool ok = solve8_damped(sums, sums + 36, lambda, dx);
        sh_flag = ok ? 1 : 0;
        for (int i = 0; i < 8; ++i) sh_trial[i] = sh_h[i] + (ok ? dx[i] : 0.0);
      }
      __syncthreads();
      const bool ok = sh_flag != 0;
      double ht[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ht[i] = sh_trial[i];
      double c1[1] = {0.0};
      if (ok) {
        for (int64_t i = tid; i < K; i += RF_THREADS) {
          if (!mk[i]) continue;
          const float2 a = s[i], b = d[i];
          const double x = ((double)a.x - csx) * ss, y = ((double)a.y - csy) * ss;
          const double u = ((double)b.x - cdx) * sd, v = ((double)b.y - cdy) * sd;
          const double w = ht[6] * x + ht[7] * y + 1.0;
          const double iw = fabs(w) > DBL_EPSILON ? 1.0 / w : 0.0;
          const double ru = (ht[0] * x + ht[1] * y + ht[2]) * iw - u;
          const double rv = (ht[3] * x + ht[4] * y + ht[5]) * iw - v;
          c1[0] += ru * ru + rv * rv;
        }
      }
      block_sum<1>(c1, scratch, sh_c1);
      if (ok && sh_c1[0] <= cost) {
        improved = true;
        lambda = fmax(lambda * 0.1, 1e-15);
      } else {
        lambda *= 10.0;
      }
      __syncthreads();
    }
    if (!improved) break;
    // accept the step; stop when it no longer moves the parameters
    double step = 0.0, scale = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      step = fmax(step, fabs(sh_trial[i] - sh_h[i]));
      scale = fmax(scale, fabs(sh_h[i]));
    }
    __syncthreads();
    if (tid < 8) sh_h[tid] = sh_trial[tid];
    __syncthreads();
    if (step <= 1e-13 * (1.0 + scale)) break;
  }

  // denormalise: H = Td^-1 * Hn * Ts, then H /= H[8]
  if (tid == 0) {
    double Hn[9];
    for (int i = 0; i < 8; ++i) Hn[i] = sh_h[i];
    Hn[8] = 1.0;
    double G[9];  // Hn * Ts, Ts = [[ss,0,-ss*csx],[0,ss,-ss*csy],[0,0,1]]
    for (int i = 0; i < 3; ++i) {
      G[3 * i] = Hn[3 * i] * ss; G[3 * i + 1] = Hn[3 * i + 1] * ss;
      G[3 * i + 2] = Hn[3 * i + 2] - ss * (Hn[3 * i] * csx + Hn[3 * i + 1] * csy);
    }
    double R[9];  // Td^-1 = [[1/sd,0,cdx],[0,1/sd,cdy],[0,0,1]]
    for (int j = 0; j < 3; ++j) {
      R[j] = G[j] / sd + cdx * G[6 + j];
      R[3 + j] = G[3 + j] / sd + cdy * G[6 + j];
      R[6 + j] = G[6 + j];
    }
    for (int i = 0; i < 9; ++i) {
      const double hv = R[i] / R[8];
      Hout[pair * 9 + i] = hv;
      sh_Hf[i] = (float)hv;
    }
  }
  __syncthreads();
  // OpenCV returns the inliers of the REFINED homography (float32 reprojection error with division, the
  // formula of its computeError); mirrored op for op by final_mask() in oracle/ransac_oracle.c.
#pragma unroll
  for (int i = 0; i < 9; ++i) hf[i] = sh_Hf[i];
  double cnt[1] = {0.0};
  for (int64_t i = tid; i < K; i += RF_THREADS) {
    const float2 a = s[i], b = d[i];
    const float ww = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(hf[6], a.x), __fmul_rn(hf[7], a.y)), 1.f));
    const float px = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(hf[0], a.x), __fmul_rn(hf[1], a.y)), hf[2]), ww);
    const float py = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(hf[3], a.x), __fmul_rn(hf[4], a.y)), hf[5]), ww);
    const float dx = __fsub_rn(px, b.x), dy = __fsub_rn(py, b.y);
    const bool in = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= thr2;
    mk[i] = in ? 1 : 0;
    cnt[0] += in ? 1.0 : 0.0;
  }
  block_sum<1>(cnt, scratch, sh_c1);
  if (tid == 0) ninl[pair] = (int)sh_c1[0];
}

static size_t ransac_smem_bytes(int nhyp, int chunk) {
  return (size_t)chunk * sizeof(float4) + (size_t)nhyp * (9 * sizeof(float) + 2 * sizeof(int));
}

static int ransac_pick_chunk(int nhyp) {
  // keep at least two CTAs per SM resident when the hypothesis table allows it
  const size_t table = (size_t)nhyp * 44;
  size_t budget = 100 * 1024;
  if (table + 1024 * sizeof(float4) > budget) budget = 220 * 1024;
  if (table + 256 * sizeof(float4) > budget) return -1;
  size_t c = (budget - table) / sizeof(float4);
  if (c > RS_CHUNK) c = RS_CHUNK;
  c = c / 32 * 32;
  return (int)c;
}

static int launch_score(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs,
                        float thresh, int nhyp, double confidence, uint64_t seed, int64_t pair0, int32_t* counts,
                        int32_t* best_hyp, int32_t* best_count, cudaStream_t st) {
  const int chunk = ransac_pick_chunk(nhyp);
  HPUSM_REQUIRE(chunk > 0, HPUSM_ERR_UNSUPPORTED, "ransac: %d hypotheses do not fit in shared memory", nhyp);
  const size_t smem = ransac_smem_bytes(nhyp, chunk);
  HPUSM_CUDA_OK(cudaFuncSetAttribute(ransac_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HPUSM_LAUNCH_TIMED(ransac_score_kernel, (unsigned)npairs, RS_THREADS, smem, st, src, dst, pair_offsets, thresh, nhyp,
               confidence, seed, pair0, chunk, counts, best_hyp, best_count);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

}  // namespace hpusm

using namespace hpusm;

extern "C" {

size_t hpusm_ransac_homography_workspace_bytes(int64_t total_matches, int64_t npairs, int nhyp) {
  (void)total_matches; (void)nhyp;
  return align_up((size_t)(npairs > 0 ? npairs : 1) * sizeof(int32_t), 256) * 2;
}

int hpusm_ransac_score_batch_at(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs,
                                int64_t total_matches, float thresh, int nhyp, uint64_t seed, int64_t pair_index0,
                                int32_t* counts, void* stream) {
  HPUSM_REQUIRE(npairs >= 0 && total_matches >= 0 && nhyp > 0 && pair_index0 >= 0, HPUSM_ERR_INVALID,
                "hpusm_ransac_score_batch: bad sizes");
  if (npairs == 0) return HPUSM_OK;
  HPUSM_REQUIRE(pair_offsets && counts && ((src && dst) || total_matches == 0), HPUSM_ERR_INVALID,
                "hpusm_ransac_score_batch: null pointer");
  HPUSM_REQUIRE(npairs < 0x40000000, HPUSM_ERR_UNSUPPORTED, "hpusm_ransac_score_batch: too many pairs for one call");
  return launch_score(src, dst, pair_offsets, npairs, thresh, nhyp, 0.0, seed, pair_index0, counts, nullptr, nullptr,
                      as_stream(stream));
}

int hpusm_ransac_score_batch(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs,
                             int64_t total_matches, float thresh, int nhyp, uint64_t seed, int32_t* counts,
                             void* stream) {
  return hpusm_ransac_score_batch_at(src, dst, pair_offsets, npairs, total_matches, thresh, nhyp, seed, 0, counts, stream);
}

int hpusm_ransac_homography_batch_at(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs,
                                     int64_t total_matches, float thresh, int nhyp, double confidence, uint64_t seed,
                                     int64_t pair_index0, int refine, double* H, uint8_t* mask, int32_t* ninl,
                                     int32_t* best_hyp, void* ws, size_t ws_bytes, void* stream) {
  HPUSM_REQUIRE(npairs >= 0 && total_matches >= 0 && nhyp > 0 && pair_index0 >= 0, HPUSM_ERR_INVALID,
                "hpusm_ransac_homography_batch: bad sizes");
  if (npairs == 0) return HPUSM_OK;
  HPUSM_REQUIRE(pair_offsets && H && ninl && (mask || total_matches == 0) && ((src && dst) || total_matches == 0),
                HPUSM_ERR_INVALID, "hpusm_ransac_homography_batch: null pointer");
  HPUSM_REQUIRE(npairs < 0x40000000, HPUSM_ERR_UNSUPPORTED, "hpusm_ransac_homography_batch: too many pairs");
  const size_t part = align_up((size_t)npairs * sizeof(int32_t), 256);
  HPUSM_REQUIRE(ws && is_aligned(ws, 256) && ws_bytes >= 2 * part, HPUSM_ERR_WORKSPACE,
                "hpusm_ransac_homography_batch: workspace needs %zu bytes, got %zu", 2 * part, ws_bytes);
  cudaStream_t st = as_stream(stream);
  int32_t* bh = best_hyp ? best_hyp : reinterpret_cast<int32_t*>(ws);
  int32_t* bc = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws) + part);
  int rc = launch_score(src, dst, pair_offsets, npairs, thresh, nhyp, confidence, seed, pair_index0, nullptr, bh, bc, st);
  if (rc != HPUSM_OK) return rc;
  HPUSM_LAUNCH(ransac_finalize_kernel, (unsigned)npairs, RF_THREADS, 0, st, src, dst, pair_offsets, thresh, seed,
               pair_index0, refine, bh, H, mask, ninl);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_ransac_homography_batch(const float* src, const float* dst, const int64_t* pair_offsets, int64_t npairs,
                                  int64_t total_matches, float thresh, int nhyp, double confidence, uint64_t seed,
                                  int refine, double* H, uint8_t* mask, int32_t* ninl, int32_t* best_hyp, void* ws,
                                  size_t ws_bytes, void* stream) {
  return hpusm_ransac_homography_batch_at(src, dst, pair_offsets, npairs, total_matches, thresh, nhyp, confidence, seed, 0,
                                          refine, H, mask, ninl, best_hyp, ws, ws_bytes, stream);
}

}  // extern "C"
