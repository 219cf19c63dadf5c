// K1q (integer tensor-core engine for GENERAL float descriptors): 15-bit fixed-point operands split into two signed
// bytes, tcgen05 kind::i8 with TWO s32 accumulators per tile, top-M candidates per list, exact FP32 re-rank with a proof
// of completeness, exact brute force for the (rare) rows the proof cannot cover.  sm_100a only.
//
// Replaces cv2.BFMatcher(cv2.NORM_L2).knnMatch(..., k=2) at reference urbanscene/features.py:59 for descriptors that are
// not small integers (SURF, RootSIFT, re-scaled SIFT ...).  The bf16 split engine (l2_tc_split.cu) needs 25 kind::f16
// MMAs of K = 16 per tile of pairs; this engine needs 13 kind::i8 MMAs of K = 32 (same pipe time per MMA).
//
//   x = delta X + e,  X = round(x / delta) in [-16256, 16256],  X = 128 H + L,  H in [-127, 127],  L in [-64, 63]
//   delta = max(max|x| / 16256, sqrt(max||t||^2 / 1.55e10))                (one scale for queries and database)
//   Xq.Xt = 16384 Hq.Ht + 128 (Hq.Lt + Lq.Ht) + Lq.Lt
//   a1 = Hq.Ht - N1   (4 MMAs + one augmentation step;  N1 = round(||t||^2 / (32768 delta^2)) <= 480000)
//   a2 = Hq.Lt + Lq.Ht                                                      (8 MMAs)
//   key = 128 a1 + a2 = (Xq.Xt - Lq.Lt) / 128 - 128 N1     and     256 delta^2 key  ~  2 q.t - ||t||^2 = K
//
// Error of the key against the true K, all terms bounded with quantities MEASURED on the data (prep kernels):
//   K - 256 delta^2 key = 2 delta Xq.et + 2 delta Xt.eq + 2 eq.et + 2 delta^2 Lq.Lt + (||t||^2 - 32768 delta^2 N1)
//   |.| <= 2 (|q| + rq) RT + 2 (TN + RT) rq + 2 rq RT + 2 lq LT + 20000 delta^2
//   rq = ||eq||, lq = delta ||Lq|| for this query;  RT, LT, TN = the maxima of ||et||, delta ||Lt||, ||t|| over the database
// The epilogue keeps the Q_M largest keys per query row, database split and 32-column group (exact integer compares, so
// the lists are exact top-M lists of the KEY).  l2_q16_finalize_kernel computes the exact FP32 distance of every listed
// row, takes the lexicographic (d2, index) top-2 and checks
//      d2(second) + E(q) < ||q||^2 - 256 delta^2 max_lists key(M-th)
// which proves that no row outside the lists is as close as the second neighbour.  Rows failing the check are redone by
// the exact FP32 engine (l2_simt.cu) after one host read of their count, so the answer is the exact FP32 one; when there
// are many of them (clustered data) they first go through the bf16 split engine, whose bound is 20 x tighter.
//
// Kernel shape: l2_tc_i8.cu's (persistent CTA per SM, TMA ring, one elected MMA issuer, query rows in TENSOR memory, three
// accumulator sets in rotation, 16 epilogue warps), with N = 64 tiles: a set = [a1 | a2] = 2 x 64 columns, the 2 x 64
// query columns [H | L] of the two query tiles fill the rest of the 512.  The (row-independent) augmentation columns of A
// live in shared memory as in hamming_tc.cu.
#include <cuda.h>
#include <limits.h>
#include <float.h>
#include <stdlib.h>
#include "common.cuh"
#include "l2_internal.cuh"
#include "tc_ptx.cuh"

namespace hpusm {

constexpr int Q_D = 128;
constexpr int Q_QB = 256;             // queries per unit (two M = 128 query tiles)
constexpr int Q_TN = 64;              // database rows per tile (MMA N)
constexpr int Q_STAGES = 8;
constexpr int Q_THREADS = 576;        // TMA producer / TMEM allocator, MMA issuer, 16 epilogue warps
constexpr int Q_CGS = 2;              // column groups of 32 per tile (one per epilogue warp of a lane quarter)
constexpr int Q_NBUF = 3;             // accumulator sets in rotation
constexpr int Q_M = 6;                // candidates kept per list
constexpr int Q_KQ = 256;             // query operand row: H 128 | L 128 (s8)
constexpr int Q_KT = 288;             // database operand row: H 128 | L 128 | 32 augmentation bytes
constexpr int Q_A_COLS = 64;          // a query tile in tensor memory: 256 bytes per row
constexpr int Q_ACC_COL0 = 128;       // accumulator set b: a1 at 128 + 128 b, a2 at 128 + 128 b + 64
constexpr int Q_SLAB = Q_TN * 128;    // [64 rows][128 B] SWIZZLE_128B slab, 8 KB
constexpr int Q_BAUG = Q_TN * 32;     // [64 rows][32 B] SWIZZLE_32B slab, 2 KB
constexpr int Q_B_BYTES = 2 * Q_SLAB + Q_BAUG;   // 18 KB per stage
constexpr int Q_AAUG_BYTES = 128 * 32;
constexpr int Q_SMEM_BYTES = Q_STAGES * Q_B_BYTES + Q_AAUG_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr uint32_t Q_SLEEP_NS = 20000;
constexpr int Q_XMAX = 16256;         // 127 * 128
constexpr int Q_N1_MAX = 480000;      // the 30 weighted augmentation digits hold 127 * 128 * 30 = 487680
constexpr float Q_TSCALE = 1.55e10f;  // 32768 * Q_N1_MAX, rounded down
constexpr int Q_KEY_NONE = -(1 << 30);   // below every real key (|key| < 3.4e8), far enough from INT_MIN to subtract from
constexpr int64_t Q_FB_ROWS = 65536;

// Instruction descriptor: D = S32, A = B = S8, both K-major, N = 64, M = 128.
constexpr uint32_t Q_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Q_TN >> 3) << 17) | ((128u >> 4) << 24);

// Device-resident parameters and statistics of one call (no host round trip).
struct QParams {
  int amax_bits;   // max |x| over queries and database (float bits; non-negative floats order like ints)
  int tmax_bits;   // max ||t||^2
  int rt2_bits;    // max ||e_t||^2
  int lt2;         // max ||L_t||^2 (integer)
  int bad;         // a non-finite value or an augmentation overflow: every row takes the exact engine
  int n_list;      // rows the proof did not cover
  float delta, inv_delta;
};

// ---- operand preparation ------------------------------------------------------------------------------------------
// pass 1: max |x| (both arrays) and max ||t||^2 (database).  One warp per row, one atomic per block.
__global__ void __launch_bounds__(256)
l2_q16_stats_kernel(const float* __restrict__ x, int64_t n, int is_db, QParams* __restrict__ P) {
  __shared__ float s_a[8], s_t[8];
  __shared__ int s_bad;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  const int64_t row = (int64_t)blockIdx.x * 8 + w;
  float a = 0.f, s = 0.f;
  bool bad = false;
  if (row < n) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * Q_D) + lane);
    const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bad |= !(fabsf(f[i]) <= FLT_MAX);
      a = fmaxf(a, fabsf(f[i]));
      s = fmaf(f[i], f[i], s);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  bad = __any_sync(0xffffffffu, bad) || !(s <= FLT_MAX);
  if (lane == 0) { s_a[w] = a; s_t[w] = s; if (bad) s_bad = 1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float am = 0.f, tm = 0.f;
    for (int i = 0; i < 8; ++i) { am = fmaxf(am, s_a[i]); tm = fmaxf(tm, s_t[i]); }
    if (s_bad) atomicOr(&P->bad, 1);
    else {
      atomicMax(&P->amax_bits, __float_as_int(am));
      if (is_db) atomicMax(&P->tmax_bits, __float_as_int(tm));
    }
  }
}

__global__ void l2_q16_params_kernel(QParams* __restrict__ P) {
  const float amax = __int_as_float(P->amax_bits), tmax = __int_as_float(P->tmax_bits);
  float d = fmaxf(amax / (float)Q_XMAX, sqrtf(tmax / Q_TSCALE)) * 1.00001f;
  // magnitudes whose squares leave the normal FP32 range: the FP32 norms the bound is built from would be wrong
  if (amax > 0.f && (amax < 1e-18f || amax > 1e18f)) P->bad = 1;
  if (!(d > 0.f) || !(d <= FLT_MAX)) d = 1.f;   // all-zero input (or flagged bad): any scale will do
  if (d < 1e-30f) d = 1e-30f;                   // keep 1 / delta^2 finite
  P->delta = d;
  P->inv_delta = 1.f / d;
}

// The fixed-point value of one component and its rounding residue; used by the prep kernel AND by the finalize kernel
// (which recomputes the query's residue norms), so both see the same numbers.
__device__ __forceinline__ int q16_quantise(float x, float delta, float inv_delta, float& e) {
  float xf = rintf(x * inv_delta);
  xf = fminf(fmaxf(xf, -(float)Q_XMAX), (float)Q_XMAX);
  e = fmaf(-xf, delta, x);
  return (int)xf;
}

// pass 2: one warp per row, float32 -> [H | L] (+ the augmentation digits of -N1 on the database side).
// Augmentation step: A = [1, 127 x 15 | 1, 127 x 15] (both 16-byte halves equal, which makes the SWIZZLE_32B placement of
// the shared-memory copy moot), B = [-r, d1 .. d15 | 0, d16 .. d30] with N1 = r + 127 (d1 + ... + d30), 0 <= d <= 128.
__global__ void __launch_bounds__(256)
l2_q16_quant_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int is_db, int8_t* __restrict__ out,
                    QParams* __restrict__ P) {
  __shared__ float s_r[8];
  __shared__ int s_l[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 8 + w;
  const float delta = P->delta, inv = P->inv_delta;
  const int K = is_db ? Q_KT : Q_KQ;
  float e2 = 0.f, s = 0.f;
  int l2 = 0;
  uint32_t hw = 0, lw = 0;
  if (row < n) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * Q_D) + lane);
    const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float e;
      const int X = q16_quantise(f[i], delta, inv, e);
      const int H = (X + 64) >> 7, L = X - 128 * H;
      hw |= (uint32_t)(H & 0xff) << (8 * i);
      lw |= (uint32_t)(L & 0xff) << (8 * i);
      e2 = fmaf(e, e, e2);
      l2 += L * L;
      s = fmaf(f[i], f[i], s);
    }
  }
  if (row < n_pad) {
    reinterpret_cast<uint32_t*>(out + row * K)[lane] = hw;
    reinterpret_cast<uint32_t*>(out + row * K + Q_D)[lane] = lw;
  }
  if (!is_db) return;
  for (int o = 16; o > 0; o >>= 1) {
    e2 += __shfl_xor_sync(0xffffffffu, e2, o);
    l2 += __shfl_xor_sync(0xffffffffu, l2, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  bool overflow = false;
  if (row < n_pad) {
    int n1 = 0;
    if (row < n) {
      const float n1f = rintf(s * inv * inv * (1.f / 32768.f));
      overflow = !(n1f <= (float)Q_N1_MAX);
      n1 = overflow ? 0 : (int)n1f;
    }
    const int r = n1 % 127, m = n1 / 127;
    int d;
    if (lane == 0) d = -r;
    else if (lane == 16) d = 0;
    else {
      const int slot = lane < 16 ? lane - 1 : lane - 2;
      int v = m - 128 * slot;
      v = v < 0 ? 0 : (v > 128 ? 128 : v);
      d = -v;
    }
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) word |= ((uint32_t)__shfl_sync(0xffffffffu, d, (4 * lane + i) & 31) & 0xffu) << (8 * i);
    if (lane < 8) reinterpret_cast<uint32_t*>(out + row * K + 2 * Q_D)[lane] = word;
  }
  if (lane == 0) { s_r[w] = e2; s_l[w] = l2; if (overflow) atomicOr(&P->bad, 1); }
  __syncthreads();
  if (threadIdx.x == 0) {
    float rm = 0.f;
    int lm = 0;
    for (int i = 0; i < 8; ++i) { rm = fmaxf(rm, s_r[i]); lm = max(lm, s_l[i]); }
    atomicMax(&P->rt2_bits, __float_as_int(rm));
    atomicMax(&P->lt2, lm);
  }
}

// ---- the contraction + fused top-M -----------------------------------------------------------------------------------
struct QList {
  int k[Q_M];   // keys, largest first; Q_KEY_NONE = empty
  int i[Q_M];
};

__device__ __forceinline__ int q_imax3(int a, int b, int c) { return max(max(a, b), c); }

// caller guarantees k > t.k[Q_M - 1]; equal keys keep the earlier entry in front
__device__ __forceinline__ void q16_insert(QList& t, int k, int col) {
  t.k[Q_M - 1] = k; t.i[Q_M - 1] = col;
#pragma unroll
  for (int j = Q_M - 1; j > 0; --j) {
    if (t.k[j] > t.k[j - 1]) {
      const int tk = t.k[j]; t.k[j] = t.k[j - 1]; t.k[j - 1] = tk;
      const int ti = t.i[j]; t.i[j] = t.i[j - 1]; t.i[j - 1] = ti;
    }
  }
}

template <int OFF>
__device__ __forceinline__ int q16_max16(const int (&v)[32]) {
  const int a = q_imax3(v[OFF + 0], v[OFF + 1], v[OFF + 2]), b = q_imax3(v[OFF + 3], v[OFF + 4], v[OFF + 5]);
  const int c = q_imax3(v[OFF + 6], v[OFF + 7], v[OFF + 8]), d = q_imax3(v[OFF + 9], v[OFF + 10], v[OFF + 11]);
  const int e = q_imax3(v[OFF + 12], v[OFF + 13], v[OFF + 14]);
  return max(q_imax3(a, b, c), q_imax3(d, e, v[OFF + 15]));
}

// The keys of 16 columns that beat the list's threshold, pulled out one at a time, largest first and the lower column
// first among equals, by a max tree over ((key - base) << 4 | 15 - j): a short loop instead of 16 unrolled insertions
// (l2_tc_i8.cu has the measurement behind that).  Keys more than 2^27 below the chunk's maximum cannot be packed; if
// the list's threshold ends up below that window and a column lies in between, it went unexamined and the row is marked
// `dirty` (its proof is made to fail, the exact engine redoes it).
template <int OFF>
__device__ __forceinline__ void q16_slow16(const int (&key)[32], int hi, int col, QList& t, bool& dirty) {
  const int base = max(t.k[Q_M - 1], hi - ((1 << 27) - 2));   // packed values stay below `last`'s initial INT_MAX
  int last = INT_MAX;
#pragma unroll 1
  for (;;) {
    int pk[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int v = key[OFF + j] - base;
      const int p = v > 0 ? ((v << 4) | (15 - j)) : -1;
      pk[OFF + j] = p < last ? p : -1;
    }
    const int best = q16_max16<OFF>(pk);
    if (best < 0) break;
    const int kb = (best >> 4) + base;
    if (kb <= t.k[Q_M - 1]) break;
    q16_insert(t, kb, col + OFF + 15 - (best & 15));
    last = best;
  }
  if (base > t.k[Q_M - 1]) {   // the list is still short of the window: did a column fall between the two?
#pragma unroll
    for (int j = 0; j < 16; ++j) dirty |= key[OFF + j] > t.k[Q_M - 1] && key[OFF + j] <= base;
  }
}

// cand [splits * Q_CGS][nq][Q_M] (-1 = empty); kth [splits * Q_CGS][nq]: the list's M-th key = bound on every row of
// the list's columns that is NOT listed (INT_MIN: all of them are listed, INT_MAX: unknown).
// DBG 0 = product.  Experiment switches (HPUSM_Q16_DBG, wrong answers, used by tools/q16_time.py to decompose the step):
// 1 = hand-over + tensor-memory loads but no scan, 2 = hand-over only, 10 = the MMA warp runs free (pace of MMA + TMA alone).
template <int DBG>
__global__ void __launch_bounds__(Q_THREADS, 1)
l2_q16_knn_kernel(const int8_t* __restrict__ qu, const __grid_constant__ CUtensorMap map_t,
                  const __grid_constant__ CUtensorMap map_ta, int64_t nq, int64_t nt, int q_blocks, int splits,
                  int64_t rows_per_split, int32_t* __restrict__ cand, int32_t* __restrict__ kth) {
  extern __shared__ uint8_t q16_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(q16_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem;                                   // [stages][8 KB H | 8 KB L | 2 KB aug]
  uint8_t* smem_aaug = smem + Q_STAGES * Q_B_BYTES;         // [128 rows][32 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aaug + Q_AAUG_BYTES);
  uint64_t* full = bars;                                    // [stages]  TMA -> MMA
  uint64_t* empty = bars + Q_STAGES;                        // [stages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * Q_STAGES;                // [NBUF]    MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * Q_STAGES + Q_NBUF;      // [NBUF]    epilogue -> MMA
  uint64_t* a_ready = bars + 2 * Q_STAGES + 2 * Q_NBUF;
  uint64_t* a_free = bars + 2 * Q_STAGES + 2 * Q_NBUF + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Q_STAGES + 2 * Q_NBUF + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Q_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int b = 0; b < Q_NBUF; ++b) { mbar_init(tmem_full + b, 1); mbar_init(tmem_empty + b, 8 * 32); }
    mbar_init(a_ready, 16);
    mbar_init(a_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // augmentation columns of the A operand, the same for every query row: [1, 127 x 15] in both 16-byte halves
  for (int i = threadIdx.x; i < Q_AAUG_BYTES / 4; i += Q_THREADS)
    reinterpret_cast<uint32_t*>(smem_aaug)[i] = (i & 3) == 0 ? 0x7F7F7F01u : 0x7F7F7F7Fu;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_units = q_blocks * splits;

  if (warp == 0) {
    // ===== TMA producer: the ring of database tiles =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_t) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ta) : "memory");
      uint32_t it = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int split = u / q_blocks;
        const int64_t t_begin = (int64_t)split * rows_per_split;
        const int64_t t_end = min(nt, t_begin + rows_per_split);
        const int tiles = (int)((t_end - t_begin + Q_TN - 1) / Q_TN);
        for (int j = 0; j < tiles; ++j, ++it) {
          const int s = it % Q_STAGES;
          mbar_wait_sleepy(empty + s, ((it / Q_STAGES) & 1) ^ 1, Q_SLEEP_NS);
          mbar_expect_tx(full + s, Q_B_BYTES);
          const int row0 = (int)(t_begin + (int64_t)j * Q_TN);
          uint8_t* bs = smem_b + s * Q_B_BYTES;
          tma_load_2d(bs, &map_t, full + s, 0, row0);
          tma_load_2d(bs + Q_SLAB, &map_t, full + s, Q_D, row0);
          tma_load_2d(bs + 2 * Q_SLAB, &map_ta, full + s, 2 * Q_D, row0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane, see l2_tc_i8.cu) =====
    uint32_t it = 0, un = 0, buf = 0, buf_phase = 0;
    const uint32_t b_addr = smem_u32(smem_b);
    const uint64_t adesc_aug = umma_desc_sw32(smem_u32(smem_aaug));
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++un) {
      const int split = u / q_blocks;
      const int64_t t_begin = (int64_t)split * rows_per_split;
      const int64_t t_end = min(nt, t_begin + rows_per_split);
      const int tiles = (int)((t_end - t_begin + Q_TN - 1) / Q_TN);
      mbar_wait(a_ready, un & 1);
      tc_fence_after();
      for (int j = 0; j < tiles; ++j, ++it) {
        const int s = it % Q_STAGES;
        mbar_wait_sleepy(full + s, (it / Q_STAGES) & 1, Q_SLEEP_NS);
        const uint64_t bdesc_h = umma_desc_sw128(b_addr + s * Q_B_BYTES);
        const uint64_t bdesc_l = umma_desc_sw128(b_addr + s * Q_B_BYTES + Q_SLAB);
        const uint64_t bdesc_aug = umma_desc_sw32(b_addr + s * Q_B_BYTES + 2 * Q_SLAB);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          if (!(DBG & 8)) mbar_wait_sleepy(tmem_empty + buf, buf_phase ^ 1, Q_SLEEP_NS);   // the epilogue has read this set out
          tc_fence_after();
          const uint32_t d1 = tmem_base + (uint32_t)(Q_ACC_COL0 + buf * 2 * Q_TN);
          const uint32_t d2 = d1 + Q_TN;
          const uint32_t a_h = tmem_base + (uint32_t)(m * Q_A_COLS), a_l = a_h + 32;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)   // a1 = Hq.Ht   (+2 in the descriptor's address field = 32 bytes along K)
              tc_mma_i8_ts(d1, a_h + k * 8, bdesc_h + (uint64_t)(2 * k), Q_IDESC, k > 0 ? 1u : 0u);
            tc_mma_i8_ss(d1, adesc_aug, bdesc_aug, Q_IDESC, 1u);           // a1 -= N1
#pragma unroll
            for (int k = 0; k < 4; ++k)   // a2 = Hq.Lt
              tc_mma_i8_ts(d2, a_h + k * 8, bdesc_l + (uint64_t)(2 * k), Q_IDESC, k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)   // a2 += Lq.Ht
              tc_mma_i8_ts(d2, a_l + k * 8, bdesc_h + (uint64_t)(2 * k), Q_IDESC, 1u);
            tc_commit(tmem_full + buf);
          }
          __syncwarp();
          if (++buf == Q_NBUF) { buf = 0; buf_phase ^= 1; }
        }
        if (elect_one()) tc_commit(empty + s);
        __syncwarp();
      }
      if (elect_one()) tc_commit(a_free);
      __syncwarp();
    }
  } else {
    // ===== epilogue: fused top-M on the integer key =====
    // warps 2-9 serve query tile 0, warps 10-17 tile 1; inside a tile two warps per TMEM lane quarter, 32 columns each.
    // piece p = 2 * tile + m lives in set p % 3; for fixed m the sets come round as 0,2,1 (m = 0) or 1,0,2 (m = 1) and
    // the barrier phase depends on the set only (l2_tc_i8.cu).
    const int ew = warp - 2, quarter = warp & 3, cg = (ew >> 2) & 1, m = ew >> 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t un = 0, buf = (uint32_t)m;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++un) {
      const int split = u / q_blocks, qb = u - split * q_blocks;
      const int64_t t_begin = (int64_t)split * rows_per_split;
      const int64_t t_end = min(nt, t_begin + rows_per_split);
      const int tiles = (int)((t_end - t_begin + Q_TN - 1) / Q_TN);
      const int64_t row = (int64_t)qb * Q_QB + m * 128 + quarter * 32 + lane;
      // -- the unit's query rows go into tensor memory: thread = row; this warp stores words [32 cg, 32 cg + 32) --
      {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = 0u;
        if (row < nq) {
          const uint4* src = reinterpret_cast<const uint4*>(qu + row * Q_KQ) + cg * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 v = __ldg(src + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
          }
        }
        mbar_wait(a_free, (un & 1) ^ 1);   // the previous unit's MMAs no longer read the old rows
        tc_fence_after();
        const uint32_t a_addr = lane_base + (uint32_t)(m * Q_A_COLS + cg * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t r8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) r8[i] = w[8 * g + i];
          tc_st8(a_addr + 8 * g, r8);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      }
      QList top;
#pragma unroll
      for (int j = 0; j < Q_M; ++j) { top.k[j] = Q_KEY_NONE; top.i[j] = -1; }
      bool dirty = false;
      for (int j = 0; j < ((DBG & 8) ? 0 : tiles); ++j) {   // DBG 8: the MMA warp runs free (pace of MMA + TMA alone)
        const int col0 = (int)(t_begin + (int64_t)j * Q_TN) + cg * 32;
        uint32_t a1[32], a2[32];
        const uint32_t phase = m == 0 ? (buf == 1u ? 1u : 0u) : (buf == 1u ? 0u : 1u);
        const uint32_t taddr = lane_base + (uint32_t)(Q_ACC_COL0 + buf * 2 * Q_TN + cg * 32);
        mbar_wait_sleepy(tmem_full + buf, phase, Q_SLEEP_NS);
        tc_fence_after();
        if ((DBG & 3) != 2) {
          tc_ld32(taddr, a1);
          tc_ld32(taddr + Q_TN, a2);
          tc_wait_ld();
        }
        tc_fence_before();
        mbar_arrive(tmem_empty + buf);   // every lane arrives (l2_tc_i8.cu)
        buf = buf == 0 ? 2u : buf - 1u;  // two pieces on: (buf + 2) % 3
        if ((DBG & 3) != 0) {
          if ((DBG & 3) == 1 && (a1[0] ^ a2[5]) == 0x12345678u) dirty = true;
          continue;
        }
        int key[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) key[i] = (int)a1[i] * 128 + (int)a2[i];
        if (col0 + 32 > (int)t_end) {    // the ragged tail of the database: columns past the last row never compete
#pragma unroll
          for (int i = 0; i < 32; ++i) key[i] = col0 + i < (int)t_end ? key[i] : Q_KEY_NONE;
        }
        const int h0 = q16_max16<0>(key), h1 = q16_max16<16>(key);
        if (max(h0, h1) > top.k[Q_M - 1]) {
          if (h0 > top.k[Q_M - 1]) q16_slow16<0>(key, h0, col0, top, dirty);
          if (h1 > top.k[Q_M - 1]) q16_slow16<16>(key, h1, col0, top, dirty);
        }
      }
      if (row < nq) {
        const int64_t part = (int64_t)(split * Q_CGS + cg);
        int32_t* o = cand + (part * nq + row) * Q_M;
#pragma unroll
        for (int j = 0; j < Q_M; ++j) o[j] = top.i[j];
        kth[part * nq + row] = dirty ? INT_MAX : (top.i[Q_M - 1] >= 0 ? top.k[Q_M - 1] : INT_MIN);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ---- exact re-rank + proof of completeness ---------------------------------------------------------------------------
// One warp per query: lanes take candidates (exact FP32 distance, increasing-k fma order = the number the FP32 engine
// and the final answer use), warp-wide lexicographic top-2, then the completeness check.  Unproven rows go to `list`.
__global__ void __launch_bounds__(256)
l2_q16_finalize_kernel(const float* __restrict__ q, int64_t nq, const float* __restrict__ t, const int32_t* __restrict__ cand,
                       const int32_t* __restrict__ kth, int nparts, QParams* __restrict__ P, int32_t* __restrict__ idx,
                       float* __restrict__ dist, int32_t* __restrict__ list) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= nq) return;
  const float delta = P->delta, inv = P->inv_delta;
  const float4 qv = __ldg(reinterpret_cast<const float4*>(q + row * Q_D) + lane);
  const float f[4] = {qv.x, qv.y, qv.z, qv.w};
  float qn = 0.f, rq2 = 0.f;
  int lq2 = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float e;
    const int X = q16_quantise(f[i], delta, inv, e);
    const int H = (X + 64) >> 7, L = X - 128 * H;
    qn = fmaf(f[i], f[i], qn);
    rq2 = fmaf(e, e, rq2);
    lq2 += L * L;
  }
  for (int o = 16; o > 0; o >>= 1) {
    qn += __shfl_xor_sync(0xffffffffu, qn, o);
    rq2 += __shfl_xor_sync(0xffffffffu, rq2, o);
    lq2 += __shfl_xor_sync(0xffffffffu, lq2, o);
  }
  Top2<float> best;
  best.d0 = best.d1 = 0.f; best.i0 = best.i1 = -1;
  int kmax = INT_MIN;
  const int C = nparts * Q_M;
  for (int c = lane; c < C; c += 32) {
    const int p = c / Q_M;
    const int32_t ci = cand[((int64_t)p * nq + row) * Q_M + (c - p * Q_M)];
    if (ci >= 0) {
      const float4* a = reinterpret_cast<const float4*>(q + row * Q_D);
      const float4* b = reinterpret_cast<const float4*>(t + (int64_t)ci * Q_D);
      float s = 0.f;
      for (int kv = 0; kv < Q_D / 4; ++kv) {
        const float4 x = __ldg(a + kv), y = __ldg(b + kv);
        float df;
        df = __fsub_rn(x.x, y.x); s = __fmaf_rn(df, df, s);
        df = __fsub_rn(x.y, y.y); s = __fmaf_rn(df, df, s);
        df = __fsub_rn(x.z, y.z); s = __fmaf_rn(df, df, s);
        df = __fsub_rn(x.w, y.w); s = __fmaf_rn(df, df, s);
      }
      top2_insert<float>(best, s, ci);
    }
  }
  for (int p = lane; p < nparts; p += 32) kmax = max(kmax, kth[(int64_t)p * nq + row]);
  for (int o = 16; o > 0; o >>= 1) {
    Top2<float> other;
    other.d0 = __shfl_xor_sync(0xffffffffu, best.d0, o); other.d1 = __shfl_xor_sync(0xffffffffu, best.d1, o);
    other.i0 = __shfl_xor_sync(0xffffffffu, best.i0, o); other.i1 = __shfl_xor_sync(0xffffffffu, best.i1, o);
    top2_merge<float>(best, other);
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  if (lane == 0) {
    bool proven;
    if (P->bad || kmax == INT_MAX) proven = false;
    else if (kmax == INT_MIN) proven = true;       // every row of every list's columns is listed
    else {
      const double dl = (double)delta, tmax = (double)__int_as_float(P->tmax_bits);
      const double qnd = (double)qn, qnorm = sqrt(qnd), tn = sqrt(tmax);
      const double rq = sqrt((double)rq2), rt = sqrt((double)__int_as_float(P->rt2_bits));
      const double lq = dl * sqrt((double)lq2), lt = dl * sqrt((double)P->lt2);
      const double E = 1.01 * (2.0 * (qnorm + rq) * rt + 2.0 * (tn + rt) * rq + 2.0 * rq * rt + 2.0 * lq * lt) +
                       20000.0 * dl * dl + 4e-5 * (qnd + tmax) + 4e-5 * qnorm * tn;
      // rows outside the lists have d2 >= qn - 256 delta^2 kmax - E; the listed second neighbour must beat that strictly
      proven = best.i1 >= 0 && (double)best.d1 + E < qnd - 256.0 * dl * dl * (double)kmax;
    }
    if (proven) {
      idx[row * 2] = best.i0; idx[row * 2 + 1] = best.i1;
      dist[row * 2] = best.i0 < 0 ? 0.f : __fsqrt_rn(best.d0);
      dist[row * 2 + 1] = best.i1 < 0 ? 0.f : __fsqrt_rn(best.d1);
    } else {
      list[atomicAdd(&P->n_list, 1)] = (int32_t)row;
    }
  }
}

// ---- cascade: many unproven rows go through the bf16 split engine (20 x tighter bound) before the exact engine --------------
__global__ void __launch_bounds__(256)
l2_q16_gather_rows_kernel(const float* __restrict__ q, const int32_t* __restrict__ list, int64_t m, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= m) return;
  reinterpret_cast<float4*>(out + i * Q_D)[lane] = __ldg(reinterpret_cast<const float4*>(q + (int64_t)list[i] * Q_D) + lane);
}

__global__ void __launch_bounds__(256)
l2_q16_scatter_rows_kernel(const int32_t* __restrict__ list, int64_t m, const int32_t* __restrict__ idx_sub,
                           const float* __restrict__ dist_sub, int32_t* __restrict__ idx, float* __restrict__ dist) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  const int64_t row = list[i];
  idx[row * 2] = idx_sub[i * 2]; idx[row * 2 + 1] = idx_sub[i * 2 + 1];
  dist[row * 2] = dist_sub[i * 2]; dist[row * 2 + 1] = dist_sub[i * 2 + 1];
}

// ---- host side ---------------------------------------------------------------------------------------------------------
static int q16_splits(int64_t nq, int64_t nt) {
  const int64_t qbs = (nq + Q_QB - 1) / Q_QB;
  const int64_t sms = sm_count();
  if (qbs >= 2 * sms) return 1;
  int64_t s = (2 * sms + qbs - 1) / qbs;
  const int64_t max_s = (nt + 2047) / 2048;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 32) s = 32;
  // every split must own at least one row after its length is rounded up to whole tiles
  while (s > 1) {
    const int64_t rps = ((nt + s - 1) / s + Q_TN - 1) / Q_TN * Q_TN;
    if ((s - 1) * rps < nt) break;
    --s;
  }
  return (int)s;
}

constexpr int64_t Q_CASCADE_MIN = 1024;   // unproven rows from which the split engine is worth its database preparation
constexpr int64_t Q_CASCADE_ROWS = 65536;  // rows per pass of the cascade

struct Q16Layout {
  size_t qb, tb, par, cand, kth, list, fb, sub_q, sub_idx, sub_dist, sub_ws, sub_ws_bytes, total;
  int splits, fb_splits;
  int64_t nt_pad, sub_rows;
};

static Q16Layout q16_layout(int64_t nq, int64_t nt) {
  Q16Layout L;
  L.nt_pad = (nt + Q_TN - 1) / Q_TN * Q_TN;
  L.splits = q16_splits(nq, nt);
  L.fb_splits = l2_simt_splits(64, nt);
  size_t off = 0;
  L.qb = off; off += align_up((size_t)nq * Q_KQ, 1024);
  L.tb = off; off += align_up((size_t)L.nt_pad * Q_KT, 1024);
  L.par = off; off += 1024;
  L.cand = off; off += align_up((size_t)L.splits * Q_CGS * nq * Q_M * sizeof(int32_t), 1024);
  L.kth = off; off += align_up((size_t)L.splits * Q_CGS * nq * sizeof(int32_t), 1024);
  L.list = off; off += align_up((size_t)nq * sizeof(int32_t), 1024);
  L.fb = off; off += align_up((size_t)L.fb_splits * (nq < Q_FB_ROWS ? nq : Q_FB_ROWS) * 2 * sizeof(int32_t), 1024);
  // cascade buffers: only when the query set can produce that many unproven rows at all
  L.sub_rows = nq >= Q_CASCADE_MIN ? (nq < Q_CASCADE_ROWS ? nq : Q_CASCADE_ROWS) : 0;
  L.sub_q = L.sub_idx = L.sub_dist = L.sub_ws = off;
  L.sub_ws_bytes = 0;
  if (L.sub_rows > 0) {
    L.sub_q = off; off += align_up((size_t)L.sub_rows * Q_D * sizeof(float), 1024);
    L.sub_idx = off; off += align_up((size_t)L.sub_rows * 2 * sizeof(int32_t), 1024);
    L.sub_dist = off; off += align_up((size_t)L.sub_rows * 2 * sizeof(float), 1024);
    L.sub_ws_bytes = align_up(l2_tc_split_workspace_bytes(L.sub_rows, nt), 1024);
    L.sub_ws = off; off += L.sub_ws_bytes;
  }
  L.total = off;
  return L;
}

size_t l2_tc_q16_workspace_bytes(int64_t nq, int64_t nt) {
  if (nq <= 0 || nt <= 0) return 256;
  return q16_layout(nq, nt).total;
}

int l2_tc_q16_knn2(const float* q, int64_t nq, const float* t, int64_t nt, int32_t* idx, float* dist, void* ws, size_t ws_bytes,
                   cudaStream_t st, int64_t* fallback_rows) {
  const Q16Layout L = q16_layout(nq, nt);
  HPUSM_REQUIRE(ws_bytes >= L.total, HPUSM_ERR_WORKSPACE, "l2 q16 engine: workspace needs %zu bytes, got %zu", L.total, ws_bytes);
  char* base = reinterpret_cast<char*>(ws);
  int8_t* qb = reinterpret_cast<int8_t*>(base + L.qb);
  int8_t* tb = reinterpret_cast<int8_t*>(base + L.tb);
  QParams* P = reinterpret_cast<QParams*>(base + L.par);
  int32_t* cand = reinterpret_cast<int32_t*>(base + L.cand);
  int32_t* kth = reinterpret_cast<int32_t*>(base + L.kth);
  int32_t* list = reinterpret_cast<int32_t*>(base + L.list);
  int32_t* fb = reinterpret_cast<int32_t*>(base + L.fb);

  HPUSM_CUDA_OK(cudaMemsetAsync(P, 0, sizeof(QParams), st));
  HPUSM_LAUNCH(l2_q16_stats_kernel, (unsigned)((nq + 7) / 8), 256, 0, st, q, nq, 0, P);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(l2_q16_stats_kernel, (unsigned)((nt + 7) / 8), 256, 0, st, t, nt, 1, P);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(l2_q16_params_kernel, 1, 1, 0, st, P);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(l2_q16_quant_kernel, (unsigned)((nq + 7) / 8), 256, 0, st, q, nq, nq, 0, qb, P);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(l2_q16_quant_kernel, (unsigned)((L.nt_pad + 7) / 8), 256, 0, st, t, nt, L.nt_pad, 1, tb, P);
  HPUSM_CHECK_LAUNCH();
  CUtensorMap map_t, map_ta;
  int rc = tc_make_map(&map_t, tb, L.nt_pad, Q_KT, 128, CU_TENSOR_MAP_SWIZZLE_128B, Q_TN, 1);
  if (rc == HPUSM_OK) rc = tc_make_map(&map_ta, tb, L.nt_pad, Q_KT, 32, CU_TENSOR_MAP_SWIZZLE_32B, Q_TN, 1);
  if (rc != HPUSM_OK) return rc;
  static std::atomic<bool> attr_done[64];
  int dev = 0;
  HPUSM_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev].load(std::memory_order_acquire)) {
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_q16_knn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_BYTES));
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_q16_knn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_BYTES));
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_q16_knn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_BYTES));
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_q16_knn_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_BYTES));
    attr_done[dev].store(true, std::memory_order_release);
  }
  const int q_blocks = (int)((nq + Q_QB - 1) / Q_QB);
  int64_t rows_per_split = (nt + L.splits - 1) / L.splits;
  rows_per_split = (rows_per_split + Q_TN - 1) / Q_TN * Q_TN;
  const int units = q_blocks * L.splits;
  const int grid = units < sm_count() ? units : sm_count();
#ifdef HPUSM_DEV_SWITCHES   // development builds only (HPUSM_EXTRA_FLAGS=-DHPUSM_DEV_SWITCHES bash build.sh): tools/q16_time.py
  const char* dbg_env = getenv("HPUSM_Q16_DBG");   // experiment switches (wrong answers), see the kernel
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
#else
  const int dbg = 0;                               // the product library ignores the variable
#endif
  if (dbg == 1)
    HPUSM_LAUNCH_TIMED(l2_q16_knn_kernel<1>, grid, Q_THREADS, Q_SMEM_BYTES, st, qb, map_t, map_ta, nq, nt, q_blocks, L.splits,
                       rows_per_split, cand, kth);
  else if (dbg == 2)
    HPUSM_LAUNCH_TIMED(l2_q16_knn_kernel<2>, grid, Q_THREADS, Q_SMEM_BYTES, st, qb, map_t, map_ta, nq, nt, q_blocks, L.splits,
                       rows_per_split, cand, kth);
  else if (dbg == 10)
    HPUSM_LAUNCH_TIMED(l2_q16_knn_kernel<10>, grid, Q_THREADS, Q_SMEM_BYTES, st, qb, map_t, map_ta, nq, nt, q_blocks, L.splits,
                       rows_per_split, cand, kth);
  else
    HPUSM_LAUNCH_TIMED(l2_q16_knn_kernel<0>, grid, Q_THREADS, Q_SMEM_BYTES, st, qb, map_t, map_ta, nq, nt, q_blocks, L.splits,
                       rows_per_split, cand, kth);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(l2_q16_finalize_kernel, (unsigned)((nq + 7) / 8), 256, 0, st, q, nq, t, cand, kth, L.splits * Q_CGS, P, idx, dist,
               list);
  HPUSM_CHECK_LAUNCH();
  // the one host read of the engine: how many rows could not be proven complete
  int h_list = 0;
  HPUSM_CUDA_OK(cudaMemcpyAsync(&h_list, &P->n_list, sizeof(int), cudaMemcpyDeviceToHost, st));
  HPUSM_CUDA_OK(cudaStreamSynchronize(st));
  *fallback_rows = h_list;
  const char* casc_env = getenv("HPUSM_Q16_CASCADE");   // "0": unproven rows go straight to the exact engine (tests, A/B)
  if (h_list >= Q_CASCADE_MIN && L.sub_rows > 0 && !(casc_env && casc_env[0] == '0')) {
    // Many rows the fixed-point bound cannot cover (clustered data: neighbours closer together than E): the bf16 split
    // engine, whose bound is 20 x tighter, takes them before anything reaches the CUDA-core engine (1.3e12 against 1.5e11
    // pairs/s).  Its own unproven rows take its own exact fallback; the answers are scattered back to their rows.
    float* sub_q = reinterpret_cast<float*>(base + L.sub_q);
    int32_t* sub_idx = reinterpret_cast<int32_t*>(base + L.sub_idx);
    float* sub_dist = reinterpret_cast<float*>(base + L.sub_dist);
    int64_t exact_rows = 0;
    for (int64_t lo = 0; lo < h_list; lo += L.sub_rows) {
      const int64_t m = h_list - lo < L.sub_rows ? h_list - lo : L.sub_rows;
      HPUSM_LAUNCH(l2_q16_gather_rows_kernel, (unsigned)((m + 7) / 8), 256, 0, st, q, list + lo, m, sub_q);
      HPUSM_CHECK_LAUNCH();
      int64_t fb2 = 0;
      rc = l2_tc_split_knn2(sub_q, m, t, nt, sub_idx, sub_dist, base + L.sub_ws, L.sub_ws_bytes, st, &fb2);
      if (rc != HPUSM_OK) return rc;
      exact_rows += fb2;
      HPUSM_LAUNCH(l2_q16_scatter_rows_kernel, (unsigned)((m + 255) / 256), 256, 0, st, list + lo, m, sub_idx, sub_dist, idx, dist);
      HPUSM_CHECK_LAUNCH();
    }
    *fallback_rows = exact_rows;   // rows that needed the exact engine in the end
    return HPUSM_OK;
  }
  for (int64_t lo = 0; lo < h_list; lo += Q_FB_ROWS) {
    const int64_t m = h_list - lo < Q_FB_ROWS ? h_list - lo : Q_FB_ROWS;
    rc = l2_simt_candidates(q, nq, t, nt, Q_D, list + lo, m, L.fb_splits, fb, st);
    if (rc != HPUSM_OK) return rc;
    rc = l2_rerank_finalize(q, nq, t, Q_D, fb, L.fb_splits, 2, list + lo, m, idx, dist, st);
    if (rc != HPUSM_OK) return rc;
  }
  return HPUSM_OK;
}

}  // namespace hpusm
