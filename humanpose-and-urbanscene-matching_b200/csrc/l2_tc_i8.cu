// K1 (integer tensor-core engine): brute-force L2 kNN (k=2) on tcgen05 kind::i8, sm_100a only.
//
// Replaces cv2.BFMatcher(cv2.NORM_L2).knnMatch(..., k=2) at reference urbanscene/features.py:59 for what SIFT actually
// produces: descriptors whose 128 values are integers 0..255 (cv2 stores them as float32).  Such rows fit the u8
// operands of the integer MMA, whose s32 accumulation is exact and which retires K=32 per instruction at twice the
// bf16 rate; a database row costs 160 bytes of operand instead of 288.
//
//   ||q - t||^2 = ||q||^2 + ||t||^2 - 2 q.t      ->  per query, order the database by  K = 2 q.t - ||t||^2  (largest first)
//
// u8 operands cannot carry the factor 2, so the accumulator holds  s = q.t + (HB - (||t||^2 >> 1))  (the bracket is
// folded into the contraction through 32 augmentation columns: A = [q | 1 255 255 ...], B = [t | d0 d1 d2 ...] with
// HB - h = d0 + 255 (d1 + ... + d31)) and  K + 2 HB = 2 s - (||t||^2 & 1).  The epilogue max-scans s; only a column
// with 2 s > K1 (the running second best) can matter, and for those few the comparison is made on the exact K with the
// parity bit taken from a bit array (each epilogue warp keeps the 64 bits of its columns in registers).  The result is
// the exact top-2 on (K, index), ties to the lower index as OpenCV resolves them; the candidates then go through
// l2_rerank_finalize (exact FP32 distances from the ORIGINAL float32 rows).
//
// The prep kernel verifies the value range for every row (integers in [0,255], ||t||^2 < 2 HB); HPUSM_L2_AUTO falls
// through to the bf16 engines (l2_tc.cu, l2_tc_split.cu) when it does not hold.
//
// Kernel shape (one persistent CTA per SM, 18 warps; a unit = 256 queries x one split of the database):
//   warp 0      TMEM allocator, then TMA producer: an 8-stage ring of database tiles ([128 rows x 160 B] u8, 20 KB:
//               a SWIZZLE_128B slab of the 128 descriptor bytes + a SWIZZLE_32B slab of the 32 augmentation bytes)
//   warp 1      MMA issuer (one ELECTED lane): per piece = (database tile, 128-query tile) 5 tcgen05.mma.kind::i8
//               (M = 128, N = 128, K = 32).  The A operand (query rows) is read from TENSOR MEMORY, so shared memory
//               serves only the database operand; three 128-column accumulators rotate.
//   warps 2-17  epilogue.  At the start of a unit they copy the unit's query rows into TMEM (tcgen05.st, thread = row).
//               Then every accumulator is pulled by eight warps at once (two per TMEM lane quarter, 64 columns each):
//               tcgen05.ld, release the accumulator, scan (VIMNMX3 trees + one compare per 64 values; the rare
//               insertion is a compact loop).  Warps 2-9 serve query tile 0, warps 10-17 tile 1; the two column
//               groups of a row share their threshold through shared memory and are separate candidate parts.
// DESIGN.md (K1) has the measurements behind each of these choices.  -DI8_TRACE compiles an in-kernel timestamp trace
// of the pipeline (tools/i8_trace.py prints it).
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>
#include <stdio.h>
#include "common.cuh"
#include "l2_internal.cuh"
#include "tc_ptx.cuh"

namespace hpusm {

constexpr int I8_D = 128;
constexpr int I8_QB = 256;            // queries per unit (two M=128 query tiles)
constexpr int I8_TN = 128;            // database rows per tile (MMA N)
constexpr int I8_STAGES = 8;
constexpr int I8_THREADS = 576;       // 2 service warps (TMA producer + TMEM allocator, MMA issuer) + 16 epilogue warps
constexpr int I8_CGS = 2;             // column groups of 64 per 128-column accumulator (one per epilogue warp of a lane quarter)
constexpr int I8_NBUF = 3;            // accumulators in rotation
constexpr int I8_KA = 160;            // operand row: 128 descriptor bytes + 32 augmentation bytes
constexpr int I8_A_COLS = 40;         // a query tile in tensor memory: 160 bytes per row = 40 32-bit columns
constexpr int I8_ACC_COL0 = 128;      // tensor memory: query tiles in columns [0, 80), accumulators at 128, 256, 384
constexpr int I8_BMAIN_BYTES = I8_TN * 128;             // [128 rows][128 u8] swizzle-128B slab, 16 KB
constexpr int I8_BAUG_BYTES = I8_TN * 32;               // [128 rows][32 u8] swizzle-32B slab, 4 KB
constexpr int I8_B_BYTES = I8_BMAIN_BYTES + I8_BAUG_BYTES;      // 20 KB per stage
constexpr int I8_SMEM_BYTES = I8_STAGES * I8_B_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ +
                              I8_QB * 4 /*shared thresholds*/;
constexpr uint32_t I8_SLEEP_NS = 20000; // suspend-time hint of the pipeline waits
constexpr int I8_HB = 2000000;        // bias of the folded half-norm; ||t||^2 < 2 HB is required (32 u8 x u8 columns hold 2.08e6)

// Instruction descriptor: D = S32, A = B = U8, both K-major, N = 128, M = 128.
constexpr uint32_t I8_IDESC = (2u << 4) | ((uint32_t)(I8_TN >> 3) << 17) | ((128u >> 4) << 24);

// ---- operand preparation --------------------------------------------------------------------------------
// One warp per row: float32 -> u8 operand row of 160 bytes and the range check (flag[0] |= 1 when a value is not an
// integer in [0,255] or a database norm exceeds 2 HB).
//   query side    (is_db == 0): [q | 1 255 255 ... 255]
//   database side (is_db == 1): [t | d0 d1 ... d31] with HB - (||t||^2 >> 1) = d0 + 255 (d1 + ... + d31), d0 < 255;
//                               bit `row` of par = ||t||^2 & 1.  Rows n..n_pad are zero: s = 0, below or tied with every real
//                               row, and their index is the highest.
__global__ void __launch_bounds__(256)
l2_i8_prep_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int is_db, uint8_t* __restrict__ xb,
                  uint32_t* __restrict__ par, int* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  float f[4] = {0.f, 0.f, 0.f, 0.f};
  bool bad = false;
  float s = 0.f;
  if (row < n) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * I8_D) + lane);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bad |= !(f[i] >= 0.f && f[i] <= 255.f) || f[i] != rintf(f[i]);
      s = fmaf(f[i], f[i], s);
    }
  }
  bad = __any_sync(0xffffffffu, bad);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int tn = bad ? 0 : (int)s;  // exact: an integer below 2^24 whenever the row passed the check
  const int h = tn >> 1;
  if (is_db && h >= I8_HB) bad = true;  // strict: HB - h >= 1 keeps every real row above the zero padding
  uint32_t packed = 0;
  if (!bad) packed = (uint32_t)f[0] | ((uint32_t)f[1] << 8) | ((uint32_t)f[2] << 16) | ((uint32_t)f[3] << 24);
  uint8_t* dst = xb + row * I8_KA;
  reinterpret_cast<uint32_t*>(dst)[lane] = packed;
  // augmentation byte `lane`
  uint32_t a;
  if (!is_db) a = lane == 0 ? 1u : 255u;
  else if (row >= n || bad) a = 0u;
  else {
    const int rem = I8_HB - h;
    if (lane == 0) a = (uint32_t)(rem % 255);
    else {
      int m = rem / 255 - 255 * (lane - 1);
      a = (uint32_t)(m < 0 ? 0 : (m > 255 ? 255 : m));
    }
  }
  uint32_t w = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) w |= __shfl_sync(0xffffffffu, a, (4 * lane + i) & 31) << (8 * i);
  if (lane < 8) reinterpret_cast<uint32_t*>(dst + I8_D)[lane] = w;
  if (lane == 0) {
    if (is_db && row < n && !bad && (tn & 1)) atomicOr(par + (row >> 5), 1u << (row & 31));  // par is zeroed by the host
    if (bad) atomicOr(flag, 1);
  }
}

// ---- the contraction + fused top-2 ------------------------------------------------------------------------
struct I8Top2 {
  int k0, k1;  // exact keys 2 s - parity of the best and second best so far
  int thr;     // k1 >> 1: a column can only matter when its s exceeds this
  int i0, i1;
};

__device__ __forceinline__ int imax3(int a, int b, int c) { return max(max(a, b), c); }

// `sh`: shared-memory address of this row's threshold, shared by the warps that scan different column groups of the row.
__device__ __forceinline__ void i8_insert(I8Top2& t, int s, int col, int parity, uint32_t sh) {
  const int k = 2 * s - parity;
  if (k > t.k1) {  // strict: candidates reach a thread in increasing index order within equal keys
    if (k > t.k0) { t.k1 = t.k0; t.i1 = t.i0; t.k0 = k; t.i0 = col; }
    else { t.k1 = k; t.i1 = col; }
    t.thr = max(t.thr, t.k1 >> 1);
    asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(sh), "r"(t.k1 >> 1) : "memory");
  }
}

// The candidates of 32 accumulator columns that beat the threshold, pulled out one at a time, largest s first and the
// lower column first among equals, by a max tree over (s << 5 | 31 - j): a short loop instead of 32 unrolled
// insertions, because rarely executed straight-line code runs at instruction-fetch speed and every warp that shares
// the accumulator hand-over waits for the slowest one.  `pbits`: parity bits (bit j = column col + j).
__device__ __forceinline__ int i8_max32(const uint32_t (&acc)[32]) {
  int m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) m[i] = imax3((int)acc[3 * i], (int)acc[3 * i + 1], (int)acc[3 * i + 2]);
  m[10] = max((int)acc[30], (int)acc[31]);
  const int a = imax3(m[0], m[1], m[2]), b = imax3(m[3], m[4], m[5]), c = imax3(m[6], m[7], m[8]);
  return max(imax3(a, b, c), max(m[9], m[10]));
}

__device__ __forceinline__ void i8_slow32(const uint32_t (&acc)[32], int col, I8Top2& t, uint32_t pbits, uint32_t sh) {
  int last = INT_MAX;
#pragma unroll 1
  for (;;) {
    int pk[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int v = (int)acc[j] * 32 + (31 - j);
      pk[j] = v < last ? v : INT_MIN;
    }
    int r[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) r[i] = imax3(pk[3 * i], pk[3 * i + 1], pk[3 * i + 2]);
    r[10] = max(pk[30], pk[31]);
    const int best = max(imax3(imax3(r[0], r[1], r[2]), imax3(r[3], r[4], r[5]), imax3(r[6], r[7], r[8])), max(r[9], r[10]));
    const int sb = best >> 5;
    if (sb <= t.thr) break;
    const int j = 31 - (best & 31);
    i8_insert(t, sb, col + j, (int)((pbits >> j) & 1u), sh);
    last = best;
  }
}

// 64 accumulator columns of one query row.  Fast path: two 3-input max trees and ONE compare.
__device__ __forceinline__ void i8_process64(const uint32_t (&acc0)[32], const uint32_t (&acc1)[32], int col, I8Top2& t, uint2 pbits,
                                             uint32_t sh) {
  const int h0 = i8_max32(acc0), h1 = i8_max32(acc1);
  if (max(h0, h1) > t.thr) {
    if (h0 > t.thr) i8_slow32(acc0, col, t, pbits.x, sh);
    if (h1 > t.thr) i8_slow32(acc1, col + 32, t, pbits.y, sh);
  }
}

// part_idx [splits * I8_CGS][nq][2];  qu [nq][160] u8 operand rows of the queries
// DBG 0 = product.  Experiment switches (HPUSM_I8_DBG, wrong answers, tools/q16_time.py --engine i8 decomposes the step):
// 1 = hand-over + tensor-memory loads but no scan, 2 = hand-over only, 10 = the MMA warp runs free (MMA + TMA alone).
template <int DBG>
__global__ void __launch_bounds__(I8_THREADS, 1)
l2_i8_knn2_kernel(const uint8_t* __restrict__ qu, const __grid_constant__ CUtensorMap map_t,
                  const __grid_constant__ CUtensorMap map_ta, int64_t nq, int64_t nt, int q_blocks, int splits,
                  int64_t rows_per_split, const uint32_t* __restrict__ par, int32_t* __restrict__ part_idx,
                  long long* __restrict__ trace, const int* __restrict__ gate) {
  // explicit HPUSM_L2_TC_I8 is fully asynchronous: the prep kernels' eligibility flag is consumed HERE instead of being
  // read back by the host; ineligible data leave the candidate lists empty (every idx -1) and the error is reported
  // through hpusm_async_status().  Uniform over the grid, before any barrier or tensor-memory allocation.
  if (gate && *gate != 0) return;
  extern __shared__ uint8_t i8_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(i8_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem;                          // [stages][16 KB | 4 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + I8_STAGES * I8_B_BYTES);
  uint64_t* full = bars;                                     // [stages]  TMA -> MMA
  uint64_t* empty = bars + I8_STAGES;                        // [stages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * I8_STAGES;                // [NBUF]    MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * I8_STAGES + I8_NBUF;     // [NBUF]    epilogue -> MMA
  uint64_t* a_ready = bars + 2 * I8_STAGES + 2 * I8_NBUF;    // the unit's query rows are in tensor memory
  uint64_t* a_free = bars + 2 * I8_STAGES + 2 * I8_NBUF + 1; // all MMAs of the unit retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * I8_STAGES + 2 * I8_NBUF + 2);
  int* thr_sh = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2 query tiles][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < I8_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int b = 0; b < I8_NBUF; ++b) { mbar_init(tmem_full + b, 1); mbar_init(tmem_empty + b, 8 * 32); }
    mbar_init(a_ready, 16);
    mbar_init(a_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
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
        const int tiles = (int)((t_end - t_begin + I8_TN - 1) / I8_TN);
        for (int j = 0; j < tiles; ++j, ++it) {
          const int s = it % I8_STAGES;
          mbar_wait_sleepy(empty + s, ((it / I8_STAGES) & 1) ^ 1, I8_SLEEP_NS);
          mbar_expect_tx(full + s, I8_B_BYTES);
          const int row0 = (int)(t_begin + (int64_t)j * I8_TN);
          uint8_t* bs = smem_b + s * I8_B_BYTES;
          tma_load_2d(bs, &map_t, full + s, 0, row0);
          tma_load_2d(bs + I8_BMAIN_BYTES, &map_ta, full + s, I8_D, row0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // The whole warp walks the loop and one ELECTED lane issues: with `elect.sync` ptxas knows a single lane is active
    // and emits the tcgen05 instructions directly; behind a plain `lane == 0` test it wraps every one of them in a
    // serialising loop, and at N = 128 (64 tensor cycles per instruction) that loop was the bottleneck.
    uint32_t it = 0, un = 0, buf = 0, buf_phase = 0;
    const uint32_t b_addr = smem_u32(smem_b);
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++un) {
      const int split = u / q_blocks;
      const int64_t t_begin = (int64_t)split * rows_per_split;
      const int64_t t_end = min(nt, t_begin + rows_per_split);
      const int tiles = (int)((t_end - t_begin + I8_TN - 1) / I8_TN);
      mbar_wait(a_ready, un & 1);
      tc_fence_after();
      for (int j = 0; j < tiles; ++j, ++it) {
        const int s = it % I8_STAGES;
        mbar_wait_sleepy(full + s, (it / I8_STAGES) & 1, I8_SLEEP_NS);
        const uint64_t bdesc = umma_desc_sw128(b_addr + s * I8_B_BYTES);
        const uint64_t bdesc_aug = umma_desc_sw32(b_addr + s * I8_B_BYTES + I8_BMAIN_BYTES);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#ifdef I8_TRACE
          const bool tr = trace && blockIdx.x == 0 && it >= 6000 && it < 6064 && lane == 0;
          if (tr) trace[((it - 6000) * 2 + m) * 8 + 0] = clock64();
#endif
          if (!(DBG & 8)) mbar_wait_sleepy(tmem_empty + buf, buf_phase ^ 1, I8_SLEEP_NS);   // the epilogue has read this accumulator out
          tc_fence_after();
#ifdef I8_TRACE
          if (tr) trace[((it - 6000) * 2 + m) * 8 + 1] = clock64();
#endif
          const uint32_t d_addr = tmem_base + (uint32_t)(I8_ACC_COL0 + buf * I8_TN);
          const uint32_t a_col = tmem_base + (uint32_t)(m * I8_A_COLS);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < I8_D / 32; ++k)   // +2 in the descriptor's address field = 32 bytes along K
              tc_mma_i8_ts(d_addr, a_col + k * 8, bdesc + (uint64_t)(2 * k), I8_IDESC, k > 0 ? 1u : 0u);
            // fifth step: the augmentation columns add HB - (||t||^2 >> 1)
            tc_mma_i8_ts(d_addr, a_col + 32, bdesc_aug, I8_IDESC, 1u);
            tc_commit(tmem_full + buf);
          }
          __syncwarp();
#ifdef I8_TRACE
          if (tr) trace[((it - 6000) * 2 + m) * 8 + 2] = clock64();
#endif
          if (++buf == I8_NBUF) { buf = 0; buf_phase ^= 1; }
        }
        if (elect_one()) tc_commit(empty + s);
        __syncwarp();
      }
      if (elect_one()) tc_commit(a_free);
      __syncwarp();
    }
  } else {
    // ===== epilogue: fused top-2 =====
    // An accumulator (one 128-query tile x 128 database rows) is only free for its next MMA block once it has been
    // read out, tcgen05.ld has a long latency and tensor memory reads run at 64 B/clk per SM sub-partition.  Three
    // accumulators rotate, so the tensor pipe always has one to fill while another is being read out, and every
    // accumulator is pulled by EIGHT warps at once (two per TMEM lane quarter, 64 columns each): they load, release
    // the accumulator, and only then scan.  Accumulators alternate between the two query tiles, so warps 0-7 always
    // serve tile 0 and warps 8-15 tile 1: a thread carries one running top-2 (its row, its 64-column group), and the
    // two groups are separate candidate parts.
    const int ew = warp - 2, quarter = warp & 3 /* the TMEM lanes a warp may touch */, cg = (ew >> 2) & 1, m = ew >> 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // piece p = 2 * tile + m lives in accumulator p % 3 and completes phase (p / 3) & 1 of its barrier: for fixed m
    // the accumulators come round as 0,2,1 (m = 0) or 1,0,2 (m = 1) and the phase depends on the accumulator only
    uint32_t un = 0, buf = (uint32_t)m;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++un) {
      const int split = u / q_blocks, qb = u - split * q_blocks;
      const int64_t t_begin = (int64_t)split * rows_per_split;
      const int64_t t_end = min(nt, t_begin + rows_per_split);
      const int tiles = (int)((t_end - t_begin + I8_TN - 1) / I8_TN);
      // -- the unit's query rows go into tensor memory: thread = row, 40 words = [q | augmentation] (MMA operand A) --
      {
        const int ahalf = cg;  // this warp stores words [20 ahalf, 20 ahalf + 20) of its rows
        const int64_t row = (int64_t)qb * I8_QB + m * 128 + quarter * 32 + lane;
        uint32_t w[20];
#pragma unroll
        for (int i = 0; i < 20; ++i) w[i] = 0u;
        if (row < nq) {
          const uint4* src = reinterpret_cast<const uint4*>(qu + row * I8_KA) + ahalf * 5;
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const uint4 v = __ldg(src + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
          }
        }
        mbar_wait(a_free, (un & 1) ^ 1);   // the previous unit's MMAs no longer read the old rows
        tc_fence_after();
        const uint32_t a_addr = lane_base + (uint32_t)(m * I8_A_COLS + ahalf * 20);
        uint32_t g0[8], g1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { g0[i] = w[i]; g1[i] = w[8 + i]; }
        tc_st8(a_addr, g0);
        tc_st8(a_addr + 8, g1);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_addr + 16), "r"(w[16]), "r"(w[17]),
                     "r"(w[18]), "r"(w[19])
                     : "memory");
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      }
      I8Top2 top;
      top.k0 = top.k1 = INT_MIN;
      top.thr = INT_MIN >> 1;
      top.i0 = top.i1 = -1;
      // The two column-group warps of a row share their thresholds: each keeps its own top-2, but a column is only
      // looked at when it could beat the better second-best key of the two (minus one, so that equal keys still
      // reach their own list and the final re-rank settles the index order).
      const uint32_t sh = smem_u32(thr_sh + m * 128 + quarter * 32 + lane);
      asm volatile("bar.sync 1, 512;" ::: "memory");   // everybody is done with the previous unit's thresholds
      if (cg == 0) thr_sh[m * 128 + quarter * 32 + lane] = INT_MIN >> 1;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      // parity bits of this warp's 64 columns, fetched one tile ahead so that an insertion never waits on memory
      uint2 pb_next = make_uint2(0u, 0u);
      if (tiles > 0) pb_next = __ldg(reinterpret_cast<const uint2*>(par + (((int)t_begin + cg * 64) >> 5)));
      for (int j = 0; j < ((DBG & 8) ? 0 : tiles); ++j) {
        const int col0 = (int)(t_begin + (int64_t)j * I8_TN) + cg * 64;
        const uint2 pb = pb_next;
        if (j + 1 < tiles) pb_next = __ldg(reinterpret_cast<const uint2*>(par + ((col0 + I8_TN) >> 5)));
        uint32_t acc0[32], acc1[32];
        int shared_thr;
        asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(shared_thr) : "r"(sh));
        const uint32_t phase = m == 0 ? (buf == 1u ? 1u : 0u) : (buf == 1u ? 0u : 1u);
#ifdef I8_TRACE
        const bool tr = trace && blockIdx.x == 0 && un == 0 && j >= 6000 && j < 6064 && lane == 0;
        if (tr) trace[1024 + ((j - 6000) * 2 + m) * 32 + 8 + (ew & 7)] = clock64();
#endif
        mbar_wait_sleepy(tmem_full + buf, phase, I8_SLEEP_NS);
        tc_fence_after();
#ifdef I8_TRACE
        if (tr) trace[1024 + ((j - 6000) * 2 + m) * 32 + 24 + (ew & 7)] = clock64();
#endif
        const uint32_t taddr = lane_base + (uint32_t)(I8_ACC_COL0 + buf * I8_TN + cg * 64);
        if ((DBG & 3) != 2) {
          tc_ld32(taddr, acc0);
          tc_ld32(taddr + 32, acc1);
          tc_wait_ld();
        }
        tc_fence_before();
        mbar_arrive(tmem_empty + buf);   // every lane arrives: one hop fewer on the hand-over than syncwarp + elected arrive
#ifdef I8_TRACE
        if (tr) trace[1024 + ((j - 6000) * 2 + m) * 32 + (ew & 7)] = clock64() + (acc0[0] & 0);
#endif
        buf = buf == 0 ? 2u : buf - 1u;   // two pieces on: (buf + 2) % 3
        top.thr = max(top.thr, shared_thr - 1);
        if ((DBG & 3) != 0) {
          if ((DBG & 3) == 1 && (acc0[0] ^ acc1[5]) == 0x12345678u) top.i0 = 7;
          continue;
        }
        i8_process64(acc0, acc1, col0, top, pb, sh);
#ifdef I8_TRACE
        if (tr) trace[1024 + ((j - 6000) * 2 + m) * 32 + 16 + (ew & 7)] = clock64() + (top.i0 & 0);
#endif
      }
      const int64_t row = (int64_t)qb * I8_QB + m * 128 + quarter * 32 + lane;
      if (row < nq) {
        int32_t* o = part_idx + ((int64_t)(split * I8_CGS + cg) * nq + row) * 2;
        o[0] = top.i0 < nt ? top.i0 : -1;  // a padding row can only surface when the share holds fewer than 2 rows
        o[1] = top.i1 < nt ? top.i1 : -1;
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

// ---- host side ---------------------------------------------------------------------------------------------
static int i8_splits(int64_t nq, int64_t nt) {
  const int64_t qbs = (nq + I8_QB - 1) / I8_QB;
  const int64_t sms = sm_count();
  if (qbs >= 2 * sms) return 1;
  int64_t s = (2 * sms + qbs - 1) / qbs;
  const int64_t max_s = (nt + 2047) / 2048;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 48) s = 48;  // the re-rank holds splits * I8_CGS * 2 <= 192 candidates per query
  // every split must own at least one row after its length is rounded up to whole tiles
  while (s > 1) {
    const int64_t rps = ((nt + s - 1) / s + I8_TN - 1) / I8_TN * I8_TN;
    if ((s - 1) * rps < nt) break;
    --s;
  }
  return (int)s;
}

struct I8Layout {
  size_t qb, tb, par, flag, part, total;
  int splits;
  int64_t nt_pad;
};

static I8Layout i8_layout(int64_t nq, int64_t nt) {
  I8Layout L;
  L.nt_pad = (nt + I8_TN - 1) / I8_TN * I8_TN;
  L.splits = i8_splits(nq, nt);
  size_t off = 0;
  L.qb = off; off += align_up((size_t)nq * I8_KA, 1024);
  L.tb = off; off += align_up((size_t)L.nt_pad * I8_KA, 1024);
  L.par = off; off += align_up((size_t)L.nt_pad / 8, 1024);
  L.flag = off; off += 1024;
  L.part = off; off += align_up((size_t)L.splits * I8_CGS * nq * 2 * sizeof(int32_t), 1024);
  L.total = off;
  return L;
}

size_t l2_tc_i8_workspace_bytes(int64_t nq, int64_t nt) {
  if (nq <= 0 || nt <= 0) return 256;
  return i8_layout(nq, nt).total;
}

// *eligible = 0: the descriptors do not fit the u8 engine and nothing was computed (AUTO falls through); with
// `require` set that is an error instead.
int l2_tc_i8_knn2(const float* q, int64_t nq, const float* t, int64_t nt, int32_t* idx, float* dist, void* ws, size_t ws_bytes,
                  cudaStream_t st, bool require, int* eligible) {
  const I8Layout L = i8_layout(nq, nt);
  HPUSM_REQUIRE(ws_bytes >= L.total, HPUSM_ERR_WORKSPACE, "l2 i8 engine: workspace needs %zu bytes, got %zu", L.total, ws_bytes);
  char* base = reinterpret_cast<char*>(ws);
  uint8_t* qb = reinterpret_cast<uint8_t*>(base + L.qb);
  uint8_t* tb = reinterpret_cast<uint8_t*>(base + L.tb);
  uint32_t* par = reinterpret_cast<uint32_t*>(base + L.par);
  int* flag = reinterpret_cast<int*>(base + L.flag);
  int32_t* part = reinterpret_cast<int32_t*>(base + L.part);

  HPUSM_CUDA_OK(cudaMemsetAsync(flag, 0, sizeof(int), st));
  HPUSM_CUDA_OK(cudaMemsetAsync(par, 0, (size_t)L.nt_pad / 8, st));
  HPUSM_LAUNCH(l2_i8_prep_kernel, (unsigned)((nq + 7) / 8), 256, 0, st, q, nq, nq, 0, qb, par, flag);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(l2_i8_prep_kernel, (unsigned)((L.nt_pad + 7) / 8), 256, 0, st, t, nt, L.nt_pad, 1, tb, par, flag);
  HPUSM_CHECK_LAUNCH();
  *eligible = 1;
  if (!require) {
    // HPUSM_L2_AUTO: the one host read of the engine -- do the descriptors fit u8 operands?  (The explicit mode never
    // synchronises: the flag gates the sweep on the device and surfaces through hpusm_async_status().)
    int h_flag = 0;
    HPUSM_CUDA_OK(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    HPUSM_CUDA_OK(cudaStreamSynchronize(st));
    *eligible = h_flag == 0;
    if (h_flag != 0) return HPUSM_OK;
  }
  HPUSM_CUDA_OK(cudaMemsetAsync(part, 0xff, (size_t)L.splits * I8_CGS * nq * 2 * sizeof(int32_t), st));
  CUtensorMap map_t, map_ta;
  int rc = tc_make_map(&map_t, tb, L.nt_pad, I8_KA, 128, CU_TENSOR_MAP_SWIZZLE_128B, I8_TN, 1);
  if (rc == HPUSM_OK) rc = tc_make_map(&map_ta, tb, L.nt_pad, I8_KA, 32, CU_TENSOR_MAP_SWIZZLE_32B, I8_TN, 1);
  if (rc != HPUSM_OK) return rc;
  static std::atomic<bool> attr_done[64];
  int dev = 0;
  HPUSM_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev].load(std::memory_order_acquire)) {
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_i8_knn2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM_BYTES));
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_i8_knn2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM_BYTES));
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_i8_knn2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM_BYTES));
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_i8_knn2_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM_BYTES));
    attr_done[dev].store(true, std::memory_order_release);
  }
  const int q_blocks = (int)((nq + I8_QB - 1) / I8_QB);
  int64_t rows_per_split = (nt + L.splits - 1) / L.splits;
  rows_per_split = (rows_per_split + I8_TN - 1) / I8_TN * I8_TN;
  const int units = q_blocks * L.splits;
  const int grid = units < sm_count() ? units : sm_count();
  long long* trace = nullptr;
#ifdef I8_TRACE
  const char* trace_path = getenv("HPUSM_I8_TRACE");
  if (trace_path) { HPUSM_CUDA_OK(cudaMalloc(&trace, 8192 * 8)); HPUSM_CUDA_OK(cudaMemset(trace, 0, 8192 * 8)); }
#endif
#ifdef HPUSM_DEV_SWITCHES   // development builds only (HPUSM_EXTRA_FLAGS=-DHPUSM_DEV_SWITCHES bash build.sh): tools/q16_time.py
  const char* dbg_env = getenv("HPUSM_I8_DBG");   // experiment switches (wrong answers), see the kernel
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
#else
  const int dbg = 0;                               // the product library ignores the variable
#endif
  const int* gate = require ? flag : nullptr;
  if (dbg == 1)
    HPUSM_LAUNCH_TIMED(l2_i8_knn2_kernel<1>, grid, I8_THREADS, I8_SMEM_BYTES, st, qb, map_t, map_ta, nq, nt, q_blocks, L.splits,
                       rows_per_split, par, part, trace, gate);
  else if (dbg == 2)
    HPUSM_LAUNCH_TIMED(l2_i8_knn2_kernel<2>, grid, I8_THREADS, I8_SMEM_BYTES, st, qb, map_t, map_ta, nq, nt, q_blocks, L.splits,
                       rows_per_split, par, part, trace, gate);
  else if (dbg == 10)
    HPUSM_LAUNCH_TIMED(l2_i8_knn2_kernel<10>, grid, I8_THREADS, I8_SMEM_BYTES, st, qb, map_t, map_ta, nq, nt, q_blocks, L.splits,
                       rows_per_split, par, part, trace, gate);
  else
    HPUSM_LAUNCH_TIMED(l2_i8_knn2_kernel<0>, grid, I8_THREADS, I8_SMEM_BYTES, st, qb, map_t, map_ta, nq, nt, q_blocks, L.splits,
                       rows_per_split, par, part, trace, gate);
  HPUSM_CHECK_LAUNCH();
  if (require) {
    int rc2 = async_report(flag, HPUSM_ERR_UNSUPPORTED, st);
    if (rc2 != HPUSM_OK) return rc2;
  }
#ifdef I8_TRACE
  if (trace) {
    static long long h[8192];
    HPUSM_CUDA_OK(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    FILE* f = fopen(trace_path, "wb");
    if (f) { fwrite(h, 1, sizeof(h), f); fclose(f); }
    cudaFree(trace);
  }
#endif
  return l2_rerank_finalize(q, nq, t, I8_D, part, L.splits * I8_CGS, 2, nullptr, 0, idx, dist, st);
}

}  // namespace hpusm
