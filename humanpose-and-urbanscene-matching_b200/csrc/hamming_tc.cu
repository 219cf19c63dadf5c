// K2t (tensor-core engine of the segmented Hamming kNN-2 + ratio): tcgen05 kind::i8 with SIGNED operands, sm_100a only.
//
// Replaces, for every (query image, database image) pair of a retrieval run, cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(
// ..., k=2) + the ratio test of features.filter_matches (reference urbanscene/features.py:59 and :45-55); same contract as
// the __popc engine in hamming.cu, bit for bit.
//
// A 256-bit descriptor becomes 256 s8 values: query bit -> +1 / -1, database bit -> -64 / +64.  Equal bits contribute -64,
// different bits +64, so the s32 accumulator holds 128 * Hamming - 16384.  One more K step (32 augmentation bytes) adds the
// database row's COLUMN inside its 128-row tile (A = 1, B = column) and lifts the rows that pad an image to whole tiles
// above every real row (A = 127, B = 127, three times):
//
//     key = 128 * Hamming + column - 16384  (+ 48387 for a padding row)
//
// All keys of a tile are different and their order IS (Hamming, index): the two nearest rows are the two smallest keys.
// The epilogue therefore needs no insertion logic (per-image top-2 lists restart every few tiles, so a guarded rare-insert
// path as in the L2 engine would fire on every tile): per 64 accumulator columns it is one 3-input min tree for the
// smallest key m, then a second tree over (key - m - 1) taken UNSIGNED -- the winner wraps to 0xFFFFFFFF and drops out --
// for the second smallest: two instructions per pair, one of them on the FMA pipe (IMAD), none of them a branch.
//
// The database operand is expanded once per call (32 -> 288 bytes per row, every image padded to whole tiles so that a
// tile never straddles two images) by ham_tc_expand_db_kernel; the tile list, its image boundaries and the split of the
// tiles over the units are computed on the device (no host round trip).
//
// Kernel shape: as l2_tc_i8.cu (one persistent CTA per SM, 18 warps; TMA producer, one elected MMA issuer, 16 epilogue
// warps; query rows read from TENSOR MEMORY, three rotating 128-column accumulators).  Tensor memory is exactly full:
// 2 query tiles x 64 columns + 3 x 128 accumulator columns = 512, so the (row-independent) augmentation columns of the A
// operand live in shared memory and the ninth MMA of a piece is the shared x shared form.  At the last tile of an image
// the two column-group warps of a row combine their lists through shared memory, apply the ratio test in double precision
// and add the survivors to score[image].
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>
#include "common.cuh"
#include "hamming_internal.cuh"
#include "tc_ptx.cuh"

namespace hpusm {

constexpr int HT_QB = 256;            // queries per unit (two M=128 query tiles)
constexpr int HT_TN = 128;            // database rows per tile (MMA N)
constexpr int HT_STAGES = 5;
constexpr int HT_THREADS = 576;       // TMA producer + MMA issuer + 16 epilogue warps
constexpr int HT_NBUF = 3;
constexpr int HT_KB = 256;            // operand bytes that carry the descriptor bits
constexpr int HT_KA = 288;            // operand row: 256 + 32 augmentation bytes
constexpr int HT_A_COLS = 64;         // a query tile in tensor memory: 256 bytes per row = 64 32-bit columns
constexpr int HT_ACC_COL0 = 128;
constexpr int HT_SLAB = HT_TN * 128;  // [128 rows][128 B] SWIZZLE_128B slab, 16 KB; a stage holds two of them
constexpr int HT_BAUG = HT_TN * 32;   // [128 rows][32 B] SWIZZLE_32B slab, 4 KB
constexpr int HT_B_BYTES = 2 * HT_SLAB + HT_BAUG;   // 36 KB per stage
constexpr int HT_AAUG_BYTES = 128 * 32;
constexpr int HT_XCH_BYTES = 2 * 2 * 128 * 16;      // [slot][query tile][row] int4 exchange between the column groups
constexpr int HT_SMEM_BYTES = HT_STAGES * HT_B_BYTES + HT_AAUG_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + HT_XCH_BYTES;
constexpr uint32_t HT_SLEEP_NS = 20000;
constexpr int HT_BIAS = 64 * 256;     // key + HT_BIAS = 128 * Hamming + column
constexpr int HT_NONE = 1 << 20;      // "no row": above every Hamming distance, padding rows included

// Instruction descriptor: D = S32, A = B = S8, both K-major, N = 128, M = 128.
constexpr uint32_t HT_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HT_TN >> 3) << 17) | ((128u >> 4) << 24);

// ---- the plan: tiles per image, their prefix sum, the tile list, the split of the tiles over the units -----------------
// tile_off [nseg + 1]: first tile of every image (exclusive scan of ceil(rows / 128)); one block.
__global__ void __launch_bounds__(1024)
ham_tc_scan_kernel(const int64_t* __restrict__ seg_offsets, int64_t nseg, int32_t* __restrict__ tile_off) {
  __shared__ int warp_sum[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nseg; base += 1024) {
    const int64_t s = base + threadIdx.x;
    int n = 0;
    if (s < nseg) {
      const int64_t len = seg_offsets[s + 1] - seg_offsets[s];
      n = len > 0 ? (int)((len + HT_TN - 1) / HT_TN) : 0;
    }
    int incl = n;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const int before = carry + (warp > 0 ? warp_sum[warp - 1] : 0) + incl - n;
    if (s < nseg) tile_off[s] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + n;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_off[nseg] = carry;
}

// tile_info[tile] = {first database row of the tile, image | last-tile-of-its-image << 31}; thread per image.
// split_tile [splits + 1]: unit boundaries in tiles, each at the first tile of an image (threads of block 0).
__global__ void __launch_bounds__(256)
ham_tc_plan_kernel(const int64_t* __restrict__ seg_offsets, int64_t nseg, const int32_t* __restrict__ tile_off, int splits,
                   int2* __restrict__ tile_info, int32_t* __restrict__ split_tile) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) {
    const int t0 = tile_off[s], t1 = tile_off[s + 1];
    const int64_t row0 = seg_offsets[s];
    for (int t = t0; t < t1; ++t)
      tile_info[t] = make_int2((int)(row0 + (int64_t)(t - t0) * HT_TN), (int)s | (t + 1 == t1 ? (int)0x80000000 : 0));
  }
  if (blockIdx.x == 0) {
    const int total = tile_off[nseg];
    for (int k = threadIdx.x; k <= splits; k += blockDim.x) {
      if (k == splits) { split_tile[k] = total; continue; }
      const int64_t target = (int64_t)total * k / splits;
      int64_t lo = 0, hi = nseg;   // first image whose first tile is >= target
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (tile_off[mid] < target) lo = mid + 1; else hi = mid;
      }
      split_tile[k] = tile_off[lo];
    }
  }
}

// ---- operand expansion ----------------------------------------------------------------------------------------------------
// 4 descriptor bits -> bit 7 of 4 bytes (bit i of x lands on bits i, i+7, i+14, i+21, i+28: all different, no carries)
__device__ __forceinline__ uint32_t ht_spread4(uint32_t x) { return ((x & 0xFu) * 0x10204081u) & 0x80808080u; }

// A block per tile of 128 operand rows (blocks stride over the tiles of splits [split_lo, split_hi)): 16 lanes per row, lane = 16 descriptor bits -> 16 operand bytes (-64 where the bit
// is set, +64 where it is clear), so a warp reads 64 contiguous descriptor bytes and writes two full 256-byte row bodies per
// store; augmentation bytes [column, 0,0,0, 0...] for a real row, [column, 127,127,127, 0...] and a zero main part for the
// rows that pad an image to whole tiles.
__global__ void __launch_bounds__(256)
ham_tc_expand_db_kernel(const uint8_t* __restrict__ t, const int64_t* __restrict__ seg_offsets, const int32_t* __restrict__ split_tile,
                        int split_lo, int split_hi, const int2* __restrict__ tile_info, uint8_t* __restrict__ tb) {
  const int tile_lo = __ldg(split_tile + split_lo), tile_hi = __ldg(split_tile + split_hi);
  const int sub = threadIdx.x & 15, r0 = threadIdx.x >> 4;   // 16 rows per pass
  for (int tile = tile_lo + blockIdx.x; tile < tile_hi; tile += gridDim.x) {
    const int2 info = __ldg(tile_info + tile);
    const int64_t seg_end = __ldg(seg_offsets + (info.y & 0x7fffffff) + 1);
    const int64_t left = seg_end - (int64_t)info.x;
    const int real_rows = left < HT_TN ? (int)left : HT_TN;
    const uint8_t* src = t + (int64_t)info.x * 32;
    uint8_t* dst = tb + (int64_t)tile * HT_TN * HT_KA;
#pragma unroll
    for (int p = 0; p < HT_TN / 16; ++p) {
      const int r = p * 16 + r0;
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      if (r < real_rows) {
        const uint32_t b = __ldg(reinterpret_cast<const uint16_t*>(src + r * 32) + sub);
        w.x = ht_spread4(b) ^ 0x40404040u;          // set bit: 0xC0 = -64, clear bit: 0x40 = +64
        w.y = ht_spread4(b >> 4) ^ 0x40404040u;
        w.z = ht_spread4(b >> 8) ^ 0x40404040u;
        w.w = ht_spread4(b >> 12) ^ 0x40404040u;
      }
      *reinterpret_cast<uint4*>(dst + r * HT_KA + sub * 16) = w;
    }
    {
      const int r = threadIdx.x >> 1;
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      if ((threadIdx.x & 1) == 0) w.x = (uint32_t)r | (r < real_rows ? 0u : 0x7F7F7F00u);
      *reinterpret_cast<uint4*>(dst + r * HT_KA + HT_KB + (threadIdx.x & 1) * 16) = w;
    }
  }
}

// queries: bit -> +1 (0x01) / -1 (0xFF); 16 lanes per row, same bit order as the database side
__global__ void __launch_bounds__(256)
ham_tc_expand_q_kernel(const uint8_t* __restrict__ q, int64_t nq, uint8_t* __restrict__ qa) {
  const int sub = threadIdx.x & 15;
  const int64_t row = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 4);
  if (row >= nq) return;
  const uint32_t b = __ldg(reinterpret_cast<const uint16_t*>(q + row * 32) + sub);
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t s = ht_spread4(b >> (4 * i)) >> 7;   // 0x01 where the bit is set
    w[i] = ~(s * 0xFFu) | s;
  }
  *reinterpret_cast<uint4*>(qa + row * HT_KB + sub * 16) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- the contraction + fused per-image top-2 + ratio test ---------------------------------------------------------------
__device__ __forceinline__ int imin3(int a, int b, int c) { return min(min(a, b), c); }
__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }
// x + c on the FMA pipe (IMAD): the ALU pipe carries the min trees
__device__ __forceinline__ uint32_t ht_shift(uint32_t x, uint32_t c) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(r) : "r"(x), "r"(c));
  return r;
}

__device__ __forceinline__ int ht_min32(const uint32_t (&acc)[32]) {
  int m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) m[i] = imin3((int)acc[3 * i], (int)acc[3 * i + 1], (int)acc[3 * i + 2]);
  m[10] = min((int)acc[30], (int)acc[31]);
  return min(imin3(imin3(m[0], m[1], m[2]), imin3(m[3], m[4], m[5]), imin3(m[6], m[7], m[8])), min(m[9], m[10]));
}
__device__ __forceinline__ uint32_t ht_min32_shifted(const uint32_t (&acc)[32], uint32_t c) {
  uint32_t m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) m[i] = umin3(ht_shift(acc[3 * i], c), ht_shift(acc[3 * i + 1], c), ht_shift(acc[3 * i + 2], c));
  m[10] = min(ht_shift(acc[30], c), ht_shift(acc[31], c));
  return min(umin3(umin3(m[0], m[1], m[2]), umin3(m[3], m[4], m[5]), umin3(m[6], m[7], m[8])), min(m[9], m[10]));
}

__global__ void __launch_bounds__(HT_THREADS, 1)
ham_tc_segmented_kernel(const uint8_t* __restrict__ qa, const __grid_constant__ CUtensorMap map_t,
                        const __grid_constant__ CUtensorMap map_ta, int64_t nq, int q_blocks, int split0, int splits,
                        const int32_t* __restrict__ split_tile, const int2* __restrict__ tile_info, double ratio,
                        int32_t* __restrict__ score, int32_t* __restrict__ best_idx) {
  extern __shared__ uint8_t ht_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ht_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem;                                   // [stages][16 KB | 16 KB | 4 KB]
  uint8_t* smem_aaug = smem + HT_STAGES * HT_B_BYTES;       // [128 rows][32 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aaug + HT_AAUG_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + HT_STAGES;
  uint64_t* tmem_full = bars + 2 * HT_STAGES;
  uint64_t* tmem_empty = bars + 2 * HT_STAGES + HT_NBUF;
  uint64_t* a_ready = bars + 2 * HT_STAGES + 2 * HT_NBUF;
  uint64_t* a_free = bars + 2 * HT_STAGES + 2 * HT_NBUF + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * HT_STAGES + 2 * HT_NBUF + 2);
  int4* xch = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2 slots][2 query tiles][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < HT_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int b = 0; b < HT_NBUF; ++b) { mbar_init(tmem_full + b, 1); mbar_init(tmem_empty + b, 8 * 32); }
    mbar_init(a_ready, 16);
    mbar_init(a_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // The augmentation columns of the A operand are the same for every query row: [1, 127, 127, 127, 0 x 12] in BOTH 16-byte
  // halves of the 32-byte row (the database side is zero in bytes 16..31), which makes the SWIZZLE_32B placement moot.
  for (int i = threadIdx.x; i < HT_AAUG_BYTES / 4; i += HT_THREADS)
    reinterpret_cast<uint32_t*>(smem_aaug)[i] = (i & 3) == 0 ? 0x7F7F7F01u : 0u;
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
    // ===== TMA producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_t) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ta) : "memory");
      uint32_t it = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int split = split0 + u / q_blocks;
        const int t0 = __ldg(split_tile + split), t1 = __ldg(split_tile + split + 1);
        for (int j = t0; j < t1; ++j, ++it) {
          const int s = it % HT_STAGES;
          mbar_wait_sleepy(empty + s, ((it / HT_STAGES) & 1) ^ 1, HT_SLEEP_NS);
          mbar_expect_tx(full + s, HT_B_BYTES);
          const int row0 = j * HT_TN;
          uint8_t* bs = smem_b + s * HT_B_BYTES;
          tma_load_2d(bs, &map_t, full + s, 0, row0);
          tma_load_2d(bs + HT_SLAB, &map_t, full + s, 128, row0);
          tma_load_2d(bs + 2 * HT_SLAB, &map_ta, full + s, HT_KB, row0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane, see l2_tc_i8.cu) =====
    uint32_t it = 0, un = 0, buf = 0, buf_phase = 0;
    const uint32_t b_addr = smem_u32(smem_b);
    const uint64_t adesc_aug = umma_desc_sw32(smem_u32(smem_aaug));
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++un) {
      const int split = split0 + u / q_blocks;
      const int t0 = __ldg(split_tile + split), t1 = __ldg(split_tile + split + 1);
      mbar_wait(a_ready, un & 1);
      tc_fence_after();
      for (int j = t0; j < t1; ++j, ++it) {
        const int s = it % HT_STAGES;
        mbar_wait_sleepy(full + s, (it / HT_STAGES) & 1, HT_SLEEP_NS);
        const uint64_t bdesc0 = umma_desc_sw128(b_addr + s * HT_B_BYTES);
        const uint64_t bdesc1 = umma_desc_sw128(b_addr + s * HT_B_BYTES + HT_SLAB);
        const uint64_t bdesc_aug = umma_desc_sw32(b_addr + s * HT_B_BYTES + 2 * HT_SLAB);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          mbar_wait_sleepy(tmem_empty + buf, buf_phase ^ 1, HT_SLEEP_NS);
          tc_fence_after();
          const uint32_t d_addr = tmem_base + (uint32_t)(HT_ACC_COL0 + buf * HT_TN);
          const uint32_t a_col = tmem_base + (uint32_t)(m * HT_A_COLS);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)   // +2 in the descriptor's address field = 32 bytes along K
              tc_mma_i8_ts(d_addr, a_col + k * 8, bdesc0 + (uint64_t)(2 * k), HT_IDESC, k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_i8_ts(d_addr, a_col + 32 + k * 8, bdesc1 + (uint64_t)(2 * k), HT_IDESC, 1u);
            tc_mma_i8_ss(d_addr, adesc_aug, bdesc_aug, HT_IDESC, 1u);   // + column (+ padding lift)
            tc_commit(tmem_full + buf);
          }
          __syncwarp();
          if (++buf == HT_NBUF) { buf = 0; buf_phase ^= 1; }
        }
        if (elect_one()) tc_commit(empty + s);
        __syncwarp();
      }
      if (elect_one()) tc_commit(a_free);
      __syncwarp();
    }
  } else {
    // ===== epilogue =====
    // warps 2-9 serve query tile 0, warps 10-17 tile 1; inside a tile two warps per TMEM lane quarter, 64 columns each.
    const int ew = warp - 2, quarter = warp & 3, cg = (ew >> 2) & 1, m = ew >> 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int pair_bar = 2 + m * 4 + quarter;    // named barrier of the two column-group warps of these 32 rows
    const int rq = m * 128 + quarter * 32 + lane;
    uint32_t un = 0, buf = (uint32_t)m, slot = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++un) {
      const int qb = u % q_blocks, split = split0 + u / q_blocks;
      const int t0 = __ldg(split_tile + split), t1 = __ldg(split_tile + split + 1);
      const int64_t row = (int64_t)qb * HT_QB + rq;
      // -- the unit's query rows go into tensor memory: thread = row; this warp stores words [32 cg, 32 cg + 32) --
      {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = 0u;
        if (row < nq) {
          const uint4* src = reinterpret_cast<const uint4*>(qa + row * HT_KB) + cg * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 v = __ldg(src + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
          }
        }
        mbar_wait(a_free, (un & 1) ^ 1);   // the previous unit's MMAs no longer read the old rows
        tc_fence_after();
        const uint32_t a_addr = lane_base + (uint32_t)(m * HT_A_COLS + cg * 32);
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
      int h0 = HT_NONE, h1 = HT_NONE, i0 = -1;   // this thread's list inside the current image (its 64-column groups)
      int2 info_next = make_int2(0, 0);
      if (t0 < t1) info_next = __ldg(tile_info + t0);
      for (int j = t0; j < t1; ++j) {
        const int2 info = info_next;
        if (j + 1 < t1) info_next = __ldg(tile_info + j + 1);
        uint32_t acc0[32], acc1[32];
        const uint32_t phase = m == 0 ? (buf == 1u ? 1u : 0u) : (buf == 1u ? 0u : 1u);
        mbar_wait_sleepy(tmem_full + buf, phase, HT_SLEEP_NS);
        tc_fence_after();
        const uint32_t taddr = lane_base + (uint32_t)(HT_ACC_COL0 + buf * HT_TN + cg * 64);
        tc_ld32(taddr, acc0);
        tc_ld32(taddr + 32, acc1);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(tmem_empty + buf);
        buf = buf == 0 ? 2u : buf - 1u;   // two pieces on: (buf + 2) % 3
        // two smallest keys of the 64 columns
        const int k1 = min(ht_min32(acc0), ht_min32(acc1));
        const uint32_t c = (uint32_t)(-(k1 + 1));
        const int k2 = k1 + 1 + (int)min(ht_min32_shifted(acc0, c), ht_min32_shifted(acc1, c));
        const int n0 = (k1 + HT_BIAS) >> 7, n1 = (k2 + HT_BIAS) >> 7;
        if (n0 < h0) {                     // strict: earlier tiles hold the lower indices
          h1 = min(h0, n1); h0 = n0; i0 = info.x + (k1 & 127);
        } else {
          h1 = min(h1, n0);
        }
        if (info.y < 0) {
          // -- last tile of the image: combine the two column groups, ratio test, count --
          int4* x = xch + (slot * 2 + m) * 128 + quarter * 32 + lane;
          if (cg == 1) *x = make_int4(h0, i0, h1, 0);
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          if (cg == 0) {
            const int4 o = *x;
            const bool other_first = o.x < h0 || (o.x == h0 && o.y < i0);
            const int b0 = other_first ? o.x : h0, bi = other_first ? o.y : i0;
            const int b1 = min(min(h1, o.z), other_first ? h0 : o.x);
            // features.py:48  len(m) == 2 and m[0].distance < m[1].distance * ratio   (Python double)
            const bool keep = row < nq && b1 <= 256 && (double)b0 < (double)b1 * ratio;
            const int seg = info.y & 0x7fffffff;
            const unsigned ballot = __ballot_sync(0xffffffffu, keep);
            if (lane == 0 && ballot) atomicAdd(score + seg, __popc(ballot));
            if (keep && best_idx) best_idx[(int64_t)seg * nq + row] = bi;
          }
          slot ^= 1u;
          h0 = h1 = HT_NONE; i0 = -1;
        }
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

// ---- host side -----------------------------------------------------------------------------------------------------------
// Default: expand the whole database, then match it (one launch each).  HPUSM_HAMMING_CHUNKS=n (2..8) cuts the database into
// n chunks of whole images (each a run of `spc` splits) and expands chunk c + 1 on an internal side stream while chunk c is
// being matched -- the expansion is HBM-write bound and needs few issue slots, the matching kernel is tensor-pipe bound and
// leaves HBM idle, two expansion blocks fit on an SM beside the matching CTA.  Measured on C3: the step is
// no shorter (8.56 ms with 8 chunks against 8.49 ms), because the board sits at its power cap either way and the
// matching kernel slows down by what the overlapped expansion draws.  Kept as an experiment switch, off by default.
struct HtLayout {
  size_t tile_off, split_tile, tile_info, qa, tb, total;
  int64_t max_tiles;
  int chunks, spc /* splits per chunk */, q_blocks;
};

static int ht_env_chunks() {
  static const int v = []() { const char* e = getenv("HPUSM_HAMMING_CHUNKS"); return e ? atoi(e) : 0; }();
  return v;
}

static HtLayout ht_layout(int64_t nq, int64_t nt, int64_t nseg) {
  HtLayout L;
  L.max_tiles = nt / HT_TN + nseg;    // sum over images of ceil(rows / 128) <= this
  L.q_blocks = (int)((nq + HT_QB - 1) / HT_QB);
  int64_t chunks = ht_env_chunks() > 0 ? ht_env_chunks() : 1;
  if (chunks > 8) chunks = 8;
  if (chunks > nseg) chunks = nseg;
  if (chunks < 1) chunks = 1;
  int64_t spc = (2 * (int64_t)sm_count() + L.q_blocks - 1) / L.q_blocks;   // two units per SM and launch
  const int64_t by_tiles = L.max_tiles / chunks / 32;                        // at least ~32 tiles per unit
  if (spc > by_tiles) spc = by_tiles;
  if (spc > nseg / chunks) spc = nseg / chunks;
  if (spc < 1) spc = 1;
  L.chunks = (int)chunks;
  L.spc = (int)spc;
  size_t off = 0;
  L.tile_off = off; off += align_up((size_t)(nseg + 1) * sizeof(int32_t), 1024);
  L.split_tile = off; off += align_up((size_t)(L.chunks * L.spc + 1) * sizeof(int32_t), 1024);
  L.tile_info = off; off += align_up((size_t)L.max_tiles * sizeof(int2), 1024);
  L.qa = off; off += align_up((size_t)nq * HT_KB, 1024);
  L.tb = off; off += align_up((size_t)L.max_tiles * HT_TN * HT_KA, 1024);
  L.total = off;
  return L;
}

bool hamming_tc_eligible(int64_t nq, int64_t nt, int64_t nseg, int nbytes) {
  // 32-byte descriptors only (64-byte rows would need 128 tensor-memory columns per query tile); enough queries to fill an
  // MMA tile, images long enough that padding them to whole tiles stays cheap, and sizes the 32-bit tile arithmetic holds
  return nbytes == 32 && nq >= 128 && nseg >= 1 && nt >= 16384 && nt / nseg >= 256 &&
         nt / HT_TN + nseg < (int64_t)(INT_MAX / HT_TN) && nq < ((int64_t)1 << 31);
}

size_t hamming_tc_workspace_bytes(int64_t nq, int64_t nt, int64_t nseg) { return ht_layout(nq, nt, nseg).total; }

// One side stream per device, created on first use (shared by all host threads: work on it is ordered by events only).
static cudaStream_t ht_side_stream(int dev) {
  static std::atomic<cudaStream_t> streams[64];
  if (dev < 0 || dev >= 64) return nullptr;
  cudaStream_t s = streams[dev].load(std::memory_order_acquire);
  if (s) return s;
  cudaStream_t made = nullptr;
  if (cudaStreamCreateWithFlags(&made, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  cudaStream_t expected = nullptr;
  if (!streams[dev].compare_exchange_strong(expected, made, std::memory_order_acq_rel)) {
    cudaStreamDestroy(made);
    return expected;
  }
  return made;
}

// score and best_idx are already cleared (score 0, best_idx -1) by the caller.
int hamming_tc_segmented(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, const int64_t* seg_offsets, int64_t nseg,
                         double ratio, int32_t* score, int32_t* best_idx, void* ws, size_t ws_bytes, cudaStream_t st) {
  const HtLayout L = ht_layout(nq, nt, nseg);
  HPUSM_REQUIRE(ws && is_aligned(ws, 256) && ws_bytes >= L.total, HPUSM_ERR_WORKSPACE,
                "hamming tensor engine: workspace needs %zu bytes (256-byte aligned), got %zu", L.total, ws_bytes);
  char* base = reinterpret_cast<char*>(ws);
  int32_t* tile_off = reinterpret_cast<int32_t*>(base + L.tile_off);
  int32_t* split_tile = reinterpret_cast<int32_t*>(base + L.split_tile);
  int2* tile_info = reinterpret_cast<int2*>(base + L.tile_info);
  uint8_t* qa = reinterpret_cast<uint8_t*>(base + L.qa);
  uint8_t* tb = reinterpret_cast<uint8_t*>(base + L.tb);
  const int splits = L.chunks * L.spc;

  CUtensorMap map_t, map_ta;
  int rc = tc_make_map(&map_t, tb, L.max_tiles * HT_TN, HT_KA, 128, CU_TENSOR_MAP_SWIZZLE_128B, HT_TN, 1);
  if (rc == HPUSM_OK) rc = tc_make_map(&map_ta, tb, L.max_tiles * HT_TN, HT_KA, 32, CU_TENSOR_MAP_SWIZZLE_32B, HT_TN, 1);
  if (rc != HPUSM_OK) return rc;
  static std::atomic<bool> attr_done[64];
  int dev = 0;
  HPUSM_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev].load(std::memory_order_acquire)) {
    HPUSM_CUDA_OK(cudaFuncSetAttribute(ham_tc_segmented_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM_BYTES));
    attr_done[dev].store(true, std::memory_order_release);
  }
  cudaStream_t side = L.chunks > 1 ? ht_side_stream(dev) : nullptr;
  HPUSM_REQUIRE(L.chunks == 1 || side, HPUSM_ERR_CUDA, "hamming tensor engine: cannot create the expansion stream");

  HPUSM_LAUNCH(ham_tc_scan_kernel, 1, 1024, 0, st, seg_offsets, nseg, tile_off);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(ham_tc_plan_kernel, (unsigned)((nseg + 255) / 256), 256, 0, st, seg_offsets, nseg, tile_off, splits, tile_info,
               split_tile);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(ham_tc_expand_q_kernel, (unsigned)((nq + 15) / 16), 256, 0, st, q, nq, qa);
  HPUSM_CHECK_LAUNCH();

  const int units = L.q_blocks * L.spc;
  const int grid = units < sm_count() ? units : sm_count();
  const unsigned expand_grid = (unsigned)(2 * sm_count());
  if (L.chunks == 1) {
    HPUSM_LAUNCH(ham_tc_expand_db_kernel, expand_grid * 4, 256, 0, st, t, seg_offsets, split_tile, 0, splits, tile_info, tb);
    HPUSM_CHECK_LAUNCH();
    HPUSM_LAUNCH_TIMED(ham_tc_segmented_kernel, grid, HT_THREADS, HT_SMEM_BYTES, st, qa, map_t, map_ta, nq, L.q_blocks, 0, L.spc,
                       split_tile, tile_info, ratio, score, best_idx);
    HPUSM_CHECK_LAUNCH();
    return HPUSM_OK;
  }
  // fork: the side stream starts after the plan; chunk c is matched on `st` once its expansion has finished
  cudaEvent_t ev[9];
  int nev = 0;
  auto new_event = [&]() -> cudaEvent_t {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ev[nev++] = e;
    return e;
  };
  int result = HPUSM_OK;
  do {
    cudaEvent_t planned = new_event();
    if (!planned || cudaEventRecord(planned, st) != cudaSuccess || cudaStreamWaitEvent(side, planned, 0) != cudaSuccess) {
      result = HPUSM_ERR_CUDA; break;
    }
    for (int c = 0; c < L.chunks && result == HPUSM_OK; ++c) {
      // the first chunk is expanded with the whole GPU, the later ones with two blocks per SM beside the matching CTAs
      ham_tc_expand_db_kernel<<<c == 0 ? expand_grid * 4 : expand_grid, 256, 0, side>>>(t, seg_offsets, split_tile, c * L.spc,
                                                                                        (c + 1) * L.spc, tile_info, tb);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      cudaEvent_t done = new_event();
      if (cudaGetLastError() != cudaSuccess || !done || cudaEventRecord(done, side) != cudaSuccess ||
          cudaStreamWaitEvent(st, done, 0) != cudaSuccess) {
        result = HPUSM_ERR_CUDA; break;
      }
      HPUSM_LAUNCH_TIMED(ham_tc_segmented_kernel, grid, HT_THREADS, HT_SMEM_BYTES, st, qa, map_t, map_ta, nq, L.q_blocks, c * L.spc,
                         L.spc, split_tile, tile_info, ratio, score, best_idx);
      if (cudaGetLastError() != cudaSuccess) result = HPUSM_ERR_CUDA;
    }
  } while (false);
  for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);   // released by the runtime once the recorded work has completed
  if (result != HPUSM_OK) set_error("hamming tensor engine: launch / event failure: %s", cudaGetErrorString(cudaPeekAtLastError()));
  return result;
}

}  // namespace hpusm
