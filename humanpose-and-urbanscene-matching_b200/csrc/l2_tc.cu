// K1 (tensor-core engine): brute-force L2 kNN (k=2) as a dense contraction on tcgen05, sm_100a only.
//
// Replaces cv2.BFMatcher(cv2.NORM_L2).knnMatch(..., k=2) at reference urbanscene/features.py:59.
//
//   ||q - t||^2 = ||q||^2 + ||t||^2 - 2 q.t      ->  per query, order the database by  s = 2 q.t - ||t||^2  (largest first)
//
// The norm is folded INTO the contraction: operand rows are 144 wide, A = [2q | 1 1 1 0..], B = [t | -p0 -p1 -p2 0..]
// with ||t||^2 = p0 + p1 + p2 split into three 8-bit-significand pieces, so the accumulator already holds s and the
// epilogue is a pure max-scan (no FMA, no norm traffic) at the price of a ninth K=16 MMA step per tile.
//
// OpenCV SIFT descriptors are integer valued (0..255) stored as float32, so bf16 operands are exact and
// the FP32 accumulation (every partial sum is an integer of magnitude <= 2^24) is exact: s is the exact integer, the
// fused top-2 on (s, index) is the exact answer, and ties fall to the lower index exactly as OpenCV resolves them.  The
// prep kernel verifies that property for every value; when it does not hold HPUSM_L2_AUTO hands the call to the
// split-operand engine (l2_tc_split.cu) after one 4-byte host read.  The candidates go through l2_rerank_finalize
// (l2_simt.cu): exact FP32 distances from the ORIGINAL float32 rows, lexicographic (d2, index) top-2, sqrtf -- the
// numbers DMatch.distance carries.
//
// Kernel shape (one persistent CTA per SM, 384 threads, warp specialised):
//   warp 0     TMA producer: the unit's 256 queries (2 x [128 x 144] bf16, 72 KB, loaded once per unit) and a 2-stage
//              ring of database tiles ([256 rows x 144] bf16, 72 KB per stage); columns 0..127 land as two
//              SWIZZLE_128B slabs, the 16 augmentation columns as one SWIZZLE_32B slab
//   warp 1     MMA issuer: per database tile 2 x 9 tcgen05.mma (M=128, N=256, K=16, kind::f16, cta_group::1), one
//              block of 9 per query tile, each into its own 256-column TMEM accumulator (2 x 256 = all 512 columns).
//              The two query tiles ping-pong: while the tensor pipe works on tile 1's accumulator the epilogue drains
//              tile 0's, and vice versa.  N=256 matters: with operands in shared memory every MMA re-reads its A
//              slice, and at N=128 shared-memory bandwidth (MMA operand reads + TMA writes = 156 B/clk against 128 B/clk)
//              capped the tensor pipe at 80 %; at N=256 the A re-reads halve (125 B/clk).
//   warp 2     TMEM allocator
//   warps 4-11 epilogue: thread = one query row (TMEM lane); tcgen05.ld 32 columns at a time, an FMNMX3 tree over the
//              32 values + one compare guards the (rare) top-2 insertion; the running top-2 lives in registers for
//              the whole database sweep, so there is no reduction anywhere.
// A unit is (256 queries) x (one split of the database); a CTA walks units round-robin so all CTAs sweep the
// database in the same order and every database tile is fetched from HBM once and then served by L2.
#include <cuda.h>
#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>
#include "common.cuh"
#include "l2_internal.cuh"
#include "tc_ptx.cuh"

namespace hpusm {

constexpr int TC_D = 128;            // descriptor length the engine is built for
constexpr int TC_QB = 256;           // queries per unit (two M=128 MMA tiles)
constexpr int TC_TN = 256;           // database rows per tile (MMA N)
constexpr int TC_STAGES = 2;
constexpr int TC_THREADS = 384;
constexpr int TC_KA = 144;                                  // operand row: 128 descriptor columns + 16 augmentation columns
constexpr int TC_HALF_BYTES = 128 * 128;                    // query side: [128 rows][64 bf16] = 16 KB swizzle-128B slab
constexpr int TC_AUG_BYTES = 128 * 32;                      // query side: [128 rows][16 bf16] = 4 KB swizzle-32B slab
constexpr int TC_TILE_BYTES = 2 * TC_HALF_BYTES + TC_AUG_BYTES;  // 36 KB: one [128 x 144] query tile
constexpr int TC_A_BYTES = 2 * TC_TILE_BYTES;               // two query tiles
constexpr int TC_BHALF_BYTES = TC_TN * 128;                 // database side: [256 rows][64 bf16] = 32 KB slab
constexpr int TC_BAUG_BYTES = TC_TN * 32;                   // database side: [256 rows][16 bf16] = 8 KB slab
constexpr int TC_B_BYTES = 2 * TC_BHALF_BYTES + TC_BAUG_BYTES;   // 72 KB per stage
constexpr int TC_SMEM_BYTES = TC_A_BYTES + TC_STAGES * TC_B_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

// Instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128, M = 128.
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_TN >> 3) << 17) | ((128u >> 4) << 24);

// ---- operand preparation --------------------------------------------------------------------------------
// One warp per row: float32 -> bf16 operand row of 144 columns and the exactness check.  A value passes when it is
// an integer in [0, 256]: then bf16 holds v and 2v exactly, every partial sum of the contraction is an integer of
// magnitude <= 2^24, and s = 2 q.t - ||t||^2 >= -||t||^2 > -2^24 stays above the key of the padding rows (mixed-sign
// integers could push s below it and lose a real neighbour to a padding row: they take the split engine).  flag[0] |= 1 otherwise.
//   query side    (is_db == 0): [2q | 1 1 1 0 ...]
//   database side (is_db == 1): [t | -p0 -p1 -p2 0 ...], ||t||^2 = p0 + p1 + p2 (bytes 0, 1, 2 of the integer norm);
//                               rows n..n_pad are padding that can never be selected (s = -(2^24 - 1)).
__global__ void __launch_bounds__(256)
l2_tc_prep_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int is_db, __nv_bfloat16* __restrict__ xb,
                  int* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  float f[4] = {0.f, 0.f, 0.f, 0.f};
  bool bad = false;
  float s = 0.f;
  if (row < n) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * TC_D) + lane);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bad |= !(f[i] >= 0.f && f[i] <= 256.f) || f[i] != rintf(f[i]);
      s = fmaf(f[i], f[i], s);
    }
  }
  const float scale = is_db ? 1.f : 2.f;
  __nv_bfloat162 lo = __floats2bfloat162_rn(scale * f[0], scale * f[1]), hi = __floats2bfloat162_rn(scale * f[2], scale * f[3]);
  uint2 packed;
  packed.x = *reinterpret_cast<uint32_t*>(&lo);
  packed.y = *reinterpret_cast<uint32_t*>(&hi);
  __nv_bfloat16* dst = xb + row * TC_KA;
  reinterpret_cast<uint2*>(dst)[lane] = packed;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane < 2) {  // the 16 augmentation columns, 8 per lane
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (lane == 0) {
      if (!is_db) { a0 = a1 = a2 = 1.f; }
      else if (row < n) {
        const int in = bad ? 0 : (int)s;  // exact: s is an integer below 2^24 whenever the row passed the check
        a0 = -(float)(in & 0xFF); a1 = -(float)(in & 0xFF00); a2 = -(float)(in & 0xFF0000);
      } else { a0 = -255.f; a1 = -65280.f; a2 = -16711680.f; }
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a0, a1), p1 = __floats2bfloat162_rn(a2, 0.f);
    uint4 w = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1), 0u, 0u);
    reinterpret_cast<uint4*>(dst + TC_D)[lane] = w;
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flag, 1);
}

// ---- the contraction + fused top-2 ------------------------------------------------------------------------
struct TcTop2 {
  float k0, k1;
  int i0, i1;
};

__device__ __forceinline__ void tc_insert(TcTop2& t, float v, int col) {
  if (v > t.k1) {  // strict: columns reach a thread in increasing index order, so ties keep the lower index
    if (v > t.k0) { t.k1 = t.k0; t.i1 = t.i0; t.k0 = v; t.i0 = col; }
    else { t.k1 = v; t.i1 = col; }
  }
}

// 32 accumulator columns of one query row; each holds s = 2 q.t - ||t||^2.  One guard per 32 values (FMNMX3 tree) in
// front of the rare insertion.
__device__ __forceinline__ void tc_process32(const uint32_t (&acc)[32], int col, TcTop2& t) {
  float m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i)
    m[i] = max3(__uint_as_float(acc[3 * i]), __uint_as_float(acc[3 * i + 1]), __uint_as_float(acc[3 * i + 2]));
  m[10] = fmaxf(__uint_as_float(acc[30]), __uint_as_float(acc[31]));
  const float a = max3(m[0], m[1], m[2]), b = max3(m[3], m[4], m[5]), c = max3(m[6], m[7], m[8]);
  const float hi = fmaxf(max3(a, b, c), fmaxf(m[9], m[10]));
  if (hi > t.k1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) tc_insert(t, __uint_as_float(acc[j]), col + j);
  }
}

// part_idx [splits][nq][2].  `gate` (device int): the kernel does nothing unless *gate == 0.
// PAIR = true: launched as clusters of two CTAs that sweep the same database tiles in lock-step on two different query
// blocks; each CTA fetches HALF of every database tile and TMA-multicasts it into both CTAs' shared memory, so the
// L2 -> SM traffic of the sweep halves.  A stage is released by the MMA commits of BOTH CTAs (multicast arrive).
template <bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1)
l2_tc_knn2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_qa,
                  const __grid_constant__ CUtensorMap map_t, const __grid_constant__ CUtensorMap map_ta, int64_t nq, int64_t nt,
                  int q_blocks, int splits, int64_t rows_per_split, int32_t* __restrict__ part_idx,
                  const int* __restrict__ gate) {
  if (*gate != 0) return;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                          // [2 tiles][16 KB | 16 KB | 4 KB]
  uint8_t* smem_b = smem + TC_A_BYTES;             // [stages][32 KB | 32 KB | 8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_A_BYTES + TC_STAGES * TC_B_BYTES);
  uint64_t* full = bars;                           // [stages]   TMA -> MMA
  uint64_t* empty = bars + TC_STAGES;              // [stages]   MMA -> TMA
  uint64_t* tmem_full = bars + 2 * TC_STAGES;      // [2 query tiles] MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * TC_STAGES + 2; // [2 query tiles] epilogue -> MMA
  uint64_t* a_full = bars + 2 * TC_STAGES + 4;     // query tiles landed
  uint64_t* a_empty = bars + 2 * TC_STAGES + 5;    // all MMAs of the unit retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, PAIR ? 2 : 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full + b, 1); mbar_init(tmem_empty + b, 4); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Work assignment.  Single: unit u = (split, query block), CTA b takes u = b, b + grid, ...  Pair: a cluster takes a
  // PAIR of query blocks (2 qp + rank) of one split; an odd tail gives the second CTA a block beyond the queries (its
  // TMA reads zeros, its rows are never written) so that it still serves its half of the database tiles.
  const int rank = PAIR ? (int)(blockIdx.x & 1) : 0;
  const int q_groups = PAIR ? (q_blocks + 1) / 2 : q_blocks;
  const int total_units = q_groups * splits;
  const int u_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int u_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_t) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qa) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ta) : "memory");
      uint32_t it = 0, un = 0;
      for (int u = u_first; u < total_units; u += u_step, ++un) {
        const int split = u / q_groups, qb = PAIR ? 2 * (u - split * q_groups) + rank : u - split * q_groups;
        const int64_t t_begin = (int64_t)split * rows_per_split;
        const int64_t t_end = min(nt, t_begin + rows_per_split);
        const int tiles = (int)((t_end - t_begin + TC_TN - 1) / TC_TN);
        mbar_wait(a_empty, (un & 1) ^ 1);
        mbar_expect_tx(a_full, TC_A_BYTES);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          uint8_t* dst = smem_a + m * TC_TILE_BYTES;
          const int row0 = qb * TC_QB + m * 128;
          tma_load_2d(dst, &map_q, a_full, 0, row0);
          tma_load_2d(dst + TC_HALF_BYTES, &map_q, a_full, 64, row0);
          tma_load_2d(dst + 2 * TC_HALF_BYTES, &map_qa, a_full, TC_D, row0);
        }
        for (int j = 0; j < tiles; ++j, ++it) {
          const int s = it % TC_STAGES;
          mbar_wait(empty + s, ((it / TC_STAGES) & 1) ^ 1);
          mbar_expect_tx(full + s, TC_B_BYTES);
          const int row0 = (int)(t_begin + (int64_t)j * TC_TN);
          uint8_t* bs = smem_b + s * TC_B_BYTES;
          if (PAIR) {  // my 128 of the tile's 256 rows, into both CTAs (the maps carry 128-row boxes here)
            const int r0 = row0 + rank * 128;
            tma_load_2d_mc(bs + rank * (TC_BHALF_BYTES / 2), &map_t, full + s, 0, r0, 3);
            tma_load_2d_mc(bs + TC_BHALF_BYTES + rank * (TC_BHALF_BYTES / 2), &map_t, full + s, 64, r0, 3);
            tma_load_2d_mc(bs + 2 * TC_BHALF_BYTES + rank * (TC_BAUG_BYTES / 2), &map_ta, full + s, TC_D, r0, 3);
          } else {
            tma_load_2d(bs, &map_t, full + s, 0, row0);
            tma_load_2d(bs + TC_BHALF_BYTES, &map_t, full + s, 64, row0);
            tma_load_2d(bs + 2 * TC_BHALF_BYTES, &map_ta, full + s, TC_D, row0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the loop, one elected lane issues (see elect_one in tc_ptx.cuh) =====
    uint32_t it = 0, un = 0;
    const uint32_t a_addr = smem_u32(smem_a), b_addr = smem_u32(smem_b);
    for (int u = u_first; u < total_units; u += u_step, ++un) {
      const int split = u / q_groups;
      const int64_t t_begin = (int64_t)split * rows_per_split;
      const int64_t t_end = min(nt, t_begin + rows_per_split);
      const int tiles = (int)((t_end - t_begin + TC_TN - 1) / TC_TN);
      mbar_wait(a_full, un & 1);
      for (int j = 0; j < tiles; ++j, ++it) {
        const int s = it % TC_STAGES;
        mbar_wait(full + s, (it / TC_STAGES) & 1);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          mbar_wait(tmem_empty + m, (it & 1) ^ 1);   // the epilogue has drained this query tile's accumulator
          tc_fence_after();
          const uint32_t d_addr = tmem_base + (uint32_t)(m * TC_TN);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_D / 16; ++k) {
              const uint32_t koff_a = (uint32_t)((k >> 2) * TC_HALF_BYTES + (k & 3) * 32);
              const uint32_t koff_b = (uint32_t)((k >> 2) * TC_BHALF_BYTES + (k & 3) * 32);
              tc_mma_bf16(d_addr, umma_desc_sw128(a_addr + m * TC_TILE_BYTES + koff_a),
                          umma_desc_sw128(b_addr + s * TC_B_BYTES + koff_b), TC_IDESC, k > 0 ? 1u : 0u);
            }
            // ninth step: the augmentation columns subtract ||t||^2
            tc_mma_bf16(d_addr, umma_desc_sw32(a_addr + m * TC_TILE_BYTES + 2 * TC_HALF_BYTES),
                        umma_desc_sw32(b_addr + s * TC_B_BYTES + 2 * TC_BHALF_BYTES), TC_IDESC, 1u);
            tc_commit(tmem_full + m);   // this query tile's accumulator is ready for its epilogue warps
          }
          __syncwarp();
        }
        if (elect_one()) {
          if (PAIR) tc_commit_mc(empty + s, 3);  // ... in BOTH CTAs: either producer may overwrite either copy
          else tc_commit(empty + s);    // the stage may be refilled once both blocks have read it
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(a_empty);           // the query tiles may be replaced
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue: fused top-2 =====
    const int ew = warp - 4, quarter = ew & 3, m = ew >> 2;
    uint32_t it = 0;
    for (int u = u_first; u < total_units; u += u_step) {
      const int split = u / q_groups, qb = PAIR ? 2 * (u - split * q_groups) + rank : u - split * q_groups;
      const int64_t t_begin = (int64_t)split * rows_per_split;
      const int64_t t_end = min(nt, t_begin + rows_per_split);
      const int tiles = (int)((t_end - t_begin + TC_TN - 1) / TC_TN);
      TcTop2 top;
      top.k0 = top.k1 = __int_as_float(0xff800000);  // -inf
      top.i0 = top.i1 = -1;
      for (int j = 0; j < tiles; ++j, ++it) {
        const int col0 = (int)(t_begin + (int64_t)j * TC_TN);
        uint32_t acc0[32], acc1[32];
        mbar_wait(tmem_full + m, it & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(m * TC_TN);
        tc_ld32(taddr, acc0);
        tc_ld32(taddr + 32, acc1);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < TC_TN / 64 - 1; ++c) {  // 64 columns per round, the next 64 in flight meanwhile
          tc_process32(acc0, col0 + 64 * c, top);
          tc_ld32(taddr + 64 * (c + 1), acc0);
          tc_process32(acc1, col0 + 64 * c + 32, top);
          tc_ld32(taddr + 64 * (c + 1) + 32, acc1);
          tc_wait_ld();
        }
        // every column of this accumulator is in registers: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + m);
        tc_process32(acc0, col0 + TC_TN - 64, top);
        tc_process32(acc1, col0 + TC_TN - 32, top);
      }
      const int64_t row = (int64_t)qb * TC_QB + m * 128 + quarter * 32 + lane;
      if (row < nq) {
        int32_t* o = part_idx + ((int64_t)split * nq + row) * 2;
        o[0] = top.i0 < nt ? top.i0 : -1;  // a padding row can only surface when the split holds fewer than 2 rows
        o[1] = top.i1 < nt ? top.i1 : -1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // nobody leaves while the peer can still multicast into this CTA
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows][cols] bf16 or u8 row-major; box = `box_cols` columns x `box_rows` rows with the given swizzle; out-of-range
// rows read 0
int tc_make_map(CUtensorMap* map, const void* base, int64_t rows, int cols, int box_cols, CUtensorMapSwizzle swizzle,
                int box_rows, int elem_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  HPUSM_REQUIRE(fn, HPUSM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * (cuuint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HPUSM_REQUIRE(r == CUDA_SUCCESS, HPUSM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return HPUSM_OK;
}

bool l2_tc_available() { return true; }

static int tc_splits(int64_t nq, int64_t nt) {
  const int64_t qbs = (nq + TC_QB - 1) / TC_QB;
  const int64_t sms = sm_count();
  if (qbs >= 2 * sms) return 1;
  int64_t s = (2 * sms + qbs - 1) / qbs;
  const int64_t max_s = (nt + 2047) / 2048;  // at least 16 tiles per unit
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  // every split must own at least one row after its length is rounded up to whole tiles
  while (s > 1) {
    const int64_t rps = ((nt + s - 1) / s + TC_TN - 1) / TC_TN * TC_TN;
    if ((s - 1) * rps < nt) break;
    --s;
  }
  return (int)s;
}

struct TcLayout {
  size_t qb, tb, flag, part, total;
  int splits_tc, splits_simt, nparts;
  int64_t nt_pad;
};

static TcLayout tc_layout(int64_t nq, int64_t nt, int mode) {
  TcLayout L;
  L.nt_pad = (nt + TC_TN - 1) / TC_TN * TC_TN;
  L.splits_tc = tc_splits(nq, nt);
  L.splits_simt = 0;  // (the FP32 engine no longer shares this buffer: general floats go to the split engine)
  L.nparts = L.splits_tc > L.splits_simt ? L.splits_tc : L.splits_simt;
  size_t off = 0;
  L.qb = off; off += align_up((size_t)nq * TC_KA * 2, 1024);
  L.tb = off; off += align_up((size_t)L.nt_pad * TC_KA * 2, 1024);
  L.flag = off; off += 1024;
  L.part = off; off += align_up((size_t)L.nparts * nq * 2 * sizeof(int32_t), 1024);
  L.total = off;
  return L;
}

size_t l2_tc_workspace_bytes(int64_t nq, int64_t nt, int d, int mode) {
  if (d != TC_D || nq <= 0 || nt <= 0) return 256;
  const size_t a = (mode == HPUSM_L2_AUTO || mode == HPUSM_L2_TC_INT) ? tc_layout(nq, nt, mode).total : 0;
  const size_t b = (mode == HPUSM_L2_AUTO || mode == HPUSM_L2_TC_SPLIT) ? l2_tc_split_workspace_bytes(nq, nt) : 0;
  const size_t c = (mode == HPUSM_L2_AUTO || mode == HPUSM_L2_TC_I8) ? l2_tc_i8_workspace_bytes(nq, nt) : 0;
  const size_t e = (mode == HPUSM_L2_AUTO || mode == HPUSM_L2_TC_Q16) ? l2_tc_q16_workspace_bytes(nq, nt) : 0;
  const size_t ab = a > b ? a : b, ce = c > e ? c : e;
  return ab > ce ? ab : ce;
}

int l2_tc_knn2(const float* q, int64_t nq, const float* t, int64_t nt, int d, int mode, int32_t* idx, float* dist, void* ws,
               size_t ws_bytes, cudaStream_t st, int* resolved_mode, int64_t* fallback_rows) {
  HPUSM_REQUIRE(d == TC_D, HPUSM_ERR_UNSUPPORTED, "tensor-core L2 engine is built for d == %d, got %d", TC_D, d);
  if (mode == HPUSM_L2_TC_SPLIT) {
    *resolved_mode = HPUSM_L2_TC_SPLIT;
    return l2_tc_split_knn2(q, nq, t, nt, idx, dist, ws, ws_bytes, st, fallback_rows);
  }
  if (mode == HPUSM_L2_TC_Q16) {
    *resolved_mode = HPUSM_L2_TC_Q16;
    return l2_tc_q16_knn2(q, nq, t, nt, idx, dist, ws, ws_bytes, st, fallback_rows);
  }
  if (mode == HPUSM_L2_AUTO || mode == HPUSM_L2_TC_I8) {
    // u8 operands on the integer tensor pipe when every value is an integer in [0,255] (what cv2 SIFT produces)
    int eligible = 0;
    const int rc8 = l2_tc_i8_knn2(q, nq, t, nt, idx, dist, ws, ws_bytes, st, mode == HPUSM_L2_TC_I8, &eligible);
    if (rc8 != HPUSM_OK) return rc8;
    if (eligible) {
      *resolved_mode = HPUSM_L2_TC_I8;
      *fallback_rows = 0;
      return HPUSM_OK;
    }
  }
  const TcLayout L = tc_layout(nq, nt, mode);
  HPUSM_REQUIRE(ws_bytes >= L.total, HPUSM_ERR_WORKSPACE, "l2 tensor engine: workspace needs %zu bytes, got %zu", L.total,
                ws_bytes);
  char* base = reinterpret_cast<char*>(ws);
  __nv_bfloat16* qb = reinterpret_cast<__nv_bfloat16*>(base + L.qb);
  __nv_bfloat16* tb = reinterpret_cast<__nv_bfloat16*>(base + L.tb);
  int* flag = reinterpret_cast<int*>(base + L.flag);
  int32_t* part = reinterpret_cast<int32_t*>(base + L.part);

  HPUSM_CUDA_OK(cudaMemsetAsync(flag, 0, sizeof(int), st));
  HPUSM_CUDA_OK(cudaMemsetAsync(part, 0xff, (size_t)L.nparts * nq * 2 * sizeof(int32_t), st));
  HPUSM_LAUNCH(l2_tc_prep_kernel, (unsigned)((nq + 7) / 8), 256, 0, st, q, nq, nq, 0, qb, flag);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(l2_tc_prep_kernel, (unsigned)((L.nt_pad + 7) / 8), 256, 0, st, t, nt, L.nt_pad, 1, tb, flag);
  HPUSM_CHECK_LAUNCH();

  if (mode == HPUSM_L2_AUTO) {
    // one host read: are the descriptors integer valued (exact single-pass engine) or general floats (split engine)?
    int h_flag = 0;
    HPUSM_CUDA_OK(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    HPUSM_CUDA_OK(cudaStreamSynchronize(st));
    if (h_flag != 0) {
      // general floats: the fixed-point integer engine (l2_tc_q16.cu); HPUSM_L2_FLOAT=split keeps the bf16 split engine
      static const bool use_split = []() { const char* e = getenv("HPUSM_L2_FLOAT"); return e && e[0] == 's'; }();
      if (use_split) {
        *resolved_mode = HPUSM_L2_TC_SPLIT;
        return l2_tc_split_knn2(q, nq, t, nt, idx, dist, ws, ws_bytes, st, fallback_rows);
      }
      *resolved_mode = HPUSM_L2_TC_Q16;
      return l2_tc_q16_knn2(q, nq, t, nt, idx, dist, ws, ws_bytes, st, fallback_rows);
    }
  }
  CUtensorMap map_q, map_qa, map_t, map_ta;
  int rc = tc_make_map(&map_q, qb, nq, TC_KA, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == HPUSM_OK) rc = tc_make_map(&map_qa, qb, nq, TC_KA, 16, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc == HPUSM_OK) rc = tc_make_map(&map_t, tb, L.nt_pad, TC_KA, 64, CU_TENSOR_MAP_SWIZZLE_128B, TC_TN);
  if (rc == HPUSM_OK) rc = tc_make_map(&map_ta, tb, L.nt_pad, TC_KA, 16, CU_TENSOR_MAP_SWIZZLE_32B, TC_TN);
  if (rc != HPUSM_OK) return rc;

  // HPUSM_L2_PAIR=1 selects the 2-CTA cluster variant (database tiles multicast to both CTAs); default: single CTAs.
  static const bool use_pair = []() { const char* e = getenv("HPUSM_L2_PAIR"); return e && e[0] == '1'; }();
  if (use_pair) {
    rc = tc_make_map(&map_t, tb, L.nt_pad, TC_KA, 64, CU_TENSOR_MAP_SWIZZLE_128B, 128);
    if (rc == HPUSM_OK) rc = tc_make_map(&map_ta, tb, L.nt_pad, TC_KA, 16, CU_TENSOR_MAP_SWIZZLE_32B, 128);
    if (rc != HPUSM_OK) return rc;
  }
  static std::atomic<bool> attr_done[64];
  int dev = 0;
  HPUSM_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev].load(std::memory_order_acquire)) {
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_tc_knn2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_tc_knn2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_done[dev].store(true, std::memory_order_release);
  }
  const int q_blocks = (int)((nq + TC_QB - 1) / TC_QB);
  int64_t rows_per_split = (nt + L.splits_tc - 1) / L.splits_tc;
  rows_per_split = (rows_per_split + TC_TN - 1) / TC_TN * TC_TN;
  if (use_pair) {
    const int units = ((q_blocks + 1) / 2) * L.splits_tc;
    int clusters = sm_count() / 2;
    if (units < clusters) clusters = units;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TC_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int splits_i = L.splits_tc;
    const int* gate = flag;
    profile_mark(st, true);
    HPUSM_CUDA_OK(cudaLaunchKernelEx(&cfg, l2_tc_knn2_kernel<true>, map_q, map_qa, map_t, map_ta, nq, nt, q_blocks, splits_i,
                                     rows_per_split, part, gate));
    profile_mark(st, false);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  } else {
    const int units = q_blocks * L.splits_tc;
    const int grid = units < sm_count() ? units : sm_count();
    HPUSM_LAUNCH_TIMED(l2_tc_knn2_kernel<false>, grid, TC_THREADS, TC_SMEM_BYTES, st, map_q, map_qa, map_t, map_ta, nq, nt, q_blocks,
                       L.splits_tc, rows_per_split, part, flag);
  }
  HPUSM_CHECK_LAUNCH();
  *resolved_mode = HPUSM_L2_TC_INT;
  *fallback_rows = 0;
  if (mode == HPUSM_L2_TC_INT) {
    // explicit mode: the range check gates the sweep on the device (every idx stays -1) and is reported asynchronously
    rc = async_report(flag, HPUSM_ERR_UNSUPPORTED, st);
    if (rc != HPUSM_OK) return rc;
  }
  return l2_rerank_finalize(q, nq, t, d, part, L.nparts, 2, nullptr, 0, idx, dist, st);
}

}  // namespace hpusm
