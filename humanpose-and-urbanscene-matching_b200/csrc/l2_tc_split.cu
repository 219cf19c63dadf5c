// K1 (tensor-core engine for GENERAL float descriptors): hi/lo bf16 operand split on tcgen05, top-4 candidates per
// database split, exact FP32 re-rank with a proof that nothing outside the candidate lists can enter the top-2, and
// an exact brute-force pass for the (rare) query rows where the proof fails.
//
// Replaces cv2.BFMatcher(cv2.NORM_L2).knnMatch(..., k=2) at reference urbanscene/features.py:59 for descriptors
// that are not integer valued (l2_tc.cu handles those exactly in one pass).
//
//   q = qh + ql + rq,  t = th + tl + rt   (qh = bf16(q), ql = bf16(q - qh), |rq| <= 2^-18 |q|)
//   s' = 2 (qh.th + qh.tl + ql.th) - ||t||^2      approximates  s = 2 q.t - ||t||^2   with
//   |s' - s| <= E(q) = 4e-5 |q| |t|max + 2e-5 |t|max^2     (dropped ql.tl / residual terms 6*2^-18 |q||t|, FP32
//                                                            accumulation, FP32 norm; constants rounded up)
// Operand rows are [hi(128) | lo(128) | aug(16)] bf16 (544 B): query [2qh | 2ql | 1 1 1 0..], database
// [th | tl | -p0 -p1 -p2 0..] with ||t||^2 = p0 + p1 + p2 an exact three-way bf16 split of the FP32 norm.  Per database
// tile the MMA warp issues 8 + 8 + 8 + 1 = 25 tcgen05.mma (hi.hi, hi.lo, lo.hi, aug) into one accumulator.
// The epilogue keeps the 4 largest s' per query row and split.  l2_split_finalize_kernel then computes the exact
// FP32 distance of every candidate, takes the lexicographic (d2, index) top-2 and checks
//      d2(second) + E(q) < ||q||^2 - s'(4th candidate of split s)      for every split s,
// which proves that no database row outside the lists is closer than the second neighbour (north_star's tolerance
// for general floats -- ties within 1e-5 relative distance -- is not even needed: the answer is the exact FP32 one).
// Rows failing the check are listed and redone by the exact FP32 engine (l2_simt.cu) after one host read of the count.
#include <float.h>
#include "common.cuh"
#include "l2_internal.cuh"
#include "tc_ptx.cuh"

namespace hpusm {

constexpr int SP_D = 128;
constexpr int SP_K = 272;                        // operand row: hi 128 | lo 128 | aug 16
constexpr int SP_QB = 128;                       // queries per unit (one M=128 tile)
constexpr int SP_TN = 128;                       // database rows per tile
constexpr int SP_STAGES = 2;
constexpr int SP_THREADS = 256;                  // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue
constexpr int SP_M = 4;                          // candidates kept per query and split
constexpr int SP_SLAB = 128 * 128;               // [128 rows][64 bf16] swizzle-128B slab, 16 KB
constexpr int SP_AUG = 128 * 32;                 // [128 rows][16 bf16] swizzle-32B slab, 4 KB
constexpr int SP_TILE = 4 * SP_SLAB + SP_AUG;    // 68 KB: hi0 hi1 lo0 lo1 aug
constexpr int SP_SMEM = SP_TILE * (1 + SP_STAGES) + 1024 + 256;
constexpr uint32_t SP_IDESC = tc_idesc_bf16(SP_TN);
constexpr int64_t SP_FB_ROWS = 65536;            // fallback rows handled per pass of the exact engine

// ---- operand preparation --------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// One warp per row.  tn_max (device float, as ordered int bits) receives max ||t||^2 over the database.
__global__ void __launch_bounds__(256)
l2_split_prep_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int is_db, __nv_bfloat16* __restrict__ xb,
                     int* __restrict__ tn_max_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  float f[4] = {0.f, 0.f, 0.f, 0.f};
  float s = 0.f;
  if (row < n) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * SP_D) + lane);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) s = fmaf(f[i], f[i], s);
  }
  const float scale = is_db ? 1.f : 2.f;  // exact: a power of two
  __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_bf16(scale * f[i], h[i], l[i]);
  __nv_bfloat16* dst = xb + row * SP_K;
  uint2 ph, pl;
  ph.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
  ph.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
  pl.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
  pl.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
  reinterpret_cast<uint2*>(dst)[lane] = ph;
  reinterpret_cast<uint2*>(dst + SP_D)[lane] = pl;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane < 2) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (lane == 0) {
      if (!is_db) { a0 = a1 = a2 = 1.f; }
      else if (row < n) {
        const float p0 = __bfloat162float(__float2bfloat16_rn(s));
        const float r1 = s - p0;                                   // exact
        const float p1 = __bfloat162float(__float2bfloat16_rn(r1));
        const float p2 = r1 - p1;                                  // exact, at most 8 significant bits
        a0 = -p0; a1 = -p1; a2 = -p2;
        atomicMax(tn_max_bits, __float_as_int(s));                 // s >= 0: integer order == float order
      } else { a0 = -3.0e38f; }                                    // padding row: never a candidate
    }
    __nv_bfloat16 b0 = __float2bfloat16_rn(a0), b1 = __float2bfloat16_rn(a1), b2 = __float2bfloat16_rn(a2);
    uint4 w = make_uint4((uint32_t)__bfloat16_as_ushort(b0) | ((uint32_t)__bfloat16_as_ushort(b1) << 16),
                         (uint32_t)__bfloat16_as_ushort(b2), 0u, 0u);
    if (lane == 1) w = make_uint4(0u, 0u, 0u, 0u);
    reinterpret_cast<uint4*>(dst + 2 * SP_D)[lane] = w;
  }
}

// ---- contraction + fused top-4 ------------------------------------------------------------------------------------
struct Top4 {
  float k[SP_M];
  int i[SP_M];
};

__device__ __forceinline__ void top4_insert(Top4& t, float v, int col) {
  if (v > t.k[3]) {  // strict: equal keys keep the earlier (lower) index
    t.k[3] = v; t.i[3] = col;
#pragma unroll
    for (int j = 3; j > 0; --j) {
      if (t.k[j] > t.k[j - 1]) {
        const float tk = t.k[j]; t.k[j] = t.k[j - 1]; t.k[j - 1] = tk;
        const int ti = t.i[j]; t.i[j] = t.i[j - 1]; t.i[j - 1] = ti;
      }
    }
  }
}

__device__ __forceinline__ void sp_process32(const uint32_t (&acc)[32], int col, Top4& t) {
  float m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i)
    m[i] = max3(__uint_as_float(acc[3 * i]), __uint_as_float(acc[3 * i + 1]), __uint_as_float(acc[3 * i + 2]));
  m[10] = fmaxf(__uint_as_float(acc[30]), __uint_as_float(acc[31]));
  const float a = max3(m[0], m[1], m[2]), b = max3(m[3], m[4], m[5]), c = max3(m[6], m[7], m[8]);
  const float hi = fmaxf(max3(a, b, c), fmaxf(m[9], m[10]));
  if (hi > t.k[3]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) top4_insert(t, __uint_as_float(acc[j]), col + j);
  }
}

// cand [splits][nq][4] int32 (-1 = empty), kth [splits][nq] float = 4th largest s' of the split (-inf if fewer).
__global__ void __launch_bounds__(SP_THREADS, 1)
l2_tc_split_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_qa,
                   const __grid_constant__ CUtensorMap map_t, const __grid_constant__ CUtensorMap map_ta, int64_t nq, int64_t nt,
                   int q_blocks, int splits, int64_t rows_per_split, int32_t* __restrict__ cand, float* __restrict__ kth) {
  extern __shared__ uint8_t sp_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sp_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                // [hi0 | hi1 | lo0 | lo1 | aug]
  uint8_t* smem_b = smem + SP_TILE;      // [stages][same]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SP_TILE * (1 + SP_STAGES));
  uint64_t* full = bars;
  uint64_t* empty = bars + SP_STAGES;
  uint64_t* tmem_full = bars + 2 * SP_STAGES;
  uint64_t* tmem_empty = bars + 2 * SP_STAGES + 2;
  uint64_t* a_full = bars + 2 * SP_STAGES + 4;
  uint64_t* a_empty = bars + 2 * SP_STAGES + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SP_STAGES + 6);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SP_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full + b, 1); mbar_init(tmem_empty + b, 4); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_units = q_blocks * splits;

  auto load_tile = [&](uint8_t* dst, const CUtensorMap* m, const CUtensorMap* ma, uint64_t* bar, int row0) {
    tma_load_2d(dst, m, bar, 0, row0);
    tma_load_2d(dst + SP_SLAB, m, bar, 64, row0);
    tma_load_2d(dst + 2 * SP_SLAB, m, bar, 128, row0);
    tma_load_2d(dst + 3 * SP_SLAB, m, bar, 192, row0);
    tma_load_2d(dst + 4 * SP_SLAB, ma, bar, 256, row0);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, un = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++un) {
        const int split = u / q_blocks, qb = u - split * q_blocks;
        const int64_t t_begin = (int64_t)split * rows_per_split;
        const int64_t t_end = min(nt, t_begin + rows_per_split);
        const int tiles = (int)((t_end - t_begin + SP_TN - 1) / SP_TN);
        mbar_wait(a_empty, (un & 1) ^ 1);
        mbar_expect_tx(a_full, SP_TILE);
        load_tile(smem_a, &map_q, &map_qa, a_full, qb * SP_QB);
        for (int j = 0; j < tiles; ++j, ++it) {
          const int s = it % SP_STAGES;
          mbar_wait(empty + s, ((it / SP_STAGES) & 1) ^ 1);
          mbar_expect_tx(full + s, SP_TILE);
          load_tile(smem_b + s * SP_TILE, &map_t, &map_ta, full + s, (int)(t_begin + (int64_t)j * SP_TN));
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp walks the loop, one elected lane issues (see elect_one in tc_ptx.cuh)
    uint32_t it = 0, un = 0;
    const uint32_t a_addr = smem_u32(smem_a), b_addr = smem_u32(smem_b);
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++un) {
      const int split = u / q_blocks;
      const int64_t t_begin = (int64_t)split * rows_per_split;
      const int64_t t_end = min(nt, t_begin + rows_per_split);
      const int tiles = (int)((t_end - t_begin + SP_TN - 1) / SP_TN);
      mbar_wait(a_full, un & 1);
      for (int j = 0; j < tiles; ++j, ++it) {
        const int s = it % SP_STAGES, b = it & 1;
        mbar_wait(full + s, (it / SP_STAGES) & 1);
        mbar_wait(tmem_empty + b, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)(b * SP_TN);
        const uint32_t bs = b_addr + s * SP_TILE;
        if (elect_one()) {
          // hi.hi, hi.lo, lo.hi: (A part, B part) = (0,0), (0,1), (1,0); a part = two 16 KB slabs
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint32_t ap = a_addr + (term == 2 ? 2 * SP_SLAB : 0);
            const uint32_t bp = bs + (term == 1 ? 2 * SP_SLAB : 0);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t koff = (uint32_t)((k >> 2) * SP_SLAB + (k & 3) * 32);
              tc_mma_bf16(d_addr, umma_desc_sw128(ap + koff), umma_desc_sw128(bp + koff), SP_IDESC, (term | k) ? 1u : 0u);
            }
          }
          tc_mma_bf16(d_addr, umma_desc_sw32(a_addr + 4 * SP_SLAB), umma_desc_sw32(bs + 4 * SP_SLAB), SP_IDESC, 1u);
          tc_commit(empty + s);
          tc_commit(tmem_full + b);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(a_empty);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int split = u / q_blocks, qb = u - split * q_blocks;
      const int64_t t_begin = (int64_t)split * rows_per_split;
      const int64_t t_end = min(nt, t_begin + rows_per_split);
      const int tiles = (int)((t_end - t_begin + SP_TN - 1) / SP_TN);
      Top4 top;
#pragma unroll
      for (int j = 0; j < SP_M; ++j) { top.k[j] = __int_as_float(0xff800000); top.i[j] = -1; }
      for (int j = 0; j < tiles; ++j, ++it) {
        const int b = it & 1;
        const int col0 = (int)(t_begin + (int64_t)j * SP_TN);
        uint32_t acc0[32], acc1[32];
        mbar_wait(tmem_full + b, (it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * SP_TN);
        tc_ld32(taddr, acc0);
        tc_ld32(taddr + 32, acc1);
        tc_wait_ld();
        sp_process32(acc0, col0, top);
        tc_ld32(taddr + 64, acc0);
        sp_process32(acc1, col0 + 32, top);
        tc_ld32(taddr + 96, acc1);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + b);
        sp_process32(acc0, col0 + 64, top);
        sp_process32(acc1, col0 + 96, top);
      }
      const int64_t row = (int64_t)qb * SP_QB + quarter * 32 + lane;
      if (row < nq) {
        int4 o;
        o.x = top.i[0] < nt ? top.i[0] : -1; o.y = top.i[1] < nt ? top.i[1] : -1;
        o.z = top.i[2] < nt ? top.i[2] : -1; o.w = top.i[3] < nt ? top.i[3] : -1;
        reinterpret_cast<int4*>(cand)[(int64_t)split * nq + row] = o;
        // bound on every row of the split that is NOT in the list; a padding row in 4th place means "all rows listed"
        kth[(int64_t)split * nq + row] = (top.i[3] >= 0 && top.i[3] < nt) ? top.k[3] : __int_as_float(0xff800000);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
  }
}

// ---- exact re-rank + proof of completeness --------------------------------------------------------------------------
// One warp per query: lanes take candidates (exact FP32 distance, increasing-k fma order = the number the FP32 engine
// and the final answer use), warp-wide lexicographic top-2, then the completeness check.  Unproven rows go to `list`.
__global__ void __launch_bounds__(256)
l2_split_finalize_kernel(const float* __restrict__ q, int64_t nq, const float* __restrict__ t, const int32_t* __restrict__ cand,
                         const float* __restrict__ kth, int nparts, const int* __restrict__ tn_max_bits, int32_t* __restrict__ idx,
                         float* __restrict__ dist, int32_t* __restrict__ list, int* __restrict__ n_list) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= nq) return;
  const float4 qv = __ldg(reinterpret_cast<const float4*>(q + row * SP_D) + lane);
  float qn = qv.x * qv.x + qv.y * qv.y + qv.z * qv.z + qv.w * qv.w;
  for (int o = 16; o > 0; o >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, o);
  Top2<float> best;
  best.d0 = best.d1 = 0.f; best.i0 = best.i1 = -1;
  float kmax = __int_as_float(0xff800000);
  const int C = nparts * SP_M;
  for (int c = lane; c < C; c += 32) {
    const int p = c / SP_M;
    const int32_t ci = cand[((int64_t)p * nq + row) * SP_M + (c - p * SP_M)];
    if (ci >= 0) {
      const float4* a = reinterpret_cast<const float4*>(q + row * SP_D);
      const float4* b = reinterpret_cast<const float4*>(t + (int64_t)ci * SP_D);
      float s = 0.f;
      for (int kv = 0; kv < SP_D / 4; ++kv) {
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
  for (int p = lane; p < nparts; p += 32) kmax = fmaxf(kmax, kth[(int64_t)p * nq + row]);
  for (int o = 16; o > 0; o >>= 1) {
    Top2<float> other;
    other.d0 = __shfl_xor_sync(0xffffffffu, best.d0, o); other.d1 = __shfl_xor_sync(0xffffffffu, best.d1, o);
    other.i0 = __shfl_xor_sync(0xffffffffu, best.i0, o); other.i1 = __shfl_xor_sync(0xffffffffu, best.i1, o);
    top2_merge<float>(best, other);
    kmax = fmaxf(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  if (lane == 0) {
    const float tn_max = __int_as_float(*tn_max_bits);
    const float E = 4e-5f * sqrtf(qn * tn_max) + 2e-5f * tn_max + 1e-6f * qn;
    // rows outside the lists have d2 >= qn - kmax - E; the listed second neighbour must beat that strictly
    const bool proven = (kmax == __int_as_float(0xff800000)) || (best.i1 >= 0 && best.d1 + E < qn - kmax);
    if (proven) {
      idx[row * 2] = best.i0; idx[row * 2 + 1] = best.i1;
      dist[row * 2] = best.i0 < 0 ? 0.f : __fsqrt_rn(best.d0);
      dist[row * 2 + 1] = best.i1 < 0 ? 0.f : __fsqrt_rn(best.d1);
    } else {
      list[atomicAdd(n_list, 1)] = (int32_t)row;
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------
static int split_splits(int64_t nq, int64_t nt) {
  const int64_t qbs = (nq + SP_QB - 1) / SP_QB;
  const int64_t sms = sm_count();
  if (qbs >= 2 * sms) return 1;
  int64_t s = (2 * sms + qbs - 1) / qbs;
  const int64_t max_s = (nt + 2047) / 2048;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 32) s = 32;
  // every split must own at least one row after its length is rounded up to whole tiles
  while (s > 1) {
    const int64_t rps = ((nt + s - 1) / s + SP_TN - 1) / SP_TN * SP_TN;
    if ((s - 1) * rps < nt) break;
    --s;
  }
  return (int)s;
}

struct SplitLayout {
  size_t qb, tb, cand, kth, scal, list, fb, total;
  int splits, fb_splits;
  int64_t nt_pad;
};

static SplitLayout split_layout(int64_t nq, int64_t nt) {
  SplitLayout L;
  L.nt_pad = (nt + SP_TN - 1) / SP_TN * SP_TN;
  L.splits = split_splits(nq, nt);
  L.fb_splits = l2_simt_splits(64, nt);  // the fallback normally has few rows: split the database for parallelism
  size_t off = 0;
  L.qb = off; off += align_up((size_t)nq * SP_K * 2, 1024);
  L.tb = off; off += align_up((size_t)L.nt_pad * SP_K * 2, 1024);
  L.cand = off; off += align_up((size_t)L.splits * nq * SP_M * sizeof(int32_t), 1024);
  L.kth = off; off += align_up((size_t)L.splits * nq * sizeof(float), 1024);
  L.scal = off; off += 1024;
  L.list = off; off += align_up((size_t)nq * sizeof(int32_t), 1024);
  L.fb = off; off += align_up((size_t)L.fb_splits * (nq < SP_FB_ROWS ? nq : SP_FB_ROWS) * 2 * sizeof(int32_t), 1024);
  L.total = off;
  return L;
}

size_t l2_tc_split_workspace_bytes(int64_t nq, int64_t nt) {
  if (nq <= 0 || nt <= 0) return 256;
  return split_layout(nq, nt).total;
}

int l2_tc_split_knn2(const float* q, int64_t nq, const float* t, int64_t nt, int32_t* idx, float* dist, void* ws, size_t ws_bytes,
                     cudaStream_t st, int64_t* fallback_rows) {
  const SplitLayout L = split_layout(nq, nt);
  HPUSM_REQUIRE(ws_bytes >= L.total, HPUSM_ERR_WORKSPACE, "l2 split engine: workspace needs %zu bytes, got %zu", L.total, ws_bytes);
  char* base = reinterpret_cast<char*>(ws);
  __nv_bfloat16* qb = reinterpret_cast<__nv_bfloat16*>(base + L.qb);
  __nv_bfloat16* tb = reinterpret_cast<__nv_bfloat16*>(base + L.tb);
  int32_t* cand = reinterpret_cast<int32_t*>(base + L.cand);
  float* kth = reinterpret_cast<float*>(base + L.kth);
  int* tn_max = reinterpret_cast<int*>(base + L.scal);
  int* n_list = tn_max + 1;
  int32_t* list = reinterpret_cast<int32_t*>(base + L.list);
  int32_t* fb = reinterpret_cast<int32_t*>(base + L.fb);

  HPUSM_CUDA_OK(cudaMemsetAsync(tn_max, 0, 2 * sizeof(int), st));
  HPUSM_LAUNCH(l2_split_prep_kernel, (unsigned)((nq + 7) / 8), 256, 0, st, q, nq, nq, 0, qb, tn_max);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(l2_split_prep_kernel, (unsigned)((L.nt_pad + 7) / 8), 256, 0, st, t, nt, L.nt_pad, 1, tb, tn_max);
  HPUSM_CHECK_LAUNCH();
  CUtensorMap map_q, map_qa, map_t, map_ta;
  int rc = tc_make_map(&map_q, qb, nq, SP_K, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == HPUSM_OK) rc = tc_make_map(&map_qa, qb, nq, SP_K, 16, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc == HPUSM_OK) rc = tc_make_map(&map_t, tb, L.nt_pad, SP_K, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == HPUSM_OK) rc = tc_make_map(&map_ta, tb, L.nt_pad, SP_K, 16, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc != HPUSM_OK) return rc;
  static std::atomic<bool> attr_done[64];
  int dev = 0;
  HPUSM_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev].load(std::memory_order_acquire)) {
    HPUSM_CUDA_OK(cudaFuncSetAttribute(l2_tc_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
    attr_done[dev].store(true, std::memory_order_release);
  }
  const int q_blocks = (int)((nq + SP_QB - 1) / SP_QB);
  int64_t rows_per_split = (nt + L.splits - 1) / L.splits;
  rows_per_split = (rows_per_split + SP_TN - 1) / SP_TN * SP_TN;
  const int units = q_blocks * L.splits;
  const int grid = units < sm_count() ? units : sm_count();
  HPUSM_LAUNCH_TIMED(l2_tc_split_kernel, grid, SP_THREADS, SP_SMEM, st, map_q, map_qa, map_t, map_ta, nq, nt, q_blocks, L.splits,
                     rows_per_split, cand, kth);
  HPUSM_CHECK_LAUNCH();
  HPUSM_LAUNCH(l2_split_finalize_kernel, (unsigned)((nq + 7) / 8), 256, 0, st, q, nq, t, cand, kth, L.splits, tn_max, idx, dist,
               list, n_list);
  HPUSM_CHECK_LAUNCH();
  // the one host read of the engine: how many rows could not be proven complete
  int h_list = 0;
  HPUSM_CUDA_OK(cudaMemcpyAsync(&h_list, n_list, sizeof(int), cudaMemcpyDeviceToHost, st));
  HPUSM_CUDA_OK(cudaStreamSynchronize(st));
  *fallback_rows = h_list;
  for (int64_t lo = 0; lo < h_list; lo += SP_FB_ROWS) {
    const int64_t m = h_list - lo < SP_FB_ROWS ? h_list - lo : SP_FB_ROWS;
    rc = l2_simt_candidates(q, nq, t, nt, SP_D, list + lo, m, L.fb_splits, fb, st);
    if (rc != HPUSM_OK) return rc;
    rc = l2_rerank_finalize(q, nq, t, SP_D, fb, L.fb_splits, 2, list + lo, m, idx, dist, st);
    if (rc != HPUSM_OK) return rc;
  }
  return HPUSM_OK;
}

}  // namespace hpusm
