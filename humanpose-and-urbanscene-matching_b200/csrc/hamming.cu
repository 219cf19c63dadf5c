// K2: brute-force Hamming kNN (k=2) for binary descriptors (ORB/AKAZE/BRISK), bit-exact.
//
// Replaces cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(..., k=2) at reference urbanscene/features.py:59.
//
// Shape of the kernel (B200): a CTA owns a tile of QPT*THREADS query descriptors held in registers
// (one 128-bit coalesced load per 16 bytes of descriptor) and sweeps a range of database descriptors
// that is staged through shared memory with cp.async double buffering.  Every thread reads the same
// database descriptor at the same time (shared-memory broadcast, conflict free) and scores it against
// its QPT private queries with XOR + POPC on the integer pipe; the running top-2 lives in registers,
// so there is no cross-thread reduction in the hot loop at all.  The database range is split across
// CTAs when there are too few query tiles to fill 148 SMs; partial top-2 lists are merged
// lexicographically on (distance, index), which reproduces OpenCV's "lowest train index wins" ties.
#include <limits.h>
#include "common.cuh"
#include "hamming_internal.cuh"

namespace hpusm {

constexpr int HAM_THREADS = 128;
constexpr int HAM_QPT = 4;                            // queries per thread
constexpr int HAM_QTILE = HAM_THREADS * HAM_QPT;      // 512 queries per CTA
constexpr int HAM_TTILE = 128;                        // database descriptors per smem stage

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// popcount of 256 bits of a ^ b with FOUR POPC instead of eight: four carry-save adders (3 words of equal weight ->
// sum + carry, one LOP3 each) fold the eight XOR words into ones {o3, x7}, twos {tw} and fours {f}.  POPC issues at
// 16 lanes/clk/SM, LOP3 at 64: 8 XOR + 8 LOP3 + 4 POPC balances the two pipes (0.25 cycles per pair on each) where the
// plain form is POPC-bound at 0.5.  Exact: d = popc(o3) + popc(x7) + 2 popc(tw) + 4 popc(f).
// (inline PTX: left to itself ptxas folds the XORs into the adders and rematerialises them, ~25 LOP3 per pair
// instead of 16, which makes the ALU pipe the new bound)
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t xor2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ int hamming256(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b) {
  const uint32_t x0 = xor2(a[0], b[0]), x1 = xor2(a[1], b[1]), x2 = xor2(a[2], b[2]), x3 = xor2(a[3], b[3]);
  const uint32_t x4 = xor2(a[4], b[4]), x5 = xor2(a[5], b[5]), x6 = xor2(a[6], b[6]), x7 = xor2(a[7], b[7]);
  const uint32_t o1 = xor3(x0, x1, x2), t1 = maj3(x0, x1, x2);
  const uint32_t o2 = xor3(x3, x4, x5), t2 = maj3(x3, x4, x5);
  const uint32_t o3 = xor3(o1, o2, x6), t3 = maj3(o1, o2, x6);
  const uint32_t tw = xor3(t1, t2, t3), f = maj3(t1, t2, t3);
  // combine with IMAD (FMA pipe, idle here) rather than LEA/IADD3 (ALU pipe, the bound)
  int d = __popc(o3) + __popc(x7), p2 = __popc(tw), p4 = __popc(f);
  asm("mad.lo.s32 %0, %1, 2, %0;" : "+r"(d) : "r"(p2));
  asm("mad.lo.s32 %0, %1, 4, %0;" : "+r"(d) : "r"(p4));
  return d;
}

struct HamTop2 {
  int d0, d1, i0, i1;
};

// Sweep database rows [t_begin, t_end) for this CTA's query tile. NW = 32-bit words per descriptor.
template <int NW>
__device__ __forceinline__ void hamming_sweep(const uint8_t* __restrict__ q, int64_t nq,
                                              const uint8_t* __restrict__ t, int64_t t_begin,
                                              int64_t t_end, int64_t q_base, HamTop2 (&top)[HAM_QPT],
                                              uint4* smem /* 2 stages */) {
  constexpr int V = NW / 4;                        // uint4 per descriptor
  constexpr int STAGE_V = HAM_TTILE * V;           // uint4 per stage
  const int tid = threadIdx.x;

  uint32_t qr[HAM_QPT][NW];
#pragma unroll
  for (int r = 0; r < HAM_QPT; ++r) {
    int64_t row = q_base + tid + (int64_t)r * HAM_THREADS;
    const uint4* src = reinterpret_cast<const uint4*>(q + row * (NW * 4));
#pragma unroll
    for (int v = 0; v < V; ++v) {
      uint4 x = (row < nq) ? __ldg(src + v) : make_uint4(0, 0, 0, 0);
      qr[r][4 * v + 0] = x.x; qr[r][4 * v + 1] = x.y; qr[r][4 * v + 2] = x.z; qr[r][4 * v + 3] = x.w;
    }
    top[r].d0 = INT_MAX; top[r].d1 = INT_MAX; top[r].i0 = -1; top[r].i1 = -1;
  }

  const int64_t n_rows = t_end - t_begin;
  if (n_rows <= 0) return;
  const int n_tiles = (int)((n_rows + HAM_TTILE - 1) / HAM_TTILE);
  const uint4* tsrc = reinterpret_cast<const uint4*>(t + t_begin * (NW * 4));

  auto issue = [&](int tile, int buf) {
    const int64_t first_v = (int64_t)tile * STAGE_V;
    const int64_t lim_v = n_rows * V;
    for (int v = tid; v < STAGE_V; v += HAM_THREADS) {
      int64_t g = first_v + v;
      if (g < lim_v) cp_async16(smem + buf * STAGE_V + v, tsrc + g);
    }
    cp_async_commit();
  };

  issue(0, 0);
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < n_tiles) {
      issue(tile + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int64_t tile_row0 = (int64_t)tile * HAM_TTILE;
    const int jmax = (int)min((int64_t)HAM_TTILE, n_rows - tile_row0);
    const uint4* sp = smem + buf * STAGE_V;
    const int idx0 = (int)(t_begin + tile_row0);
#pragma unroll 2
    for (int j = 0; j < jmax; ++j) {
      uint32_t w[NW];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        uint4 x = sp[j * V + v];
        w[4 * v + 0] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
      }
      int d[HAM_QPT];
      bool hit = false;
#pragma unroll
      for (int r = 0; r < HAM_QPT; ++r) {
        d[r] = 0;
#pragma unroll
        for (int k = 0; k < NW; k += 8) d[r] += hamming256(&qr[r][k], &w[k]);
        hit |= d[r] < top[r].d1;
      }
      // One warp-uniform branch per database row guards the (rare, after the first rows) top-2 update; left to
      // itself the compiler predicates the update and spends ~8 ALU-pipe selects per pair on it.
      if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
        for (int r = 0; r < HAM_QPT; ++r) {
          if (d[r] < top[r].d1) {  // strict '<' keeps the lower index on ties
            if (d[r] < top[r].d0) {
              top[r].d1 = top[r].d0; top[r].i1 = top[r].i0; top[r].d0 = d[r]; top[r].i0 = idx0 + j;
            } else {
              top[r].d1 = d[r]; top[r].i1 = idx0 + j;
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

// grid = (query tiles, splits).  Split s covers database rows [s*rows_per_split, ...).
template <int NW>
__global__ void __launch_bounds__(HAM_THREADS)
hamming_knn2_kernel(const uint8_t* __restrict__ q, int64_t nq, const uint8_t* __restrict__ t, int64_t nt,
                    int64_t rows_per_split, int32_t* __restrict__ out_idx, int32_t* __restrict__ out_dist) {
  __shared__ uint4 smem[2 * HAM_TTILE * (NW / 4)];
  const int64_t q_base = (int64_t)blockIdx.x * HAM_QTILE;
  const int split = blockIdx.y;
  const int64_t t_begin = (int64_t)split * rows_per_split;
  const int64_t t_end = min(nt, t_begin + rows_per_split);
  HamTop2 top[HAM_QPT];
  hamming_sweep<NW>(q, nq, t, t_begin, t_end, q_base, top, smem);
  int32_t* oi = out_idx + (int64_t)split * nq * 2;
  int32_t* od = out_dist + (int64_t)split * nq * 2;
#pragma unroll
  for (int r = 0; r < HAM_QPT; ++r) {
    int64_t row = q_base + threadIdx.x + (int64_t)r * HAM_THREADS;
    if (row < nq) {
      reinterpret_cast<int2*>(oi)[row] = make_int2(top[r].i0, top[r].i1);
      reinterpret_cast<int2*>(od)[row] =
          make_int2(top[r].i0 < 0 ? 0 : top[r].d0, top[r].i1 < 0 ? 0 : top[r].d1);
    }
  }
}

// grid = (query tiles, segments): one database image per blockIdx.y; fused Lowe ratio + survivor count.
template <int NW>
__global__ void __launch_bounds__(HAM_THREADS)
hamming_knn2_segmented_kernel(const uint8_t* __restrict__ q, int64_t nq, const uint8_t* __restrict__ t,
                              const int64_t* __restrict__ seg_offsets, double ratio,
                              int32_t* __restrict__ score, int32_t* __restrict__ best_idx) {
  __shared__ uint4 smem[2 * HAM_TTILE * (NW / 4)];
  __shared__ int warp_counts[HAM_THREADS / 32];
  const int64_t q_base = (int64_t)blockIdx.x * HAM_QTILE;
  const int64_t seg = blockIdx.y;
  const int64_t t_begin = seg_offsets[seg];
  const int64_t t_end = seg_offsets[seg + 1];
  HamTop2 top[HAM_QPT];
  hamming_sweep<NW>(q, nq, t, t_begin, t_end, q_base, top, smem);
  int survivors = 0;
#pragma unroll
  for (int r = 0; r < HAM_QPT; ++r) {
    int64_t row = q_base + threadIdx.x + (int64_t)r * HAM_THREADS;
    bool keep = false;
    if (row < nq) {
      // features.py:48  len(m) == 2 and m[0].distance < m[1].distance * ratio   (Python double)
      keep = top[r].i1 >= 0 && (double)top[r].d0 < (double)top[r].d1 * ratio;
      if (best_idx) best_idx[seg * nq + row] = keep ? top[r].i0 : -1;
    }
    survivors += keep ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) survivors += __shfl_xor_sync(0xffffffffu, survivors, o);
  if ((threadIdx.x & 31) == 0) warp_counts[threadIdx.x >> 5] = survivors;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
#pragma unroll
    for (int w = 0; w < HAM_THREADS / 32; ++w) total += warp_counts[w];
    if (gridDim.x == 1) score[seg] = total;
    else if (total) atomicAdd(score + seg, total);
  }
}

template <typename D>
__global__ void knn2_merge_kernel(const int32_t* __restrict__ part_idx, const D* __restrict__ part_dist,
                                  int nparts, int64_t nq, int32_t* __restrict__ idx, D* __restrict__ dist) {
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nq) return;
  Top2<D> best;
  best.d0 = best.d1 = D(0); best.i0 = best.i1 = -1;
  for (int p = 0; p < nparts; ++p) {
    const int64_t o = ((int64_t)p * nq + row) * 2;
    Top2<D> c;
    c.i0 = part_idx[o]; c.i1 = part_idx[o + 1]; c.d0 = part_dist[o]; c.d1 = part_dist[o + 1];
    top2_merge<D>(best, c);
  }
  idx[row * 2] = best.i0; idx[row * 2 + 1] = best.i1;
  dist[row * 2] = best.i0 < 0 ? D(0) : best.d0;
  dist[row * 2 + 1] = best.i1 < 0 ? D(0) : best.d1;
}

template <typename D>
__global__ void ratio_filter_kernel(const D* __restrict__ dist, const int32_t* __restrict__ idx, int64_t nq,
                                    double ratio, uint8_t* __restrict__ keep, int32_t* __restrict__ count) {
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool k = false;
  if (row < nq) {
    k = idx[row * 2 + 1] >= 0 && (double)dist[row * 2] < (double)dist[row * 2 + 1] * ratio;
    keep[row] = k ? 1 : 0;
  }
  unsigned b = __ballot_sync(0xffffffffu, k);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

// Split heuristic shared by the workspace query and the launcher.
static int hamming_splits(int64_t nq, int64_t nt) {
  const int64_t tiles = (nq + HAM_QTILE - 1) / HAM_QTILE;
  const int64_t want = (int64_t)sm_count() * 8;  // CTAs wanted to keep 148 SMs busy
  int64_t s = (want + tiles - 1) / tiles;
  const int64_t max_s = (nt + 2047) / 2048;      // at least 2048 database rows per split
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 65535) s = 65535;
  return (int)s;
}

}  // namespace hpusm

using namespace hpusm;

extern "C" {

size_t hpusm_knn2_hamming_workspace_bytes(int64_t nq, int64_t nt, int nbytes) {
  (void)nbytes;
  if (nq <= 0 || nt <= 0) return 256;
  const int s = hamming_splits(nq, nt);
  if (s == 1) return 256;
  return align_up((size_t)s * nq * 2 * sizeof(int32_t), 256) * 2 + 256;
}

int hpusm_knn2_hamming(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int nbytes,
                       int32_t* idx, int32_t* dist, void* ws, size_t ws_bytes, void* stream) {
  HPUSM_REQUIRE(nq >= 0 && nt >= 0, HPUSM_ERR_INVALID, "hpusm_knn2_hamming: negative size");
  HPUSM_REQUIRE(nbytes == 32 || nbytes == 64, HPUSM_ERR_UNSUPPORTED,
                "hpusm_knn2_hamming: descriptor width %d bytes (built for 32 and 64)", nbytes);
  if (nq == 0) return HPUSM_OK;
  HPUSM_REQUIRE(q && idx && dist && (t || nt == 0), HPUSM_ERR_INVALID, "hpusm_knn2_hamming: null pointer");
  HPUSM_REQUIRE(is_aligned(q, 16) && is_aligned(t, 16), HPUSM_ERR_INVALID,
                "hpusm_knn2_hamming: descriptor arrays must be 16-byte aligned");
  HPUSM_REQUIRE(nt < (int64_t)INT_MAX && nq < ((int64_t)1 << 40), HPUSM_ERR_UNSUPPORTED,
                "hpusm_knn2_hamming: size out of range");
  cudaStream_t st = as_stream(stream);
  const int64_t tiles = (nq + HAM_QTILE - 1) / HAM_QTILE;
  HPUSM_REQUIRE(tiles <= 0x7fffffff, HPUSM_ERR_UNSUPPORTED, "hpusm_knn2_hamming: too many query tiles");
  const int splits = nt == 0 ? 1 : hamming_splits(nq, nt);
  int32_t *pi = idx, *pd = dist;
  if (splits > 1) {
    const size_t part = align_up((size_t)splits * nq * 2 * sizeof(int32_t), 256);
    HPUSM_REQUIRE(ws && is_aligned(ws, 256) && ws_bytes >= 2 * part, HPUSM_ERR_WORKSPACE,
                  "hpusm_knn2_hamming: workspace needs %zu bytes, got %zu", 2 * part, ws_bytes);
    pi = reinterpret_cast<int32_t*>(ws);
    pd = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws) + part);
  }
  const int64_t rows_per_split = nt == 0 ? 1 : (nt + splits - 1) / splits;
  dim3 grid((unsigned)tiles, (unsigned)splits);
  if (nbytes == 32)
    HPUSM_LAUNCH_TIMED(hamming_knn2_kernel<8>, grid, HAM_THREADS, 0, st, q, nq, t, nt, rows_per_split, pi, pd);
  else
    HPUSM_LAUNCH_TIMED(hamming_knn2_kernel<16>, grid, HAM_THREADS, 0, st, q, nq, t, nt, rows_per_split, pi, pd);
  HPUSM_CHECK_LAUNCH();
  if (splits > 1) {
    HPUSM_LAUNCH(knn2_merge_kernel<int32_t>, (unsigned)((nq + 255) / 256), 256, 0, st, pi, pd, splits, nq,
                 idx, dist);
    HPUSM_CHECK_LAUNCH();
  }
  return HPUSM_OK;
}

// Which engine a segmented call runs on: 0 = unsupported request, else HPUSM_HAMMING_POPC / HPUSM_HAMMING_TC.
static int hamming_resolve_engine(int engine, int64_t nq, int64_t nt, int64_t nseg, int nbytes) {
  if (engine == HPUSM_HAMMING_POPC) return HPUSM_HAMMING_POPC;
  const bool ok = hamming_tc_eligible(nq, nt, nseg, nbytes);
  if (engine == HPUSM_HAMMING_TC) return ok ? HPUSM_HAMMING_TC : 0;
  return ok ? HPUSM_HAMMING_TC : HPUSM_HAMMING_POPC;
}

size_t hpusm_knn2_hamming_segmented_workspace_bytes(int64_t nq, int64_t nt, int64_t nseg, int nbytes, int engine) {
  if (nq <= 0 || nt <= 0 || nseg <= 0) return 256;
  if (hamming_resolve_engine(engine, nq, nt, nseg, nbytes) == HPUSM_HAMMING_TC) return hamming_tc_workspace_bytes(nq, nt, nseg);
  return 256;
}

int hpusm_knn2_hamming_segmented(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt,
                                 const int64_t* seg_offsets, int64_t nseg, int nbytes, double ratio, int engine,
                                 int32_t* score, int32_t* best_idx, void* ws, size_t ws_bytes,
                                 void* stream) {
  HPUSM_REQUIRE(nq >= 0 && nt >= 0 && nseg >= 0, HPUSM_ERR_INVALID,
                "hpusm_knn2_hamming_segmented: negative size");
  HPUSM_REQUIRE(nbytes == 32 || nbytes == 64, HPUSM_ERR_UNSUPPORTED,
                "hpusm_knn2_hamming_segmented: descriptor width %d bytes (built for 32 and 64)", nbytes);
  HPUSM_REQUIRE(engine == HPUSM_HAMMING_AUTO || engine == HPUSM_HAMMING_POPC || engine == HPUSM_HAMMING_TC, HPUSM_ERR_INVALID,
                "hpusm_knn2_hamming_segmented: unknown engine %d", engine);
  if (nseg == 0) return HPUSM_OK;
  HPUSM_REQUIRE(seg_offsets && score && (q || nq == 0) && (t || nt == 0), HPUSM_ERR_INVALID,
                "hpusm_knn2_hamming_segmented: null pointer");
  HPUSM_REQUIRE(is_aligned(q, 16) && is_aligned(t, 16), HPUSM_ERR_INVALID,
                "hpusm_knn2_hamming_segmented: descriptor arrays must be 16-byte aligned");
  HPUSM_REQUIRE(nt < (int64_t)INT_MAX && nseg <= 65535 * (int64_t)32768, HPUSM_ERR_UNSUPPORTED,
                "hpusm_knn2_hamming_segmented: size out of range");
  cudaStream_t st = as_stream(stream);
  HPUSM_CUDA_OK(cudaMemsetAsync(score, 0, (size_t)nseg * sizeof(int32_t), st));
  if (nq == 0) return HPUSM_OK;
  const int resolved = nt == 0 ? HPUSM_HAMMING_POPC : hamming_resolve_engine(engine, nq, nt, nseg, nbytes);
  HPUSM_REQUIRE(resolved != 0, HPUSM_ERR_UNSUPPORTED,
                "hpusm_knn2_hamming_segmented: the tensor-core engine takes 32-byte descriptors, >= 128 queries, >= 16384 database "
                "rows and images of >= 256 rows on average (got nbytes %d, nq %lld, nt %lld, nseg %lld)", nbytes, (long long)nq,
                (long long)nt, (long long)nseg);
  if (resolved == HPUSM_HAMMING_TC) {
    if (best_idx) HPUSM_CUDA_OK(cudaMemsetAsync(best_idx, 0xff, (size_t)nseg * nq * sizeof(int32_t), st));
    return hamming_tc_segmented(q, nq, t, nt, seg_offsets, nseg, ratio, score, best_idx, ws, ws_bytes, st);
  }
  const int64_t tiles = (nq + HAM_QTILE - 1) / HAM_QTILE;
  // gridDim.y is limited to 65535: walk the segments in slabs.
  for (int64_t s0 = 0; s0 < nseg; s0 += 65535) {
    const int64_t ns = nseg - s0 < 65535 ? nseg - s0 : 65535;
    dim3 grid((unsigned)tiles, (unsigned)ns);
    int32_t* bi = best_idx ? best_idx + s0 * nq : nullptr;
    if (nbytes == 32)
      HPUSM_LAUNCH_TIMED(hamming_knn2_segmented_kernel<8>, grid, HAM_THREADS, 0, st, q, nq, t, seg_offsets + s0,
                   ratio, score + s0, bi);
    else
      HPUSM_LAUNCH_TIMED(hamming_knn2_segmented_kernel<16>, grid, HAM_THREADS, 0, st, q, nq, t, seg_offsets + s0,
                   ratio, score + s0, bi);
    HPUSM_CHECK_LAUNCH();
  }
  return HPUSM_OK;
}

int hpusm_knn2_merge_f32(const int32_t* part_idx, const float* part_dist, int nparts, int64_t nq,
                         int32_t* idx, float* dist, void* stream) {
  HPUSM_REQUIRE(nparts >= 1 && nq >= 0, HPUSM_ERR_INVALID, "hpusm_knn2_merge_f32: bad sizes");
  if (nq == 0) return HPUSM_OK;
  HPUSM_REQUIRE(part_idx && part_dist && idx && dist, HPUSM_ERR_INVALID, "hpusm_knn2_merge_f32: null pointer");
  HPUSM_LAUNCH(knn2_merge_kernel<float>, (unsigned)((nq + 255) / 256), 256, 0, as_stream(stream), part_idx,
               part_dist, nparts, nq, idx, dist);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_knn2_merge_i32(const int32_t* part_idx, const int32_t* part_dist, int nparts, int64_t nq,
                         int32_t* idx, int32_t* dist, void* stream) {
  HPUSM_REQUIRE(nparts >= 1 && nq >= 0, HPUSM_ERR_INVALID, "hpusm_knn2_merge_i32: bad sizes");
  if (nq == 0) return HPUSM_OK;
  HPUSM_REQUIRE(part_idx && part_dist && idx && dist, HPUSM_ERR_INVALID, "hpusm_knn2_merge_i32: null pointer");
  HPUSM_LAUNCH(knn2_merge_kernel<int32_t>, (unsigned)((nq + 255) / 256), 256, 0, as_stream(stream), part_idx,
               part_dist, nparts, nq, idx, dist);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

int hpusm_ratio_filter(const void* dist, const int32_t* idx, int64_t nq, int dist_is_int, double ratio,
                       uint8_t* keep, int32_t* count, void* stream) {
  HPUSM_REQUIRE(nq >= 0, HPUSM_ERR_INVALID, "hpusm_ratio_filter: negative size");
  HPUSM_REQUIRE(count, HPUSM_ERR_INVALID, "hpusm_ratio_filter: null count");
  cudaStream_t st = as_stream(stream);
  HPUSM_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
  if (nq == 0) return HPUSM_OK;
  HPUSM_REQUIRE(dist && idx && keep, HPUSM_ERR_INVALID, "hpusm_ratio_filter: null pointer");
  const unsigned grid = (unsigned)((nq + 255) / 256);
  if (dist_is_int)
    HPUSM_LAUNCH(ratio_filter_kernel<int32_t>, grid, 256, 0, st, reinterpret_cast<const int32_t*>(dist), idx, nq,
                 ratio, keep, count);
  else
    HPUSM_LAUNCH(ratio_filter_kernel<float>, grid, 256, 0, st, reinterpret_cast<const float*>(dist), idx, nq,
                 ratio, keep, count);
  HPUSM_CHECK_LAUNCH();
  return HPUSM_OK;
}

}  // extern "C"
