// Subsystem (3), tensor-core path: mutual nearest neighbours on tcgen05 (sm_100a).
//
// Replaces sim = A @ B.T; max(dim=1); max(dim=0) of the reference matchers
// (evaluations/hpatches/evaluation.py:27-38, losses/preprocess_utils.py:795-803)
// without ever writing the N x M similarity matrix.
//
// Per direction (rows of X against all rows of Y; run for (A,B) and (B,A)):
//   * operands are rounded once to bf16 (prep kernel; it also records |x|, the
//     norm of the rounding error |x~ - x| per row, and their maxima);
//   * a persistent, warp-specialised kernel walks work units (256 rows of X,
//     a range of 128-row Y tiles): warp 0 feeds shared memory with TMA
//     (SWIZZLE_128B boxes), warp 1 issues tcgen05.mma (M=128, N=128, K=16, bf16,
//     fp32 accumulators in TMEM, two accumulator stages = all 512 columns),
//     warps 4..19 drain TMEM with tcgen05.ld and reduce every 8 consecutive
//     columns of a row to their maximum, stored as fp16 in a chunk-maximum
//     table [rows][M/8] (1/16 of the bytes of the similarity matrix, which is
//     itself never written);
//   * the rescoring kernel takes the table row of each x, finds the chunks whose
//     maximum is within delta of the row maximum, where
//     delta_i = 2 (|e_xi| max|y~| + |x_i| max|e_y|) + slack bounds twice the error
//     of an approximate similarity (Cauchy-Schwarz on x~.y~ - x.y plus the fp16
//     rounding of the table), so the true argmax always lies in such a chunk;
//     it re-evaluates those chunks in float32, then every column within the
//     float32 error band exactly (float64 accumulation -- the value the exact
//     SIMT kernel uses) and takes the argmax, first index on ties.
// The result therefore never depends on the approximation.
#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "mnn_common.cuh"
#include "tc_common.cuh"

namespace posfeat {

constexpr int kD = 128;
constexpr int kXRows = 256;   // rows of X per work unit: one cta_group::2 MMA tile, 128 rows per CTA of the pair
constexpr int kYRows = 256;   // rows of Y per tile (MMA N); each CTA of the pair stages 128 of them
constexpr int kStages = 5;    // Y ring depth (per CTA: 5 x 32 KB)
constexpr int kSubBytes = 128 * 64 * 2;  // one [128 rows x 64 K] bf16 swizzled sub-tile
constexpr int kChunk = 8;     // candidate granularity (columns)
constexpr int kMaxSplits = 8;
constexpr int kEpiWarps = 16;
constexpr int kTcThreads = (4 + kEpiWarps) * 32;
constexpr float kAccSlack = 3.0e-5f;   // fp32 accumulation error of the tensor core, relative to |x||y|
constexpr float kEps32 = 1.0e-5f;      // float32 dot-product error bound (>= 128 * 2^-24), relative to |x||y|
constexpr float kHalfSlack = 9.8e-4f;  // 2 * 2^-11: fp16 rounding of two table entries of magnitude <= 1
constexpr float kG8Slack = 1.0f - 9.0f / 1024.0f;   // group entries: fp16 rounding + 3 replaced mantissa bits < 8.5 ulps <= 8.5 * 2^-10 relative

constexpr int kSmemX = 0;                      // 2 buffers x 2 K-halves
constexpr int kSmemY = 4 * kSubBytes;          // kStages x 2 K-halves
constexpr int kSmemBar = kSmemY + kStages * 2 * kSubBytes;
constexpr int kSmemTotal = kSmemBar + 256;
constexpr int kSmemAlloc = kSmemTotal + 1024;

// per-matrix scalars produced by the prep kernel (float bits, combined with atomicMax)

struct DirParams {            // all arrays are batched over pairs: index = pair * stride + ...
  const float* xnorm;          // [pairs][NXpad] |x_i|
  const float* xerr;           // [pairs][NXpad] |x~_i - x_i|
  const MatStats* xstats;      // [pairs][2] -> element pair*2 + which (set up per direction)
  const MatStats* ystats;
  __half* table;               // [pairs][NXpad][pitch] chunk maxima, scaled by 1/(max|x| max|y|)
  int pitch;                   // y_tiles * 32 chunks per row
  int NX, NY, NXpad, NYpad;
  int splits, tiles_per_split, y_tiles;
  int row_blocks;              // NXpad / 256
};
// Matches-only path (kLists epilogue): instead of the chunk-maximum table the epilogue leaves
//   * per (row, 64-column quarter j of the 256-column tiles, column split): a 128-byte LIST of the chunks that came
//     within delta of the row's running maximum when they were produced -- a superset of the chunks within delta
//     of the final row maximum (the running maximum only grows), a handful of entries instead of a 2 KB table row;
//   * per (8 consecutive rows, chunk): the GROUP ENTRY half2(max1 | leader, max2) of the 8 rows' chunk maxima
//     (non-negative, scaled, fp16 rounded up to a multiple of 8 ulps with the leader's row-in-group in the three
//     low bits), 1/64 of the similarity matrix's element count.  A column chunk's competitor rows are then the
//     leaders of the groups with max1 >= threshold, or all 8 rows of a group whose max2 reaches it as well.
constexpr int kListSlots = 16;          // slot 0: {count, running max}; slots 1..15: {first chunk of a tile quarter | chunk mask << 24, quarter maximum}
struct ListParams {
  uint2* lists;                // [pairs][NXpad][4 * splits][kListSlots]
  unsigned* g8;                // [pairs][pitch][groups]
  int groups;                  // NXpad / 8
};
struct TcParams {
  DirParams d[2];
  int pairs;
  int units0, units_pair;      // units of direction 0 / of both directions, per pair
  int units_total;             // pairs * units_pair
  int debug;   // POSFEAT_TC_DEBUG bits (bring-up only)
  ListParams L;                // kLists kernels only
};

// 1 / (max|x| max|y|): keeps every table entry inside [-1, 1]
__device__ __forceinline__ float table_scale(const DirParams& d, int pair) {
  const float m = __uint_as_float(d.xstats[2 * pair].max_norm) * __uint_as_float(d.ystats[2 * pair].max_norm);
  return m > 0.f ? 1.f / m : 1.f;      // an all-zero operand: every similarity is 0, any finite scale will do
}

__device__ __forceinline__ float row_delta(const DirParams& d, int pair, int row) {
  const MatStats& ys = d.ystats[2 * pair];
  const float yb = __uint_as_float(ys.max_norm_bf), ye = __uint_as_float(ys.max_err);
  const float yn = __uint_as_float(ys.max_norm);
  const size_t r = (size_t)pair * d.NXpad + row;
  const float xn = d.xnorm[r];
  return 2.f * (d.xerr[r] * yb + xn * ye + kAccSlack * xn * yn);
}

// ---------------------------------------------------------------- PTX helpers
// ---- CTA-pair (cta_group::2) variants
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> leader CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// both CTAs load into their own shared memory; the bytes are accounted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once all prior MMAs retire) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3) : "memory");
}
// arrive on the barrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n}"
      ::"r"(bar), "r"(rank) : "memory");
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=256 (CTA pair), N=256
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kYRows >> 3) << 17) | ((256u >> 4) << 24);

struct UnitInfo {
  int pair, dir, rb, sp, t0, t1;
};
__device__ __forceinline__ UnitInfo decode_unit(const TcParams& p, int u) {
  UnitInfo q;
  q.pair = u / p.units_pair;
  int v = u - q.pair * p.units_pair;
  q.dir = v >= p.units0 ? 1 : 0;
  if (q.dir) v -= p.units0;
  const DirParams& d = p.d[q.dir];
  q.rb = v / d.splits;
  q.sp = v - q.rb * d.splits;
  q.t0 = q.sp * d.tiles_per_split;
  q.t1 = min(d.y_tiles, q.t0 + d.tiles_per_split);
  return q;
}

__device__ __forceinline__ float max8(const float* f) {
  return fmaxf(fmaxf(fmaxf(f[0], f[1]), f[2]), fmaxf(fmaxf(fmaxf(f[3], f[4]), f[5]), fmaxf(f[6], f[7])));
}

// 32 accumulator columns of the row owned by this thread -> four scaled chunk
// maxima packed as 2 x half2
template <bool kRagged>
__device__ __forceinline__ uint2 reduce32(const uint32_t (&v)[32], int col0, int NY, float scale) {
  float m[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = __uint_as_float(v[c * 8 + j]);
      if (kRagged && col0 + c * 8 + j >= NY) f[j] = -INFINITY;   // zero padding rows of Y
    }
    m[c] = max8(f) * scale;
  }
  const __half2 lo = __floats2half2_rn(m[0], m[1]), hi = __floats2half2_rn(m[2], m[3]);
  uint2 r;
  r.x = *reinterpret_cast<const unsigned*>(&lo);
  r.y = *reinterpret_cast<const unsigned*>(&hi);
  return r;
}

// 32 accumulator columns of the row owned by this thread -> four (unscaled) chunk maxima
template <bool kRagged, int kOff>
__device__ __forceinline__ void chunkmax32(const uint32_t (&v)[32], int col0, int NY, float (&m)[8]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = __uint_as_float(v[c * 8 + j]);
      if (kRagged && col0 + c * 8 + j >= NY) f[j] = -INFINITY;   // zero padding rows of Y
    }
    m[kOff + c] = max8(f);
  }
}

__device__ __forceinline__ __half2 as_h2(unsigned u) { return *reinterpret_cast<const __half2*>(&u); }
__device__ __forceinline__ unsigned as_u32(__half2 h) { return *reinterpret_cast<const unsigned*>(&h); }

// The 8 chunk maxima of this thread's row -> the group entry of ONE chunk per lane: after three transposing
// exchanges inside the 8-lane group (rows 8g .. 8g+7 of the warp's 32 rows), lane l holds
// half2(max1 | leader, max2) of chunk (l & 7) over the group's rows.  Values are scaled into [0, 1], clamped at
// zero and rounded to fp16 by one conversion, and the three low bits are replaced by the row's index in the group:
// for non-negative halves the bit pattern orders like the value, so the maximum carries its row along.  A stored
// value differs from the scaled similarity by less than 8.5 fp16 ulps; the thresholds it is compared with are
// lowered by that much (kG8Slack).
__device__ __forceinline__ unsigned g8_reduce(const float (&m)[8], float scale, int lane) {
  const unsigned lb = (unsigned)(lane & 7) * 0x00010001u;
  unsigned E[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned h;    // packed fp16 pair of the two scaled values, negative values clamped to zero by the conversion
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(m[2 * k + 1] * scale), "f"(m[2 * k] * scale));
    E[k] = (h & 0xFFF8FFF8u) | lb;
  }
  constexpr unsigned kFull = 0xffffffffu;
  // exchange over lane bit 2: keep chunks 4*b2 .. 4*b2+3
  const bool b2 = lane & 4;
  const unsigned k0 = b2 ? E[2] : E[0], k1 = b2 ? E[3] : E[1];
  const unsigned r0 = __shfl_xor_sync(kFull, b2 ? E[0] : E[2], 4), r1 = __shfl_xor_sync(kFull, b2 ? E[1] : E[3], 4);
  const __half2 a1_0 = __hmax2(as_h2(k0), as_h2(r0)), a2_0 = __hmin2(as_h2(k0), as_h2(r0));
  const __half2 a1_1 = __hmax2(as_h2(k1), as_h2(r1)), a2_1 = __hmin2(as_h2(k1), as_h2(r1));
  // exchange over lane bit 1: keep chunks 4*b2 + 2*b1 + {0, 1}
  const bool b1 = lane & 2;
  const __half2 q1 = b1 ? a1_1 : a1_0, q2 = b1 ? a2_1 : a2_0;
  const __half2 o1 = as_h2(__shfl_xor_sync(kFull, as_u32(b1 ? a1_0 : a1_1), 2));
  const __half2 o2 = as_h2(__shfl_xor_sync(kFull, as_u32(b1 ? a2_0 : a2_1), 2));
  const __half2 n1 = __hmax2(q1, o1), n2 = __hmax2(__hmin2(q1, o1), __hmax2(q2, o2));
  // exchange over lane bit 0: keep chunk 4*b2 + 2*b1 + b0 as half2(first, second)
  const bool b0 = lane & 1;
  const unsigned keep = __byte_perm(as_u32(n1), as_u32(n2), b0 ? 0x7632 : 0x5410);
  const unsigned recv = __shfl_xor_sync(kFull, __byte_perm(as_u32(n1), as_u32(n2), b0 ? 0x5410 : 0x7632), 1);
  const __half2 mx = __hmax2(as_h2(keep), as_h2(recv)), mn = __hmin2(as_h2(keep), as_h2(recv));
  // first = mx.lo, second = max(mn.lo, mx.hi)
  return as_u32(__hmax2(as_h2(__byte_perm(as_u32(mn), 0u, 0x1010)), mx));
}

// A full list: the running maximum only grows, so earlier entries below the present threshold are dead -- drop them
// (this thread's own stores, re-read) before giving up on the row.  Out of line: it almost never runs.
__device__ __noinline__ int list_compact(uint2* line, int cnt, float thr) {
  if (cnt > kListSlots - 1) return cnt;          // already overflowed
  int k = 0;
  for (int sl = 1; sl < kListSlots; ++sl) {
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(line + sl);
    if (__uint_as_float((unsigned)(e >> 32)) >= thr) {
      *reinterpret_cast<volatile unsigned long long*>(line + 1 + k) = e;
      ++k;
    }
  }
  return k;
}

template <bool kLists>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
mnn_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = sbase + kSmemBar;
  const uint32_t bar_x_full = bar0, bar_x_empty = bar0 + 16;                       // [2] each
  const uint32_t bar_y_full = bar0 + 32, bar_y_empty = bar_y_full + 8 * kStages;   // [kStages] each
  const uint32_t bar_acc_full = bar_y_empty + 8 * kStages, bar_acc_empty = bar_acc_full + 16;  // [2] each
  const uint32_t tmem_slot = bar_acc_empty + 16;
  unsigned char* smem_gen = smem_raw + (sbase - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + kSmemBar + 32 + 16 * kStages + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();        // 0 = leader (issues the MMAs), 1 = peer
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_x_full + 8 * b, 1);
      mbar_init(bar_x_empty + 8 * b, 1);
      mbar_init(bar_acc_full + 8 * b, 1);
      mbar_init(bar_acc_empty + 8 * b, 2 * kEpiWarps);  // one arrival per epilogue warp of BOTH CTAs (leader's copy is used)
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_y_full + 8 * s, 1);
      mbar_init(bar_y_empty + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // same warp in both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();      // barriers of both CTAs initialised, TMEM allocated, before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer (both CTAs: own half of X and of every Y tile) =====
    int ys = 0, yph = 0, it = 0;
    for (int u = cluster_id; u < p.units_total; u += n_clusters, ++it) {
      const UnitInfo q = decode_unit(p, u);
      const CUtensorMap* xmap = q.dir ? &mapB : &mapA;
      const CUtensorMap* ymap = q.dir ? &mapA : &mapB;
      const int xb = it & 1;
      mbar_wait(bar_x_empty + 8 * xb, ((it >> 1) & 1) ^ 1);
      if (rank == 0) mbar_expect_tx(bar_x_full + 8 * xb, 2 * 2 * kSubBytes);
#pragma unroll
      for (int kh = 0; kh < 2; ++kh)
        tma_load_2d_pair(sbase + kSmemX + (xb * 2 + kh) * kSubBytes, xmap, kh * 64,
                         q.pair * p.d[q.dir].NXpad + q.rb * kXRows + (int)rank * 128, bar_x_full + 8 * xb);
      for (int t = q.t0; t < q.t1; ++t) {
        mbar_wait(bar_y_empty + 8 * ys, yph ^ 1);
        if (p.debug & 8) {   // bring-up: no Y traffic, MMA runs on stale shared memory
          if (rank == 0) mbar_arrive(bar_y_full + 8 * ys);
        } else {
          if (rank == 0) mbar_expect_tx(bar_y_full + 8 * ys, 2 * 2 * kSubBytes);
#pragma unroll
          for (int kh = 0; kh < 2; ++kh)
            tma_load_2d_pair(sbase + kSmemY + (ys * 2 + kh) * kSubBytes, ymap, kh * 64,
                             q.pair * p.d[q.dir].NYpad + t * kYRows + (int)rank * 128, bar_y_full + 8 * ys);
        }
        if (++ys == kStages) { ys = 0; yph ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ===== MMA issuer (one thread of the leader CTA drives both tensor cores) =====
    int ys = 0, yph = 0, as = 0, aph = 0, it = 0;
    for (int u = cluster_id; u < p.units_total; u += n_clusters, ++it) {
      const UnitInfo q = decode_unit(p, u);
      const int xb = it & 1;
      mbar_wait(bar_x_full + 8 * xb, (it >> 1) & 1);
      for (int t = q.t0; t < q.t1; ++t) {
        if (!(p.debug & 16)) mbar_wait(bar_acc_empty + 8 * as, aph ^ 1);   // bit 16 (bring-up): ignore the epilogue
        mbar_wait(bar_y_full + 8 * ys, yph);
        tc_fence_after();
        if (!(p.debug & 4)) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            const uint64_t ad = make_sw128_desc(sbase + kSmemX + (xb * 2 + kh) * kSubBytes);
            const uint64_t bd = make_sw128_desc(sbase + kSmemY + (ys * 2 + kh) * kSubBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16_pair(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), kIdesc, (kh | k) ? 1u : 0u);
          }
        }
        tc_commit_pair(bar_y_empty + 8 * ys);     // both producers may refill this stage once the MMAs retire
        tc_commit_pair(bar_acc_full + 8 * as);    // accumulators ready in both CTAs
        if (t == q.t1 - 1) tc_commit_pair(bar_x_empty + 8 * xb);
        if (++ys == kStages) { ys = 0; yph ^= 1; }
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4 && kLists) {
    // ===== epilogue, matches-only path: TMEM -> registers -> chunk maxima -> row candidate lists + group entries =====
    const int quarter = warp & 3, j = (warp - 4) >> 2;
    const int row_in_block = (int)rank * 128 + quarter * 32 + lane;
    int as = 0, aph = 0;
    for (int u = cluster_id; u < p.units_total; u += n_clusters) {
      const UnitInfo q = decode_unit(p, u);
      const DirParams& d = p.d[0];
      const int row = q.rb * kXRows + row_in_block;
      const float scale = table_scale(d, q.pair);
      const int NY = d.NY;
      const bool real_row = row < d.NX;                  // padding rows: no list (their operand rows are zero)
      const float delta = real_row ? row_delta(d, q.pair, row) : 0.f;
      uint2* line = p.L.lists + ((((size_t)q.pair * d.NXpad + row) * 4 + j) * d.splits + q.sp) * kListSlots;
      const size_t gstep = (size_t)32 * p.L.groups;      // one tile further = 32 chunks further
      unsigned* gdst = p.L.g8 + ((size_t)q.pair * d.pitch + (size_t)q.t0 * 32 + j * 8 + (lane & 7)) * p.L.groups +
                       (q.rb * 32 + (int)rank * 16 + quarter * 4 + (lane >> 3));
      float rm = -INFINITY;                              // running maximum of this thread's share of the row
      float thr_cur = (real_row && !(p.debug & 256)) ? -INFINITY : INFINITY;   // rm - delta; padding rows never emit
      int cnt = 0;
      int chunk0 = q.t0 * 32 + j * 8;
      for (int t = q.t0; t < q.t1; ++t, gdst += gstep, chunk0 += 32) {
        mbar_wait(bar_acc_full + 8 * as, aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(as * 256 + j * 64) + ((uint32_t)(quarter * 32) << 16);
        const int c0 = t * kYRows + j * 64;
        const bool full = c0 + 64 <= NY;
        uint32_t v[32];
        float m[8];
        tc_ld32(taddr, v);
        tc_wait_ld();
        if (full) chunkmax32<false, 0>(v, c0, NY, m); else chunkmax32<true, 0>(v, c0, NY, m);
        tc_ld32(taddr + 32, v);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(bar_acc_empty + 8 * as, 0);
        if (++as == 2) { as = 0; aph ^= 1; }
        if (full) chunkmax32<false, 4>(v, c0 + 32, NY, m); else chunkmax32<true, 4>(v, c0 + 32, NY, m);
        // row side: one entry per tile quarter that comes within delta of the running maximum, with the mask of its
        // chunks that do (the branch is taken by most warps on most tiles -- 32 independent rows -- so it is kept short)
        const float tm = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
        if (tm >= thr_cur) {
          rm = fmaxf(rm, tm);
          thr_cur = rm - delta;
          unsigned mask = 0;
#pragma unroll
          for (int c = 0; c < 8; ++c) mask |= m[c] >= thr_cur ? (1u << (24 + c)) : 0u;
          if (cnt >= kListSlots - 1) cnt = list_compact(line, cnt, thr_cur);
          if (cnt < kListSlots - 1) line[1 + cnt] = make_uint2((unsigned)chunk0 | mask, __float_as_uint(tm));
          ++cnt;
        }
        // column side: group entries of the warp's four 8-row groups for the tile's 8 chunks of this quarter
        if (!(p.debug & 128)) {
          const unsigned ge = g8_reduce(m, scale, lane);
          if (!(p.debug & 64)) *gdst = ge;
        }
      }
      if (real_row) line[0] = make_uint2((unsigned)cnt, __float_as_uint(rm));   // cnt >= kListSlots: the list overflowed
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): TMEM -> registers -> per-chunk maxima -> fp16 table =====
    // warp -> (TMEM lane quarter, 64-column quarter of every 256-column Y tile)
    const int quarter = warp & 3, j = (warp - 4) >> 2;
    const int row_in_block = (int)rank * 128 + quarter * 32 + lane;
    int as = 0, aph = 0;
    for (int u = cluster_id; u < p.units_total; u += n_clusters) {
      const UnitInfo q = decode_unit(p, u);
      const DirParams& d = p.d[q.dir];
      const int row = q.rb * kXRows + row_in_block;
      // everything the tile loop needs, read once per unit (indexed kernel-parameter loads are slow)
      const float scale = table_scale(d, q.pair);
      const int NY = d.NY;
      const size_t nxpad = (size_t)d.NXpad;
      __half* trow = d.table + ((size_t)q.pair * nxpad + row) * d.pitch + j * 8 + (size_t)q.t0 * 32;
      const bool skipR = p.debug & 64;
      for (int t = q.t0; t < q.t1; ++t, trow += 32) {
        mbar_wait(bar_acc_full + 8 * as, aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(as * 256 + j * 64) + ((uint32_t)(quarter * 32) << 16);
        const int c0 = t * kYRows + j * 64;
        // 32 columns at a time (keeps the live register set small: no spills in this loop)
        const bool full = c0 + 64 <= NY;
        uint32_t v[32];
        uint2 lo, hi;
        tc_ld32(taddr, v);
        tc_wait_ld();
        lo = full ? reduce32<false>(v, c0, NY, scale) : reduce32<true>(v, c0, NY, scale);
        tc_ld32(taddr + 32, v);
        tc_wait_ld();
        // this thread's accumulator slice has been read: release the TMEM stage (leader's barrier)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(bar_acc_empty + 8 * as, 0);
        if (++as == 2) { as = 0; aph ^= 1; }
        hi = full ? reduce32<false>(v, c0 + 32, NY, scale) : reduce32<true>(v, c0 + 32, NY, scale);
        if (!skipR) *reinterpret_cast<uint4*>(trow) = make_uint4(lo.x, lo.y, hi.x, hi.y);
      }
    }
  }

  // no CTA may exit (or free TMEM) while its peer can still signal it
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------ prep
// float32 rows -> bf16 [pairs][rows_pad][128] (zero padded) for both matrices of
// every pair in one launch; per row |x| and |x~ - x|; per matrix the maxima.
struct PrepArgs {
  const float* A; int64_t lda, strideA; int N, Np;
  const float* B; int64_t ldb, strideB; int M, Mp;
  __nv_bfloat16 *Ab, *Bb;
  float *anorm, *aerr, *bnorm, *berr;
  MatStats* stats;   // [pairs][2]
  int pairs;
};

constexpr int kPrepRowsPerWarp = 8;
__global__ void __launch_bounds__(256)
tc_prep_kernel(const PrepArgs a) {
  // a block = 64 consecutive rows of one matrix of one pair (row counts are multiples of 256)
  __shared__ float s_max[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows_pair = a.Np + a.Mp;
  const long long g0 = (long long)blockIdx.x * 64;
  const int pair = (int)(g0 / rows_pair);
  int row0 = (int)(g0 - (long long)pair * rows_pair);
  const bool second = row0 >= a.Np;
  if (second) row0 -= a.Np;
  const float* X = second ? a.B + pair * a.strideB : a.A + pair * a.strideA;
  const int NX = second ? a.M : a.N, NXp = second ? a.Mp : a.Np;
  const int64_t ldx = second ? a.ldb : a.lda;
  __nv_bfloat16* Xb = second ? a.Bb : a.Ab;
  float* xnorm = second ? a.bnorm : a.anorm;
  float* xerr = second ? a.berr : a.aerr;
  float m_n = 0.f, m_b = 0.f, m_e = 0.f;
  float4 v[kPrepRowsPerWarp];
#pragma unroll
  for (int k = 0; k < kPrepRowsPerWarp; ++k) {
    const int row = row0 + warp * kPrepRowsPerWarp + k;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < NX) {
      const float* src = X + (int64_t)row * ldx + lane * 4;
      v[k].x = __ldg(src); v[k].y = __ldg(src + 1); v[k].z = __ldg(src + 2); v[k].w = __ldg(src + 3);
    }
  }
#pragma unroll
  for (int k = 0; k < kPrepRowsPerWarp; ++k) {
    const int row = row0 + warp * kPrepRowsPerWarp + k;
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v[k].x, v[k].y), hi = __floats2bfloat162_rn(v[k].z, v[k].w);
    const float2 rl = __bfloat1622float2(lo), rh = __bfloat1622float2(hi);
    float ss = v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
    float sb = rl.x * rl.x + rl.y * rl.y + rh.x * rh.x + rh.y * rh.y;
    const float ex = rl.x - v[k].x, ey = rl.y - v[k].y, ez = rh.x - v[k].z, ew = rh.y - v[k].w;
    float se = ex * ex + ey * ey + ez * ez + ew * ew;
    ss = warp_sum(ss); sb = warp_sum(sb); se = warp_sum(se);
    const float up = 1.000001f;   // never under-estimate a norm (float32 rounding of the sums)
    const float nrm = sqrtf(ss) * up, nrb = sqrtf(sb) * up, nre = sqrtf(se) * up;
    uint2 pk;
    pk.x = *reinterpret_cast<const unsigned*>(&lo);
    pk.y = *reinterpret_cast<const unsigned*>(&hi);
    const size_t r = (size_t)pair * NXp + row;
    *reinterpret_cast<uint2*>(Xb + r * kD + lane * 4) = pk;
    if (lane == 0) { xnorm[r] = nrm; xerr[r] = nre; }
    m_n = fmaxf(m_n, nrm); m_b = fmaxf(m_b, nrb); m_e = fmaxf(m_e, nre);   // zero rows contribute 0
  }
  if (lane == 0) { s_max[0][warp] = m_n; s_max[1][warp] = m_b; s_max[2][warp] = m_e; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) m = fmaxf(m, s_max[threadIdx.x][k]);
    unsigned* dst = &a.stats[2 * pair + (second ? 1 : 0)].max_norm + threadIdx.x;   // max_norm, max_norm_bf, max_err
    if (__float_as_uint(m) > *(volatile unsigned*)dst) atomicMax(dst, __float_as_uint(m));
  }
}

__device__ __forceinline__ int float_to_ordered_int(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_int_to_float(int o) {
  return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff);
}

// ------------------------------------------------------------------ rescoring
// One warp per row, both directions in one launch.  The row of the chunk-maximum
// table gives the approximate row maximum F and the candidate chunks (>= F -
// delta).  Candidates are evaluated in float32 (error <= kEps32 |x||y|); every
// column within twice that band of the running float32 maximum is evaluated
// exactly (float32 operands, float64 accumulation); the argmax of the exact
// values wins, first index on ties.
struct RescoreArgs {
  DirParams d;
  const float* X; int64_t ldx, strideX;   // pair stride in elements
  const float* Y; int64_t ldy, strideY;
  int32_t* nn;                            // [pairs][NX]
  float* best;                            // optional [pairs][NX]: exact best similarity, rounded down
  float* top2;                            // kTop2 only: [pairs][NX][2] best and second best similarity (float32)
  // matches-only path (may be NULL): per-chunk minimum verification threshold and the mutual flags
  int* tmin;                              // [pairs][nchunks] ordered ints, pre-set to 0x7f7f7f7f
  float* best8;                           // [pairs][NX][8] float32 similarities to the 8 columns of the winner's chunk
  unsigned char* mutual;                  // [pairs][NX], set to 1 here
  int nchunks;
};

// kTop2 (ratio-test matchers): the second largest similarity of the row is needed as well.  With T2 the
// second largest table entry, every chunk whose entry reaches T2 - delta is rescored: the chunk of the true
// maximum and the chunk holding the best value outside it both satisfy that (two chunks have entries >= T2,
// one of them is not the maximum's chunk, and entries are within delta/2 of the true chunk maxima), and
// the runner-up is either in the maximum's chunk or is that other chunk's maximum.
// kTwoDir = false: only a0's rows exist (matches-only path); the argument block is then addressed statically
// instead of through indexed constant-bank loads.
// rows differ in cost (number of candidate chunks): small CTAs keep SM slots from idling behind a slow row
constexpr int kResWarps = 2;
template <bool kTop2, bool kTwoDir>
__global__ void __launch_bounds__(kResWarps * 32, 32 / kResWarps)   // 64 registers, 32 warps per SM: they hide the dependent table / Y-row loads
tc_rescore_kernel(const RescoreArgs a0, const RescoreArgs a1, const int pairs) {
  __shared__ __align__(16) float xs[kResWarps][kD];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_pair = a0.d.NX + (kTwoDir ? a1.d.NX : 0);
  const int pair = blockIdx.y;                     // grid: (row blocks of a pair, pairs)
  int row = blockIdx.x * kResWarps + w;
  if (row >= rows_pair) return;
  const bool second = kTwoDir && row >= a0.d.NX;
  if (second) row -= a0.d.NX;
  const RescoreArgs& a = second ? a1 : a0;
  const DirParams& d = a.d;
  const float* __restrict__ Y = a.Y + pair * a.strideY;
  const int64_t ldy = a.ldy;
  // ---- pass 1 over the table row: F = max.  The row is walked in blocks of 128
  // uint4 (1024 chunks); each lane holds 4 uint4 of a block, so rows of up to 1024
  // chunks (M <= 8192) stay in registers for the candidate pass.
  const uint4* trow = reinterpret_cast<const uint4*>(d.table + ((size_t)pair * d.NXpad + row) * d.pitch);
  const int nvec = d.pitch >> 3;
  const int nblk = (nvec + 127) >> 7;
  const unsigned kNegInf2 = 0xFC00FC00u;   // half2(-inf, -inf)
  uint4 tv[4];
  auto load_block = [&](int blk) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = blk * 128 + r * 32 + lane;
      tv[r] = i < nvec ? __ldg(trow + i) : make_uint4(kNegInf2, kNegInf2, kNegInf2, kNegInf2);
    }
  };
  __half2 hm = *reinterpret_cast<const __half2*>(&kNegInf2);
  __half2 hm2 = hm;                        // kTop2: element-wise second largest of the two half streams
  load_block(0);
  {
    // the x row is fetched while the table loads are in flight
    const float* xr = a.X + pair * a.strideX + (int64_t)row * a.ldx;
    float4 xv;
    if ((((uintptr_t)xr) & 15) == 0) {
      xv = __ldg(reinterpret_cast<const float4*>(xr) + lane);
    } else {
      xv.x = __ldg(xr + lane * 4); xv.y = __ldg(xr + lane * 4 + 1); xv.z = __ldg(xr + lane * 4 + 2); xv.w = __ldg(xr + lane * 4 + 3);
    }
    *reinterpret_cast<float4*>(&xs[w][lane * 4]) = xv;
  }
  for (int blk = 0; blk < nblk; ++blk) {
    if (blk > 0) load_block(blk);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (kTop2) {
        const unsigned wds[4] = {tv[r].x, tv[r].y, tv[r].z, tv[r].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __half2 v = *reinterpret_cast<const __half2*>(&wds[k]);
          hm2 = __hmax2(hm2, __hmin2(hm, v));
          hm = __hmax2(hm, v);
        }
      } else {
        hm = __hmax2(hm, __hmax2(__hmax2(*reinterpret_cast<const __half2*>(&tv[r].x), *reinterpret_cast<const __half2*>(&tv[r].y)),
                                 __hmax2(*reinterpret_cast<const __half2*>(&tv[r].z), *reinterpret_cast<const __half2*>(&tv[r].w))));
      }
    }
  }
  float F, T2 = 0.f;
  if (kTop2) {
    // top two of the lane's four values, then a warp merge of (first, second) pairs
    const float l1 = __low2float(hm), h1 = __high2float(hm), l2 = __low2float(hm2), h2 = __high2float(hm2);
    float m1 = fmaxf(l1, h1), m2 = fmaxf(fminf(l1, h1), fmaxf(l2, h2));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float o1 = __shfl_xor_sync(0xffffffffu, m1, o), o2 = __shfl_xor_sync(0xffffffffu, m2, o);
      m2 = fmaxf(fminf(m1, o1), fmaxf(m2, o2));
      m1 = fmaxf(m1, o1);
    }
    F = m1; T2 = m2;
  } else {
    F = warp_max(fmaxf(__low2float(hm), __high2float(hm)));
  }
  const float xn = d.xnorm[(size_t)pair * d.NXpad + row], ymax = __uint_as_float(d.ystats[2 * pair].max_norm);
  const float thr = (kTop2 ? T2 : F) - (row_delta(d, pair, row) * table_scale(d, pair) + kHalfSlack);
  const __half2 thr2 = __float2half2_rn(__half2float(__float2half_rd(thr)));   // rounded down: never drops a candidate
  const float band = 2.f * kEps32 * xn * ymax;
  const bool vec_ok = (ldy % 4 == 0) && (((uintptr_t)Y & 15) == 0);
  __syncwarp();

  float m32 = -INFINITY;       // running float32 maximum over everything seen
  float r2 = -INFINITY;        // kTop2: running float32 runner-up (m32 is the running first)
  float best_s32 = 0.f;        // this lane's column of the float32 similarities of the winner's chunk
  double bestv = -INFINITY;    // exact best
  int besti = 0x7fffffff;
  const float4* x4 = reinterpret_cast<const float4*>(&xs[w][0]);

  auto exact_col = [&](int col) {   // whole warp: exact <x, y_col>
    const float* yr = Y + (int64_t)col * ldy + lane * 4;
    const float4 xv = x4[lane];
    double acc = (double)xv.x * (double)__ldg(yr);
    acc = fma((double)xv.y, (double)__ldg(yr + 1), acc);
    acc = fma((double)xv.z, (double)__ldg(yr + 2), acc);
    acc = fma((double)xv.w, (double)__ldg(yr + 3), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (acc > bestv || (acc == bestv && col < besti)) { bestv = acc; besti = col; }
  };

  // float32 similarities of x to the 8 columns of a chunk.  Every row of Y is read by the
  // whole warp (32 x 16 B = one coalesced 512-byte row per instruction), each lane keeps the
  // partial sums over its 4 components, and a transposing butterfly (9 shuffles) leaves the
  // total of column ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1) in every lane.
  const float4 xme = x4[lane];
  const int myc = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  auto rescore = [&](int col0) {
    float pr[kChunk];
    if (vec_ok) {
      float4 yv[kChunk];
#pragma unroll
      for (int r = 0; r < kChunk; ++r) {
        const int col = min(col0 + r, d.NY - 1);          // clamped: masked below
        yv[r] = __ldg(reinterpret_cast<const float4*>(Y + (int64_t)col * ldy) + lane);
      }
#pragma unroll
      for (int r = 0; r < kChunk; ++r)
        pr[r] = fmaf(xme.w, yv[r].w, fmaf(xme.z, yv[r].z, fmaf(xme.y, yv[r].y, xme.x * yv[r].x)));
    } else {
#pragma unroll
      for (int r = 0; r < kChunk; ++r) {
        const float* yr = Y + (int64_t)min(col0 + r, d.NY - 1) * ldy + lane * 4;
        pr[r] = fmaf(xme.w, __ldg(yr + 3), fmaf(xme.z, __ldg(yr + 2), fmaf(xme.y, __ldg(yr + 1), xme.x * __ldg(yr))));
      }
    }
    float q4[4], q2[2], s32;
    {
      const bool hi = lane & 16;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float got = __shfl_xor_sync(0xffffffffu, hi ? pr[r] : pr[r + 4], 16);
        q4[r] = (hi ? pr[r + 4] : pr[r]) + got;
      }
    }
    {
      const bool hi = lane & 8;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float got = __shfl_xor_sync(0xffffffffu, hi ? q4[r] : q4[r + 2], 8);
        q2[r] = (hi ? q4[r + 2] : q4[r]) + got;
      }
    }
    {
      const bool hi = lane & 4;
      const float got = __shfl_xor_sync(0xffffffffu, hi ? q2[0] : q2[1], 4);
      s32 = (hi ? q2[1] : q2[0]) + got;
    }
    s32 += __shfl_xor_sync(0xffffffffu, s32, 2);
    s32 += __shfl_xor_sync(0xffffffffu, s32, 1);
    if (col0 + myc >= d.NY) s32 = -INFINITY;
    float cm = s32;
    if (kTop2) {
      float c2 = -INFINITY;
#pragma unroll
      for (int o = 16; o >= 4; o >>= 1) {
        const float o1 = __shfl_xor_sync(0xffffffffu, cm, o), o2 = __shfl_xor_sync(0xffffffffu, c2, o);
        c2 = fmaxf(fminf(cm, o1), fmaxf(c2, o2));
        cm = fmaxf(cm, o1);
      }
      r2 = fmaxf(fminf(m32, cm), fmaxf(r2, c2));
    } else {
#pragma unroll
      for (int o = 16; o >= 4; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    }
    m32 = fmaxf(m32, cm);
    unsigned need = __ballot_sync(0xffffffffu, (lane & 3) == 0 && s32 >= m32 - band);
    while (need) {
      const int l = __ffs(need) - 1;
      need &= need - 1;
      exact_col(col0 + ((l >> 4) & 1) * 4 + ((l >> 3) & 1) * 2 + ((l >> 2) & 1));
    }
    if ((besti >> 3) == (col0 >> 3)) best_s32 = s32;     // the winner moved into this chunk (a chunk is visited once)
  };

  // ---- pass 2: candidate chunks = table entries >= thr
  for (int blk = 0; blk < nblk; ++blk) {
    if (nblk > 1) load_block(blk);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      // candidates are rare: test the lane's 8 entries as a whole first, expand bit positions only on a hit
      const unsigned m0 = __hge2_mask(*reinterpret_cast<const __half2*>(&tv[r].x), thr2);   // 0xffff per half
      const unsigned m1 = __hge2_mask(*reinterpret_cast<const __half2*>(&tv[r].y), thr2);
      const unsigned m2 = __hge2_mask(*reinterpret_cast<const __half2*>(&tv[r].z), thr2);
      const unsigned m3 = __hge2_mask(*reinterpret_cast<const __half2*>(&tv[r].w), thr2);
      unsigned mask = 0;
      if (m0 | m1 | m2 | m3)
        mask = ((m0 & 1u) | ((m0 >> 16) & 1u) << 1) | ((m1 & 1u) | ((m1 >> 16) & 1u) << 1) << 2 |
               ((m2 & 1u) | ((m2 >> 16) & 1u) << 1) << 4 | ((m3 & 1u) | ((m3 >> 16) & 1u) << 1) << 6;
      unsigned any = __ballot_sync(0xffffffffu, mask != 0);
      while (any) {
        const int l = __ffs(any) - 1;
        any &= any - 1;
        unsigned mk = __shfl_sync(0xffffffffu, mask, l);
        while (mk) {
          const int k = __ffs(mk) - 1;
          mk &= mk - 1;
          rescore(((blk * 128 + r * 32 + l) * 8 + k) * kChunk);
        }
      }
    }
  }
  if (a.best8 && (lane & 3) == 0) a.best8[((size_t)pair * d.NX + row) * kChunk + myc] = best_s32;
  if (lane == 0) {
    const int bj = besti == 0x7fffffff ? 0 : besti;
    a.nn[(size_t)pair * d.NX + row] = bj;
    if (a.best) a.best[(size_t)pair * d.NX + row] = __double2float_rd(bestv);
    if (kTop2) {
      float* t2 = a.top2 + 2 * ((size_t)pair * d.NX + row);
      t2[0] = (float)bestv;                 // the reference's sim is float32
      t2[1] = r2;
    }
    if (a.tmin) {
      // threshold a competitor's table entry must reach to possibly beat this row at column bj
      const MatStats& sx = d.xstats[2 * pair];
      const MatStats& sy = d.ystats[2 * pair];
      const float eps_max = __uint_as_float(sx.max_err) * __uint_as_float(sy.max_norm_bf) +
                            __uint_as_float(sx.max_norm) * __uint_as_float(sy.max_err) +
                            kAccSlack * __uint_as_float(sx.max_norm) * __uint_as_float(sy.max_norm);
      const float thr_v = (__double2float_rd(bestv) - eps_max) * table_scale(d, pair) - 0.5f * kHalfSlack - 1e-6f;
      atomicMin(a.tmin + (size_t)pair * a.nchunks + (bj >> 3), float_to_ordered_int(thr_v));
      a.mutual[(size_t)pair * d.NX + row] = 1;
    }
  }
}

// ------------------------------------------------------------------ matches-only path
// When the caller does not need nn21 the second direction is not computed at all.
// i and j = nn12[i] are mutual iff no row i' has S[i'][j] > S[i][j] (or == with i' < i).
// The chunk-maximum table bounds S[i'][j] from above for every i' (entry of row i' in chunk
// j/8), so only rows whose entry comes within the error bound of S[i][j] can beat i:
//   1. the rescoring kernel leaves, per chunk c, tmin[c] = the smallest such threshold over
//      the rows whose nearest neighbour lies in c (atomicMin);
//   2. tc_scan_kernel streams the table once more (coalesced) and appends every row with
//      entry >= tmin[c] to the competitor list of chunk c -- in a well-matched pair these
//      are the ~8 rows matched to the chunk's columns plus the odd noise row;
//   3. tc_verify_kernel computes, per chunk, the exact similarities of the competitors to
//      the 8 columns and settles every member (first index on exact ties).
constexpr int kVerMaxChunks = 8192;   // M <= 65536 on this path
constexpr int kCompCap = 256;         // competitor slots per chunk (overflow -> exhaustive fallback)

struct ScanArgs {
  DirParams d;
  const int* tmin;        // [pairs][nchunks] ordered-int thresholds (scaled units)
  int* comp_cnt;          // [pairs][nchunks]
  int* comp;              // [pairs][nchunks][kCompCap]
  int nchunks;
  int rows_per_block;
};

// grid (row blocks, pairs); every thread compares 8 table entries (one uint4) with the 8
// thresholds of its chunks
__global__ void __launch_bounds__(256)
tc_scan_kernel(const ScanArgs a) {
  extern __shared__ __align__(16) __half s_thr[];     // [pitch] thresholds, rounded down
  const DirParams& d = a.d;
  const int pair = blockIdx.y;
  const int pitch = d.pitch, nvec = pitch >> 3;
  for (int c = threadIdx.x; c < pitch; c += blockDim.x) {
    float t = INFINITY;                                  // chunks without members / padding: never
    if (c < a.nchunks) {
      const int o = a.tmin[(size_t)pair * a.nchunks + c];
      if (o != 0x7f7f7f7f) t = ordered_int_to_float(o);
    }
    s_thr[c] = __float2half_rd(t);
  }
  __syncthreads();
  // a warp walks rows, its lanes the vectors of a row (no index division); four rows in flight per lane.
  // Hits are rare (about one table entry in a thousand), so a vector is first tested as a whole.
  const int rows_per_block = a.rows_per_block;
  const int row0 = blockIdx.x * rows_per_block;
  const int row1 = min(d.NX, row0 + rows_per_block);
  const uint4* thr4 = reinterpret_cast<const uint4*>(s_thr);
  const uint4* tbase = reinterpret_cast<const uint4*>(d.table + (size_t)pair * d.NXpad * pitch);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  constexpr int kU = 4;
  const uint4 kNever = make_uint4(0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u);
  for (int v = lane; v < nvec; v += 32) {
    const uint4 th = thr4[v];
    const __half2 t0 = *reinterpret_cast<const __half2*>(&th.x), t1 = *reinterpret_cast<const __half2*>(&th.y);
    const __half2 t2 = *reinterpret_cast<const __half2*>(&th.z), t3 = *reinterpret_cast<const __half2*>(&th.w);
    for (int rb = row0 + warp * kU; rb < row1; rb += nwarps * kU) {
      uint4 u[kU];
#pragma unroll
      for (int q = 0; q < kU; ++q) u[q] = rb + q < row1 ? __ldg(tbase + (size_t)(rb + q) * nvec + v) : kNever;
#pragma unroll
      for (int q = 0; q < kU; ++q) {
        const unsigned m0 = __hge2_mask(*reinterpret_cast<const __half2*>(&u[q].x), t0);
        const unsigned m1 = __hge2_mask(*reinterpret_cast<const __half2*>(&u[q].y), t1);
        const unsigned m2 = __hge2_mask(*reinterpret_cast<const __half2*>(&u[q].z), t2);
        const unsigned m3 = __hge2_mask(*reinterpret_cast<const __half2*>(&u[q].w), t3);
        if ((m0 | m1 | m2 | m3) == 0) continue;
        unsigned mask = ((m0 & 1u) | ((m0 >> 16) & 1u) << 1) | ((m1 & 1u) | ((m1 >> 16) & 1u) << 1) << 2 |
                        ((m2 & 1u) | ((m2 >> 16) & 1u) << 1) << 4 | ((m3 & 1u) | ((m3 >> 16) & 1u) << 1) << 6;
        while (mask) {
          const int h = __ffs(mask) - 1;
          mask &= mask - 1;
          const int c = v * 8 + h;
          const size_t slot = (size_t)pair * a.nchunks + c;
          const int pos = atomicAdd(a.comp_cnt + slot, 1);
          if (pos < kCompCap) a.comp[slot * kCompCap + pos] = rb + q;
        }
      }
    }
  }
}

struct VerifyArgs {
  DirParams d;
  const float* X; int64_t ldx, strideX;
  const float* Y; int64_t ldy, strideY;
  const int32_t* nn12;       // [pairs][NX]
  const int* comp_cnt;       // [pairs][nchunks]
  const int* comp;           // [pairs][nchunks][kCompCap]
  const float* best8;        // [pairs][NX][8] from the rescoring kernel (valid for the row's own chunk)
  unsigned char* mutual;     // [pairs][NX]
  int nchunks;
  // group-entry form (kG8): competitors are derived here from the chunk's row of group entries
  const int* tmin;           // [pairs][nchunks]
  const unsigned* g8;        // [pairs][pitch][groups]
  int groups;
  unsigned long long* dbg;   // POSFEAT_MNN_DEBUG counters (NULL otherwise)
  int mode;                  // bring-up timing (POSFEAT_TC_DEBUG bits 0x10000 / 0x20000 / 0x40000 / 0x80000), 0 otherwise
};

// float32 similarities of one row x (float4 per lane) to the 8 columns of a chunk staged in
// shared memory (float4 per lane and column), by the whole warp.  A transposing butterfly
// (9 shuffles) leaves column ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1) in every lane;
// the 8 results are written to e_out[0..7].
__device__ __forceinline__ void warp_chunk_dots_f32(const float4 xv, const float4* __restrict__ ysm, int lane,
                                                    float* __restrict__ e_out) {
  float p[kChunk];
#pragma unroll
  for (int r = 0; r < kChunk; ++r) {
    const float4 y = ysm[r * 32 + lane];
    p[r] = fmaf(xv.w, y.w, fmaf(xv.z, y.z, fmaf(xv.y, y.y, xv.x * y.x)));
  }
  float q4[4], q2[2], q1;
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int r = 0; r < 4; ++r) q4[r] = (hi ? p[r + 4] : p[r]) + __shfl_xor_sync(0xffffffffu, hi ? p[r] : p[r + 4], 16);
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int r = 0; r < 2; ++r) q2[r] = (hi ? q4[r + 2] : q4[r]) + __shfl_xor_sync(0xffffffffu, hi ? q4[r] : q4[r + 2], 8);
  }
  {
    const bool hi = lane & 4;
    q1 = (hi ? q2[1] : q2[0]) + __shfl_xor_sync(0xffffffffu, hi ? q2[0] : q2[1], 4);
  }
  q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
  q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
  if ((lane & 3) == 0) e_out[((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = q1;
}

// exact <x_row, y_col> (float32 operands, float64 accumulation) by the whole warp; the same
// association tree for every call, so identical inputs give identical values
__device__ __forceinline__ double warp_exact_dot(const float* __restrict__ xr, const float* __restrict__ yr, int lane) {
  double acc = (double)__ldg(xr + lane * 4) * (double)__ldg(yr + lane * 4);
  acc = fma((double)__ldg(xr + lane * 4 + 1), (double)__ldg(yr + lane * 4 + 1), acc);
  acc = fma((double)__ldg(xr + lane * 4 + 2), (double)__ldg(yr + lane * 4 + 2), acc);
  acc = fma((double)__ldg(xr + lane * 4 + 3), (double)__ldg(yr + lane * 4 + 3), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

constexpr int kVerWarps = 1;            // (4-warp CTAs: the CTA launch cost drops 43 -> 33 us per 64 pairs, but a CTA then lives as
                                        // long as its longest chunk: 152 -> 159 us on matched pairs, 502 -> 582 us on unrelated ones)
// One WARP per (chunk c of Y columns, pair).  R = competitor rows of the chunk (a superset of
// the members, the rows whose nearest neighbour lies in c).  The competitors' similarities to
// the chunk's 8 columns are evaluated in float32; a member loses its match if another
// competitor is larger in the member's column.  Comparisons inside the float32 error band are
// settled exactly (float64), ties by the lower row index.
// kG8: the competitor list is built here, in shared memory, from the chunk's row of group entries (one coalesced
// read of groups * 4 bytes, 4 KB at N = 8192) instead of by a scan kernel over a table: a group whose max1 reaches
// the chunk's threshold contributes its leader row, or all of its 8 rows when max2 reaches the threshold too.
// Measured and dropped: evaluating the non-member rows four at a time (eight lanes per row, 7-shuffle reduction, a
// third of the instructions) or requesting four rows ahead -- both need more than 64 registers or spill, and the
// kernel is bound by the warps in flight: 24 / 20 CTAs per SM cost 10 % / 25 % on well-matched pairs
// (tools/verify_debug_sweep.py: 150 -> 166 -> 189 us per 64 pairs) and gain at most 5 % on unrelated ones.
template <bool kG8>
__global__ void __launch_bounds__(kVerWarps * 32, 32 / kVerWarps)
tc_verify_kernel(const VerifyArgs a, const int getenv_dbg) {
  __shared__ __align__(16) float4 s_y[kVerWarps][kChunk * 32];   // the 8 columns of the chunk
  __shared__ float s_e[kVerWarps][32][kChunk];                    // competitor matrix of short lists
  __shared__ float s_tmp[kVerWarps][kChunk];
  __shared__ int s_rows[kG8 ? kVerWarps * kCompCap : 1];
  __shared__ unsigned short s_ent[kG8 ? kVerWarps * kCompCap : 1];   // group entry (fp16 bits) behind each competitor row
  __shared__ int s_mi[kG8 ? kVerWarps * 32 : 1], s_mj[kG8 ? kVerWarps * 32 : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.y, c = blockIdx.x * kVerWarps + warp;
  const DirParams& d = a.d;
  if (c >= a.nchunks) return;
  const size_t slot = (size_t)pair * a.nchunks + c;
  int total = 0;
  const int* rows = nullptr;
  const int32_t* nn = a.nn12 + (size_t)pair * d.NX;
  int* myrows = &s_rows[kG8 ? warp * kCompCap : 0];
  unsigned short* myent = &s_ent[kG8 ? warp * kCompCap : 0];
  const uint4* g = kG8 ? reinterpret_cast<const uint4*>(a.g8 + ((size_t)pair * d.pitch + c) * a.groups) : nullptr;
  const int nvec = kG8 ? a.groups >> 2 : 0;                       // groups is a multiple of 32
  // Competitor list of the chunk from its row of group entries: a group whose max1 reaches thr contributes its leader
  // row, or all 8 rows when max2 reaches it too.  Every lane appends the rows of its own hot groups (a warp scan of the
  // per-lane counts gives the slots).  mode 0: all rows; 1: only members (rows whose nearest neighbour lies in this
  // chunk); 2: only non-members.  e0 holds the first block of entries when it was requested ahead.  Returns the
  // number of rows (which may exceed the capacity: the caller then narrows the request).
  auto build = [&](const __half2 thr2, const int mode, uint4 (&e)[4], bool have_first) -> int {
    int n = 0;
    for (int v0 = 0; v0 < nvec; v0 += 128) {
      if (v0 > 0 || !have_first) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int v = v0 + q * 32 + lane;
          e[q] = v < nvec ? __ldg(g + v) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned m0 = __hge2_mask(as_h2(e[q].x), thr2), m1 = __hge2_mask(as_h2(e[q].y), thr2);
        const unsigned m2 = __hge2_mask(as_h2(e[q].z), thr2), m3 = __hge2_mask(as_h2(e[q].w), thr2);
        // low half = max1 (hot), high half = max2 (two or more rows reach the threshold)
        unsigned info = 0;
        if ((m0 | m1 | m2 | m3) & 0xffffu)
          info = (m0 & 1u) | (m1 & 1u) << 1 | (m2 & 1u) << 2 | (m3 & 1u) << 3 |
                 ((m0 >> 16) & 1u) << 4 | ((m1 >> 16) & 1u) << 5 | ((m2 >> 16) & 1u) << 6 | ((m3 >> 16) & 1u) << 7 |
                 (e[q].x & 7u) << 8 | (e[q].y & 7u) << 11 | (e[q].z & 7u) << 14 | (e[q].w & 7u) << 17;
        if (!__any_sync(0xffffffffu, info != 0)) continue;
        const int g0 = (v0 + q * 32 + lane) * 4;
        // bit 8k+j of `take`: row j of this lane's hot group k goes onto the list
        unsigned take = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (!((info >> k) & 1u)) continue;
          const int r0 = (g0 + k) * 8;
          unsigned bits = ((info >> (4 + k)) & 1u) ? 0xffu : (1u << ((info >> (8 + 3 * k)) & 7u));
          if (r0 + 8 > d.NX) bits &= r0 < d.NX ? (1u << (d.NX - r0)) - 1u : 0u;
          if (mode != 0) {
            for (unsigned bb = bits; bb; bb &= bb - 1) {
              const int j = __ffs(bb) - 1;
              const bool mem = (__ldg(nn + r0 + j) >> 3) == c;
              if (mem != (mode == 1)) bits &= ~(1u << j);
            }
          }
          take |= bits << (8 * k);
        }
        const int cnt = __popc(take);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        int pos = n + inc - cnt;
        n += __shfl_sync(0xffffffffu, inc, 31);
        const unsigned ew[4] = {e[q].x, e[q].y, e[q].z, e[q].w};
        for (unsigned tk = take; tk; tk &= tk - 1, ++pos) {
          const int bit = __ffs(tk) - 1;
          if (pos < kCompCap) {
            myrows[pos] = (g0 + (bit >> 3)) * 8 + (bit & 7);
            myent[pos] = (unsigned short)(ew[bit >> 3] & 0xffffu);          // max1: bounds every row of the group
          }
        }
      }
    }
    __syncwarp();
    return n;
  };
  __half2 thr_chunk = __float2half2_rn(0.f);
  if (kG8) {
    // the first block of group entries is requested together with the threshold (almost every chunk has members):
    // one memory round trip less on this kernel's dependent chain
    uint4 e[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int v = q * 32 + lane;
      e[q] = v < nvec ? __ldg(g + v) : make_uint4(0u, 0u, 0u, 0u);
    }
    const int o = a.tmin[slot];
    if (o == 0x7f7f7f7f) return;                                  // no row has its nearest neighbour in this chunk
    thr_chunk = __float2half2_rn(__half2float(__float2half_rd(ordered_int_to_float(o))));
    if (a.mode & 0x80000) return;               // bring-up timing: threshold + first block of group entries only
    total = build(thr_chunk, 0, e, true);
    rows = myrows;
    if (a.dbg && lane == 0) {
      atomicAdd(a.dbg + 4, 1ull);
      atomicAdd(a.dbg + 5, (unsigned long long)total);
      atomicAdd(a.dbg + 6, total > kCompCap ? 1ull : 0ull);
      atomicAdd(a.dbg + 7, total > 32 ? 1ull : 0ull);
    }
    if (total == 0) return;
  } else {
    total = a.comp_cnt[slot];
    if (total == 0) return;
    rows = a.comp + slot * kCompCap;
  }
  const float* Xp = a.X + pair * a.strideX;
  const float* Yp = a.Y + pair * a.strideY;
  const float4* yv = &s_y[warp][0];
  const bool vec_x = (a.ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(Xp) & 15) == 0;   // 128-bit row loads
  const bool vec_y = (a.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(Yp) & 15) == 0;
  auto stage_y = [&]() {                 // the chunk's 8 columns -> shared memory (only needed to evaluate rows here)
#pragma unroll
    for (int r = 0; r < kChunk; ++r) {
      const int jj = c * kChunk + r;
      const float* yr = Yp + (int64_t)jj * a.ldy + lane * 4;
      float4 yv4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (jj < d.NY) yv4 = vec_y ? __ldg(reinterpret_cast<const float4*>(yr)) : make_float4(__ldg(yr), __ldg(yr + 1), __ldg(yr + 2), __ldg(yr + 3));
      s_y[warp][r * 32 + lane] = yv4;
    }
    __syncwarp();
  };
  auto load_row = [&](int row) {
    const float* xr = Xp + (int64_t)row * a.ldx + lane * 4;
    if (vec_x) return __ldg(reinterpret_cast<const float4*>(xr));
    return make_float4(__ldg(xr), __ldg(xr + 1), __ldg(xr + 2), __ldg(xr + 3));
  };
  // two float32 similarities whose difference is below this band are compared exactly
  const float band = 2.f * kEps32 * __uint_as_float(d.xstats[2 * pair].max_norm) * __uint_as_float(d.ystats[2 * pair].max_norm);
  // member (in lane `member lane`) against competitor row ir with float32 value `other`
  auto settle = [&](bool member, int iq, int jq, float mine, int ir, float other, bool& lost) {
    const bool clear_win = member && ir != iq && other > mine + band;
    unsigned amb = __ballot_sync(0xffffffffu, member && ir != iq && !clear_win && other >= mine - band);
    if (clear_win) lost = true;
    while (amb) {                                   // rare: exact comparison, one member at a time
      const int l = __ffs(amb) - 1;
      amb &= amb - 1;
      const int il = __shfl_sync(0xffffffffu, iq, l), jl = __shfl_sync(0xffffffffu, jq, l);
      const float* yr = Yp + (int64_t)jl * a.ldy;
      const double em = warp_exact_dot(Xp + (int64_t)il * a.ldx, yr, lane);
      const double eo = warp_exact_dot(Xp + (int64_t)ir * a.ldx, yr, lane);
      if (lane == l && (eo > em || (eo == em && ir < il))) lost = true;
    }
  };

  // a list over capacity (descriptors of neighbouring keypoints are correlated enough on real maps to put a few hundred
  // rows above a weak member's threshold now and then) is narrowed instead of dropping the chunk into the exhaustive
  // fallback (30 ms for one warp): first only the members are listed, and after they have been settled among
  // themselves only the non-members that reach a SURVIVING member's threshold
  bool narrowed = false;
  if (kG8 && total > kCompCap) {
    uint4 e[4];
    total = build(thr_chunk, 1, e, false);
    narrowed = true;
  }
  if (kG8 && (a.mode & 0x10000)) return;       // bring-up timing: threshold + group-entry scan + list only
  if (kG8 && total <= kCompCap) {
    // Member-centric form.  Members (rows whose nearest neighbour lies in this chunk) are first settled among
    // themselves from the 8 values the rescoring kernel left for each of them -- in a well-matched pair a weak
    // member (a row without a true partner, which drags the chunk's threshold down and pulls dozens of noise rows
    // into the competitor list) loses right there to the column's real partner.  Only the surviving members'
    // thresholds decide which non-member rows still have to be evaluated: a row whose group entry (an upper
    // bound of its similarity to every column of the chunk) lies below all of them cannot beat anyone.
    const unsigned short* ents = &s_ent[warp * kCompCap];
    int* mi = &s_mi[warp * 32];
    int* mj = &s_mj[warp * 32];
    int nmem = 0;
    for (int q0 = 0; q0 < total; q0 += 32) {
      const int q = q0 + lane;
      const int iq = q < total ? rows[q] : -1;
      const int jq = iq >= 0 ? __ldg(nn + iq) : -1;
      const bool mem = iq >= 0 && (jq >> 3) == c;
      const unsigned bal = __ballot_sync(0xffffffffu, mem);
      if (mem) {
        const int pos = nmem + __popc(bal & ((1u << lane) - 1u));
        if (pos < 32) { mi[pos] = iq; mj[pos] = jq; }
      }
      nmem += __popc(bal);
    }
    __syncwarp();
    if (nmem == 0 || (a.mode & 0x40000)) return;       // (bring-up timing: members collected, nothing settled)
    if (nmem <= 32) {
      const bool member = lane < nmem;
      const int iq = member ? mi[lane] : -1, jq = member ? mj[lane] : -1;
      if (member) {
        const float4* b8 = reinterpret_cast<const float4*>(a.best8 + ((size_t)pair * d.NX + iq) * kChunk);
        *reinterpret_cast<float4*>(&s_e[warp][lane][0]) = __ldg(b8);
        *reinterpret_cast<float4*>(&s_e[warp][lane][4]) = __ldg(b8 + 1);
      }
      __syncwarp();
      const int cc = member ? jq - c * kChunk : 0;
      const float mine = member ? s_e[warp][lane][cc] : 0.f;
      bool lost = false;
      for (int r = 0; r < nmem; ++r) settle(member, iq, jq, mine, mi[r], s_e[warp][r][cc], lost);
      // which non-members can still matter
      const MatStats sx = d.xstats[2 * pair], sy = d.ystats[2 * pair];
      const float xmax = __uint_as_float(sx.max_norm), ymax = __uint_as_float(sy.max_norm);
      const float eps_max = __uint_as_float(sx.max_err) * __uint_as_float(sy.max_norm_bf) + xmax * __uint_as_float(sy.max_err) +
                            kAccSlack * xmax * ymax;
      const float mm = xmax * ymax;
      const float scale = mm > 0.f ? 1.f / mm : 1.f;
      // exact s(i, j) >= mine - band / 2; same threshold construction as in the rescoring kernel
      float tl = INFINITY;
      if (member && !lost) {
        tl = (mine - 0.5f * band - eps_max) * scale - 1e-6f;
        if (tl > 0.f) tl *= kG8Slack;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tl = fminf(tl, __shfl_xor_sync(0xffffffffu, tl, o));
      bool give_up = false;
      if (a.mode & 0x20000) tl = INFINITY;       // bring-up timing: no non-member is evaluated
      if (tl < INFINITY) {
        if (narrowed) {            // the list held the members only: now the non-members that can still matter
          uint4 e[4];
          total = build(__float2half2_rn(__half2float(__float2half_rd(tl))), 2, e, false);
          give_up = total > kCompCap;
        }
        bool staged = false;
        for (int q0 = 0; !give_up && q0 < total; q0 += 32) {
          const int q = q0 + lane;
          const int ir_l = q < total ? rows[q] : -1;
          bool need = false;
          if (ir_l >= 0) {
            const float e = __half2float(__ushort_as_half(ents[q]));
            need = e >= tl && (narrowed || (__ldg(nn + ir_l) >> 3) != c);
          }
          unsigned todo = __ballot_sync(0xffffffffu, need);
          if (todo && !staged) { stage_y(); staged = true; }
          if (a.dbg && lane == 0) atomicAdd(a.dbg + 10, (unsigned long long)__popc(todo));
          while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const int ir = __shfl_sync(0xffffffffu, ir_l, l);
            __syncwarp();
            warp_chunk_dots_f32(load_row(ir), yv, lane, &s_tmp[warp][0]);
            __syncwarp();
            settle(member, iq, jq, mine, ir, s_tmp[warp][cc], lost);
          }
        }
      }
      if (give_up) total = kCompCap + 1;      // still too many: the exhaustive path below decides everything anew
      else {
      if (member && lost) a.mutual[(size_t)pair * d.NX + iq] = 0;
      return;
      }
    } else if (narrowed) {
      total = kCompCap + 1;                    // more than 32 members and a list over capacity: exhaustive path
    }
    // more than 32 members (many rows share a nearest neighbour): the general paths below
  }
  if (total <= 32) {
    // common case: the whole competitor matrix fits in shared memory; one pass over the rows
    const int iq = lane < total ? rows[lane] : -1;
    const int jq = iq >= 0 ? __ldg(nn + iq) : -1;
    const bool member = iq >= 0 && (jq >> 3) == c;
    if (!__any_sync(0xffffffffu, member)) return;
    // rows matched into this chunk were already evaluated against its 8 columns by the rescoring kernel
    if (member) {
      const float4* b8 = reinterpret_cast<const float4*>(a.best8 + ((size_t)pair * d.NX + iq) * kChunk);
      *reinterpret_cast<float4*>(&s_e[warp][lane][0]) = __ldg(b8);
      *reinterpret_cast<float4*>(&s_e[warp][lane][4]) = __ldg(b8 + 1);
    }
    // the few other rows whose table entry reached the threshold are evaluated here
    unsigned others = __ballot_sync(0xffffffffu, iq >= 0 && !member);
    if (others) stage_y();
    while (others) {
      float4 xr[4];
      int lk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        lk[k] = others ? __ffs(others) - 1 : -1;
        if (others) others &= others - 1;
        if (lk[k] >= 0) xr[k] = load_row(__shfl_sync(0xffffffffu, iq, lk[k]));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (lk[k] >= 0) warp_chunk_dots_f32(xr[k], yv, lane, &s_e[warp][lk[k]][0]);
    }
    __syncwarp();
    const int cc = member ? jq - c * kChunk : 0;
    const float mine = s_e[warp][lane & 31][cc];
    bool lost = false;
    for (int r = 0; r < total; ++r) {
      const int ir = __shfl_sync(0xffffffffu, iq, r);
      settle(member, iq, jq, mine, ir, s_e[warp][r][cc], lost);
    }
    if (member && lost) a.mutual[(size_t)pair * d.NX + iq] = 0;
    return;
  }
  stage_y();
  if (total <= kCompCap) {
    // long list: members in batches of 32 (one per lane), competitors streamed
    for (int q0 = 0; q0 < total; q0 += 32) {
      const int q = q0 + lane;
      const int iq = q < total ? rows[q] : -1;
      const int jq = iq >= 0 ? __ldg(nn + iq) : -1;
      const bool member = iq >= 0 && (jq >> 3) == c;
      const unsigned mem = __ballot_sync(0xffffffffu, member);
      if (!mem) continue;
      const int cc = member ? jq - c * kChunk : 0;
      float mine = 0.f;
      for (unsigned mm = mem; mm;) {                      // the members' own values
        const int l = __ffs(mm) - 1;
        mm &= mm - 1;
        __syncwarp();
        warp_chunk_dots_f32(load_row(__shfl_sync(0xffffffffu, iq, l)), yv, lane, &s_tmp[warp][0]);
        __syncwarp();
        if (lane == l) mine = s_tmp[warp][cc];
      }
      bool lost = false;
      for (int r0 = 0; r0 < total; r0 += 4) {             // every competitor, four row loads in flight
        float4 xr[4];
        int ir[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ir[k] = rows[min(r0 + k, total - 1)];
          xr[k] = load_row(ir[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (r0 + k >= total) break;
          __syncwarp();
          warp_chunk_dots_f32(xr[k], yv, lane, &s_tmp[warp][0]);
          __syncwarp();
          settle(member, iq, jq, mine, ir[k], s_tmp[warp][cc], lost);
        }
      }
      if (member && lost) a.mutual[(size_t)pair * d.NX + iq] = 0;
    }
    return;
  }

  if (lane == 0 && getenv_dbg) printf("verify fallback pair %d chunk %d total %d\n", pair, c, total);
  // ---- fallback (competitor list overflowed, e.g. many duplicated descriptors): exhaustive.
  // Members are found by scanning nn12; every row of X is checked against them.
  for (int m0 = 0; m0 < d.NX; m0 += 32) {
    const int i = m0 + lane;
    const int ji = i < d.NX ? __ldg(nn + i) : -1;
    const bool member = ji >= 0 && (ji >> 3) == c;
    const unsigned mem = __ballot_sync(0xffffffffu, member);
    if (!mem) continue;
    const int cc = member ? ji - c * kChunk : 0;
    float mine = 0.f;
    for (unsigned mm = mem; mm;) {
      const int l = __ffs(mm) - 1;
      mm &= mm - 1;
      __syncwarp();
      warp_chunk_dots_f32(load_row(m0 + l), yv, lane, &s_tmp[warp][0]);
      __syncwarp();
      if (lane == l) mine = s_tmp[warp][cc];
    }
    bool lost = false;
    for (int ip = 0; ip < d.NX; ++ip) {
      __syncwarp();
      warp_chunk_dots_f32(load_row(ip), yv, lane, &s_tmp[warp][0]);
      __syncwarp();
      settle(member, i, ji, mine, ip, s_tmp[warp][cc], lost);
    }
    if (member && lost) a.mutual[(size_t)pair * d.NX + i] = 0;
  }
}

// ------------------------------------------------------------------ matches-only path, list / group-entry form
// Rescoring from the candidate lists the kLists epilogue left (one warp per row).  The row's 4 * splits lists
// (128 bytes each) hold every chunk that came within delta of a running maximum; entries >= F - delta, F the
// largest value listed, are the candidates -- the same set the table formulation finds, from 512 bytes instead of
// a 2 KB table row and without the two passes over it.  The rest (float32 re-evaluation of the candidate chunks,
// exact float64 values inside the float32 error band, first index on ties, per-chunk verification threshold)
// is the algorithm of tc_rescore_kernel.  A row whose list overflowed is rescored exhaustively.
struct RescoreListArgs {
  DirParams d;
  const uint2* lists;                     // [pairs][NXpad][4 * splits][kListSlots]
  const float* X; int64_t ldx, strideX;
  const float* Y; int64_t ldy, strideY;
  int32_t* nn;                            // [pairs][NX]
  float* best8;                           // [pairs][NX][8]
  int* tmin;                              // [pairs][nchunks] ordered ints, pre-set to 0x7f7f7f7f
  unsigned char* mutual;                  // [pairs][NX], set to 1 here
  int nchunks;
  unsigned long long* dbg;                // POSFEAT_MNN_DEBUG counters (NULL otherwise)
};

constexpr int kRlWarps = 2;
// One row by one warp: float32 pass over the candidate chunks, exact pass when two columns come within the float32
// error band, exhaustive walk when a list overflowed.  (A variant working on four rows per warp in lockstep, eight
// lanes per row, measured SLOWER -- 613 vs 392 us per 64 pairs.  ncu on this kernel: ~420-560 warp instructions per
// row of which ~100 are the dot products, 57 % issue utilisation at the 32 warps per SM its registers allow, L2 at
// 30 %: it is bound by its own instruction count and by latency, which is what the chunk-ordered form below attacks.)
__device__ __forceinline__ void rescore_row_lists(const RescoreListArgs& a, const int pair, const int row,
                                                  float (&xs)[kRlWarps][kD], const int w, const int lane) {
  const DirParams& d = a.d;
  const float* __restrict__ Y = a.Y + pair * a.strideY;
  const int64_t ldy = a.ldy;
  const size_t prow = (size_t)pair * d.NXpad + row;
  // the row's lists: 4 * splits lines of 8 uint4; lane l reads vector (l & 7) of line (l >> 3) + 4 * it
  const uint4* lines = reinterpret_cast<const uint4*>(a.lists + prow * 4 * d.splits * kListSlots);
  const int nlines = 4 * d.splits;
  uint4 lv = __ldg(lines + lane);                  // nlines >= 4: the first four lines always exist
  {
    const float* xr = a.X + pair * a.strideX + (int64_t)row * a.ldx;
    float4 xv;
    if ((((uintptr_t)xr) & 15) == 0) {
      xv = __ldg(reinterpret_cast<const float4*>(xr) + lane);
    } else {
      xv.x = __ldg(xr + lane * 4); xv.y = __ldg(xr + lane * 4 + 1); xv.z = __ldg(xr + lane * 4 + 2); xv.w = __ldg(xr + lane * 4 + 3);
    }
    *reinterpret_cast<float4*>(&xs[w][lane * 4]) = xv;
  }
  const MatStats ys = d.ystats[2 * pair], xst = d.xstats[2 * pair];
  const float ymax = __uint_as_float(ys.max_norm);
  const float xn = d.xnorm[prow];
  const float delta = 2.f * (d.xerr[prow] * __uint_as_float(ys.max_norm_bf) + xn * __uint_as_float(ys.max_err) +
                             kAccSlack * xn * ymax);
  const float band = 2.f * kEps32 * xn * ymax;
  const bool vec_ok = (ldy % 4 == 0) && (((uintptr_t)Y & 15) == 0) && ldy < (1ll << 31);
  const int ld4 = (int)(ldy >> 2);                                   // row pitch in float4 units (vec_ok)
  const float4* y4 = reinterpret_cast<const float4*>(Y) + lane;
  const size_t rowg = (size_t)pair * d.NX + row;                      // index of the row's outputs
  __syncwarp();

  float m32 = -INFINITY;       // running float32 maximum over everything seen
  float best_s32 = 0.f;        // this lane's column of the float32 similarities of the winner's chunk
  double bestv = -INFINITY;    // exact best (exact pass only)
  int besti = 0x7fffffff;
  bool amb = false;            // two columns within the float32 error band of the maximum: the exact pass decides
  bool exact = false;
  const float4* x4 = reinterpret_cast<const float4*>(&xs[w][0]);
  auto exact_col = [&](int col) {   // whole warp: exact <x, y_col>
    const float* yr = Y + (int64_t)col * ldy + lane * 4;
    const float4 xv = x4[lane];
    double acc = (double)xv.x * (double)__ldg(yr);
    acc = fma((double)xv.y, (double)__ldg(yr + 1), acc);
    acc = fma((double)xv.z, (double)__ldg(yr + 2), acc);
    acc = fma((double)xv.w, (double)__ldg(yr + 3), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (acc > bestv || (acc == bestv && col < besti)) { bestv = acc; besti = col; }
  };
  const float4 xme = x4[lane];
  const int myc = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  const int NY = d.NY;
  auto rescore = [&](int col0) {
    float pr[kChunk];
    if (vec_ok) {
      float4 yv[kChunk];
      // one 64-bit product for the chunk, then 32-bit row steps (the address arithmetic of eight independent
      // 64-bit row offsets was a fifth of this kernel's instructions)
      const float4* yp = y4 + (size_t)(unsigned)col0 * (size_t)ld4;
      if (col0 + kChunk <= NY) {
#pragma unroll
        for (int r = 0; r < kChunk; ++r) yv[r] = __ldg(yp + r * ld4);
      } else {
#pragma unroll
        for (int r = 0; r < kChunk; ++r) yv[r] = __ldg(yp + min(r, NY - 1 - col0) * ld4);      // clamped: masked below
      }
#pragma unroll
      for (int r = 0; r < kChunk; ++r)
        pr[r] = fmaf(xme.w, yv[r].w, fmaf(xme.z, yv[r].z, fmaf(xme.y, yv[r].y, xme.x * yv[r].x)));
    } else {
#pragma unroll
      for (int r = 0; r < kChunk; ++r) {
        const float* yr = Y + (int64_t)min(col0 + r, NY - 1) * ldy + lane * 4;
        pr[r] = fmaf(xme.w, __ldg(yr + 3), fmaf(xme.z, __ldg(yr + 2), fmaf(xme.y, __ldg(yr + 1), xme.x * __ldg(yr))));
      }
    }
    float q4[4], q2[2], s32;
    {
      const bool hi = lane & 16;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float got = __shfl_xor_sync(0xffffffffu, hi ? pr[r] : pr[r + 4], 16);
        q4[r] = (hi ? pr[r + 4] : pr[r]) + got;
      }
    }
    {
      const bool hi = lane & 8;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float got = __shfl_xor_sync(0xffffffffu, hi ? q4[r] : q4[r + 2], 8);
        q2[r] = (hi ? q4[r + 2] : q4[r]) + got;
      }
    }
    {
      const bool hi = lane & 4;
      const float got = __shfl_xor_sync(0xffffffffu, hi ? q2[0] : q2[1], 4);
      s32 = (hi ? q2[1] : q2[0]) + got;
    }
    s32 += __shfl_xor_sync(0xffffffffu, s32, 2);
    s32 += __shfl_xor_sync(0xffffffffu, s32, 1);
    if (col0 + myc >= NY) s32 = -INFINITY;
    float cm = s32;
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    if (exact) {
      // exact pass (rare): every column within the float32 error band of the running maximum is evaluated in
      // float64; the argmax of the exact values wins, first index on ties
      m32 = fmaxf(m32, cm);
      unsigned need = __ballot_sync(0xffffffffu, (lane & 3) == 0 && s32 >= m32 - band);
      while (need) {
        const int l = __ffs(need) - 1;
        need &= need - 1;
        exact_col(col0 + (l >> 2));
      }
      if ((besti >> 3) == (col0 >> 3)) best_s32 = s32;   // the winner moved into this chunk (a chunk is visited once)
    } else {
      // float32 pass: the winner is the float32 maximum unless a second column comes within the error band
      const unsigned near = __ballot_sync(0xffffffffu, (lane & 3) == 0 && s32 >= cm - band);
      amb |= (near & (near - 1)) != 0 || (cm <= m32 ? cm >= m32 - band : m32 >= cm - band);
      if (cm > m32) {
        m32 = cm;
        besti = col0 + ((__ffs(__ballot_sync(0xffffffffu, (lane & 3) == 0 && s32 == cm)) - 1) >> 2);
        best_s32 = s32;
      }
    }
  };

  // candidates = listed entries >= F - delta (every chunk when a list overflowed)
  auto walk = [&](float thr, bool overflow) {
    if (overflow) {
      for (int c = 0; c < a.nchunks; ++c) rescore(c * kChunk);
      return 0;
    }
    int ncand = 0;
    for (int l0 = 0; l0 < nlines; l0 += 4) {
      const uint4 lw = __ldg(lines + l0 * 8 + lane);
      const int cnt = (int)__shfl_sync(0xffffffffu, lw.x, lane & ~7);
      const int s0 = 2 * (lane & 7);
      const bool c0 = s0 >= 1 && s0 <= cnt && __uint_as_float(lw.y) >= thr;
      const bool c1 = s0 + 1 <= cnt && __uint_as_float(lw.w) >= thr;
      unsigned any = __ballot_sync(0xffffffffu, c0 || c1);
      while (any) {
        const int l = __ffs(any) - 1;
        any &= any - 1;
        const int f0 = __shfl_sync(0xffffffffu, (int)c0, l), f1 = __shfl_sync(0xffffffffu, (int)c1, l);
        const unsigned ch0 = __shfl_sync(0xffffffffu, lw.x, l), ch1 = __shfl_sync(0xffffffffu, lw.z, l);
        // an entry names a tile quarter (8 chunks from chunk0) and the chunks of it that were within delta
        if (f0) for (unsigned mk = ch0 >> 24; mk; mk &= mk - 1) { rescore((int)((ch0 & 0xffffffu) + __ffs(mk) - 1) * kChunk); ++ncand; }
        if (f1) for (unsigned mk = ch1 >> 24; mk; mk &= mk - 1) { rescore((int)((ch1 & 0xffffffu) + __ffs(mk) - 1) * kChunk); ++ncand; }
      }
    }
    return ncand;
  };

  // ---- F = largest listed value; overflow flags
  bool overflow = false;
  float F = -INFINITY;
  int dbg_cand = 0, dbg_entries = 0;
  for (int l0 = 0; l0 < nlines; l0 += 4) {
    if (l0 > 0) lv = __ldg(lines + l0 * 8 + lane);
    const int cnt = (int)__shfl_sync(0xffffffffu, lv.x, lane & ~7);        // slot 0 of this lane's line
    overflow |= cnt >= kListSlots;
    if ((lane & 7) == 0) dbg_entries += cnt;
    const int s0 = 2 * (lane & 7);                                        // slots held by this lane: s0, s0 + 1
    if (s0 >= 1 && s0 <= cnt) F = fmaxf(F, __uint_as_float(lv.y));
    if (s0 + 1 <= cnt) F = fmaxf(F, __uint_as_float(lv.w));
  }
  overflow = __any_sync(0xffffffffu, overflow);
  F = warp_max(F);
  const float thr = F - delta;
  dbg_cand = walk(thr, overflow);
  if (amb) {            // warp uniform
    exact = true;
    m32 = -INFINITY; besti = 0x7fffffff; best_s32 = 0.f;
    walk(thr, overflow);
  }
  if ((lane & 3) == 0) a.best8[rowg * kChunk + myc] = best_s32;
  if (lane == 0) {
    const int bj = besti == 0x7fffffff ? 0 : besti;
    a.nn[rowg] = bj;
    // threshold a competitor's group entry must reach to possibly beat this row at column bj
    const float xmax = __uint_as_float(xst.max_norm);
    const float eps_max = __uint_as_float(xst.max_err) * __uint_as_float(ys.max_norm_bf) + xmax * __uint_as_float(ys.max_err) +
                          kAccSlack * xmax * ymax;
    const float mm = xmax * ymax;
    const float scale = mm > 0.f ? 1.f / mm : 1.f;
    // exact s(i, bj) >= m32 - band / 2 when only the float32 pass ran
    const float vlow = exact ? __double2float_rd(bestv) : m32 - 0.5f * band;
    float thr_v = (vlow - eps_max) * scale - 1e-6f;
    if (thr_v > 0.f) thr_v *= kG8Slack;           // what a stored group entry may lack (see g8_reduce)
    atomicMin(a.tmin + (size_t)pair * a.nchunks + (bj >> 3), float_to_ordered_int(thr_v));
    a.mutual[rowg] = 1;
  }
  if (a.dbg) {
    dbg_entries = (int)warp_sum((float)dbg_entries);
    if (lane == 0) {
      atomicAdd(a.dbg + 0, 1ull);
      atomicAdd(a.dbg + 1, overflow ? 1ull : 0ull);
      atomicAdd(a.dbg + 8, amb ? 1ull : 0ull);
      atomicAdd(a.dbg + 2, (unsigned long long)dbg_cand);
      atomicAdd(a.dbg + 3, (unsigned long long)dbg_entries);
    }
  }
}

__global__ void __launch_bounds__(kRlWarps * 32, 32 / kRlWarps)
tc_rescore_lists_kernel(const RescoreListArgs a) {
  __shared__ __align__(16) float xs[kRlWarps][kD];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kRlWarps + w;
  if (row >= a.d.NX) return;
  rescore_row_lists(a, blockIdx.y, row, xs, w, lane);
}

// Large calls move the per-row control flow from a warp onto a LANE and order the arithmetic by column chunk (below).
// In the kernel above a row costs ~420 warp instructions of which ~100 are the eight dot products: the rest is list
// walking, thresholds and outputs executed by 32 lanes for one row (ncu: 57 % issue utilisation at the 32 warps per
// SM the registers allow), and every (row, candidate chunk) streams 4 KB of Y out of L2.  With fewer than ~64k rows
// in a call there are not enough 32-row warps to fill the machine and the kernel above stays the faster one
// (4096 rows: 13 us against 46 us for a lane-per-row kernel).
constexpr int kRqWarps = 4;
constexpr int kRqCap = 6;               // candidates per row the float32 pass handles; more: general routine
constexpr int kRqGroups = 4;            // bucket entries in flight per warp (8 lanes each)

__device__ __noinline__ void rescore_row_lists_slow(const RescoreListArgs& a, const int pair, const int row,
                                                    float (&xs)[kRlWarps][kD], const int lane) {
  rescore_row_lists(a, pair, row, xs, 0, lane);
}

// ------------------------------------------------------------------ rescoring by COLUMN CHUNK (large calls)
// A chunk of 8 columns is a candidate of ~8-11 rows, so the work ordered by chunk reads every chunk of Y once (by row
// it is 3 GB per 64 pairs of 8192 out of L2):
//   1  tc_rescore_enum_kernel    lane = row: F = max of the lines' running maxima, threshold, list walk; every candidate (row, k-th
//                                candidate of the row) is appended to its chunk's bucket (one atomicAdd);
//   2  tc_rescore_chunk_kernel   warp = chunk: Y chunk staged in shared memory, bucket entries four at a time
//                                (eight lanes per row: lane t holds 16 of the 128 dimensions, 8 x 16 FMAs, a 7-shuffle
//                                transposing reduction leaves column t of the chunk on lane t), one 48-byte record
//                                per entry: the 8 float32 similarities, their maximum, its column, "two columns
//                                within the float32 band";
//   3  tc_rescore_merge_kernel   lane = row: merges the row's records in list order, writes nn12 / best8 / chunk
//                                threshold / mutual flag.
// Rows the float32 pass cannot settle, overflowed lists, rows with more than kRqCap candidates and rows that hit a
// full bucket go through rescore_row_lists (in kernel 1 or 3): one row in several hundred.
constexpr int kBucketCap = 64;           // entries per chunk bucket (the comp array of the table form is reused: kCompCap ints)
constexpr int kRecFloats = 12;           // s8[8], cm, column, multi, pad
static_assert(kBucketCap <= kCompCap, "bucket storage is the table form's competitor array");

struct RescoreBucketArgs {
  RescoreListArgs r;
  int* bcnt;                              // [pairs][nchunks], zeroed per call
  unsigned* bucket;                       // [pairs][nchunks][kCompCap] (first kBucketCap used): row | k << 28
  float* rec;                             // [pairs][NX][kRqCap][kRecFloats]
  int* rowinfo;                           // [pairs][NX]: number of candidates, or -1 = settled by the general routine
};

__global__ void __launch_bounds__(kRqWarps * 32, 8)
tc_rescore_enum_kernel(const __grid_constant__ RescoreBucketArgs b) {
  __shared__ __align__(16) float s_x[kRqWarps][kRlWarps][kD];
  const RescoreListArgs& a = b.r;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.y;
  const DirParams& d = a.d;
  const int row0 = (blockIdx.x * kRqWarps + w) * 32;
  if (row0 >= d.NX) return;
  const int row = row0 + lane;
  const bool live = row < d.NX;
  const size_t prow = (size_t)pair * d.NXpad + (live ? row : row0);
  const int nlines = 4 * d.splits;
  const uint2* lines = a.lists + prow * nlines * kListSlots;
  const MatStats ys = d.ystats[2 * pair];
  const float ymax = __uint_as_float(ys.max_norm);
  const float xn = d.xnorm[prow];
  const float delta = 2.f * (d.xerr[prow] * __uint_as_float(ys.max_norm_bf) + xn * __uint_as_float(ys.max_err) +
                             kAccSlack * xn * ymax);
  float F = -INFINITY;
  bool slow = false;
  int entries = 0;
  // four lines per row (no column split: every large call): the first sector of each line -- header and three
  // entries -- is requested up front, eight loads in flight instead of a chain of round trips
  const bool four = nlines == 4;
  uint4 f0[4], f1[4];
  if (four) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const uint4* lp = reinterpret_cast<const uint4*>(lines + l * kListSlots);
      f0[l] = __ldg(lp); f1[l] = __ldg(lp + 1);
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      slow |= (int)f0[l].x >= kListSlots;
      entries += (int)f0[l].x;
      if ((int)f0[l].x > 0) F = fmaxf(F, __uint_as_float(f0[l].y));
    }
  } else {
    for (int l = 0; l < nlines; ++l) {
      const uint2 h = __ldg(lines + l * kListSlots);
      slow |= (int)h.x >= kListSlots;                     // the list overflowed
      entries += (int)h.x;
      if ((int)h.x > 0) F = fmaxf(F, __uint_as_float(h.y));
    }
  }
  const float thr = F - delta;
  int ncand = 0;
  int cand[kRqCap];
#pragma unroll
  for (int k = 0; k < kRqCap; ++k) cand[k] = 0;
  if (live && !slow) {
    auto consider = [&](unsigned ex, unsigned ey) {
      if (__uint_as_float(ey) >= thr) {
        for (unsigned mk = ex >> 24; mk; mk &= mk - 1) {
          const int chunk = (int)((ex & 0xffffffu) + __ffs(mk) - 1);
#pragma unroll
          for (int k = 0; k < kRqCap; ++k) if (k == ncand) cand[k] = chunk;      // (registers: no dynamic indexing)
          ++ncand;
        }
      }
    };
    for (int l = 0; l < nlines; ++l) {
      const uint4* lp = reinterpret_cast<const uint4*>(lines + l * kListSlots);
      uint4 v0, v1;                                       // slots 0 (header), 1 | 2, 3: one sector
      if (four) {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (q == l) { v0 = f0[q]; v1 = f1[q]; }
      } else {
        v0 = __ldg(lp); v1 = __ldg(lp + 1);
      }
      const int cnt = (int)v0.x;
      if (cnt >= 1) consider(v0.z, v0.w);
      if (cnt >= 2) consider(v1.x, v1.y);
      if (cnt >= 3) consider(v1.z, v1.w);
      for (int base = 4; base <= cnt; base += 4) {
        v0 = __ldg(lp + (base >> 1)); v1 = __ldg(lp + (base >> 1) + 1);
        consider(v0.x, v0.y);
        if (cnt >= base + 1) consider(v0.z, v0.w);
        if (cnt >= base + 2) consider(v1.x, v1.y);
        if (cnt >= base + 3) consider(v1.z, v1.w);
      }
    }
    slow |= ncand > kRqCap;
    if (!slow) {
      // the bucket slots of all candidates are requested together (an atomic with a result is a memory round trip)
      int* bc = b.bcnt + (size_t)pair * a.nchunks;
      unsigned* bk = b.bucket + (size_t)pair * a.nchunks * kCompCap;
      int pos[kRqCap];
#pragma unroll
      for (int k = 0; k < kRqCap; ++k) pos[k] = k < ncand ? atomicAdd(bc + cand[k], 1) : 0;
#pragma unroll
      for (int k = 0; k < kRqCap; ++k) {
        if (k < ncand) {
          if (pos[k] < kBucketCap) bk[(size_t)cand[k] * kCompCap + pos[k]] = (unsigned)row | ((unsigned)k << 28);
          else slow = true;                                // bucket full: the general routine takes the row
        }
      }
    }
  }
  if (!live) { slow = false; ncand = 0; }
  if (live) b.rowinfo[(size_t)pair * d.NX + row] = slow ? -1 : ncand;
  if (a.dbg) {
    const bool fast = live && !slow;
    const int nrows = __popc(__ballot_sync(0xffffffffu, fast));
    const int nc = (int)warp_sum(fast ? (float)ncand : 0.f), ne = (int)warp_sum(fast ? (float)entries : 0.f);
    const int nslow = __popc(__ballot_sync(0xffffffffu, slow));
    if (lane == 0) {
      atomicAdd(a.dbg + 0, (unsigned long long)nrows);
      atomicAdd(a.dbg + 2, (unsigned long long)nc);
      atomicAdd(a.dbg + 3, (unsigned long long)ne);
      atomicAdd(a.dbg + 9, (unsigned long long)nslow);
    }
  }
  unsigned rest = __ballot_sync(0xffffffffu, slow);
  while (rest) {
    const int o = __ffs(rest) - 1;
    rest &= rest - 1;
    __syncwarp();
    rescore_row_lists_slow(a, pair, row0 + o, s_x[w], lane);
  }
}

constexpr int kRcWarps = 4;
__global__ void __launch_bounds__(kRcWarps * 32, 6)
tc_rescore_chunk_kernel(const __grid_constant__ RescoreBucketArgs b) {
  __shared__ __align__(16) float4 s_y[kRcWarps][kChunk * 32];
  const RescoreListArgs& a = b.r;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.y;
  const int c = blockIdx.x * kRcWarps + w;
  if (c >= a.nchunks) return;
  const DirParams& d = a.d;
  const int n_raw = b.bcnt[(size_t)pair * a.nchunks + c];
  const unsigned* bk = b.bucket + ((size_t)pair * a.nchunks + c) * kCompCap;
  const int NY = d.NY, col0 = c * kChunk;
  const unsigned ent_first = __ldg(bk + (lane >> 3));        // requested with the count and the chunk (valid memory either way)
  {
    const float4* y4 = reinterpret_cast<const float4*>(a.Y + pair * a.strideY) + lane;
    const int ld4 = (int)(a.ldy >> 2);
#pragma unroll
    for (int r = 0; r < kChunk; ++r) s_y[w][r * 32 + lane] = __ldg(y4 + (size_t)min(col0 + r, NY - 1) * ld4);   // past the end: masked below
  }
  __syncwarp();
  const int n = min(n_raw, kBucketCap);
  if (n == 0) return;
  const float ymax = __uint_as_float(d.ystats[2 * pair].max_norm);
  const int t = lane & 7, gi = lane >> 3;
  const int ldx4 = (int)(a.ldx >> 2);
  const float4* x4 = reinterpret_cast<const float4*>(a.X + pair * a.strideX) + t;
  const float4* ys4 = &s_y[w][t];
  // software pipeline: the entries and rows of the next batch are requested before the current one is reduced
  unsigned ent_n = gi < n ? ent_first : __ldg(bk + n - 1);
  float4 xn4[4];
  float xnorm_n;
  {
    const int row = (int)(ent_n & 0x0fffffffu);
    const float4* xp = x4 + (size_t)(unsigned)row * (size_t)ldx4;
#pragma unroll
    for (int i = 0; i < 4; ++i) xn4[i] = __ldg(xp + 8 * i);
    xnorm_n = __ldg(d.xnorm + (size_t)pair * d.NXpad + row);
  }
  for (int e0 = 0; e0 < n; e0 += kRqGroups) {
    const int e = e0 + gi;
    const bool act = e < n;
    const unsigned ent = ent_n;
    const int row = (int)(ent & 0x0fffffffu), k = (int)(ent >> 28);
    float4 xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = xn4[i];
    const float band = 2.f * kEps32 * xnorm_n * ymax;
    if (e0 + kRqGroups < n) {
      const int en = e0 + kRqGroups + gi;
      ent_n = __ldg(bk + (en < n ? en : n - 1));
      const int rown = (int)(ent_n & 0x0fffffffu);
      const float4* xp = x4 + (size_t)(unsigned)rown * (size_t)ldx4;
#pragma unroll
      for (int i = 0; i < 4; ++i) xn4[i] = __ldg(xp + 8 * i);
      xnorm_n = __ldg(d.xnorm + (size_t)pair * d.NXpad + rown);
    }
    float pr[kChunk];
#pragma unroll
    for (int r = 0; r < kChunk; ++r) {
      float4 yv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) yv[i] = ys4[r * 32 + 8 * i];
      float acc = xv[0].x * yv[0].x;
      acc = fmaf(xv[0].y, yv[0].y, acc); acc = fmaf(xv[0].z, yv[0].z, acc); acc = fmaf(xv[0].w, yv[0].w, acc);
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        acc = fmaf(xv[i].x, yv[i].x, acc); acc = fmaf(xv[i].y, yv[i].y, acc);
        acc = fmaf(xv[i].z, yv[i].z, acc); acc = fmaf(xv[i].w, yv[i].w, acc);
      }
      pr[r] = acc;
    }
    float q4[4], q2[2], sc;
    {
      const bool hi = t & 4;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float got = __shfl_xor_sync(0xffffffffu, hi ? pr[r] : pr[r + 4], 4);
        q4[r] = (hi ? pr[r + 4] : pr[r]) + got;
      }
    }
    {
      const bool hi = t & 2;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float got = __shfl_xor_sync(0xffffffffu, hi ? q4[r] : q4[r + 2], 2);
        q2[r] = (hi ? q4[r + 2] : q4[r]) + got;
      }
    }
    {
      const bool hi = t & 1;
      const float got = __shfl_xor_sync(0xffffffffu, hi ? q2[0] : q2[1], 1);
      sc = (hi ? q2[1] : q2[0]) + got;                  // lane t: column col0 + t
    }
    if (col0 + t >= NY) sc = -INFINITY;
    float cm = sc;
#pragma unroll
    for (int sh = 4; sh >= 1; sh >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, sh));
    const unsigned near = (__ballot_sync(0xffffffffu, sc >= cm - band) >> (lane & 24)) & 0xffu;
    const unsigned at = (__ballot_sync(0xffffffffu, sc == cm) >> (lane & 24)) & 0xffu;
    if (act) {
      float* rec = b.rec + (((size_t)pair * d.NX + row) * kRqCap + k) * kRecFloats;
      rec[t] = sc;
      if (t == 0) {
        rec[8] = cm;
        rec[9] = __int_as_float(col0 + __ffs(at) - 1);
        rec[10] = __int_as_float((near & (near - 1)) != 0 ? 1 : 0);
      }
    }
  }
}

__global__ void __launch_bounds__(kRqWarps * 32, 8)
tc_rescore_merge_kernel(const __grid_constant__ RescoreBucketArgs b) {
  __shared__ __align__(16) float s_x[kRqWarps][kRlWarps][kD];
  const RescoreListArgs& a = b.r;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.y;
  const DirParams& d = a.d;
  const int row0 = (blockIdx.x * kRqWarps + w) * 32;
  if (row0 >= d.NX) return;
  const int row = row0 + lane;
  const bool live = row < d.NX;
  const size_t rowg = (size_t)pair * d.NX + (live ? row : row0);
  const int info = live ? b.rowinfo[rowg] : -1;
  const MatStats ys = d.ystats[2 * pair], xst = d.xstats[2 * pair];
  const float ymax = __uint_as_float(ys.max_norm);
  const float band = 2.f * kEps32 * d.xnorm[(size_t)pair * d.NXpad + (live ? row : row0)] * ymax;
  float m32 = -INFINITY;
  int besti = 0x7fffffff, bestk = 0;
  bool amb = false;
  const float* rec = b.rec + rowg * kRqCap * kRecFloats;
  // the first record (most rows have exactly one) is requested together with the row's info, not after it; the
  // loads are of initialised or stale-but-mapped workspace memory and only used when info says so
  const float4 h0 = __ldg(reinterpret_cast<const float4*>(rec + 8));
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(rec)), s1 = __ldg(reinterpret_cast<const float4*>(rec) + 1);
  for (int k = 0; k < info; ++k) {
    const float4 h = k == 0 ? h0 : __ldg(reinterpret_cast<const float4*>(rec + k * kRecFloats + 8));
    const float c2 = h.x;
    amb |= __float_as_int(h.z) != 0 || (c2 <= m32 ? c2 >= m32 - band : m32 >= c2 - band);
    if (c2 > m32) { m32 = c2; besti = __float_as_int(h.y); bestk = k; }
  }
  if (info >= 0 && !amb) {
    float4* dst = reinterpret_cast<float4*>(a.best8 + rowg * kChunk);
    if (besti == 0x7fffffff) {          // no candidate at all (cannot happen with a well-formed list): defined outputs
      dst[0] = make_float4(0.f, 0.f, 0.f, 0.f); dst[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      if (bestk == 0) { dst[0] = s0; dst[1] = s1; }
      else {
        const float4* src = reinterpret_cast<const float4*>(rec + bestk * kRecFloats);
        dst[0] = __ldg(src); dst[1] = __ldg(src + 1);
      }
    }
    const int bj = besti == 0x7fffffff ? 0 : besti;
    a.nn[rowg] = bj;
    const float xmax = __uint_as_float(xst.max_norm);
    const float eps_max = __uint_as_float(xst.max_err) * __uint_as_float(ys.max_norm_bf) + xmax * __uint_as_float(ys.max_err) +
                          kAccSlack * xmax * ymax;
    const float mm = xmax * ymax;
    const float scale = mm > 0.f ? 1.f / mm : 1.f;
    const float vlow = m32 - 0.5f * band;
    float thr_v = (vlow - eps_max) * scale - 1e-6f;
    if (thr_v > 0.f) thr_v *= kG8Slack;
    atomicMin(a.tmin + (size_t)pair * a.nchunks + (bj >> 3), float_to_ordered_int(thr_v));
    a.mutual[rowg] = 1;
  }
  // two columns within the float32 error band: the exact pass of the general routine decides
  unsigned rest = __ballot_sync(0xffffffffu, info >= 0 && amb);
  if (a.dbg && lane == 0 && rest) atomicAdd(a.dbg + 9, (unsigned long long)__popc(rest));
  while (rest) {
    const int o = __ffs(rest) - 1;
    rest &= rest - 1;
    __syncwarp();
    rescore_row_lists_slow(a, pair, row0 + o, s_x[w], lane);
  }
}

// ordered compaction of the rows flagged mutual (ascending i), one CTA per pair.  A pair of more than 8192 rows takes
// several rounds of the block scan: the next round's rows are requested before the current round is scanned and
// stored (without that every round is a memory round trip of its own: 91 us for one 65536-row pair).
__global__ void __launch_bounds__(1024)
tc_compact_flags_kernel(const int32_t* __restrict__ nn12_, const unsigned char* __restrict__ mutual_, int N,
                        int64_t* __restrict__ matches_, int32_t* __restrict__ n_matches) {
  __shared__ int s_warp[32];
  __shared__ int s_total[2];
  constexpr int kIt = 8;
  const int pair = blockIdx.x;
  const int32_t* nn12 = nn12_ + (size_t)pair * N;
  const unsigned char* mutual = mutual_ + (size_t)pair * N;
  int64_t* matches = matches_ + (size_t)pair * N * 2;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // 8 consecutive rows per thread: two 16-byte loads of nn12 and one 8-byte load of the flags when the pair's arrays
  // are aligned for them
  const bool vec = (N & 7) == 0 && ((reinterpret_cast<uintptr_t>(nn12) & 15) == 0) && ((reinterpret_cast<uintptr_t>(mutual) & 7) == 0);
  auto fetch = [&](int i0, int (&j)[kIt], unsigned& keep) {
    const int first = i0 + tid * kIt;
    keep = 0;
    if (vec && first + kIt <= N) {
      const int4 a = __ldg(reinterpret_cast<const int4*>(nn12 + first)), b = __ldg(reinterpret_cast<const int4*>(nn12 + first) + 1);
      const uint2 m = __ldg(reinterpret_cast<const uint2*>(mutual + first));
      j[0] = a.x; j[1] = a.y; j[2] = a.z; j[3] = a.w; j[4] = b.x; j[5] = b.y; j[6] = b.z; j[7] = b.w;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        keep |= ((m.x >> (8 * k)) & 0xffu) ? (1u << k) : 0u;
        keep |= ((m.y >> (8 * k)) & 0xffu) ? (1u << (4 + k)) : 0u;
      }
    } else {
#pragma unroll
      for (int k = 0; k < kIt; ++k) {
        const bool in = first + k < N;
        j[k] = in ? nn12[first + k] : 0;
        keep |= (in && mutual[first + k]) ? (1u << k) : 0u;
      }
    }
  };
  int base = 0;
  int jn[kIt];
  unsigned keepn;
  fetch(0, jn, keepn);
  for (int i0 = 0, round = 0; i0 < N; i0 += 1024 * kIt, ++round) {
    const int first = i0 + tid * kIt;
    int j[kIt];
#pragma unroll
    for (int k = 0; k < kIt; ++k) j[k] = jn[k];
    const unsigned keep = keepn;
    if (i0 + 1024 * kIt < N) fetch(i0 + 1024 * kIt, jn, keepn);
    const int mine = __popc(keep);
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      const int v = s_warp[lane];
      int winc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      s_warp[lane] = winc - v;
      if (lane == 31) s_total[round & 1] = winc;
    }
    __syncthreads();
    int pos = base + s_warp[wid] + inc - mine;
    base += s_total[round & 1];              // (double buffered: the next round's total is written after the next barrier)
#pragma unroll
    for (int k = 0; k < kIt; ++k)
      if (keep & (1u << k)) {
        matches[2 * (int64_t)pos] = first + k;
        matches[2 * (int64_t)pos + 1] = j[k];
        ++pos;
      }
    __syncthreads();                         // s_warp is rewritten by the next round
  }
  if (tid == 0) n_matches[pair] = base;
}

int launch_compact_flags(const int32_t* nn12, const unsigned char* flags, int P, int N, int64_t* matches,
                         int32_t* n_matches, cudaStream_t stream) {
  tc_compact_flags_kernel<<<P, 1024, 0, stream>>>(nn12, flags, N, matches, n_matches);
  PF_LAUNCH_CHECK("tc_compact_flags_kernel");
  return POSFEAT_OK;
}

// ------------------------------------------------------------------ host side
static int make_map(CUtensorMap* map, void* base, int rows_pad) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(POSFEAT_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)kD, (cuuint64_t)rows_pad};
  cuuint64_t gstride[1] = {(cuuint64_t)kD * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(POSFEAT_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return POSFEAT_OK;
}

static inline int pad_rows(int n) { return (n + kXRows - 1) / kXRows * kXRows; }

struct TcWs {
  __nv_bfloat16 *Ab, *Bb;
  float *anorm, *aerr, *bnorm, *berr;
  MatStats* stats;  // [pairs][2]: A, B
  __half* table[2];
  int pitch[2];
  // matches-only path
  float* best;
  float* best8;
  int *tmin, *comp_cnt, *comp;
  unsigned char* mutual;
  uint2* lists;
  unsigned* g8;
  unsigned long long* dbg;
  int list_splits;
  size_t total;
};

static void choose_splits(int P, int rb0, int yt0, int rb1, int yt1, int G, int* s0, int* s1);

// column splits of the matches-only launch for this problem (the lists are sized by it)
static int one_dir_splits(int P, int N, int M) {
  const int Np = (N + kXRows - 1) / kXRows * kXRows;
  int s0, s1;
  choose_splits(P, Np / kXRows, (M + kYRows - 1) / kYRows, 0, (N + kYRows - 1) / kYRows, sm_count() / 2, &s0, &s1);
  return s0;
}

static TcWs carve_tc(void* base, int P, int N, int M) {
  TcWs w;
  const int list_splits = one_dir_splits(P, N, M);
  const size_t Np = pad_rows(N), Mp = pad_rows(M);
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 1024);
    return r;
  };
  w.Ab = (__nv_bfloat16*)take(sizeof(__nv_bfloat16) * P * Np * kD);
  w.Bb = (__nv_bfloat16*)take(sizeof(__nv_bfloat16) * P * Mp * kD);
  w.anorm = (float*)take(sizeof(float) * P * Np);
  w.aerr = (float*)take(sizeof(float) * P * Np);
  w.bnorm = (float*)take(sizeof(float) * P * Mp);
  w.berr = (float*)take(sizeof(float) * P * Mp);
  w.stats = (MatStats*)take(sizeof(MatStats) * 2 * P);
  // chunk-maximum tables: rows of X times (tiles of Y) * 32 chunks, fp16
  w.pitch[0] = ((M + kYRows - 1) / kYRows) * (kYRows / kChunk);
  w.pitch[1] = ((N + kYRows - 1) / kYRows) * (kYRows / kChunk);
  w.table[0] = (__half*)take(sizeof(__half) * P * Np * w.pitch[0]);
  w.table[1] = (__half*)take(sizeof(__half) * P * Mp * w.pitch[1]);
  // matches-only path
  w.best = (float*)take(sizeof(float) * P * (size_t)N);
  w.best8 = (float*)take(sizeof(float) * P * (size_t)N * kChunk);
  const size_t nch = (size_t)(M + kChunk - 1) / kChunk;
  w.tmin = (int*)take(sizeof(int) * P * nch * 2);            // tmin followed by comp_cnt
  w.comp_cnt = w.tmin + P * nch;
  w.comp = (int*)take(sizeof(int) * P * nch * kCompCap);
  w.mutual = (unsigned char*)take((size_t)P * N);
  // list / group-entry form of the matches-only path
  w.list_splits = list_splits;
  w.lists = (uint2*)take(sizeof(uint2) * kListSlots * (size_t)P * Np * 4 * list_splits);
  w.g8 = (unsigned*)take(sizeof(unsigned) * (size_t)P * w.pitch[0] * (Np / 8));
  w.dbg = (unsigned long long*)take(sizeof(unsigned long long) * 16);
  w.total = off;
  return w;
}

int tc_prep_sink(void* ws, size_t ws_bytes, int P, int N, int M, PrepSink* sink, cudaStream_t stream) {
  TcWs w = carve_tc(ws, P, N, M);
  if (ws_bytes < w.total) return set_error(POSFEAT_EWORKSPACE, "mnn tc workspace: need %zu bytes, got %zu", w.total, ws_bytes);
  if (((uintptr_t)ws & 255) != 0) return set_error(POSFEAT_EINVAL, "mnn workspace must be 256-byte aligned");
  const int Np = pad_rows(N), Mp = pad_rows(M);
  PF_CUDA(cudaMemsetAsync(w.stats, 0, sizeof(MatStats) * 2 * P, stream));
  if (Np != N || Mp != M) {   // padding rows must read as zero operands
    PF_CUDA(cudaMemsetAsync(w.Ab, 0, (size_t)((char*)w.stats - (char*)w.Ab), stream));
  }
  *sink = PrepSink{w.Ab, w.Bb, w.anorm, w.aerr, w.bnorm, w.berr, w.stats, Np, Mp};
  return POSFEAT_OK;
}

bool tc_supported(int N, int M, int D) { return D == kD && N >= 1 && M >= 1; }
size_t tc_workspace_bytes(int P, int N, int M) { return carve_tc(nullptr, P, N, M).total; }

// choose the number of column splits so the persistent grid runs full waves
static void choose_splits(int P, int rb0, int yt0, int rb1, int yt1, int G, int* s0, int* s1) {
  double best = -1.0;
  *s0 = *s1 = 1;
  for (int S = 1; S <= kMaxSplits; ++S) {
    auto used = [&](int yt) {
      const int tps = (yt + S - 1) / S;
      return (yt + tps - 1) / tps;
    };
    const int u0 = used(yt0), u1 = used(yt1);
    const long total = (long)P * ((long)rb0 * u0 + (long)rb1 * u1);
    const long waves = (total + G - 1) / G;
    // cost model: waves are quantised; every unit pays a small cost for its (double buffered) X block
    const double tiles = (double)P * ((double)rb0 * yt0 + (double)rb1 * yt1);
    const double per_wave = tiles / total + 0.4;
    const double score = 1.0 / (waves * per_wave);
    if (score > best * 1.02) { best = score; *s0 = u0; *s1 = u1; }
  }
}

int mnn_tc(const float* A, int64_t strideA, int N, int64_t lda, const float* Bm, int64_t strideB, int M, int64_t ldb,
           int D, int P, int32_t* nn12, int32_t* nn21, int64_t* matches, int32_t* n_matches, void* ws, size_t ws_bytes,
           cudaStream_t stream, float* top12, float* top21, bool prepared) {
  if (D != kD) return set_error(POSFEAT_EUNSUPPORTED, "tensor-core matcher needs D == 128 (got D=%d)", D);
  TcWs w = carve_tc(ws, P, N, M);
  if (ws_bytes < w.total) return set_error(POSFEAT_EWORKSPACE, "mnn tc workspace: need %zu bytes, got %zu", w.total, ws_bytes);
  if (((uintptr_t)ws & 255) != 0) return set_error(POSFEAT_EINVAL, "mnn workspace must be 256-byte aligned");
  const int Np = pad_rows(N), Mp = pad_rows(M);
  if ((long long)P * Np >= (1ll << 31) || (long long)P * Mp >= (1ll << 31))
    return set_error(POSFEAT_EINVAL, "batched matcher: pairs * rows exceeds 2^31");
  // matches-only calls (nn21 == NULL): candidate lists + group entries (default), or the chunk-maximum table with a
  // scan kernel (POSFEAT_MNN_TABLE=1, kept for A/B measurements), which is limited to kVerMaxChunks column chunks:
  // refuse BEFORE anything is queued (the two-direction path needs an nn21 buffer to write to)
  const bool want_one = nn21 == nullptr && top12 == nullptr;
  const char* env_table = getenv("POSFEAT_MNN_TABLE");
  const bool use_lists = want_one && !(env_table && atoi(env_table) != 0);
  if (want_one && !use_lists && w.pitch[0] > kVerMaxChunks)
    return set_error(POSFEAT_EUNSUPPORTED, "matches-only matcher (nn21 == NULL) in table form supports M <= %d; pass an nn21 buffer",
                     kVerMaxChunks * kChunk);

  if (!prepared) {   // otherwise the sampler already left bf16 rows, norms and maxima in the workspace (tc_prep_sink)
    PF_CUDA(cudaMemsetAsync(w.stats, 0, sizeof(MatStats) * 2 * P, stream));
    PrepArgs pa{A, lda, strideA, N, Np, Bm, ldb, strideB, M, Mp, w.Ab, w.Bb, w.anorm, w.aerr, w.bnorm, w.berr, w.stats, P};
    const long long prep_warps = (long long)P * (Np + Mp);
    prof_begin(PROF_MNN_PREP, stream);
    tc_prep_kernel<<<(unsigned)(prep_warps / 64), 256, 0, stream>>>(pa);
    prof_end(PROF_MNN_PREP, stream);
    PF_LAUNCH_CHECK("tc_prep_kernel");
  }

  CUtensorMap mapA, mapB;
  if (int e = make_map(&mapA, w.Ab, P * Np)) return e;
  if (int e = make_map(&mapB, w.Bb, P * Mp)) return e;

  TcParams p;
  const int G = sm_count() / 2;   // persistent CTA pairs
  const int rb0 = Np / kXRows, rb1 = Mp / kXRows;
  const int yt0 = (M + kYRows - 1) / kYRows, yt1 = (N + kYRows - 1) / kYRows;
  int s0, s1;
  // nn21 == NULL: matches only -> the second direction is replaced by the column verification
  const bool one_dir = want_one;
  choose_splits(P, rb0, yt0, one_dir ? 0 : rb1, yt1, G, &s0, &s1);
  if (use_lists && s0 != w.list_splits) return set_error(POSFEAT_EINVAL, "internal: list split count changed (%d vs %d)", s0, w.list_splits);
  auto fill = [&](DirParams& d, int dir, int NX, int NY, int NXpad, int NYpad, int yt, int S) {
    d.xnorm = dir ? w.bnorm : w.anorm;
    d.xerr = dir ? w.berr : w.aerr;
    d.xstats = w.stats + (dir ? 1 : 0);
    d.ystats = w.stats + (dir ? 0 : 1);
    d.table = w.table[dir];
    d.pitch = w.pitch[dir];
    d.NX = NX; d.NY = NY; d.NXpad = NXpad; d.NYpad = NYpad;
    d.y_tiles = yt;
    d.tiles_per_split = (yt + S - 1) / S;
    d.splits = (yt + d.tiles_per_split - 1) / d.tiles_per_split;
    d.row_blocks = NXpad / kXRows;
  };
  fill(p.d[0], 0, N, M, Np, Mp, yt0, s0);
  fill(p.d[1], 1, M, N, Mp, Np, yt1, s1);
  p.pairs = P;
  p.units0 = rb0 * p.d[0].splits;
  p.units_pair = p.units0 + (one_dir ? 0 : rb1 * p.d[1].splits);
  p.units_total = P * p.units_pair;
  {
    const char* dbg = getenv("POSFEAT_TC_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }

  p.L = ListParams{w.lists, w.g8, Np / 8};
  const int grid = 2 * (p.units_total < G ? p.units_total : G);
  prof_begin(PROF_MNN_TC, stream);
  if (use_lists) {
    PF_CUDA(cudaFuncSetAttribute(mnn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    mnn_tc_kernel<true><<<grid, kTcThreads, kSmemAlloc, stream>>>(mapA, mapB, p);
  } else {
    PF_CUDA(cudaFuncSetAttribute(mnn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    mnn_tc_kernel<false><<<grid, kTcThreads, kSmemAlloc, stream>>>(mapA, mapB, p);
  }
  prof_end(PROF_MNN_TC, stream);
  PF_LAUNCH_CHECK("mnn_tc_kernel");
  if (p.debug & 0x1000) return POSFEAT_OK;     // bring-up timing of the tensor kernel alone (results are not produced)

  const int nchunks = (M + kChunk - 1) / kChunk;
  if (P > 65535) return set_error(POSFEAT_EINVAL, "batched matcher: at most 65535 pairs per call");
  if (use_lists) {
    PF_CUDA(cudaMemsetAsync(w.tmin, 0x7f, sizeof(int) * (size_t)P * nchunks, stream));
    const bool dbg_on = getenv("POSFEAT_MNN_DEBUG") != nullptr;
    if (dbg_on) PF_CUDA(cudaMemsetAsync(w.dbg, 0, sizeof(unsigned long long) * 16, stream));
    RescoreListArgs ra{p.d[0], w.lists, A, lda, strideA, Bm, ldb, strideB, nn12, w.best8, w.tmin, w.mutual, nchunks,
                       dbg_on ? w.dbg : nullptr};
    prof_begin(PROF_MNN_RESCORE, stream);
    // large calls: the chunk-ordered form (16-byte vector loads: aligned operands only; the records live in the table
    // form's first table, 2 * pitch bytes per row); POSFEAT_MNN_RESCORE_WARP=1 keeps the warp-per-row kernel for A/B runs
    const bool bucket_env = getenv("POSFEAT_MNN_RESCORE_WARP") == nullptr;
    const bool bucket_form = bucket_env && (size_t)P * N >= 65536 && N < (1 << 28) &&
                             lda % 4 == 0 && ldb % 4 == 0 && strideA % 4 == 0 && strideB % 4 == 0 &&
                             (((uintptr_t)A | (uintptr_t)Bm) & 15) == 0 && lda < (1ll << 31) && ldb < (1ll << 31) &&
                             (size_t)w.pitch[0] * sizeof(__half) >= sizeof(float) * kRqCap * kRecFloats;
    if (bucket_form) {
      PF_CUDA(cudaMemsetAsync(w.comp_cnt, 0, sizeof(int) * (size_t)P * nchunks, stream));
      RescoreBucketArgs ba{ra, w.comp_cnt, (unsigned*)w.comp, (float*)w.table[0], (int*)w.best};
      const dim3 grows((unsigned)((N + kRqWarps * 32 - 1) / (kRqWarps * 32)), (unsigned)P);
      tc_rescore_enum_kernel<<<grows, kRqWarps * 32, 0, stream>>>(ba);
      PF_LAUNCH_CHECK("tc_rescore_enum_kernel");
      tc_rescore_chunk_kernel<<<dim3((unsigned)((nchunks + kRcWarps - 1) / kRcWarps), (unsigned)P), kRcWarps * 32, 0, stream>>>(ba);
      PF_LAUNCH_CHECK("tc_rescore_chunk_kernel");
      tc_rescore_merge_kernel<<<grows, kRqWarps * 32, 0, stream>>>(ba);
    } else
      tc_rescore_lists_kernel<<<dim3((unsigned)((N + kRlWarps - 1) / kRlWarps), (unsigned)P), kRlWarps * 32, 0, stream>>>(ra);
    prof_end(PROF_MNN_RESCORE, stream);
    PF_LAUNCH_CHECK(bucket_form ? "tc_rescore_merge_kernel" : "tc_rescore_lists_kernel");
    if (p.debug & 0x2000) return POSFEAT_OK;   // bring-up timing: stop after the rescoring kernel (no results)
    VerifyArgs va{p.d[0], A, lda, strideA, Bm, ldb, strideB, nn12, nullptr, nullptr, w.best8, w.mutual, nchunks, w.tmin, w.g8, Np / 8,
                  dbg_on ? w.dbg : nullptr, p.debug & 0xf0000};
    prof_begin(PROF_MNN_VERIFY, stream);
    tc_verify_kernel<true><<<dim3((nchunks + kVerWarps - 1) / kVerWarps, P), kVerWarps * 32, 0, stream>>>(va, getenv("POSFEAT_VERIFY_DEBUG") ? 1 : 0);
    prof_end(PROF_MNN_VERIFY, stream);
    PF_LAUNCH_CHECK("tc_verify_kernel<g8>");
    prof_begin(PROF_MNN_COMPACT, stream);
    tc_compact_flags_kernel<<<P, 1024, 0, stream>>>(nn12, w.mutual, N, matches, n_matches);
    prof_end(PROF_MNN_COMPACT, stream);
    PF_LAUNCH_CHECK("tc_compact_flags_kernel");
    if (dbg_on) {   // diagnostics only: synchronises
      unsigned long long h[11];
      PF_CUDA(cudaMemcpyAsync(h, w.dbg, sizeof(h), cudaMemcpyDeviceToHost, stream));
      PF_CUDA(cudaStreamSynchronize(stream));
      fprintf(stderr, "posfeat mnn lists: P=%d N=%d M=%d splits=%d | rows %llu overflow %llu exact-pass %llu general-path %llu candidates %llu list-entries %llu | "
                      "chunks %llu competitors %llu overflow-chunks %llu long-chunks(>32) %llu non-member-evaluations %llu\n", P, N, M, s0, h[0], h[1], h[8], h[9],
              h[2], h[3], h[4], h[5], h[6], h[7], h[10]);
    }
    return POSFEAT_OK;
  }
  if (one_dir) {
    PF_CUDA(cudaMemsetAsync(w.tmin, 0x7f, sizeof(int) * (size_t)P * nchunks, stream));
    PF_CUDA(cudaMemsetAsync(w.comp_cnt, 0, sizeof(int) * (size_t)P * nchunks, stream));
  }
  RescoreArgs r0{p.d[0], A, lda, strideA, Bm, ldb, strideB, nn12, one_dir ? w.best : nullptr, top12,
                 one_dir ? w.tmin : nullptr, one_dir ? w.best8 : nullptr, w.mutual, nchunks};
  RescoreArgs r1{p.d[1], Bm, ldb, strideB, A, lda, strideA, nn21, nullptr, top21, nullptr, nullptr, nullptr, 0};
  if (one_dir) r1.d.NX = 0;
  const dim3 resc_grid((unsigned)((N + (one_dir ? 0 : M) + kResWarps - 1) / kResWarps), (unsigned)P);
  prof_begin(PROF_MNN_RESCORE, stream);
  if (top12) tc_rescore_kernel<true, true><<<resc_grid, kResWarps * 32, 0, stream>>>(r0, r1, P);
  else if (one_dir) tc_rescore_kernel<false, false><<<resc_grid, kResWarps * 32, 0, stream>>>(r0, r1, P);
  else tc_rescore_kernel<false, true><<<resc_grid, kResWarps * 32, 0, stream>>>(r0, r1, P);
  prof_end(PROF_MNN_RESCORE, stream);
  PF_LAUNCH_CHECK("tc_rescore_kernel");
  if (top12) return POSFEAT_OK;      // ratio-test callers apply their own acceptance rule to (nn, top2)
  if (!one_dir) {
    return launch_mutual_compact_batched(nn12, nn21, P, N, M, matches, n_matches, stream);
  }
  // enough CTAs to fill the chip also when a single small pair is matched
  int scan_rows = 64;
  while (scan_rows > 4 && (int64_t)((N + scan_rows - 1) / scan_rows) * P < 2 * 148) scan_rows >>= 1;
  ScanArgs sa{p.d[0], w.tmin, w.comp_cnt, w.comp, nchunks, scan_rows};
  prof_begin(PROF_MNN_SCAN, stream);
  tc_scan_kernel<<<dim3((N + scan_rows - 1) / scan_rows, P), 256, sizeof(__half) * w.pitch[0], stream>>>(sa);
  prof_end(PROF_MNN_SCAN, stream);
  PF_LAUNCH_CHECK("tc_scan_kernel");
  VerifyArgs va{p.d[0], A, lda, strideA, Bm, ldb, strideB, nn12, w.comp_cnt, w.comp, w.best8, w.mutual, nchunks, nullptr, nullptr, 0, nullptr, 0};
  prof_begin(PROF_MNN_VERIFY, stream);
  tc_verify_kernel<false><<<dim3((nchunks + kVerWarps - 1) / kVerWarps, P), kVerWarps * 32, 0, stream>>>(va, getenv("POSFEAT_VERIFY_DEBUG") ? 1 : 0);
  prof_end(PROF_MNN_VERIFY, stream);
  PF_LAUNCH_CHECK("tc_verify_kernel");
  prof_begin(PROF_MNN_COMPACT, stream);
  tc_compact_flags_kernel<<<P, 1024, 0, stream>>>(nn12, w.mutual, N, matches, n_matches);
  prof_end(PROF_MNN_COMPACT, stream);
  PF_LAUNCH_CHECK("tc_compact_flags_kernel");
  return POSFEAT_OK;
}

}  // namespace posfeat
