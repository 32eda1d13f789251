// Subsystem (1): score-map keypoint selection for sm_100a.
//
// Replaces generate_kpts_single(stable=True), reference
// losses/preprocess_utils.py:215-267:
//   nms (:449-464)  -> nms_candidates_kernel (shared-memory halo tile, closed-form
//                      "first maximum in scan order" predicate, warp-ballot
//                      compaction of survivors into 64-bit keys)
//   thr (:232-240)  -> fused into the same kernel (thr_val prepared on device)
//   topk (:264)     -> select_kernel: radix select over the candidate keys only,
//                      bitonic sort of the winners in shared memory
//   centroid/score (:243-247), gather (:266-267) -> evaluated only at the winners
//   gen_grid (:84-87, :217-221) -> computed in-kernel (bit exact linspace)
// HBM-bound: the score map is read once (4 B/pixel); survivors cost 8 B each.
#include <stdlib.h>

#include "common.cuh"

namespace posfeat {

typedef unsigned long long u64;

constexpr int kTileW = 128;
constexpr int kTileH = 16;
constexpr int kNmsThreads = 256;
constexpr int kSelThreads = 1024;
constexpr int kSortSmemKeys = 16384;      // 128 KB of 64-bit keys
constexpr int kFillBitmapWords = 2048;    // filler search covers the first 65536 pixels
constexpr int kFillMax = 4096;
constexpr int kDigitBits = 11;   // radix-select digit width (2048-bin shared histogram)
constexpr int kBucketMax = 1024;  // boundary bucket small enough to be sorted on its own
constexpr int kRadixMinKeys = 2048;  // fewer candidates than this are simply sorted; more go through the histogram (rank path)
constexpr int kUnroll = 8;       // independent candidate loads in flight per thread

constexpr int kCntStride = 32;   // int32 slots between the per-image candidate counters
constexpr int kCntKeyMax = 1;    // slots of the counter line: max of the candidates' score bits and max of their
constexpr int kCntKeyNegMin = 2; // complement (0 = not recorded by the NMS kernel that ran)
struct DetectWs {
  int32_t* cand_count;   // [B * kCntStride] positive-score survivors written to cand (one 128-byte line per image:
                         // same-line atomics serialise in the L2 atomic unit)
  float* thr_val;        // [B]
  double* red_sum;       // [B]
  unsigned* red_max;     // [B] order-preserving uint of the max
  int32_t* status;       // [1] device-side error flag
  u64* cand;             // [B, cand_cap]
  u64* sortbuf;          // [B, sort_cap] (only when cap_pts > kSortSmemKeys)
  int64_t cand_cap;
  int64_t sort_cap;
  size_t total;
};

static inline int64_t next_pow2(int64_t v) {
  int64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

static DetectWs carve(void* base, int B, int H, int W, int cap_pts) {
  DetectWs w;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.cand_count = (int32_t*)take(sizeof(int32_t) * B * kCntStride);
  w.thr_val = (float*)take(sizeof(float) * B);
  w.red_sum = (double*)take(sizeof(double) * B);
  w.red_max = (unsigned*)take(sizeof(unsigned) * B);
  w.status = (int32_t*)take(sizeof(int32_t));
  w.cand_cap = (int64_t)(H - 2) * (W - 2);
  w.cand = (u64*)take(sizeof(u64) * (size_t)B * w.cand_cap);
  w.sort_cap = cap_pts > kSortSmemKeys ? next_pow2(cap_pts) : 0;
  w.sortbuf = (u64*)take(sizeof(u64) * (size_t)B * w.sort_cap);
  w.total = off;
  return w;
}

__device__ __forceinline__ unsigned float_to_ordered(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- threshold preparation ------------------------------------------------
__global__ void reduce_interior_kernel(const float* __restrict__ score, int H, int W, int64_t sb,
                                       int64_t sy, double* red_sum, unsigned* red_max) {
  const int b = blockIdx.y;
  const int hi = H - 2, wi = W - 2;
  const int64_t total = (int64_t)hi * wi;
  const float* img = score + b * sb;
  double s = 0.0;
  float m = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int y = (int)(i / wi), x = (int)(i - (int64_t)y * wi);
    float v = __ldg(img + (int64_t)(y + 1) * sy + x + 1);
    s += (double)v;
    m = fmaxf(m, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(red_sum + b, s);
    atomicMax(red_max + b, float_to_ordered(m));
  }
}

__global__ void finalize_thr_kernel(int B, int thr_mode, float thr, int64_t n_interior,
                                    const double* red_sum, const unsigned* red_max, float* thr_val) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float t;
  if (thr_mode == POSFEAT_THR_ABS) {
    t = thr * 1.0f;
  } else if (thr_mode == POSFEAT_THR_MAX) {
    t = thr * ordered_to_float(red_max[b]);
  } else if (thr_mode == POSFEAT_THR_MEAN) {
    t = thr * (float)(red_sum[b] / (double)n_interior);
  } else {
    t = -INFINITY;
  }
  thr_val[b] = t;
}

// ---- phase 1: NMS + threshold + compaction --------------------------------
// One CTA = kTileH x kTileW interior pixels.  The padded tile (halo r, reflect
// about the INTERIOR map, as F.pad(mode='reflect') on kp_map[:,:,1:-1,1:-1])
// lives in shared memory.  keep(y,x) <=> c > v for every window entry that
// precedes the centre in row-major scan order, c >= v for every later entry.
__global__ void __launch_bounds__(kNmsThreads)
nms_candidates_kernel(const float* __restrict__ score, int H, int W, int64_t sb, int64_t sy,
                      int nms_mode, int r, int has_thr, const float* __restrict__ thr_val,
                      int32_t* __restrict__ counts, int32_t* __restrict__ cand_count,
                      u64* __restrict__ cand, int64_t cand_cap) {
  extern __shared__ float tile[];
  __shared__ u64 s_list[kTileH * kTileW];
  __shared__ int s_n, s_nall, s_base;

  const int b = blockIdx.z;
  const int hi = H - 2, wi = W - 2;
  const int ty0 = blockIdx.y * kTileH, tx0 = blockIdx.x * kTileW;
  const int pw = kTileW + 2 * r, ph = kTileH + 2 * r;
  const int pitch = pw | 1;  // odd pitch: column walks hit distinct banks
  const float* img = score + b * sb;

  if (threadIdx.x == 0) { s_n = 0; s_nall = 0; }
  for (int i = threadIdx.x; i < ph * pw; i += kNmsThreads) {
    int py = i / pw, px = i - py * pw;
    int iy = ty0 + py - r, ix = tx0 + px - r;
    float v = -INFINITY;
    if (iy < hi + r && ix < wi + r) {
      iy = reflect_idx(iy, hi);
      ix = reflect_idx(ix, wi);
      v = __ldg(img + (int64_t)(iy + 1) * sy + ix + 1);
    }
    tile[py * pitch + px] = v;
  }
  __syncthreads();

  const float tv = has_thr ? thr_val[b] : 0.f;
  const int lane = threadIdx.x & 31;
  int n_all = 0;
  for (int i = threadIdx.x; i < kTileH * kTileW; i += kNmsThreads) {
    const int ly = i / kTileW, lx = i - ly * kTileW;
    const int y = ty0 + ly, x = tx0 + lx;
    bool keep = (y < hi) && (x < wi);
    float c = 0.f;
    if (keep) {
      c = tile[(ly + r) * pitch + lx + r];
      if (has_thr) keep = c > tv;
    }
    if (keep && nms_mode == POSFEAT_NMS_HARD) {
      const float* base = tile + ly * pitch + lx;  // window top-left
      for (int dy = 0; dy <= 2 * r && keep; ++dy) {
        const float* row = base + dy * pitch;
        if (dy < r) {
          for (int dx = 0; dx <= 2 * r; ++dx) keep &= c > row[dx];
        } else if (dy == r) {
          for (int dx = 0; dx < r; ++dx) keep &= c > row[dx];
          for (int dx = r + 1; dx <= 2 * r; ++dx) keep &= c >= row[dx];
        } else {
          for (int dx = 0; dx <= 2 * r; ++dx) keep &= c >= row[dx];
        }
      }
    }
    if (keep && nms_mode == POSFEAT_NMS_SOFT) {
      // soft_nms (:431-447): alpha = softplus(c - avg_pool(reflect-padded window)); the top-k key
      // is (thr_mask * alpha) * score (:240, :263).  Window summed row-major in fp32 like avg_pool2d.
      const float* base = tile + ly * pitch + lx;
      float sum = 0.f;
      for (int dy = 0; dy <= 2 * r; ++dy)
        for (int dx = 0; dx <= 2 * r; ++dx) sum = __fadd_rn(sum, base[dy * pitch + dx]);
      const float a = __fsub_rn(c, __fdiv_rn(sum, (float)((2 * r + 1) * (2 * r + 1))));
      const float alpha = a > 20.f ? a : log1pf(expf(a));
      c = __fmul_rn(alpha, c);
    }
    n_all += keep ? 1 : 0;
    const bool emit = keep && c > 0.f;
    const unsigned bal = __ballot_sync(0xffffffffu, emit);
    if (bal) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_n, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (emit) {
        const unsigned idx = (unsigned)y * (unsigned)wi + (unsigned)x;
        s_list[base + __popc(bal & ((1u << lane) - 1u))] =
            ((u64)__float_as_uint(c) << 32) | (u64)(0xffffffffu - idx);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) n_all += __shfl_xor_sync(0xffffffffu, n_all, o);
  if (lane == 0 && n_all) atomicAdd(&s_nall, n_all);
  __syncthreads();
  if (threadIdx.x == 0) {
    s_base = s_n ? atomicAdd(cand_count + b * kCntStride, s_n) : 0;
    if (s_nall) atomicAdd(counts + b, s_nall);
  }
  __syncthreads();
  const int n = s_n;
  u64* dst = cand + (int64_t)b * cand_cap + s_base;
  for (int i = threadIdx.x; i < n; i += kNmsThreads) dst[i] = s_list[i];
}

// ---- phase 1, fast path: register sliding window, one warp per vertical strip ----
// For radius R <= 3 (and for use_nms=False, R = 0) a warp owns a strip of
// 32 - 2R interior columns (R halo lanes on each side) and walks down kStripRows
// rows.  Every input row is loaded once (coalesced), horizontal window maxima
// come from warp shuffles, the vertical window lives in registers:
//   keep(y,x) <=> c >  max over rows above of hmax(row)      (they precede in scan order)
//              && c >  max of the R pixels to the left       (same row, precede)
//              && c >= max of the R pixels to the right      (same row, follow)
//              && c >= max over rows below of hmax(row)      (follow)
// Kept pixels are pairwise non-adjacent, so a strip yields at most
// ceil(rows/2)*ceil(cols/2) <= 512 survivors: they are staged in a per-warp
// shared-memory list and flushed with ONE global atomic per warp.
constexpr int kStripWarps = 8;
constexpr int kStripList = 512;

template <int R>
__global__ void __launch_bounds__(kStripWarps * 32)
nms_strip_kernel(const float* __restrict__ score, int H, int W, int64_t sb, int64_t sy, int rows_per_strip,
                 int has_thr, const float* __restrict__ thr_val, int32_t* __restrict__ counts,
                 int32_t* __restrict__ cand_count, u64* __restrict__ cand, int64_t cand_cap) {
  __shared__ u64 s_list[kStripWarps][kStripList];
  constexpr int kUse = 32 - 2 * R;
  constexpr unsigned kFull = 0xffffffffu;
  const int b = blockIdx.z;
  const int hi = H - 2, wi = W - 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x = (blockIdx.x * kStripWarps + warp) * kUse + lane - R;   // interior column of this lane
  const int y0 = blockIdx.y * rows_per_strip;
  const int y1 = min(hi, y0 + rows_per_strip);
  if ((blockIdx.x * kStripWarps + warp) * kUse >= wi) return;          // whole warp outside the map
  const bool col_loadable = x < wi + R;                                // inside the reflect-padded extent
  const int xr = reflect_idx(min(x, wi + R - 1), wi);
  const bool col_eval = lane >= R && lane < 32 - R && x < wi;
  const float* img = score + b * sb + xr + 1;
  const float tv = has_thr ? thr_val[b] : 0.f;

  // sliding state: hm[k] = horizontal window max of input row (yy - 2R + k); centre row = index R
  float hm[2 * R + 1], cv[R + 1], lv[R + 1], rv[R + 1];
#pragma unroll
  for (int k = 0; k < 2 * R + 1; ++k) hm[k] = -INFINITY;
#pragma unroll
  for (int k = 0; k < R + 1; ++k) { cv[k] = -INFINITY; lv[k] = -INFINITY; rv[k] = -INFINITY; }

  int cnt = 0, n_all = 0;
  constexpr int kBatch = 8;
  for (int yb = y0 - R; yb < y1 + R; yb += kBatch) {
    float in[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int yy = yb + k;
      float v = -INFINITY;
      if (col_loadable && yy < hi + R && yy < y1 + R) v = __ldg(img + (int64_t)(reflect_idx(yy, hi) + 1) * sy);
      in[k] = v;
    }
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int yy = yb + k;              // input row just loaded; the centre row is yy - R
      const float v = in[k];
      float L = -INFINITY, Rm = -INFINITY;
#pragma unroll
      for (int d = 1; d <= R; ++d) {
        L = fmaxf(L, __shfl_up_sync(kFull, v, d));
        Rm = fmaxf(Rm, __shfl_down_sync(kFull, v, d));
      }
      // shift the window down by one row
#pragma unroll
      for (int q = 0; q < 2 * R; ++q) hm[q] = hm[q + 1];
      hm[2 * R] = fmaxf(fmaxf(L, Rm), v);
#pragma unroll
      for (int q = 0; q < R; ++q) { cv[q] = cv[q + 1]; lv[q] = lv[q + 1]; rv[q] = rv[q + 1]; }
      cv[R] = v; lv[R] = L; rv[R] = Rm;
      const int yc = yy - R;
      const float c = cv[0];
      bool keep = col_eval && yc >= y0 && yc < y1;
      if (has_thr) keep = keep && c > tv;
      if (R > 0) {
        float above = -INFINITY, below = -INFINITY;
#pragma unroll
        for (int q = 0; q < R; ++q) { above = fmaxf(above, hm[q]); below = fmaxf(below, hm[R + 1 + q]); }
        keep = keep && c > above && c > lv[0] && c >= rv[0] && c >= below;
      }
      const unsigned ball = __ballot_sync(kFull, keep);
      n_all += __popc(ball);
      const bool emit = keep && c > 0.f;
      const unsigned bal = __ballot_sync(kFull, emit);
      if (bal) {
        if (emit) {
          const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
          const unsigned idx = (unsigned)yc * (unsigned)wi + (unsigned)x;
          if (pos < kStripList) s_list[warp][pos] = ((u64)__float_as_uint(c) << 32) | (u64)(0xffffffffu - idx);
        }
        cnt += __popc(bal);
      }
    }
  }
  cnt = min(cnt, kStripList);   // cannot trigger: survivors of a strip are pairwise non-adjacent
  __syncwarp();
  int base = 0;
  if (lane == 0) {
    if (cnt) base = atomicAdd(cand_count + b * kCntStride, cnt);
    if (n_all) atomicAdd(counts + b, n_all);
  }
  base = __shfl_sync(kFull, base, 0);
  u64* dst = cand + (int64_t)b * cand_cap + base;
  for (int i = lane; i < cnt; i += 32) dst[i] = s_list[warp][i];
}

// Radius-1 fast path (the HPatches / training configuration): FOUR pixels per lane, survivors kept
// as a bit mask.
// A warp task covers 128 map columns [S, S+128) (S = 124 * column-warp) and evaluates the 124
// columns S+1 .. S+124 over kQuadRows centre rows.  Per row and lane: one 128-bit load (kVec) or
// four clamped scalar loads, two shuffles, four 3-input maxima for the horizontal window, and per
// pixel   keep <=> c > max3(above, left, thr)  &&  c >= max(right, below)   (scan-order tie rule).
// The keep bits of an 8-row batch fill one 32-bit register per lane (bit 4k+j = row k, pixel j);
// after the batch the warp scans the bit counts, takes ONE global atomic, and every lane appends
// its survivors (score re-read through L1).  No shared memory, no per-row stores.
// The reflect-101 border of the interior grid is applied when the row / column is fetched:
// interior column -1 is map column 0 -> map column 2, interior column wi is map column W-1 ->
// map column W-3, and the same for rows.
constexpr int kQuadRows = 32;
constexpr int kQuadBatch = 8;
constexpr int kQuadCols = 124;
constexpr int kQuadWarps = 8;

template <bool kVec, bool kPosThr>
__global__ void __launch_bounds__(kQuadWarps * 32, 4)
nms_quad_r1_kernel(const float* __restrict__ score, int H, int W, int64_t sb, int64_t sy, int has_thr,
                   const float* __restrict__ thr_val, int32_t* __restrict__ counts,
                   int32_t* __restrict__ cand_count, u64* __restrict__ cand, int64_t cand_cap, int ncw,
                   int ntasks) {
  constexpr unsigned kFull = 0xffffffffu;
  // image index varies fastest over the grid: CTAs running at the same time append to different
  // candidate lists, so their atomics do not queue on one address
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int task = blockIdx.y * kQuadWarps + warp;
  if (task >= ntasks) return;
  const int cw = task % ncw, strip = task / ncw;
  const int hi = H - 2, wi = W - 2;
  const int x0 = cw * kQuadCols + 4 * lane;              // map column of this lane's pixel 0
  const int y0 = strip * kQuadRows, y1 = min(hi, y0 + kQuadRows);
  const float* img = score + b * sb;
  const float tv = has_thr ? thr_val[b] : -INFINITY;

  // column addressing, fixed for the whole strip
  const float* colp;                                     // kVec: 16-byte aligned base of the lane's quad
  int co0 = 0, co1 = 0, co2 = 0, co3 = 0;                // !kVec: per-pixel (reflected, clamped) column offsets
  bool fix_l = false, fix_r = false;
  if (kVec) {
    colp = img + (x0 + 3 < W ? x0 : 0);                  // lanes past the right edge read a harmless quad
    fix_l = x0 == 0;                                     // map column 0  := map column 2
    fix_r = x0 + 3 == W - 1;                             // map column W-1 := map column W-3 (W % 4 == 0 here)
  } else {
    auto col = [&](int x) { return x <= 0 ? 2 : (x >= W - 1 ? W - 3 : x); };
    co0 = col(x0); co1 = col(x0 + 1); co2 = col(x0 + 2); co3 = col(x0 + 3);
    colp = img;
  }
  auto load_at = [&](const float* rp) -> float4 {        // rp: lane base + map row offset
    float4 v;
    if (kVec) {
      v = __ldg(reinterpret_cast<const float4*>(rp));
      if (fix_l) v.x = v.z;
      if (fix_r) v.w = v.y;
    } else {
      v.x = __ldg(rp + co0); v.y = __ldg(rp + co1); v.z = __ldg(rp + co2); v.w = __ldg(rp + co3);
    }
    return v;
  };
  auto load_row = [&](int yi) -> float4 {                // interior row yi in [-1, hi + 7], reflect-101 / clamped
    const int yr = yi < 0 ? 1 : (yi < hi ? yi : max(2 * hi - 2 - yi, 0));
    return load_at(colp + (int64_t)(yr + 1) * sy);
  };
  // which of the lane's four pixels are evaluated: map columns S+1 .. S+124, and <= W-2
  unsigned colmask = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = x0 + j, rel = 4 * lane + j;
    if (rel >= 1 && rel <= kQuadCols && x <= W - 2) colmask |= 1u << j;
  }
  const unsigned lanevalid = colmask * 0x11111111u;

  // prologue: row above the strip and the first centre row
  float a0, a1, a2, a3, c0, c1, c2, c3, cl, cr, h0, h1, h2, h3;
  {
    const float4 p = load_row(y0 - 1), q = load_row(y0);
    float l = __shfl_up_sync(kFull, p.w, 1), r = __shfl_down_sync(kFull, p.x, 1);
    a0 = fmaxf(fmaxf(l, p.x), p.y); a1 = fmaxf(fmaxf(p.x, p.y), p.z);
    a2 = fmaxf(fmaxf(p.y, p.z), p.w); a3 = fmaxf(fmaxf(p.z, p.w), r);
    cl = __shfl_up_sync(kFull, q.w, 1); cr = __shfl_down_sync(kFull, q.x, 1);
    c0 = q.x; c1 = q.y; c2 = q.z; c3 = q.w;
    h0 = fmaxf(fmaxf(cl, c0), c1); h1 = fmaxf(fmaxf(c0, c1), c2);
    h2 = fmaxf(fmaxf(c1, c2), c3); h3 = fmaxf(fmaxf(c2, c3), cr);
  }
  // Rows arrive in half batches of four; two half batches fill one 32-bit survivor mask.  Few
  // registers per thread matter more here than deep per-warp prefetch: the kernel needs ~14
  // instructions per pixel, so it only keeps up with HBM when many warps interleave their load
  // and compare phases (measured: 24 warps with register double buffering were slower).
  constexpr int kHalf = kQuadBatch / 2;
  auto load_half = [&](float4 (&buf)[kHalf], int yfirst) {       // new rows yfirst .. yfirst+3 (interior index)
    if (yfirst + kHalf - 1 < hi) {                               // warp uniform: all of them are interior rows
      const float* rp = colp + (int64_t)(yfirst + 1) * sy;
#pragma unroll
      for (int k = 0; k < kHalf; ++k) buf[k] = load_at(rp + k * sy);
    } else {
#pragma unroll
      for (int k = 0; k < kHalf; ++k) buf[k] = load_row(yfirst + k);
    }
  };
  unsigned mask = 0, maskp = 0;
  auto eval_half = [&](const float4 (&buf)[kHalf], int shift) {
#pragma unroll
    for (int k = 0; k < kHalf; ++k) {
      const float4 n = buf[k];
      const float l = __shfl_up_sync(kFull, n.w, 1), r = __shfl_down_sync(kFull, n.x, 1);
      const float n0 = fmaxf(fmaxf(l, n.x), n.y), n1 = fmaxf(fmaxf(n.x, n.y), n.z);
      const float n2 = fmaxf(fmaxf(n.y, n.z), n.w), n3 = fmaxf(fmaxf(n.z, n.w), r);
      const bool k0 = c0 > fmaxf(fmaxf(a0, cl), tv) && c0 >= fmaxf(c1, n0);
      const bool k1 = c1 > fmaxf(fmaxf(a1, c0), tv) && c1 >= fmaxf(c2, n1);
      const bool k2 = c2 > fmaxf(fmaxf(a2, c1), tv) && c2 >= fmaxf(c3, n2);
      const bool k3 = c3 > fmaxf(fmaxf(a3, c2), tv) && c3 >= fmaxf(cr, n3);
      const int sh = shift + 4 * k;
      if (kPosThr) {
        // chained predicate + one predicated OR per pixel (the compiler otherwise emits two SETP, two SEL and an add)
        auto keep_bit = [&](float c, float m_gt, float m_ge, unsigned bit) {
          asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\tsetp.ge.and.f32 p, %1, %3, p;\n\t@p or.b32 %0, %0, %4;\n\t}"
              : "+r"(mask) : "f"(c), "f"(m_gt), "f"(m_ge), "r"(bit));
        };
        keep_bit(c0, fmaxf(fmaxf(a0, cl), tv), fmaxf(c1, n0), 1u << sh);
        keep_bit(c1, fmaxf(fmaxf(a1, c0), tv), fmaxf(c2, n1), 2u << sh);
        keep_bit(c2, fmaxf(fmaxf(a2, c1), tv), fmaxf(c3, n2), 4u << sh);
        keep_bit(c3, fmaxf(fmaxf(a3, c2), tv), fmaxf(cr, n3), 8u << sh);
      } else {
      if (k0) mask |= 1u << sh;
      if (k1) mask |= 2u << sh;
      if (k2) mask |= 4u << sh;
      if (k3) mask |= 8u << sh;
      }
      if (!kPosThr) {
        if (k0 && c0 > 0.f) maskp |= 1u << sh;
        if (k1 && c1 > 0.f) maskp |= 2u << sh;
        if (k2 && c2 > 0.f) maskp |= 4u << sh;
        if (k3 && c3 > 0.f) maskp |= 8u << sh;
      }
      a0 = h0; a1 = h1; a2 = h2; a3 = h3;
      c0 = n.x; c1 = n.y; c2 = n.z; c3 = n.w; cl = l; cr = r;
      h0 = n0; h1 = n1; h2 = n2; h3 = n3;
    }
  };
  const int sy32 = (int)sy;                                      // host checks the row stride fits
  int n_all = 0;
  float4 A[kHalf];
  unsigned kbits_max = 0u, kbits_nmin = 0u;              // largest score bits / largest ~score bits among emitted survivors
#pragma unroll 1
  for (int yb = y0; yb < y1; yb += kQuadBatch) {         // centre rows yb .. yb+7, new rows yb+1 .. yb+8
    mask = 0; maskp = 0;
    load_half(A, yb + 1);
    eval_half(A, 0);
    load_half(A, yb + 1 + kHalf);
    eval_half(A, 4 * kHalf);
    const int vr = y1 - yb;                              // centre rows of this group inside the strip
    const unsigned valid = lanevalid & (vr >= kQuadBatch ? kFull : ((1u << (4 * vr)) - 1u));
    mask &= valid;
    unsigned emit = kPosThr ? mask : (maskp & valid);
    n_all += __popc(mask);
    const int cnt = __popc(emit);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    const int total = __shfl_sync(kFull, inc, 31);
    int base = 0;
    if (total && lane == 31) base = atomicAdd(cand_count + b * kCntStride, total);
    if (total) {                                         // warp uniform
      base = __shfl_sync(kFull, base, 31);
      u64* dst = cand + (int64_t)b * cand_cap + base + inc - cnt;
      const float* vbase = img + (int64_t)(yb + 1) * sy + x0;
      const unsigned ibase = 0xffffffffu - ((unsigned)yb * (unsigned)wi + (unsigned)(x0 - 1));
      while (emit) {                                     // two survivors per trip: both score reads in flight together
        const int b0 = __ffs(emit) - 1;
        emit &= emit - 1;
        const bool two = emit != 0;
        const int b1 = two ? __ffs(emit) - 1 : b0;
        emit &= emit - 1;
        const float v0 = __ldg(vbase + ((b0 >> 2) * sy32 + (b0 & 3)));   // just streamed through L1 by this warp
        const float v1 = __ldg(vbase + ((b1 >> 2) * sy32 + (b1 & 3)));
        dst[0] = ((u64)__float_as_uint(v0) << 32) | (u64)(ibase - (unsigned)((b0 >> 2) * wi + (b0 & 3)));
        if (two) dst[1] = ((u64)__float_as_uint(v1) << 32) | (u64)(ibase - (unsigned)((b1 >> 2) * wi + (b1 & 3)));
        dst += 2;
        // key range of the image (v1 == v0 when there is no second survivor)
        kbits_max = max(kbits_max, max(__float_as_uint(v0), __float_as_uint(v1)));
        kbits_nmin = max(kbits_nmin, max(~__float_as_uint(v0), ~__float_as_uint(v1)));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_all += __shfl_xor_sync(kFull, n_all, o);
  if (lane == 0 && n_all) atomicAdd(counts + b, n_all);
  // the selection kernel starts its radix walk at the first bit in which the image's keys differ: leave the
  // range of the score bits beside the candidate counter (same zero-initialised 128-byte line) so that it does
  // not need a pass over the candidates to find it
  kbits_max = __reduce_max_sync(kFull, kbits_max);
  kbits_nmin = __reduce_max_sync(kFull, kbits_nmin);
  if (lane == 0 && kbits_nmin) {
    atomicMax(reinterpret_cast<unsigned*>(cand_count) + b * kCntStride + kCntKeyMax, kbits_max);
    atomicMax(reinterpret_cast<unsigned*>(cand_count) + b * kCntStride + kCntKeyNegMin, kbits_nmin);
  }
}

// ---- phase 2: select + sort + centroid ------------------------------------
__device__ void bitonic_steps(u64* a, int P, int k, int j_from, int j_to, int64_t gbase) {
  // descending bitonic network steps j = j_from, j_from/2, ..., j_to for stage k
  // on the P elements a[0..P); gbase = global index of a[0] (direction bit).
  for (int j = j_from; j >= j_to; j >>= 1) {
    for (int p = threadIdx.x; p < P / 2; p += blockDim.x) {
      const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
      const int l = i | j;
      const u64 x = a[i], y = a[l];
      const bool up = (((int64_t)i + gbase) & k) != 0;  // ascending block of the network
      if (up ? (x > y) : (x < y)) { a[i] = y; a[l] = x; }
    }
    __syncthreads();
  }
}

// Shared-memory bitonic network (descending), 4 independent compare-exchanges
// per thread in flight.  `s` must point into shared memory.
__device__ __forceinline__ void bitonic_steps_smem(u64* s, int P, int k, int j_from, int j_to, int64_t gbase) {
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(s);
  const int npairs = P >> 1;
  for (int j = j_from; j >= j_to; j >>= 1) {
    for (int base = 0; base < npairs; base += 4 * kSelThreads) {
      u64 x[4], y[4];
      unsigned ai[4], al[4];
      bool act[4], up[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = base + u * kSelThreads + (int)threadIdx.x;
        act[u] = p < npairs;
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        ai[u] = sbase + 8u * (unsigned)i;
        al[u] = ai[u] + 8u * (unsigned)j;
        up[u] = (((int64_t)i + gbase) & k) != 0;
        if (act[u]) {
          asm volatile("ld.shared.u64 %0, [%1];" : "=l"(x[u]) : "r"(ai[u]));
          asm volatile("ld.shared.u64 %0, [%1];" : "=l"(y[u]) : "r"(al[u]));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (act[u]) {   // uniform except in the last partial group: always store, no divergent swap branch
          const bool sw = up[u] ? (x[u] > y[u]) : (x[u] < y[u]);
          const u64 a = sw ? y[u] : x[u], bq = sw ? x[u] : y[u];
          asm volatile("st.shared.u64 [%0], %1;" ::"r"(ai[u]), "l"(a) : "memory");
          asm volatile("st.shared.u64 [%0], %1;" ::"r"(al[u]), "l"(bq) : "memory");
        }
      }
    }
    __syncthreads();
  }
}

// Sort P (power of two) keys descending.  If the keys live in shared memory (P <=
// CH) the whole network runs there; otherwise steps with stride >= CH run in
// global memory and the rest chunk by chunk through the shared buffer `s`.
__device__ void bitonic_sort_desc(u64* g, int P, u64* s, int CH, bool in_smem) {
  if (in_smem) {
    for (int k = 2; k <= P; k <<= 1) bitonic_steps_smem(g, P, k, k >> 1, 1, 0);
    return;
  }
  const int nchunk = P / CH;
  for (int c = 0; c < nchunk; ++c) {
    for (int i = threadIdx.x; i < CH; i += blockDim.x) s[i] = g[(int64_t)c * CH + i];
    __syncthreads();
    for (int k = 2; k <= CH; k <<= 1) bitonic_steps_smem(s, CH, k, k >> 1, 1, (int64_t)c * CH);
    for (int i = threadIdx.x; i < CH; i += blockDim.x) g[(int64_t)c * CH + i] = s[i];
    __syncthreads();
  }
  for (int k = CH << 1; k <= P; k <<= 1) {
    bitonic_steps(g, P, k, k >> 1, CH, 0);
    for (int c = 0; c < nchunk; ++c) {
      for (int i = threadIdx.x; i < CH; i += blockDim.x) s[i] = g[(int64_t)c * CH + i];
      __syncthreads();
      bitonic_steps_smem(s, CH, k, CH >> 1, 1, (int64_t)c * CH);
      for (int i = threadIdx.x; i < CH; i += blockDim.x) g[(int64_t)c * CH + i] = s[i];
      __syncthreads();
    }
  }
}

// ---- register / shuffle / shared-memory hybrid bitonic sort (descending) ----------------------
// P = E * kSelThreads keys, E per thread in registers.  Layout A: thread t holds indices t*E + e, so
// network steps with stride j < E are register compare-exchanges, E <= j < 32E are warp shuffles.  The
// five strides that cross warps (j >= 32E) are done in layout B, reached through a padded shared-memory
// transpose: register bits = index bits [10, 10+eb), lane bits = index bits [0, eb) and [eb+5, 10),
// warp = index bits [eb, eb+5) -- there every one of those strides is a register or shuffle step too.
// Shared memory is touched twice per stage k >= 64E (10 round trips at P = 8192) instead of once per
// step (91): the plain network is bound by shared-memory bandwidth (32 B per compare-exchange).
__device__ __forceinline__ int sort_pad(int i) { return i + (i >> 4) + ((i >> 8) << 3); }

template <int E, int JE>
__device__ __forceinline__ void cx_reg(u64 (&r)[E], int base, int rshift, int k) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if ((e & JE) == 0) {
      const bool desc = ((base | (e << rshift)) & k) == 0;
      const u64 a = r[e], b = r[e | JE];
      const bool sw = desc ? (a < b) : (a > b);
      r[e] = sw ? b : a;
      r[e | JE] = sw ? a : b;
    }
  }
}
// all register steps with element stride <= JE (JE, JE/2, ..., 1)
template <int E, int JE>
__device__ __forceinline__ void cx_reg_from(u64 (&r)[E], int base, int rshift, int k) {
  if constexpr (JE >= 1) {
    cx_reg<E, JE>(r, base, rshift, k);
    cx_reg_from<E, JE / 2>(r, base, rshift, k);
  }
}
// register steps of layout B for stage k: element strides (k/2 >> 10) ... 1, i.e. only those <= JE_MAX
template <int E, int JE>
__device__ __forceinline__ void cx_reg_upto(u64 (&r)[E], int base, int rshift, int k, int je_max) {
  if constexpr (JE >= 1) {
    if (JE <= je_max) cx_reg<E, JE>(r, base, rshift, k);
    cx_reg_upto<E, JE / 2>(r, base, rshift, k, je_max);
  }
}
template <int E>
__device__ __forceinline__ void cx_shfl(u64 (&r)[E], int base, int rshift, int k, int j, int lane_mask) {
  const bool lower = (base & j) == 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const u64 a = r[e];
    const u64 b = __shfl_xor_sync(0xffffffffu, a, lane_mask);
    const bool desc = ((base | (e << rshift)) & k) == 0;
    const bool keep_max = lower == desc;
    const bool b_gt = b > a;
    r[e] = (keep_max == b_gt) ? b : a;
  }
}

// buf: P unordered keys in shared memory (buf == s); s: kSortSmemKeys slots.  Sorted keys end up in buf[0..P).
template <int E>
__device__ void bitonic_sort_desc_reg(u64* buf, u64* s) {
  constexpr int EB = E == 2 ? 1 : E == 4 ? 2 : E == 8 ? 3 : 4;
  constexpr int P = E * kSelThreads;
  static_assert(kSelThreads == 1024 && (1 << EB) == E, "layout arithmetic assumes 32 warps and E in {2,4,8,16}");
  static_assert(P + (P >> 4) + ((P >> 8) << 3) <= kSortSmemKeys, "padded transpose buffer must fit");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64 r[E];
#pragma unroll
  for (int e = 0; e < E; ++e) r[e] = buf[e * kSelThreads + tid];   // any bijection will do: the input is unordered
  __syncthreads();
  const int baseA = tid * E;
  const int baseB = (lane & (E - 1)) | (warp << EB) | ((lane >> EB) << (EB + 5));
  // stages inside one thread
#pragma unroll
  for (int k = 2; k <= E; k <<= 1) {
    if (k == 2) cx_reg_from<E, 1>(r, baseA, 0, k);
    if (k == 4) cx_reg_from<E, (E >= 4 ? 2 : 0)>(r, baseA, 0, k);
    if (k == 8) cx_reg_from<E, (E >= 8 ? 4 : 0)>(r, baseA, 0, k);
    if (k == 16) cx_reg_from<E, (E >= 16 ? 8 : 0)>(r, baseA, 0, k);
  }
  // stages inside one warp
  for (int k = 2 * E; k <= 32 * E; k <<= 1) {
    for (int j = k >> 1; j >= E; j >>= 1) cx_shfl<E>(r, baseA, 0, k, j, j >> EB);
    cx_reg_from<E, E / 2>(r, baseA, 0, k);
  }
  // stages across warps
  u64* sp = s;
  for (int k = 64 * E; k <= P; k <<= 1) {
#pragma unroll
    for (int e = 0; e < E; ++e) sp[sort_pad(baseA + e)] = r[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e) r[e] = sp[sort_pad(baseB | (e << 10))];
    __syncthreads();
    cx_reg_upto<E, E / 2>(r, baseB, 10, k, (k >> 1) >> 10);
    for (int j = min(k >> 1, 512); j >= 32 * E; j >>= 1) cx_shfl<E>(r, baseB, 10, k, j, (j >> (EB + 5)) << EB);
#pragma unroll
    for (int e = 0; e < E; ++e) sp[sort_pad(baseB | (e << 10))] = r[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e) r[e] = sp[sort_pad(baseA + e)];
    __syncthreads();
    for (int j = 16 * E; j >= E; j >>= 1) cx_shfl<E>(r, baseA, 0, k, j, j >> EB);
    cx_reg_from<E, E / 2>(r, baseA, 0, k);
  }
  // sorted, layout A -> buf[0..P) (through the padded buffer: a direct store would be an 8-way bank conflict)
#pragma unroll
  for (int e = 0; e < E; ++e) sp[sort_pad(baseA + e)] = r[e];
  __syncthreads();
#pragma unroll
  for (int e = 0; e < E; ++e) r[e] = sp[sort_pad(e * kSelThreads + tid)];
  __syncthreads();
#pragma unroll
  for (int e = 0; e < E; ++e) buf[e * kSelThreads + tid] = r[e];
  __syncthreads();
}

__global__ void __launch_bounds__(kSelThreads)
select_kernel(const float* __restrict__ score, int B, int H, int W, int64_t sb, int64_t sy,
              int num_pts, int min_pts, int cap_pts, int n_fixed,
              const int32_t* __restrict__ counts, const int32_t* __restrict__ cand_count,
              const u64* __restrict__ cand, int64_t cand_cap, u64* __restrict__ sortbuf,
              int64_t sort_cap, int32_t* __restrict__ status, int32_t* __restrict__ n_out,
              int64_t* __restrict__ idx_out, float* __restrict__ kps_out,
              float* __restrict__ kpscore_out, int fullmap, int debug) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* s_keys = (u64*)smem_raw;  // kSortSmemKeys
  __shared__ unsigned s_hist[1 << kDigitBits];
  __shared__ unsigned s_off[1 << kDigitBits];      // rank path: keys in higher bins (= first output slot of the bin)
  __shared__ int s_maxbin;
  __shared__ unsigned s_bitmap[kFillBitmapWords];
  __shared__ unsigned s_fill[kFillMax];
  __shared__ int s_cnt, s_cnt_lo;
  __shared__ int s_wsum[kSelThreads / 32];
  __shared__ u64 s_prefix;
  __shared__ int s_above, s_G, s_done, s_shift0;
  __shared__ u64 s_red[2][kSelThreads / 32];

  const int b = blockIdx.x;
  const int hi = H - 2, wi = W - 2;
  const int64_t n_interior = (int64_t)hi * wi;
  const int tid = threadIdx.x, lane = tid & 31;
  long long tk[8];
  int ntk = 0;
#define PF_TICK() do { if (debug && b == 0 && tid == 0 && ntk < 8) tk[ntk++] = clock64(); } while (0)
  PF_TICK();

  // n: :249-261 of the reference (min over the batch, then the 128 floor)
  int n;
  if (n_fixed >= 0) {
    n = n_fixed;
  } else {
    int m = 0x7fffffff;
    for (int i = 0; i < B; ++i) m = min(m, counts[i]);
    n = num_pts > 0 ? min(num_pts, m) : m;
    n = max(n, min_pts);
  }
  if (b == 0 && tid == 0) *n_out = n;
  if (n == 0) return;
  if (n > cap_pts || (int64_t)n > n_interior) {
    if (tid == 0) atomicMax(status, n > cap_pts ? 1 : 2);
    return;
  }
  const int C = (int)min((int64_t)cand_count[b * kCntStride], cand_cap);
  const int n_real = min(n, C);
  const u64* keys = cand + (int64_t)b * cand_cap;

  // ---- lower bound LB with n_real <= #{key >= LB} = G, G small enough to sort
  u64 LB = 0;
  int G = C;
  // after the radix walk: keys with (key >> hi_shift) > hi_prefix are certain winners (n_above of
  // them), keys with (key >> hi_shift) == hi_prefix form the boundary bucket
  bool split = false;
  int hi_shift = 0, n_above = 0;
  u64 hi_prefix = 0;
  int passes = 0, rank_shift = 0, rank_bucket = 0;
  u64 rank_base = 0;
  if (C > kRadixMinKeys) {
    // common leading bits (scores of one map share sign/exponent bits): start the
    // radix walk at the first differing bit so the histogram bins actually spread
    u64 kmin = ~0ull, kmax = 0ull;
    const unsigned rec_max = (unsigned)cand_count[b * kCntStride + kCntKeyMax];
    const unsigned rec_nmin = (unsigned)cand_count[b * kCntStride + kCntKeyNegMin];
    const bool recorded = rec_nmin != 0u && (int64_t)cand_count[b * kCntStride] <= cand_cap;   // block uniform
    if (recorded) {            // range of the score bits left by the NMS kernel: no pass over the candidates
      kmin = (u64)(~rec_nmin) << 32;
      kmax = ((u64)rec_max << 32) | 0xffffffffull;
    }
    for (int base = 0; !recorded && base < C; base += kSelThreads * kUnroll) {
      u64 k[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int i = base + u * kSelThreads + tid;
        k[u] = i < C ? keys[i] : keys[0];
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        kmin = k[u] < kmin ? k[u] : kmin;
        kmax = k[u] > kmax ? k[u] : kmax;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const u64 a = __shfl_xor_sync(0xffffffffu, kmin, o), bb = __shfl_xor_sync(0xffffffffu, kmax, o);
      kmin = a < kmin ? a : kmin;
      kmax = bb > kmax ? bb : kmax;
    }
    if (lane == 0) { s_red[0][tid >> 5] = kmin; s_red[1][tid >> 5] = kmax; }
    __syncthreads();
    if (tid == 0) {
      for (int k = 1; k < kSelThreads / 32; ++k) {
        kmin = s_red[0][k] < kmin ? s_red[0][k] : kmin;
        kmax = s_red[1][k] > kmax ? s_red[1][k] : kmax;
      }
      // bucket = [base, base + 2^width): starts as the smallest power-of-two window over [kmin, kmax] anchored at
      // kmin, so the 2048 bins of the first pass spread over the keys' actual RANGE (a bit field that starts at
      // the highest differing bit lands on the exponent and leaves a boundary bucket of thousands of keys)
      const u64 range = kmax - kmin;
      s_shift0 = range ? 64 - __clzll((long long)range) : 1;                       // width in bits, 1 .. 63
      s_prefix = kmin;                                                              // base
      s_above = 0;
      s_done = 0;
    }
    __syncthreads();
    PF_TICK();
    int width = s_shift0;                           // current bucket: keys k with (k - base) >> width == 0
    int shift = 0;
    if (tid == 0) s_maxbin = 0;
    for (;;) {
      ++passes;
      for (int i = tid; i < (1 << kDigitBits); i += kSelThreads) s_hist[i] = 0;
      __syncthreads();
      const u64 base_key = s_prefix;
      shift = width > kDigitBits ? width - kDigitBits : 0;         // digit = (k - base) >> shift, < 2^(width - shift)
      const int nb = 1 << (width - shift);
      for (int base = 0; base < C; base += kSelThreads * kUnroll) {
        u64 k[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int i = base + u * kSelThreads + tid;
          k[u] = i < C ? keys[i] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int i = base + u * kSelThreads + tid;
          const u64 off = k[u] - base_key;                          // keys below the bucket wrap to huge offsets
          if (i < C && (off >> width) == 0) atomicAdd(&s_hist[(unsigned)(off >> shift)], 1u);
        }
      }
      __syncthreads();
      // find the digit d with  above(d) < n_real <= above(d) + hist[d]  (above = keys in higher
      // bins): block-wide suffix sum, two bins per thread
      {
        const int b1 = 2 * tid + 1, b0 = 2 * tid;
        const int h1 = b1 < nb ? (int)s_hist[b1] : 0, h0 = b0 < nb ? (int)s_hist[b0] : 0;
        // inclusive scan over threads in DESCENDING tid order
        int v = h0 + h1, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t2 = __shfl_down_sync(0xffffffffu, inc, o);
          if (lane + o < 32) inc += t2;
        }
        if (lane == 0) s_wsum[tid >> 5] = inc;          // total of this warp
        __syncthreads();
        int higher = 0;                                   // sum over warps with larger index
        for (int wq = (tid >> 5) + 1; wq < kSelThreads / 32; ++wq) higher += s_wsum[wq];
        const int above1 = s_above + higher + (inc - v);  // keys in bins > b1
        const int above0 = above1 + h1;                   // keys in bins > b0
        if (b1 < nb) s_off[b1] = (unsigned)above1;
        if (b0 < nb) s_off[b0] = (unsigned)above0;
        __syncthreads();                                  // everyone has read s_above
        const bool hit1 = b1 < nb && above1 < n_real && n_real <= above1 + h1;
        const bool hit0 = b0 < nb && !hit1 && above0 < n_real && (n_real <= above0 + h0 || b0 == 0);
        if (hit1 || hit0) {
          const int d = hit1 ? b1 : b0;
          const int cum = hit1 ? above1 : above0;
          s_prefix = base_key + ((u64)d << shift);        // the boundary bin becomes the bucket
          s_above = cum;
          s_G = cum + (hit1 ? h1 : h0);
          const int bucket = hit1 ? h1 : h0;
          if ((s_G <= kSortSmemKeys && (bucket <= kBucketMax || cum + bucket == n_real)) || shift == 0) s_done = 1;
        }
      }
      __syncthreads();
      if (s_done && passes == 1) {
        // largest bin above the boundary bin (the rank path orders every bin by counting, quadratic in its size)
        const int d = (int)((s_prefix - base_key) >> shift);
        const int b1 = 2 * tid + 1, b0 = 2 * tid;
        int m = 0;
        if (b1 < nb && b1 > d) m = max(m, (int)s_hist[b1]);
        if (b0 < nb && b0 > d) m = max(m, (int)s_hist[b0]);
        m = __reduce_max_sync(0xffffffffu, m);
        if (lane == 0 && m) atomicMax(&s_maxbin, m);
        rank_base = base_key;
        rank_shift = shift;
        rank_bucket = (int)s_hist[d];
        __syncthreads();
      }
      width = shift;
      if (s_done) break;
    }
    LB = s_prefix;
    G = s_G;
    hi_shift = width;                               // boundary bucket: (k - LB) >> hi_shift == 0
    hi_prefix = 0;
    n_above = s_above;
    split = true;
  }
  const bool in_smem = G <= kSortSmemKeys;
  int P = 2;
  while (P < G) P <<= 1;
  u64* buf = in_smem ? s_keys : sortbuf + (int64_t)b * sort_cap;
  if (!in_smem && (int64_t)P > sort_cap) {
    if (tid == 0) atomicMax(status, 3);
    return;
  }
  PF_TICK();
  // ---- rank path (one histogram pass sufficed, every bin small): the histogram already says where each bin of
  // winners starts in the output, so the gather drops every key into its bin's slot range and the order inside a
  // bin comes from counting the larger keys of that bin -- no sorting network at all (the two bitonic sorts were
  // 40 % of this kernel).  Keys are distinct (the pixel index is part of them), so the ranks are a permutation.
  constexpr int kRankPerThread = 9;
  const bool ranked = split && passes == 1 && in_smem && G <= kRankPerThread * kSelThreads && s_maxbin <= 512 &&
                      rank_bucket <= kBucketMax;
  if (ranked) {
    for (int i = tid; i < (1 << kDigitBits); i += kSelThreads) s_hist[i] = 0;     // becomes the per-bin cursor, then the count
    __syncthreads();
    for (int base = 0; base < C; base += kSelThreads * kUnroll) {
      u64 k[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int i = base + u * kSelThreads + tid;
        k[u] = i < C ? keys[i] : 0ull;
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int i = base + u * kSelThreads + tid;
        if (i < C && k[u] >= LB) {
          const unsigned bin = (unsigned)((k[u] - rank_base) >> rank_shift);
          buf[s_off[bin] + atomicAdd(&s_hist[bin], 1u)] = k[u];
        }
      }
    }
    __syncthreads();
    PF_TICK();
    u64 mine[kRankPerThread];
    int fin[kRankPerThread];
#pragma unroll
    for (int e = 0; e < kRankPerThread; ++e) {
      const int p = e * kSelThreads + tid;
      fin[e] = -1;
      if (p < G) {
        const u64 key = buf[p];
        const unsigned bin = (unsigned)((key - rank_base) >> rank_shift);
        const int start = (int)s_off[bin], cnt = (int)s_hist[bin];
        int r = 0;
        for (int q = 0; q < cnt; ++q) r += buf[start + q] > key ? 1 : 0;
        mine[e] = key;
        fin[e] = start + r;
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kRankPerThread; ++e)
      if (fin[e] >= 0) buf[fin[e]] = mine[e];
    __syncthreads();
    G = n_real;          // buf[0, n_real) holds the winners in final order
    PF_TICK();
    PF_TICK();
  } else {
  // ---- gather keys >= LB: certain winners from the front, boundary-bucket keys from the back
  if (tid == 0) { s_cnt = 0; s_cnt_lo = 0; }
  __syncthreads();
  for (int base = 0; base < C; base += kSelThreads * kUnroll) {
    u64 k[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = base + u * kSelThreads + tid;
      k[u] = i < C ? keys[i] : 0ull;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = base + u * kSelThreads + tid;
      const bool take = i < C && k[u] >= LB;
      const bool lo = take && split && ((k[u] - LB) >> hi_shift) == hi_prefix;
      const bool hi = take && !lo;
      const unsigned bal_hi = __ballot_sync(0xffffffffu, hi);
      const unsigned bal_lo = __ballot_sync(0xffffffffu, lo);
      if (bal_hi) {
        int pos = 0;
        if (lane == 0) pos = atomicAdd(&s_cnt, __popc(bal_hi));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (hi) buf[pos + __popc(bal_hi & ((1u << lane) - 1u))] = k[u];
      }
      if (bal_lo) {
        int pos = 0;
        if (lane == 0) pos = atomicAdd(&s_cnt_lo, __popc(bal_lo));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (lo) buf[G - 1 - (pos + __popc(bal_lo & ((1u << lane) - 1u)))] = k[u];
      }
    }
  }
  __syncthreads();
  PF_TICK();
  // ---- if the boundary bucket is small, sort it alone and keep its top (n_real - n_above):
  // the final sort then runs on exactly n_real keys (half the network for n = 8192)
  if (split && in_smem && G > n_real) {
    const int pop = G - n_above;
    int P2 = 2;
    while (P2 < pop) P2 <<= 1;
    if (pop <= kBucketMax && n_above + P2 <= kSortSmemKeys) {
      for (int i = G + tid; i < n_above + P2; i += kSelThreads) buf[i] = 0;
      __syncthreads();
      for (int k = 2; k <= P2; k <<= 1) bitonic_steps_smem(buf + n_above, P2, k, k >> 1, 1, 0);
      G = n_real;   // buf[0, n_real) now holds exactly the winners (unordered front + sorted bucket top)
      P = 2;
      while (P < G) P <<= 1;
    }
  }
  for (int i = G + tid; i < P; i += kSelThreads) buf[i] = 0;
  __syncthreads();
  PF_TICK();
  if (in_smem && P == 8 * kSelThreads) bitonic_sort_desc_reg<8>(buf, s_keys);
  else if (in_smem && P == 4 * kSelThreads) bitonic_sort_desc_reg<4>(buf, s_keys);
  else if (in_smem && P == 2 * kSelThreads) bitonic_sort_desc_reg<2>(buf, s_keys);
  else bitonic_sort_desc(buf, P, s_keys, kSortSmemKeys, in_smem);
  PF_TICK();
  }   // !ranked

  // ---- filler for rows [n_real, n): lowest-index pixels that are not winners
  const int f = n - n_real;
  if (f > 0) {
    if (f > kFillMax || n > kFillBitmapWords * 32) {
      if (tid == 0) atomicMax(status, 4);
      return;
    }
    for (int i = tid; i < kFillBitmapWords; i += kSelThreads) s_bitmap[i] = 0;
    __syncthreads();
    for (int i = tid; i < n_real; i += kSelThreads) {
      const unsigned idx = 0xffffffffu - (unsigned)buf[i];
      if (idx < (unsigned)n) atomicOr(&s_bitmap[idx >> 5], 1u << (idx & 31));
    }
    __syncthreads();
    if (tid == 0) {
      int cnt = 0;
      for (int wd = 0; wd < (n + 31) / 32 && cnt < f; ++wd) {
        unsigned free_bits = ~s_bitmap[wd];
        while (free_bits && cnt < f) {
          const int bit = __ffs(free_bits) - 1;
          free_bits &= free_bits - 1;
          const int p = wd * 32 + bit;
          if (p < n) s_fill[cnt++] = (unsigned)p;
        }
      }
    }
    __syncthreads();
  }

  // ---- winners' interior indices; centroid and score follow in keypoint_outputs_kernel (their 3x3 reads are
  // scattered over the map: one CTA per image keeps too few of them in flight)
  for (int i = tid; i < n; i += kSelThreads) {
    const unsigned idx = i < n_real ? 0xffffffffu - (unsigned)buf[i] : s_fill[i - n_real];
    idx_out[(int64_t)b * cap_pts + i] = (int64_t)idx;
  }
  PF_TICK();
  if (debug && b == 0 && tid == 0) {
    printf("select_kernel C=%d n=%d n_real=%d G=%d P=%d n_above=%d split=%d in_smem=%d shift=%d ticks:", C, n, n_real, G, P, n_above, (int)split, (int)in_smem, hi_shift);
    for (int i = 1; i < ntk; ++i) printf(" %lld", tk[i] - tk[i - 1]);
    printf("\n");
  }
#undef PF_TICK
}

// centroid (3x3 score-weighted mean of the bit-exact linspace grid, :243-246), 3x3 maximum (:247) and the
// gathers (:266-267) at the winners only; one thread per keypoint, the whole batch in one grid
__global__ void __launch_bounds__(256)
keypoint_outputs_kernel(const float* __restrict__ score, int H, int W, int64_t sb, int64_t sy, int cap_pts,
                        const int32_t* __restrict__ status, const int32_t* __restrict__ n_dev,
                        const int64_t* __restrict__ idx_in, float* __restrict__ kps_out,
                        float* __restrict__ kpscore_out, int fullmap) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (*status != 0 || i >= *n_dev) return;
  const int hi = H - 2, wi = W - 2;
  const int64_t o = (int64_t)b * cap_pts + i;
  const unsigned idx = (unsigned)idx_in[o];
  const int y0 = idx / wi, x0 = idx - y0 * wi;
  const float* img = score + b * sb;
  if (fullmap) {
    // generate_kpts_single_noavg (:280-336): the pixel's own grid coordinate and score; (H, W) here
    // are the virtual padded sizes, the real map is the hi x wi "interior"
    kps_out[2 * o + 0] = linspace_pm1(x0, wi, 2.0f / (float)(wi - 1));
    kps_out[2 * o + 1] = linspace_pm1(y0, hi, 2.0f / (float)(hi - 1));
    kpscore_out[o] = __ldg(img + (int64_t)(y0 + 1) * sy + x0 + 1);
    return;
  }
  const float stepx = 2.0f / (float)(W - 1), stepy = 2.0f / (float)(H - 1);
  float p[9];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const float* row = img + (int64_t)(y0 + dy) * sy + x0;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) p[dy * 3 + dx] = __ldg(row + dx);
  }
  float sx = 0.f, sy_ = 0.f, sw = 0.f, mx = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const float gy = linspace_pm1(y0 + dy, H, stepy);
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const float v = p[dy * 3 + dx];
      const float gx = linspace_pm1(x0 + dx, W, stepx);
      sx = __fadd_rn(sx, __fmul_rn(v, gx));
      sy_ = __fadd_rn(sy_, __fmul_rn(v, gy));
      sw = __fadd_rn(sw, v);
      mx = fmaxf(mx, v);
    }
  }
  const float wgt = __fdiv_rn(sw, 9.0f);
  kps_out[2 * o + 0] = __fdiv_rn(__fdiv_rn(sx, 9.0f), wgt);
  kps_out[2 * o + 1] = __fdiv_rn(__fdiv_rn(sy_, 9.0f), wgt);
  kpscore_out[o] = mx;
}

// POSFEAT_DETECT_FULLMAP: the whole map plays the role of the interior.  The kernels only ever read
// interior pixels (reflect padding is about the interior), so the map is presented as the interior of a
// virtual (H+2) x (W+2) map whose origin lies one row and one column before the real one.
static int check_common(const float*& score, int B, int& H, int& W, int64_t sy, int mode) {
  PF_CHECK_ARG(score != nullptr, "score is NULL");
  PF_CHECK_ARG(B >= 1, "B must be >= 1 (got %d)", B);
  PF_CHECK_ARG(sy >= W, "row stride %lld smaller than W=%d", (long long)sy, W);
  if (mode & POSFEAT_DETECT_FULLMAP) {
    score -= sy + 1;
    H += 2;
    W += 2;
  }
  PF_CHECK_ARG(H >= 4 && W >= 4, "score map must be at least 4x4 (got %dx%d)", H, W);
  PF_CHECK_ARG((int64_t)(H - 2) * (W - 2) < 0xffffffffLL, "interior grid too large for 32-bit indices");
  return 0;
}

}  // namespace posfeat

using namespace posfeat;

extern "C" size_t posfeat_detect_workspace_bytes(int B, int H, int W, int cap_pts) {
  if (B < 1 || H < 4 || W < 4 || cap_pts < 1) return 0;
  return carve(nullptr, B, H, W, cap_pts).total;
}

extern "C" int posfeat_detect_candidates_f32(const float* score, int B, int H, int W, int64_t stride_b,
                                             int64_t stride_y, int nms_mode, int radius, int thr_mode,
                                             float thr, int32_t* counts, void* workspace, size_t ws_bytes,
                                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_common(score, B, H, W, stride_y, nms_mode)) return e;
  nms_mode &= ~POSFEAT_DETECT_FULLMAP;
  PF_CHECK_ARG(nms_mode >= POSFEAT_NMS_NONE && nms_mode <= POSFEAT_NMS_SOFT, "unsupported nms_mode %d", nms_mode);
  PF_CHECK_ARG(nms_mode != POSFEAT_NMS_SOFT || thr_mode != POSFEAT_THR_NONE,
               "soft NMS needs a threshold (the reference's thr_mask is undefined otherwise)");
  PF_CHECK_ARG(thr_mode >= POSFEAT_THR_NONE && thr_mode <= POSFEAT_THR_MEAN, "unsupported thr_mode %d", thr_mode);
  PF_CHECK_ARG(radius >= 0 && radius <= 16, "nms radius %d outside [0,16]", radius);
  PF_CHECK_ARG(nms_mode == POSFEAT_NMS_NONE || (radius <= H - 3 && radius <= W - 3),
               "reflect padding needs radius < interior size");
  PF_CHECK_ARG(counts && workspace, "counts/workspace is NULL");
  DetectWs w = carve(workspace, B, H, W, 1);
  if (ws_bytes < w.total) return set_error(POSFEAT_EWORKSPACE, "detect workspace: need %zu bytes, got %zu", w.total, ws_bytes);
  const int r = nms_mode != POSFEAT_NMS_NONE ? radius : 0;

  // zero the small header (cand_count, thr_val, red_sum, red_max, status) and counts
  PF_CUDA(cudaMemsetAsync(workspace, 0, (size_t)((char*)w.cand - (char*)workspace), stream));
  PF_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * B, stream));
  const int has_thr = thr_mode != POSFEAT_THR_NONE;
  if (has_thr) {
    if (thr_mode == POSFEAT_THR_MAX || thr_mode == POSFEAT_THR_MEAN) {
      dim3 g(64, B);
      reduce_interior_kernel<<<g, 256, 0, stream>>>(score, H, W, stride_b, stride_y, w.red_sum, w.red_max);
      PF_LAUNCH_CHECK("reduce_interior_kernel");
    }
    finalize_thr_kernel<<<(B + 127) / 128, 128, 0, stream>>>(B, thr_mode, thr, (int64_t)(H - 2) * (W - 2), w.red_sum,
                                                              w.red_max, w.thr_val);
    PF_LAUNCH_CHECK("finalize_thr_kernel");
  }
  ProfScope prof(PROF_NMS, stream);
  if (r <= 3 && nms_mode != POSFEAT_NMS_SOFT) {
    // register sliding-window kernel; without NMS every pixel may survive, so strips are 16 rows tall
    const int rows = nms_mode == POSFEAT_NMS_HARD ? 64 : 16;
    const int use = 32 - 2 * r;
    dim3 grid((W - 2 + kStripWarps * use - 1) / (kStripWarps * use), (H - 2 + rows - 1) / rows, B);
#define PF_STRIP(RR)                                                                                              \
  nms_strip_kernel<RR><<<grid, kStripWarps * 32, 0, stream>>>(score, H, W, stride_b, stride_y, rows, has_thr,     \
                                                               w.thr_val, counts, w.cand_count, w.cand, w.cand_cap)
    if (r == 1) {
      const int ncw = (W - 2 + kQuadCols - 1) / kQuadCols, nstrips = (H - 2 + kQuadRows - 1) / kQuadRows;
      const int ntasks = ncw * nstrips;
      dim3 gq(B, (ntasks + kQuadWarps - 1) / kQuadWarps);
      PF_CHECK_ARG(stride_y < (1 << 27), "row stride too large");
      const bool vec = (W % 4 == 0) && (stride_y % 4 == 0) && (stride_b % 4 == 0) && (((uintptr_t)score & 15) == 0);
      const bool pos = thr_mode == POSFEAT_THR_ABS && thr >= 0.f;
#define PF_QUAD(V, P)                                                                                             \
  nms_quad_r1_kernel<V, P><<<gq, kQuadWarps * 32, 0, stream>>>(score, H, W, stride_b, stride_y, has_thr, w.thr_val, \
                                                               counts, w.cand_count, w.cand, w.cand_cap, ncw, ntasks)
      if (vec) { if (pos) PF_QUAD(true, true); else PF_QUAD(true, false); }
      else { if (pos) PF_QUAD(false, true); else PF_QUAD(false, false); }
#undef PF_QUAD
    } else if (r == 0) PF_STRIP(0);
    else if (r == 2) PF_STRIP(2);
    else PF_STRIP(3);
#undef PF_STRIP
    PF_LAUNCH_CHECK("nms_strip_kernel");
    return POSFEAT_OK;
  }
  const int pw = kTileW + 2 * r, ph = kTileH + 2 * r;
  const size_t smem = sizeof(float) * (size_t)ph * (pw | 1);
  dim3 grid((W - 2 + kTileW - 1) / kTileW, (H - 2 + kTileH - 1) / kTileH, B);
  nms_candidates_kernel<<<grid, kNmsThreads, smem, stream>>>(score, H, W, stride_b, stride_y, nms_mode, r, has_thr,
                                                             w.thr_val, counts, w.cand_count, w.cand, w.cand_cap);
  PF_LAUNCH_CHECK("nms_candidates_kernel");
  return POSFEAT_OK;
}

extern "C" int posfeat_detect_select_f32(const float* score, int B, int H, int W, int64_t stride_b, int64_t stride_y,
                                         int mode, int num_pts, int min_pts, int cap_pts, int n_fixed, const int32_t* counts,
                                         int32_t* n_out, int64_t* idx_out, float* kps_out, float* kpscore_out,
                                         void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_common(score, B, H, W, stride_y, mode)) return e;
  PF_CHECK_ARG(cap_pts >= 1 && num_pts >= 0 && min_pts >= 0, "bad num_pts/min_pts/cap_pts");
  PF_CHECK_ARG(n_fixed <= cap_pts, "n_fixed %d exceeds cap_pts %d", n_fixed, cap_pts);
  PF_CHECK_ARG(counts && n_out && idx_out && kps_out && kpscore_out && workspace, "NULL output pointer");
  DetectWs w = carve(workspace, B, H, W, cap_pts);
  if (ws_bytes < w.total) return set_error(POSFEAT_EWORKSPACE, "detect workspace: need %zu bytes, got %zu", w.total, ws_bytes);
  const size_t smem = sizeof(u64) * kSortSmemKeys;
  // per device/context attribute: set on every call (cheap), never cached in a process-global flag
  PF_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(PROF_SELECT, stream);
  select_kernel<<<B, kSelThreads, smem, stream>>>(score, B, H, W, stride_b, stride_y, num_pts, min_pts, cap_pts, n_fixed,
                                                   counts, w.cand_count, w.cand, w.cand_cap, w.sortbuf, w.sort_cap,
                                                   w.status, n_out, idx_out, kps_out, kpscore_out,
                                                   (mode & POSFEAT_DETECT_FULLMAP) ? 1 : 0,
                                                   getenv("POSFEAT_SELECT_DEBUG") ? 1 : 0);
  prof_end(PROF_SELECT, stream);
  PF_LAUNCH_CHECK("select_kernel");
  ProfScope prof(PROF_KPOUT, stream);
  keypoint_outputs_kernel<<<dim3((cap_pts + 255) / 256, B), 256, 0, stream>>>(
      score, H, W, stride_b, stride_y, cap_pts, w.status, n_out, idx_out, kps_out, kpscore_out,
      (mode & POSFEAT_DETECT_FULLMAP) ? 1 : 0);
  PF_LAUNCH_CHECK("keypoint_outputs_kernel");
  return POSFEAT_OK;
}

extern "C" int posfeat_detect_topk_f32(const float* score, int B, int H, int W, int64_t stride_b, int64_t stride_y,
                                       int nms_mode, int radius, int thr_mode, float thr, int num_pts, int min_pts,
                                       int cap_pts, int32_t* counts, int32_t* n_out, int64_t* idx_out, float* kps_out,
                                       float* kpscore_out, void* workspace, size_t ws_bytes, void* stream) {
  const int pad = (nms_mode & POSFEAT_DETECT_FULLMAP) ? 2 : 0;
  size_t need = posfeat_detect_workspace_bytes(B, H + pad, W + pad, cap_pts);
  if (ws_bytes < need) return set_error(POSFEAT_EWORKSPACE, "detect workspace: need %zu bytes, got %zu", need, ws_bytes);
  int e = posfeat_detect_candidates_f32(score, B, H, W, stride_b, stride_y, nms_mode, radius, thr_mode, thr, counts,
                                        workspace, ws_bytes, stream);
  if (e) return e;
  return posfeat_detect_select_f32(score, B, H, W, stride_b, stride_y, nms_mode, num_pts, min_pts, cap_pts, -1, counts, n_out,
                                   idx_out, kps_out, kpscore_out, workspace, ws_bytes, stream);
}

// Reads the device status flag of the last select (host sync on `stream`).
extern "C" int posfeat_detect_status(void* workspace, int B, int H, int W, int cap_pts, void* stream_) {
  return posfeat_detect_finish(workspace, B, H, W, cap_pts, nullptr, nullptr, stream_);
}

// Status and, optionally, the keypoint count n (device int32 written by the select kernel) in ONE host
// round trip: both copies are queued before the single stream synchronisation.
extern "C" int posfeat_detect_finish(void* workspace, int B, int H, int W, int cap_pts, const int32_t* n_dev,
                                     int32_t* n_host, void* stream_) {
  DetectWs w = carve(workspace, B, H, W, cap_pts);
  int32_t st = 0;
  PF_CUDA(cudaMemcpyAsync(&st, w.status, sizeof(st), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
  if (n_dev && n_host) PF_CUDA(cudaMemcpyAsync(n_host, n_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
  PF_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  if (st == 0) return POSFEAT_OK;
  const char* why = st == 1 ? "n exceeds cap_pts" : st == 2 ? "n exceeds the number of interior pixels (topk k out of range)"
                    : st == 3 ? "sort buffer too small" : "too many filler rows";
  return set_error(POSFEAT_EINVAL, "detect select failed on device: %s", why);
}
