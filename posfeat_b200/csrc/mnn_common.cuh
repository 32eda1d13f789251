// Declarations shared by the matcher translation units.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace posfeat {

// per-matrix maxima kept by the tensor-core matcher (float bit patterns of non-negative values)
struct __align__(16) MatStats {
  unsigned max_norm;      // max_j |y_j|
  unsigned max_norm_bf;   // max_j |y~_j|
  unsigned max_err;       // max_j |y~_j - y_j|
  unsigned pad;
};

// Where a producer of descriptor rows (the sampler) leaves the tensor-core matcher's operands, so that
// the matcher's own rounding pass (tc_prep_kernel) can be skipped.  All pointers lie inside the
// workspace of posfeat_mnn_batched_f32 for the same (P, N, M); pair p = images (2p, 2p+1).
struct PrepSink {
  __nv_bfloat16 *Ab, *Bb;                // [pairs][Np|Mp][128] bf16
  float *anorm, *aerr, *bnorm, *berr;    // [pairs][Np|Mp]
  MatStats* stats;                       // [pairs][2]
  int Np, Mp;
};
int tc_prep_sink(void* ws, size_t ws_bytes, int P, int N, int M, PrepSink* sink, cudaStream_t stream);

// One descriptor row (D == 128: float4 per lane, whole warp) -> bf16 operand row, |x|, |x~ - x| and the
// running maxima.  Same arithmetic as tc_prep_kernel.
__device__ __forceinline__ void prep_sink_row(const PrepSink& s, int pair, bool second, int row, int lane, float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  const float2 rl = __bfloat1622float2(lo), rh = __bfloat1622float2(hi);
  float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  float sb = rl.x * rl.x + rl.y * rl.y + rh.x * rh.x + rh.y * rh.y;
  const float ex = rl.x - v.x, ey = rl.y - v.y, ez = rh.x - v.z, ew = rh.y - v.w;
  float se = ex * ex + ey * ey + ez * ez + ew * ew;
  ss = warp_sum(ss); sb = warp_sum(sb); se = warp_sum(se);
  const float up = 1.000001f;   // never under-estimate a norm (float32 rounding of the sums)
  const float nrm = sqrtf(ss) * up, nrb = sqrtf(sb) * up, nre = sqrtf(se) * up;
  uint2 pk;
  pk.x = *reinterpret_cast<const unsigned*>(&lo);
  pk.y = *reinterpret_cast<const unsigned*>(&hi);
  const size_t r = (size_t)pair * (second ? s.Mp : s.Np) + row;
  *reinterpret_cast<uint2*>((second ? s.Bb : s.Ab) + r * 128 + lane * 4) = pk;
  if (lane == 0) {
    (second ? s.bnorm : s.anorm)[r] = nrm;
    (second ? s.berr : s.aerr)[r] = nre;
  }
  if (lane < 3) {
    // guarded maxima: after the first few rows almost every warp only reads.  The read goes through L1
    // (a stale value only means one atomic more), so it does not add an L2 round trip to every row.
    const float m = lane == 0 ? nrm : (lane == 1 ? nrb : nre);
    unsigned* dst = &s.stats[2 * pair + (second ? 1 : 0)].max_norm + lane;
    unsigned cur;
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(cur) : "l"(dst));
    if (__float_as_uint(m) > cur) atomicMax(dst, __float_as_uint(m));
  }
}

size_t simt_workspace_bytes(int N, int M);
int mnn_simt(const float* A, int N, int64_t lda, const float* Bm, int M, int64_t ldb, int D, int32_t* nn12,
             int32_t* nn21, void* ws, cudaStream_t stream);
int run_rowbest_simt(const float* X, int NX, int64_t ldx, const float* Y, int NY, int64_t ldy, int D,
                     int32_t* nn, float* top2, void* ws, cudaStream_t stream);
// both directions, optionally with the top-2 similarities per row ([rows][2] float32) for the ratio test
int mnn_simt_top2(const float* A, int N, int64_t lda, const float* Bm, int M, int64_t ldb, int D, int32_t* nn12,
                  int32_t* nn21, float* top12, float* top21, void* ws, cudaStream_t stream);
int launch_ratio_flags(const int32_t* nn12, const int32_t* nn21, const float* top12, const float* top21, int N, int M,
                       float ratio, int mutual, unsigned char* flags, cudaStream_t stream);
int launch_compact_flags(const int32_t* nn12, const unsigned char* flags, int P, int N, int64_t* matches,
                         int32_t* n_matches, cudaStream_t stream);
int launch_mutual_compact(const int32_t* nn12, const int32_t* nn21, int N, int M, int64_t* matches,
                          int32_t* n_matches, cudaStream_t stream);

// tensor-core path (mnn_tc.cu)
bool tc_supported(int N, int M, int D);
size_t tc_workspace_bytes(int P, int N, int M);
// batched over P pairs: pair p uses A + p*strideA, Bm + p*strideB (elements), nn12 + p*N, nn21 + p*M
// nn21 may be NULL (matches only); the call also produces matches / n_matches
int mnn_tc(const float* A, int64_t strideA, int N, int64_t lda, const float* Bm, int64_t strideB, int M, int64_t ldb,
           int D, int P, int32_t* nn12, int32_t* nn21, int64_t* matches, int32_t* n_matches, void* ws, size_t ws_bytes,
           cudaStream_t stream, float* top12 = nullptr, float* top21 = nullptr, bool prepared = false);
int launch_mutual_compact_batched(const int32_t* nn12, const int32_t* nn21, int P, int N, int M, int64_t* matches,
                                  int32_t* n_matches, cudaStream_t stream);

}  // namespace posfeat
