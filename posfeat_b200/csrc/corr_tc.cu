// Subsystem (4), tensor-core path: dense correlation + softmax expectation on tcgen05 (sm_100a).
//
//     out[b,i,:] = sum_j softmax_j(scale * <q_i, k_j>) * v[j,:]        lse[b,i] = logsumexp_j(...)
// (get_expected_correspondence_locs, losses/preprocess_utils.py:55-82; the grid<->grid stage of
// Preprocess_Line2Window.forward, losses/preprocess.py:59-81), float32 results.
//
// The logits must be float32-accurate (the north star asks for 1e-5 relative on the expectation), so
// the contraction runs as split TF32 ("3xTF32"): every operand x is stored as hi = tf32(x) (round to
// nearest) and lo = x - hi, and  <q,k> ~= qh.kh + qh.kl + ql.kh  is accumulated in float32 in TMEM;
// the dropped ql.kl term and the rounding of lo are O(2^-22) relative to |q||k|.
//
// One CTA = 128 query rows x a range of 128-key tiles:
//   warp 0   TMA producer: Q (hi, lo; all of D) once, then per key tile and per 32-column slice of D
//            a [128 keys x 128 B] box of K hi and K lo into a two-stage ring (SWIZZLE_128B)
//   warp 1   MMA issuer: per slice 4 K-steps x 3 split terms of tcgen05.mma kind::tf32 (M=128,
//            N=128, K=8) into one of two TMEM accumulator stages
//   warp 2   TMEM allocation
//   warps 4-11  epilogue: thread <-> query row (tcgen05.ld 32x32b), two warps per TMEM lane quarter
//            split the tile's columns; online softmax in the log2 domain, the C <= 4 value columns
//            are accumulated in registers -- the [B,n,m] probability tensor never exists
// Key ranges are split over CTAs so that the grid fills the chip; a small merge kernel combines the
// per-split (max, sum, weighted sums).
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace posfeat {

constexpr int kCtM = 128;            // query rows per CTA (MMA M)
constexpr int kCtN = 128;            // keys per tile (MMA N)
constexpr int kCtAtomBytes = 128 * 128;     // [128 rows x 32 tf32] swizzled slice
constexpr int kCtStages = 3;         // K ring depth (each stage: hi + lo slice = 32 KB)
constexpr int kCtEpiWarps = 8;
constexpr int kCtThreads = (4 + kCtEpiWarps) * 32;
constexpr int kCtMaxAtoms = 4;       // D <= 128

constexpr int kCtSmemQ = 0;                                          // [2 hi/lo][4 slices][16 KB]
constexpr int kCtSmemK = 2 * kCtMaxAtoms * kCtAtomBytes;             // [stages][2 hi/lo][16 KB]
constexpr int kCtSmemMerge = kCtSmemK;    // [128 rows][8 floats], reuses ring stage 0 once every MMA has retired
constexpr int kCtSmemBar = kCtSmemK + kCtStages * 2 * kCtAtomBytes;
constexpr int kCtSmemTotal = kCtSmemBar + 128;
static_assert(kCtSmemTotal + 1024 <= 232448, "shared memory budget");
constexpr int kCtSmemAlloc = kCtSmemTotal + 1024;

// kind::tf32 instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=128
constexpr uint32_t kCtIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kCtN >> 3) << 17) | ((uint32_t)(kCtM >> 4) << 24);

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct CorrTcParams {
  int B, n, m, D, C;
  int n_pad, m_pad;          // rows per batch in the split buffers (multiples of 128)
  int atoms;                 // ceil(D / 32)
  int q_tiles, k_tiles, splits, tiles_per_split;
  float sl2;                 // scale * log2(e)
  const float4* v4;          // [v_batched ? B : 1][m_pad] value rows padded to 4 columns
  int v_batched;
  float* part;               // [B][q_tiles][splits][128][8]: m, l, acc0..3
  // kMode 1 (backward, W materialisation): W = P o (g.v - g.out), written as tf32 hi/lo, plain and transposed
  const float* g;            // [B][n][C]
  const float* o;            // [B][n][C] forward output
  const float* lse;          // [B][n]
  float *w_hi, *w_lo;        // [B][n_pad][m_pad]
  float *wt_hi, *wt_lo;      // [B][m_pad][n_pad]
  // kMode 2 (DiskLoss dual-softmax reward sums, losses/kploss.py:158-182)
  const float4* rowtab;      // [B][n_pad][2]: {lse, la, lb, lc}, {x, y, logp, accept} of the row side
  const float4* coltab;      // [B][m_pad][2]: the same for the column side
  float temp;                // affinity = temp * s - temp
  float thr_own, thr_other;  // reward thresholds on |own line . other point| and |other line . own point|
  float good_reward, bad_reward;
  int dynamic_reward;        // 0: constant_reward, 1: dynamic_reward
  float4* part4;             // [B][q_tiles][splits][128]: {reinforce, reward-weighted p, sum p, max p}
};

// ---- split: x -> hi = tf32(x), lo = x - hi, zero padded to [rows_pad][Dp] per batch (Dp = 32 * slices).
// One thread per float4 of a padded row; a warp covers one 128-column row (or 4 rows of 32 columns).
__device__ __forceinline__ void corr_split_body(const float* __restrict__ src, int B, int rows, int D, int rows_pad, int Dp,
                                                float* __restrict__ hi, float* __restrict__ lo, int blk, int nblk) {
  const int vpr = Dp >> 2;                                   // float4 per padded row
  const int64_t total = (int64_t)B * rows_pad * vpr;
  const bool vec_ok = (D & 3) == 0 && ((uintptr_t)src & 15) == 0;
  for (int64_t i = (int64_t)blk * blockDim.x + threadIdx.x; i < total; i += (int64_t)nblk * blockDim.x) {
    const int cv = (int)(i % vpr);
    const int64_t rr = i / vpr;
    const int r = (int)(rr % rows_pad), b = (int)(rr / rows_pad);
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < rows) {
      const float* sp = src + ((int64_t)b * rows + r) * D + cv * 4;
      if (vec_ok && cv * 4 + 3 < D) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(sp));
        x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (cv * 4 + j < D) x[j] = __ldg(sp + j);
      }
    }
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t hb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x[j]));
      h[j] = __uint_as_float(hb);
      l[j] = x[j] - h[j];
    }
    reinterpret_cast<float4*>(hi)[i] = make_float4(h[0], h[1], h[2], h[3]);
    reinterpret_cast<float4*>(lo)[i] = make_float4(l[0], l[1], l[2], l[3]);
  }
}

__global__ void corr_split_kernel(const float* __restrict__ src, int B, int rows, int D, int rows_pad, int Dp,
                                  float* __restrict__ hi, float* __restrict__ lo) {
  corr_split_body(src, B, rows, D, rows_pad, Dp, hi, lo, blockIdx.x, gridDim.x);
}

__device__ __forceinline__ void corr_padv_body(const float* __restrict__ v, int vb, int m, int C, int m_pad,
                                               float4* __restrict__ v4, int blk, int nblk) {
  const int64_t total = (int64_t)vb * m_pad;
  for (int64_t i = (int64_t)blk * blockDim.x + threadIdx.x; i < total; i += (int64_t)nblk * blockDim.x) {
    const int r = (int)(i % m_pad), b = (int)(i / m_pad);
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < m)
      for (int c = 0; c < C; ++c) t[c] = __ldg(v + ((int64_t)b * m + r) * C + c);
    v4[i] = make_float4(t[0], t[1], t[2], t[3]);
  }
}

__global__ void corr_padv_kernel(const float* __restrict__ v, int vb, int m, int C, int m_pad, float4* __restrict__ v4) {
  corr_padv_body(v, vb, m, C, m_pad, v4, blockIdx.x, gridDim.x);
}

// forward operand preparation in ONE launch (the small dense problems -- coarse 30x40 maps -- are launch bound):
// blocks [0, gq) split q, [gq, gq + gk) split k, the rest pad the value table
struct CorrPrepArgs {
  const float *q, *k, *v;
  float *qh, *ql, *kh, *kl;
  float4* v4;
  int B, n, m, D, n_pad, m_pad, Dp, vb, C, gq, gk, gv;
};
__global__ void __launch_bounds__(256) corr_prep_fwd_kernel(const CorrPrepArgs a) {
  const int b = blockIdx.x;
  if (b < a.gq) corr_split_body(a.q, a.B, a.n, a.D, a.n_pad, a.Dp, a.qh, a.ql, b, a.gq);
  else if (b < a.gq + a.gk) corr_split_body(a.k, a.B, a.m, a.D, a.m_pad, a.Dp, a.kh, a.kl, b - a.gq, a.gk);
  else corr_padv_body(a.v, a.vb, a.m, a.C, a.m_pad, a.v4, b - a.gq - a.gk, a.gv);
}

// kMode 0: forward (online softmax expectation).  kMode 1: backward, first step -- the same contraction, but
// the epilogue turns every logit into W_ij = exp(scale s_ij - lse_i) (g_i . v_j - g_i . out_i) and stores it
// (tf32 hi/lo, row-major and transposed) as the operand of the two gradient GEMMs.
template <int kMode>
__global__ void __launch_bounds__(kCtThreads, 1)
corr_tc_fwd_kernel(const __grid_constant__ CUtensorMap mapQh, const __grid_constant__ CUtensorMap mapQl,
                   const __grid_constant__ CUtensorMap mapKh, const __grid_constant__ CUtensorMap mapKl,
                   const CorrTcParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar0 = sbase + kCtSmemBar;
  const uint32_t bar_q_full = bar0;
  const uint32_t bar_k_full = bar0 + 8, bar_k_empty = bar_k_full + 8 * kCtStages;
  const uint32_t bar_acc_full = bar_k_empty + 8 * kCtStages, bar_acc_empty = bar_acc_full + 16;
  const uint32_t tmem_slot = bar_acc_empty + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - sbase));
  float* merge = reinterpret_cast<float*>(smem_gen + kCtSmemMerge);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, qt = blockIdx.y, b = blockIdx.z;
  const int t0 = split * p.tiles_per_split, t1 = min(p.k_tiles, t0 + p.tiles_per_split);
  const int atoms = p.atoms;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapKh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapKl) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_q_full, 1);
    for (int s = 0; s < kCtStages; ++s) {
      mbar_init(bar_k_full + 8 * s, 1);
      mbar_init(bar_k_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, kCtEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (t0 < t1) {
    if (warp == 0 && lane == 0) {
      // ===== TMA producer =====
      const int qrow = b * p.n_pad + qt * kCtM;
      mbar_expect_tx(bar_q_full, (uint32_t)(2 * atoms * kCtAtomBytes));
      for (int a = 0; a < atoms; ++a) {
        tma_load_2d(sbase + kCtSmemQ + a * kCtAtomBytes, &mapQh, a * 32, qrow, bar_q_full);
        tma_load_2d(sbase + kCtSmemQ + (kCtMaxAtoms + a) * kCtAtomBytes, &mapQl, a * 32, qrow, bar_q_full);
      }
      int s = 0, ph = 0;
      for (int t = t0; t < t1; ++t) {
        const int krow = b * p.m_pad + t * kCtN;
        for (int a = 0; a < atoms; ++a) {
          mbar_wait(bar_k_empty + 8 * s, ph ^ 1);
          mbar_expect_tx(bar_k_full + 8 * s, 2 * kCtAtomBytes);
          tma_load_2d(sbase + kCtSmemK + (s * 2 + 0) * kCtAtomBytes, &mapKh, a * 32, krow, bar_k_full + 8 * s);
          tma_load_2d(sbase + kCtSmemK + (s * 2 + 1) * kCtAtomBytes, &mapKl, a * 32, krow, bar_k_full + 8 * s);
          if (++s == kCtStages) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===== MMA issuer =====
      mbar_wait(bar_q_full, 0);
      int s = 0, ph = 0, as = 0, aph = 0;
      for (int t = t0; t < t1; ++t) {
        mbar_wait(bar_acc_empty + 8 * as, aph ^ 1);
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kCtN);
        for (int a = 0; a < atoms; ++a) {
          mbar_wait(bar_k_full + 8 * s, ph);
          tc_fence_after();
          const uint64_t qh = make_sw128_desc(sbase + kCtSmemQ + a * kCtAtomBytes);
          const uint64_t ql = make_sw128_desc(sbase + kCtSmemQ + (kCtMaxAtoms + a) * kCtAtomBytes);
          const uint64_t kh = make_sw128_desc(sbase + kCtSmemK + (s * 2 + 0) * kCtAtomBytes);
          const uint64_t kl = make_sw128_desc(sbase + kCtSmemK + (s * 2 + 1) * kCtAtomBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {      // 4 K-steps of 8 tf32 (32 bytes) inside the 128-byte slice
            const uint64_t o = (uint64_t)(k * 2);
            tc_mma_tf32(d_tmem, ql + o, kh + o, kCtIdesc, (a | k) ? 1u : 0u);   // small terms first
            tc_mma_tf32(d_tmem, qh + o, kl + o, kCtIdesc, 1u);
            tc_mma_tf32(d_tmem, qh + o, kh + o, kCtIdesc, 1u);
          }
          tc_commit(bar_k_empty + 8 * s);
          if (++s == kCtStages) { s = 0; ph ^= 1; }
        }
        tc_commit(bar_acc_full + 8 * as);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    } else if (warp >= 4) {
      // ===== epilogue: thread <-> query row; the two warps of a lane quarter split the columns =====
      const int e = warp - 4, quarter = warp & 3, half = e >> 2;
      const int row = quarter * 32 + lane;
      const float4* vb = p.v4 + (p.v_batched ? (size_t)b * p.m_pad : 0);
      const float sl2 = p.sl2;
      const int m = p.m;
      if (kMode == 2) {
        // ---- DiskLoss: p = softmax_row(A) * softmax_col(A), A = temp * s - temp; per row the sums
        //      sum_j acc_ij r_ij p_ij (log p_ij + logp_i + logp_j),  sum_j acc_ij r_ij p_ij,  sum_j p_ij,  max_j p_ij
        const int qi = qt * kCtM + row;
        const float4* rt = p.rowtab + ((size_t)b * p.n_pad + qi) * 2;
        const float4 r0 = __ldg(rt), r1 = __ldg(rt + 1);       // padding rows carry lse = +inf -> p = 0
        const float4* ct = p.coltab + (size_t)b * p.m_pad * 2;
        const float temp = p.temp;
        float reinf = 0.f, rsum = 0.f, psum = 0.f, pmax = 0.f;
        int as = 0, aph = 0;
        for (int t = t0; t < t1; ++t) {
          mbar_wait(bar_acc_full + 8 * as, aph);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (uint32_t)(as * kCtN + half * 64) + ((uint32_t)(quarter * 32) << 16);
          const int c0 = t * kCtN + half * 64;
#pragma unroll 1
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t v[32];
            tc_ld32(taddr + ch * 32, v);
            tc_wait_ld();
            if (ch == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
            }
            const float4* cp = ct + (size_t)(c0 + ch * 32) * 2;
#pragma unroll
            for (int j = 0; j < 32; ++j) {       // fully unrolled: v[] must stay in registers
              const float4 q0 = __ldg(cp + 2 * j), q1 = __ldg(cp + 2 * j + 1);   // broadcast loads
              const float a = fmaf(__uint_as_float(v[j]), temp, -temp);
              const float lp = (a - r0.x) + (a - q0.x);                          // log p_ij
              const float pj = ex2_approx(lp * 1.4426950408889634f);
              const float d1 = fabsf(__fadd_rn(__fmaf_rn(r0.z, q1.y, __fmul_rn(r0.y, q1.x)), r0.w));   // own line . other point
              const float d2 = fabsf(__fadd_rn(__fmaf_rn(q0.z, r1.y, __fmul_rn(q0.y, r1.x)), q0.w));   // other line . own point
              float rw;
              if (p.dynamic_reward)
                rw = fmaxf(__expf(-d1 / p.thr_own) + __expf(-d2 / p.thr_other) - 0.7357588823428847f, p.bad_reward);
              else
                rw = (d1 < p.thr_own && d2 < p.thr_other) ? p.good_reward : p.bad_reward;
              if (r1.w != 0.f && q1.w != 0.f) {          // accepted pair (padding columns carry accept = 0, lp = -inf)
                const float wv = rw * pj;
                reinf = fmaf(wv, lp + r1.z + q1.z, reinf);
                rsum += wv;
              }
              psum += pj;
              pmax = fmaxf(pmax, pj);
            }
          }
          if (++as == 2) { as = 0; aph ^= 1; }
        }
        if (half == 1) {
          float* d = merge + row * 8;
          d[0] = reinf; d[1] = rsum; d[2] = psum; d[3] = pmax;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kCtEpiWarps * 32) : "memory");
        if (half == 0) {
          const float* d = merge + row * 8;
          p.part4[(((size_t)b * p.q_tiles + qt) * p.splits + split) * kCtM + row] =
              make_float4(reinf + d[0], rsum + d[1], psum + d[2], fmaxf(pmax, d[3]));
        }
      } else if (kMode == 1) {
        // ---- backward: W tiles ----
        const int qi = qt * kCtM + row;                      // query index inside the batch
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, go = 0.f, lse2 = INFINITY;   // padding rows: W = 0
        if (qi < p.n) {
          const float* gp = p.g + ((size_t)b * p.n + qi) * p.C;
          const float* op = p.o + ((size_t)b * p.n + qi) * p.C;
          float gg[4] = {0.f, 0.f, 0.f, 0.f};
          for (int c = 0; c < p.C; ++c) { gg[c] = __ldg(gp + c); go = fmaf(gg[c], __ldg(op + c), go); }
          g0 = gg[0]; g1 = gg[1]; g2 = gg[2]; g3 = gg[3];
          lse2 = __ldg(p.lse + (size_t)b * p.n + qi) * 1.4426950408889634f;
        }
        float* wrow_hi = p.w_hi + ((size_t)b * p.n_pad + qi) * p.m_pad;
        float* wrow_lo = p.w_lo + ((size_t)b * p.n_pad + qi) * p.m_pad;
        int as = 0, aph = 0;
        for (int t = t0; t < t1; ++t) {
          mbar_wait(bar_acc_full + 8 * as, aph);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (uint32_t)(as * kCtN + half * 64) + ((uint32_t)(quarter * 32) << 16);
          const int c0 = t * kCtN + half * 64;
#pragma unroll 1
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t v[32];
            tc_ld32(taddr + ch * 32, v);
            tc_wait_ld();
            if (ch == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
            }
            const int cb = c0 + ch * 32;
            const float4* vp = vb + cb;
            float* th = p.wt_hi + ((size_t)b * p.m_pad + cb) * p.n_pad + qi;
            float* tl = p.wt_lo + ((size_t)b * p.m_pad + cb) * p.n_pad + qi;
            float hi[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float pj = ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -lse2));
              const float4 vv = __ldg(vp + j);
              const float dp = fmaf(g3, vv.w, fmaf(g2, vv.z, fmaf(g1, vv.y, g0 * vv.x))) - go;
              const float wv = pj * dp;
              uint32_t hb;
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(wv));
              hi[j] = __uint_as_float(hb);
              const float lo = wv - hi[j];
              v[j] = __float_as_uint(lo);
              th[(size_t)j * p.n_pad] = hi[j];           // transposed copy: the 32 lanes (rows) write 128 contiguous bytes
              tl[(size_t)j * p.n_pad] = lo;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              *reinterpret_cast<float4*>(wrow_hi + cb + j) = make_float4(hi[j], hi[j + 1], hi[j + 2], hi[j + 3]);
              *reinterpret_cast<float4*>(wrow_lo + cb + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                         __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
          }
          if (++as == 2) { as = 0; aph ^= 1; }
        }
      } else {
      float mrun = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int as = 0, aph = 0;
      for (int t = t0; t < t1; ++t) {
        mbar_wait(bar_acc_full + 8 * as, aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(as * kCtN + half * 64) + ((uint32_t)(quarter * 32) << 16);
        const int c0 = t * kCtN + half * 64;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t v[32];
          tc_ld32(taddr + ch * 32, v);
          tc_wait_ld();
          if (ch == 1) {     // accumulator slice fully read: hand the TMEM stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
          }
          const int cb = c0 + ch * 32;
          float s2[32];
          float cm = -INFINITY;
          if (cb + 32 <= m) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { s2[j] = __uint_as_float(v[j]) * sl2; cm = fmaxf(cm, s2[j]); }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              s2[j] = (cb + j < m) ? __uint_as_float(v[j]) * sl2 : -INFINITY;
              cm = fmaxf(cm, s2[j]);
            }
          }
          const float mnew = fmaxf(mrun, cm);
          const float msafe = mnew == -INFINITY ? 0.f : mnew;
          const float corr = ex2_approx(mrun - msafe);      // mrun = -inf -> 0
          l *= corr; a0 *= corr; a1 *= corr; a2 *= corr; a3 *= corr;
          const float4* vp = vb + cb;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float pj = ex2_approx(s2[j] - msafe);
            const float4 vv = __ldg(vp + j);                // same address in every lane: one broadcast load
            l += pj;
            a0 = fmaf(pj, vv.x, a0); a1 = fmaf(pj, vv.y, a1); a2 = fmaf(pj, vv.z, a2); a3 = fmaf(pj, vv.w, a3);
          }
          mrun = mnew;
        }
        if (++as == 2) { as = 0; aph ^= 1; }
      }
      // merge the two column halves of a row, then write the split's partial result
      if (half == 1) {
        float* d = merge + row * 8;
        d[0] = mrun; d[1] = l; d[2] = a0; d[3] = a1; d[4] = a2; d[5] = a3;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kCtEpiWarps * 32) : "memory");
      if (half == 0) {
        const float* d = merge + row * 8;
        const float mo = d[0];
        const float mm = fmaxf(mrun, mo);
        const float ms = mm == -INFINITY ? 0.f : mm;
        const float ca = ex2_approx(mrun - ms), cbb = ex2_approx(mo - ms);
        float* out = p.part + ((((size_t)b * p.q_tiles + qt) * p.splits + split) * kCtM + row) * 8;
        reinterpret_cast<float4*>(out)[0] = make_float4(mm, l * ca + d[1] * cbb, a0 * ca + d[2] * cbb, a1 * ca + d[3] * cbb);
        reinterpret_cast<float4*>(out)[1] = make_float4(a2 * ca + d[4] * cbb, a3 * ca + d[5] * cbb, 0.f, 0.f);
      }
      }   // kMode == 0
    }
  } else if (kMode == 0 && warp >= 4 && warp < 8) {
    // a split without tiles (cannot happen with the host's split choice, kept for safety): neutral partial
    const int row = (warp & 3) * 32 + lane;
    float* out = p.part + ((((size_t)b * p.q_tiles + qt) * p.splits + split) * kCtM + row) * 8;
    reinterpret_cast<float4*>(out)[0] = make_float4(-INFINITY, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(out)[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// out = (sum_s acc_s 2^(m_s - M)) / (sum_s l_s 2^(m_s - M)),  lse = (M + log2 L) ln 2
__global__ void corr_tc_merge_kernel(const CorrTcParams p, float* __restrict__ out, float* __restrict__ lse) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)p.B * p.n) return;
  const int b = (int)(i / p.n), r = (int)(i % p.n);
  const int qt = r / kCtM, row = r % kCtM;
  const float* base = p.part + (((size_t)b * p.q_tiles + qt) * p.splits * kCtM + row) * 8;
  float M = -INFINITY;
  for (int s = 0; s < p.splits; ++s) M = fmaxf(M, base[(size_t)s * kCtM * 8]);
  float L = 0.f, a[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < p.splits; ++s) {
    const float* d = base + (size_t)s * kCtM * 8;
    const float w = exp2f(d[0] - M);
    L = fmaf(d[1], w, L);
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c] = fmaf(d[2 + c], w, a[c]);
  }
  const float inv = 1.f / L;
  for (int c = 0; c < p.C; ++c) out[i * p.C + c] = a[c] * inv;
  lse[i] = (M + log2f(L)) * 0.6931471805599453f;
}

// ---- transposed split: src [B][rows][D] -> hiT / loT [B][Dp][rows_pad] (rows contiguous), zero padded.
// 32 x 32 tiles through shared memory so that both the reads (along D) and the writes (along rows) coalesce.
__global__ void corr_split_t_kernel(const float* __restrict__ src, int rows, int D, int rows_pad, int Dp,
                                    float* __restrict__ hiT, float* __restrict__ loT) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 256 threads: 8 rows per pass
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + tx;
    tile[k][tx] = (r < rows && c < D) ? __ldg(src + ((size_t)b * rows + r) * D + c) : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + tx;                              // output row c (a channel), column r
    if (c < Dp && r < rows_pad) {
      const float x = tile[tx][k];
      uint32_t hb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
      const float h = __uint_as_float(hb);
      const size_t o = ((size_t)b * Dp + c) * rows_pad + r;
      hiT[o] = h;
      loT[o] = x - h;
    }
  }
}

// ---- split-K 3xTF32 GEMM for the gradients:  C[b][M][N] (+)= scale * A[b][M][K] * Bt[b][N][K]^T
// A, Bt are pre-split into tf32 hi/lo and K-major (K contiguous); N = Dp <= 128.  One CTA = one 128-row tile of C
// and a contiguous range of 32-wide K slices; the accumulator lives in TMEM for the whole range and is written
// once (to the output, or to a per-split partial that corr_tc_reduce_kernel sums).
constexpr int kGmStages = 3;
constexpr int kGmStageBytes = 4 * kCtAtomBytes;                    // A hi, A lo, B hi, B lo
constexpr int kGmSmemBar = kGmStages * kGmStageBytes;
constexpr int kGmSmemAlloc = kGmSmemBar + 128 + 1024;
constexpr int kGmThreads = 8 * 32;                                 // warps 0,1,2: TMA, MMA, TMEM; warps 4-7: epilogue

struct GemmParams {
  int M_pad;             // rows of A (and C) per batch, multiple of 128
  int N;                 // Dp
  int m_valid;           // rows of C that exist
  int n_valid;           // columns of C that exist (D)
  int slices;            // K / 32
  int splits, slices_per_split;
  float scale;
  float* out;            // splits == 1: [B][m_valid][n_valid];  else partials [splits][B][M_pad][N]
  int B;
};

__global__ void __launch_bounds__(kGmThreads, 1)
corr_tc_gemm_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                    const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                    const GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar0 = sbase + kGmSmemBar;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * kGmStages, bar_acc = bar_empty + 8 * kGmStages;
  const uint32_t tmem_slot = bar_acc + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - sbase));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, mt = blockIdx.y, b = blockIdx.z;
  const int s0 = split * p.slices_per_split, s1 = min(p.slices, s0 + p.slices_per_split);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kGmStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t b_bytes = (uint32_t)p.N * 128u;                  // one B slice: N rows x 128 bytes

  if (warp == 0 && lane == 0) {
    int st = 0, ph = 0;
    const int arow = b * p.M_pad + mt * 128, brow = b * p.N;
    for (int sl = s0; sl < s1; ++sl) {
      mbar_wait(bar_empty + 8 * st, ph ^ 1);
      mbar_expect_tx(bar_full + 8 * st, 2 * kCtAtomBytes + 2 * b_bytes);
      const uint32_t base = sbase + st * kGmStageBytes;
      tma_load_2d(base, &mapAh, sl * 32, arow, bar_full + 8 * st);
      tma_load_2d(base + kCtAtomBytes, &mapAl, sl * 32, arow, bar_full + 8 * st);
      tma_load_2d(base + 2 * kCtAtomBytes, &mapBh, sl * 32, brow, bar_full + 8 * st);
      tma_load_2d(base + 3 * kCtAtomBytes, &mapBl, sl * 32, brow, bar_full + 8 * st);
      if (++st == kGmStages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int st = 0, ph = 0;
    for (int sl = s0; sl < s1; ++sl) {
      mbar_wait(bar_full + 8 * st, ph);
      tc_fence_after();
      const uint32_t base = sbase + st * kGmStageBytes;
      const uint64_t ah = make_sw128_desc(base), al = make_sw128_desc(base + kCtAtomBytes);
      const uint64_t bh = make_sw128_desc(base + 2 * kCtAtomBytes), bl = make_sw128_desc(base + 3 * kCtAtomBytes);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t o = (uint64_t)(k * 2);
        tc_mma_tf32(tmem_base, al + o, bh + o, idesc, (sl > s0 || k) ? 1u : 0u);
        tc_mma_tf32(tmem_base, ah + o, bl + o, idesc, 1u);
        tc_mma_tf32(tmem_base, ah + o, bh + o, idesc, 1u);
      }
      tc_commit(bar_empty + 8 * st);
      if (++st == kGmStages) { st = 0; ph ^= 1; }
    }
    tc_commit(bar_acc);
  } else if (warp >= 4) {
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const int r = mt * 128 + row;
    if (s0 < s1) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
    }
    for (int c0 = 0; c0 < p.N; c0 += 32) {
      uint32_t v[32];
      if (s0 < s1) {
        tc_ld32(tmem_base + (uint32_t)c0 + ((uint32_t)(quarter * 32) << 16), v);
        tc_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (p.splits == 1) {
        if (r < p.m_valid) {
          float* dst = p.out + ((size_t)b * p.m_valid + r) * p.n_valid + c0;
          if ((p.n_valid & 3) == 0 && c0 + 32 <= p.n_valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]) * p.scale, __uint_as_float(v[j + 1]) * p.scale,
                                                                __uint_as_float(v[j + 2]) * p.scale, __uint_as_float(v[j + 3]) * p.scale);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < p.n_valid) dst[j] = __uint_as_float(v[j]) * p.scale;
          }
        }
      } else {
        float* dst = p.out + ((((size_t)split * p.B + b) * p.M_pad + r) * p.N) + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
  }
}

// out[b][r][c] = scale * sum_s part[s][b][r][c]
__global__ void corr_tc_reduce_kernel(const float* __restrict__ part, int splits, int B, int M_pad, int N, int m_valid,
                                      int n_valid, float scale, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * m_valid * n_valid) return;
  const int c = (int)(i % n_valid);
  const int64_t rr = i / n_valid;
  const int r = (int)(rr % m_valid), b = (int)(rr / m_valid);
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += part[(((size_t)s * B + b) * M_pad + r) * N + c];
  out[i] = acc * scale;
}

// ------------------------------------------------------------------ host side
static int make_map_f32(CUtensorMap* map, void* base, int64_t Dp, int64_t rows, int box_rows = 128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(POSFEAT_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)Dp, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)Dp * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(POSFEAT_ECUDA, "cuTensorMapEncodeTiled (f32) failed with CUresult %d", (int)r);
  return POSFEAT_OK;
}

static inline int pad128(int x) { return (x + 127) / 128 * 128; }

struct CorrTcWs {
  float *qh, *ql, *kh, *kl;
  float4* v4;
  float* part;
  size_t total;
};

static void corr_tc_shape(int B, int n, int m, int* q_tiles, int* k_tiles, int* splits, int* tps) {
  *q_tiles = pad128(n) / 128;
  *k_tiles = pad128(m) / 128;
  const int units = B * *q_tiles;
  // one CTA per SM: choose the number of key splits that wastes the least of the last wave
  // (cost ~ ceil(units * s / 148) / s), preferring fewer splits on ties, at least 4 key tiles per CTA
  int s = 1;
  double best = 1e30;
  for (int c = 1; c <= std::min(*k_tiles, 16); ++c) {
    if (c > 1 && *k_tiles / c < 4) break;
    const double cost = (double)((units * c + 147) / 148) / c;
    if (cost < best - 1e-9) { best = cost; s = c; }
  }
  *tps = (*k_tiles + s - 1) / s;
  *splits = (*k_tiles + *tps - 1) / *tps;          // no empty split
}

static CorrTcWs carve_corr_tc(void* base, int B, int n, int m, int D, int vb) {
  CorrTcWs w{};
  const int Dp = (D + 31) / 32 * 32, np = pad128(n), mp = pad128(m);
  int qt, kt, sp, tps;
  corr_tc_shape(B, n, m, &qt, &kt, &sp, &tps);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (char*)base + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  w.qh = (float*)take(sizeof(float) * (size_t)B * np * Dp);
  w.ql = (float*)take(sizeof(float) * (size_t)B * np * Dp);
  w.kh = (float*)take(sizeof(float) * (size_t)B * mp * Dp);
  w.kl = (float*)take(sizeof(float) * (size_t)B * mp * Dp);
  w.v4 = (float4*)take(sizeof(float4) * (size_t)vb * mp);
  w.part = (float*)take(sizeof(float) * (size_t)B * qt * sp * kCtM * 8);
  w.total = off;
  return w;
}

// The tensor-core path pays a fixed cost (operand split, merge); tiny problems stay on the SIMT kernel.
bool corr_tc_eligible(int B, int n, int m, int D, int C) {
  return D >= 8 && D <= 128 && C >= 1 && C <= 4 && m >= 128 && (int64_t)B * n * m >= (1 << 18);
}

size_t corr_tc_workspace_bytes(int B, int n, int m, int D, int C) {
  if (!corr_tc_eligible(B, n, m, D, C)) return 0;
  return carve_corr_tc(nullptr, B, n, m, D, B).total;
}

int corr_tc_fwd(const float* q, const float* k, const float* v, int v_batched, int B, int n, int m, int D, int C,
                float scale, float* out, float* lse, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const int vb = v_batched ? B : 1;
  CorrTcWs w = carve_corr_tc(ws, B, n, m, D, vb);
  PF_CHECK_ARG(ws && ((uintptr_t)ws & 255) == 0, "corr workspace must be 256-byte aligned");
  if (ws_bytes < w.total) return set_error(POSFEAT_EWORKSPACE, "corr workspace: need %zu bytes, got %zu", w.total, ws_bytes);
  CorrTcParams p{};
  p.B = B; p.n = n; p.m = m; p.D = D; p.C = C;
  p.n_pad = pad128(n); p.m_pad = pad128(m);
  p.atoms = (D + 31) / 32;
  corr_tc_shape(B, n, m, &p.q_tiles, &p.k_tiles, &p.splits, &p.tiles_per_split);
  p.sl2 = scale * 1.4426950408889634f;
  p.v4 = w.v4; p.v_batched = v_batched; p.part = w.part;
  const int Dp = p.atoms * 32;
  ProfScope prof(PROF_CORR_FWD, stream);
  {
    CorrPrepArgs pa{q, k, v, w.qh, w.ql, w.kh, w.kl, w.v4, B, n, m, D, p.n_pad, p.m_pad, Dp, vb, C, 0, 0, 0};
    pa.gq = (int)std::min<int64_t>(((int64_t)B * p.n_pad * (Dp / 4) + 255) / 256, 148 * 16);
    pa.gk = (int)std::min<int64_t>(((int64_t)B * p.m_pad * (Dp / 4) + 255) / 256, 148 * 16);
    pa.gv = (int)std::min<int64_t>(((int64_t)vb * p.m_pad + 255) / 256, 148 * 4);
    corr_prep_fwd_kernel<<<pa.gq + pa.gk + pa.gv, 256, 0, stream>>>(pa);
    PF_LAUNCH_CHECK("corr_prep_fwd_kernel");
  }
  CUtensorMap mqh, mql, mkh, mkl;
  if (int e = make_map_f32(&mqh, w.qh, Dp, (int64_t)B * p.n_pad)) return e;
  if (int e = make_map_f32(&mql, w.ql, Dp, (int64_t)B * p.n_pad)) return e;
  if (int e = make_map_f32(&mkh, w.kh, Dp, (int64_t)B * p.m_pad)) return e;
  if (int e = make_map_f32(&mkl, w.kl, Dp, (int64_t)B * p.m_pad)) return e;
  // per device/context attribute: set on every call (cheap), never cached in a process-global flag
  PF_CUDA(cudaFuncSetAttribute(corr_tc_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtSmemAlloc));
  dim3 grid(p.splits, p.q_tiles, B);
  corr_tc_fwd_kernel<0><<<grid, kCtThreads, kCtSmemAlloc, stream>>>(mqh, mql, mkh, mkl, p);
  PF_LAUNCH_CHECK("corr_tc_fwd_kernel");
  corr_tc_merge_kernel<<<(int)(((int64_t)B * n + 255) / 256), 256, 0, stream>>>(p, out, lse);
  PF_LAUNCH_CHECK("corr_tc_merge_kernel");
  return POSFEAT_OK;
}

// ---------------------------------------------------------------- backward on the tensor cores
struct CorrTcBwdWs {
  float *qh, *ql, *kh, *kl;          // [B][n_pad|m_pad][Dp]
  float *qth, *qtl, *kth, *ktl;      // [B][Dp][n_pad|m_pad]
  float4* v4;
  float *wh, *wl, *wth, *wtl;        // [B][n_pad][m_pad], [B][m_pad][n_pad]
  float* part;                       // split-K partials of the larger of the two GEMMs
  size_t total;
};

static void gemm_shape(int B, int m_tiles, int slices, int* splits, int* sps) {
  const int units = B * m_tiles;
  int s = std::max(1, std::min((2 * 148) / std::max(units, 1), slices));
  *sps = (slices + s - 1) / s;
  *splits = (slices + *sps - 1) / *sps;
}

static CorrTcBwdWs carve_corr_tc_bwd(void* base, int B, int n, int m, int D, int vb) {
  CorrTcBwdWs w{};
  const size_t Dp = (D + 31) / 32 * 32, np = pad128(n), mp = pad128(m);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (char*)base + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  w.qh = (float*)take(4 * B * np * Dp); w.ql = (float*)take(4 * B * np * Dp);
  w.kh = (float*)take(4 * B * mp * Dp); w.kl = (float*)take(4 * B * mp * Dp);
  w.qth = (float*)take(4 * B * np * Dp); w.qtl = (float*)take(4 * B * np * Dp);
  w.kth = (float*)take(4 * B * mp * Dp); w.ktl = (float*)take(4 * B * mp * Dp);
  w.v4 = (float4*)take(sizeof(float4) * (size_t)vb * mp);
  w.wh = (float*)take(4 * B * np * mp); w.wl = (float*)take(4 * B * np * mp);
  w.wth = (float*)take(4 * B * np * mp); w.wtl = (float*)take(4 * B * np * mp);
  int sq, spq, sk, spk;
  gemm_shape(B, (int)(np / 128), (int)(mp / 32), &sq, &spq);      // g_q: K dimension = m
  gemm_shape(B, (int)(mp / 128), (int)(np / 32), &sk, &spk);      // g_k: K dimension = n
  const size_t pq = sq > 1 ? (size_t)sq * B * np * Dp : 0, pk = sk > 1 ? (size_t)sk * B * mp * Dp : 0;
  w.part = (float*)take(4 * std::max(pq, pk));
  w.total = off;
  return w;
}

// The backward path launches ten kernels; below ~2M logits the SIMT kernels finish sooner.  (The crossover was first
// measured at 8M with an eager, host-bound caller; in the training step -- eager and as a CUDA graph -- the grid<->grid
// problems of 2M logits and the coarse dense one of 4.9M are 2x faster here: 3.49 -> 3.28 ms per graphed step.
// POSFEAT_CORR_TC_BWD_LOG2 moves the threshold for A/B runs.)
bool corr_tc_bwd_eligible(int B, int n, int m, int D, int C) {
  static const int lg = [] { const char* e = getenv("POSFEAT_CORR_TC_BWD_LOG2"); return e ? atoi(e) : 21; }();
  return corr_tc_eligible(B, n, m, D, C) && (int64_t)B * n * m >= ((int64_t)1 << lg);
}

size_t corr_tc_bwd_workspace_bytes(int B, int n, int m, int D, int C) {
  if (!corr_tc_bwd_eligible(B, n, m, D, C)) return 0;
  return carve_corr_tc_bwd(nullptr, B, n, m, D, B).total;
}

static int run_gemm(float* Ah, float* Al, float* Bh, float* Bl, int B, int M_pad, int Kdim, int Dp, int m_valid, int n_valid,
                    float scale, float* out, float* part, cudaStream_t stream) {
  GemmParams g{};
  g.M_pad = M_pad; g.N = Dp; g.m_valid = m_valid; g.n_valid = n_valid; g.slices = Kdim / 32; g.B = B;
  gemm_shape(B, M_pad / 128, g.slices, &g.splits, &g.slices_per_split);
  g.scale = g.splits == 1 ? scale : 1.f;
  g.out = g.splits == 1 ? out : part;
  CUtensorMap mah, mal, mbh, mbl;
  if (int e = make_map_f32(&mah, Ah, Kdim, (int64_t)B * M_pad)) return e;
  if (int e = make_map_f32(&mal, Al, Kdim, (int64_t)B * M_pad)) return e;
  if (int e = make_map_f32(&mbh, Bh, Kdim, (int64_t)B * Dp, Dp)) return e;
  if (int e = make_map_f32(&mbl, Bl, Kdim, (int64_t)B * Dp, Dp)) return e;
  // per device/context attribute: set on every call (cheap), never cached in a process-global flag
  PF_CUDA(cudaFuncSetAttribute(corr_tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGmSmemAlloc));
  corr_tc_gemm_kernel<<<dim3(g.splits, M_pad / 128, B), kGmThreads, kGmSmemAlloc, stream>>>(mah, mal, mbh, mbl, g);
  PF_LAUNCH_CHECK("corr_tc_gemm_kernel");
  if (g.splits > 1) {
    const int64_t total = (int64_t)B * m_valid * n_valid;
    corr_tc_reduce_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(part, g.splits, B, M_pad, Dp, m_valid, n_valid, scale, out);
    PF_LAUNCH_CHECK("corr_tc_reduce_kernel");
  }
  return POSFEAT_OK;
}

// g_q = scale * W K,  g_k = scale * W^T Q  with  W = P o (g.v - g.out)  (see corr_tc_fwd_kernel<1>)
int corr_tc_bwd(const float* q, const float* k, const float* v, int v_batched, int B, int n, int m, int D, int C,
                float scale, const float* out, const float* lse, const float* g_out, float* g_q, float* g_k, void* ws,
                size_t ws_bytes, cudaStream_t stream) {
  const int vb = v_batched ? B : 1;
  CorrTcBwdWs w = carve_corr_tc_bwd(ws, B, n, m, D, vb);
  PF_CHECK_ARG(ws && ((uintptr_t)ws & 255) == 0, "corr workspace must be 256-byte aligned");
  if (ws_bytes < w.total) return set_error(POSFEAT_EWORKSPACE, "corr backward workspace: need %zu bytes, got %zu", w.total, ws_bytes);
  CorrTcParams p{};
  p.B = B; p.n = n; p.m = m; p.D = D; p.C = C;
  p.n_pad = pad128(n); p.m_pad = pad128(m);
  p.atoms = (D + 31) / 32;
  corr_tc_shape(B, n, m, &p.q_tiles, &p.k_tiles, &p.splits, &p.tiles_per_split);
  p.sl2 = scale * 1.4426950408889634f;
  p.v4 = w.v4; p.v_batched = v_batched;
  p.g = g_out; p.o = out; p.lse = lse;
  p.w_hi = w.wh; p.w_lo = w.wl; p.wt_hi = w.wth; p.wt_lo = w.wtl;
  const int Dp = p.atoms * 32;
  ProfScope prof(PROF_CORR_BWD, stream);
  corr_split_kernel<<<(int)std::min<int64_t>(((int64_t)B * p.n_pad * (Dp / 4) + 255) / 256, 148 * 32), 256, 0, stream>>>(q, B, n, D, p.n_pad, Dp, w.qh, w.ql);
  PF_LAUNCH_CHECK("corr_split_kernel(q)");
  corr_split_kernel<<<(int)std::min<int64_t>(((int64_t)B * p.m_pad * (Dp / 4) + 255) / 256, 148 * 32), 256, 0, stream>>>(k, B, m, D, p.m_pad, Dp, w.kh, w.kl);
  PF_LAUNCH_CHECK("corr_split_kernel(k)");
  if (g_k) {
    corr_split_t_kernel<<<dim3(p.n_pad / 32, Dp / 32, B), 256, 0, stream>>>(q, n, D, p.n_pad, Dp, w.qth, w.qtl);
    PF_LAUNCH_CHECK("corr_split_t_kernel(q)");
  }
  if (g_q) {
    corr_split_t_kernel<<<dim3(p.m_pad / 32, Dp / 32, B), 256, 0, stream>>>(k, m, D, p.m_pad, Dp, w.kth, w.ktl);
    PF_LAUNCH_CHECK("corr_split_t_kernel(k)");
  }
  corr_padv_kernel<<<(int)std::min<int64_t>(((int64_t)vb * p.m_pad + 255) / 256, 148 * 8), 256, 0, stream>>>(v, vb, m, C, p.m_pad, w.v4);
  PF_LAUNCH_CHECK("corr_padv_kernel");
  CUtensorMap mqh, mql, mkh, mkl;
  if (int e = make_map_f32(&mqh, w.qh, Dp, (int64_t)B * p.n_pad)) return e;
  if (int e = make_map_f32(&mql, w.ql, Dp, (int64_t)B * p.n_pad)) return e;
  if (int e = make_map_f32(&mkh, w.kh, Dp, (int64_t)B * p.m_pad)) return e;
  if (int e = make_map_f32(&mkl, w.kl, Dp, (int64_t)B * p.m_pad)) return e;
  // per device/context attribute: set on every call (cheap), never cached in a process-global flag
  PF_CUDA(cudaFuncSetAttribute(corr_tc_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtSmemAlloc));
  corr_tc_fwd_kernel<1><<<dim3(p.splits, p.q_tiles, B), kCtThreads, kCtSmemAlloc, stream>>>(mqh, mql, mkh, mkl, p);
  PF_LAUNCH_CHECK("corr_tc_fwd_kernel<W>");
  if (g_q)
    if (int e = run_gemm(w.wh, w.wl, w.kth, w.ktl, B, p.n_pad, p.m_pad, Dp, n, D, scale, g_q, w.part, stream)) return e;
  if (g_k)
    if (int e = run_gemm(w.wth, w.wtl, w.qth, w.qtl, B, p.m_pad, p.n_pad, Dp, m, D, scale, g_k, w.part, stream)) return e;
  return POSFEAT_OK;
}

// ---------------------------------------------------------------- DiskLoss dual-softmax reward sums
__global__ void corr_padtab_kernel(const float* __restrict__ tab, int B, int rows, int rows_pad, float4* __restrict__ out) {
  const int64_t total = (int64_t)B * rows_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i % rows_pad), b = (int)(i / rows_pad);
    float4 a = make_float4(INFINITY, 0.f, 0.f, 0.f), c = make_float4(0.f, 0.f, 0.f, 0.f);   // padding: lse = +inf -> p = 0
    if (r < rows) {
      const float* t = tab + ((int64_t)b * rows + r) * 8;
      a = make_float4(t[0], t[1], t[2], t[3]);
      c = make_float4(t[4], t[5], t[6], t[7]);
    }
    out[2 * i] = a;
    out[2 * i + 1] = c;
  }
}

__global__ void corr_disk_merge_kernel(const float4* __restrict__ part, int B, int n, int q_tiles, int splits,
                                       float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * n) return;
  const int b = (int)(i / n), r = (int)(i % n);
  const float4* base = part + (((size_t)b * q_tiles + r / kCtM) * splits) * kCtM + r % kCtM;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = base[(size_t)s * kCtM];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w = fmaxf(acc.w, v.w);
  }
  out[i] = acc;
}

struct CorrDiskWs {
  float *qh, *ql, *kh, *kl;
  float4 *rowtab, *coltab, *part4;
  size_t total;
};

static CorrDiskWs carve_corr_disk(void* base, int B, int n, int m, int D) {
  CorrDiskWs w{};
  const size_t Dp = (D + 31) / 32 * 32, np = pad128(n), mp = pad128(m);
  int qt, kt, sp, tps;
  corr_tc_shape(B, n, m, &qt, &kt, &sp, &tps);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (char*)base + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  w.qh = (float*)take(4 * B * np * Dp); w.ql = (float*)take(4 * B * np * Dp);
  w.kh = (float*)take(4 * B * mp * Dp); w.kl = (float*)take(4 * B * mp * Dp);
  w.rowtab = (float4*)take(sizeof(float4) * 2 * B * np);
  w.coltab = (float4*)take(sizeof(float4) * 2 * B * mp);
  w.part4 = (float4*)take(sizeof(float4) * (size_t)B * qt * sp * kCtM);
  w.total = off;
  return w;
}

size_t corr_disk_workspace_bytes(int B, int n, int m, int D) { return carve_corr_disk(nullptr, B, n, m, D).total; }

// rows_out [B][n] float4 = {sum_j acc r p (log p + logp_i + logp_j), sum_j acc r p, sum_j p, max_j p}
int corr_disk_rows(const float* q, const float* k, const float* rowtab, const float* coltab, int B, int n, int m, int D,
                   float temp, float thr_own, float thr_other, float good_reward, float bad_reward, int dynamic_reward,
                   float* rows_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
  CorrDiskWs w = carve_corr_disk(ws, B, n, m, D);
  PF_CHECK_ARG(ws && ((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
  if (ws_bytes < w.total) return set_error(POSFEAT_EWORKSPACE, "dual softmax workspace: need %zu bytes, got %zu", w.total, ws_bytes);
  CorrTcParams p{};
  p.B = B; p.n = n; p.m = m; p.D = D; p.C = 1;
  p.n_pad = pad128(n); p.m_pad = pad128(m);
  p.atoms = (D + 31) / 32;
  corr_tc_shape(B, n, m, &p.q_tiles, &p.k_tiles, &p.splits, &p.tiles_per_split);
  p.rowtab = w.rowtab; p.coltab = w.coltab; p.part4 = w.part4;
  p.temp = temp; p.thr_own = thr_own; p.thr_other = thr_other;
  p.good_reward = good_reward; p.bad_reward = bad_reward; p.dynamic_reward = dynamic_reward;
  const int Dp = p.atoms * 32;
  ProfScope prof(PROF_CORR_FWD, stream);
  corr_split_kernel<<<(int)std::min<int64_t>(((int64_t)B * p.n_pad * (Dp / 4) + 255) / 256, 148 * 32), 256, 0, stream>>>(q, B, n, D, p.n_pad, Dp, w.qh, w.ql);
  PF_LAUNCH_CHECK("corr_split_kernel(q)");
  corr_split_kernel<<<(int)std::min<int64_t>(((int64_t)B * p.m_pad * (Dp / 4) + 255) / 256, 148 * 32), 256, 0, stream>>>(k, B, m, D, p.m_pad, Dp, w.kh, w.kl);
  PF_LAUNCH_CHECK("corr_split_kernel(k)");
  corr_padtab_kernel<<<(int)std::min<int64_t>(((int64_t)B * p.n_pad + 255) / 256, 148 * 8), 256, 0, stream>>>(rowtab, B, n, p.n_pad, w.rowtab);
  PF_LAUNCH_CHECK("corr_padtab_kernel(rows)");
  corr_padtab_kernel<<<(int)std::min<int64_t>(((int64_t)B * p.m_pad + 255) / 256, 148 * 8), 256, 0, stream>>>(coltab, B, m, p.m_pad, w.coltab);
  PF_LAUNCH_CHECK("corr_padtab_kernel(cols)");
  CUtensorMap mqh, mql, mkh, mkl;
  if (int e = make_map_f32(&mqh, w.qh, Dp, (int64_t)B * p.n_pad)) return e;
  if (int e = make_map_f32(&mql, w.ql, Dp, (int64_t)B * p.n_pad)) return e;
  if (int e = make_map_f32(&mkh, w.kh, Dp, (int64_t)B * p.m_pad)) return e;
  if (int e = make_map_f32(&mkl, w.kl, Dp, (int64_t)B * p.m_pad)) return e;
  // per device/context attribute: set on every call (cheap), never cached in a process-global flag
  PF_CUDA(cudaFuncSetAttribute(corr_tc_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtSmemAlloc));
  corr_tc_fwd_kernel<2><<<dim3(p.splits, p.q_tiles, B), kCtThreads, kCtSmemAlloc, stream>>>(mqh, mql, mkh, mkl, p);
  PF_LAUNCH_CHECK("corr_tc_fwd_kernel<disk>");
  corr_disk_merge_kernel<<<(int)(((int64_t)B * n + 255) / 256), 256, 0, stream>>>(w.part4, B, n, p.q_tiles, p.splits,
                                                                                 reinterpret_cast<float4*>(rows_out));
  PF_LAUNCH_CHECK("corr_disk_merge_kernel");
  return POSFEAT_OK;
}

}  // namespace posfeat
