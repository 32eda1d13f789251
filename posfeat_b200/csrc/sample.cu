// Subsystem (2): bilinear descriptor sampling fused with L2 normalisation.
//
// Replaces sample_feat_by_coord, reference losses/preprocess_utils.py:40-53:
//   F.grid_sample(x, coord_n[:, :, None], bilinear, padding 'zeros',
//                 align_corners=False)  ->  F.normalize(dim=C)  -> [B, n, C].
// One warp per keypoint.  The descriptor map is addressed through element
// strides, so NCHW (the reference layout) and NHWC (channels_last backbones)
// use the same entry point; with unit channel stride the four taps are read
// as 128-bit vectors (4 x 512 B per keypoint at D=128).  Gather-bound: the
// algorithmic traffic is n*(4*D*4 + D*4 + 8) bytes.
#include <cuda_bf16.h>

#include "common.cuh"
#include "mnn_common.cuh"

namespace posfeat {

struct Taps {
  int x0, y0;
  float w00, w01, w10, w11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1), zero when outside
  bool in00, in01, in10, in11;
};

__device__ __forceinline__ Taps make_taps(float gx, float gy, int h, int w) {
  // grid_sampler_unnormalize (align_corners=False): ((g + 1) * size - 1) / 2
  // (explicit _rn ops: no FMA contraction, same roundings as ATen's CUDA sampler)
  const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)w), 1.f), 0.5f);
  const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)h), 1.f), 0.5f);
  const float fx = floorf(ix), fy = floorf(iy);
  Taps t;
  // clamp before the int conversion so absurd coordinates cannot overflow
  t.x0 = (int)fminf(fmaxf(fx, -2.f), (float)w);
  t.y0 = (int)fminf(fmaxf(fy, -2.f), (float)h);
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix;
  const float wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  const bool xin0 = t.x0 >= 0 && t.x0 < w, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < w;
  const bool yin0 = t.y0 >= 0 && t.y0 < h, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < h;
  t.in00 = xin0 && yin0; t.in01 = xin1 && yin0; t.in10 = xin0 && yin1; t.in11 = xin1 && yin1;
  t.w00 = __fmul_rn(wy0, wx0); t.w01 = __fmul_rn(wy0, wx1);
  t.w10 = __fmul_rn(wy1, wx0); t.w11 = __fmul_rn(wy1, wx1);
  return t;
}

// generic strides: lane handles channels lane, lane+32, ...
template <int CPL>  // channels per lane (D <= 32*CPL)
__global__ void __launch_bounds__(256)
sample_strided_kernel(const float* __restrict__ fmap, int D, int h, int w, int64_t sb, int64_t sc,
                      int64_t sy, int64_t sx, const float* __restrict__ coord, int n,
                      const int32_t* __restrict__ n_valid, int do_norm, float* __restrict__ out,
                      __nv_bfloat16* __restrict__ out_bf16) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nv = n_valid ? min(*n_valid, n) : n;
  if (p >= nv) return;
  const float2 g = *reinterpret_cast<const float2*>(coord + ((int64_t)b * n + p) * 2);
  const Taps t = make_taps(g.x, g.y, h, w);
  const float* base = fmap + b * sb + (int64_t)t.y0 * sy + (int64_t)t.x0 * sx;
  float v[CPL];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
    if (c < D) {
      const float* pc = base + (int64_t)c * sc;
      if (t.in00) a00 = __ldg(pc);
      if (t.in01) a01 = __ldg(pc + sx);
      if (t.in10) a10 = __ldg(pc + sy);
      if (t.in11) a11 = __ldg(pc + sy + sx);
    }
    v[j] = a00 * t.w00 + a01 * t.w01 + a10 * t.w10 + a11 * t.w11;
    ss += v[j] * v[j];
  }
  float inv = 1.f;
  if (do_norm) {
    ss = warp_sum(ss);
    inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  }
  float* o = out + ((int64_t)b * n + p) * D;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    if (c < D) {
      const float r = do_norm ? v[j] * inv : v[j];
      o[c] = r;
      if (out_bf16) out_bf16[((int64_t)b * n + p) * D + c] = __float2bfloat16_rn(r);
    }
  }
}

// unit channel stride (NHWC), D % 4 == 0, 16-byte aligned pixels: float4 taps
template <int VPL, bool kSink>  // float4 per lane (D <= 128*VPL)
__global__ void __launch_bounds__(256)
sample_nhwc_kernel(const float* __restrict__ fmap, int D, int h, int w, int64_t sb, int64_t sy,
                   int64_t sx, const float* __restrict__ coord, int n,
                   const int32_t* __restrict__ n_valid, int do_norm, float* __restrict__ out,
                   __nv_bfloat16* __restrict__ out_bf16, const PrepSink sink) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nv = n_valid ? min(*n_valid, n) : n;
  if (p >= nv) return;
  const float2 g = *reinterpret_cast<const float2*>(coord + ((int64_t)b * n + p) * 2);
  const Taps t = make_taps(g.x, g.y, h, w);
  const float* base = fmap + b * sb + (int64_t)t.y0 * sy + (int64_t)t.x0 * sx;
  float4 v[VPL];
  float ss = 0.f;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c = 4 * (lane + 32 * j);
    float4 a00 = z, a01 = z, a10 = z, a11 = z;
    if (c < D) {
      const float* pc = base + c;
      if (t.in00) a00 = __ldg(reinterpret_cast<const float4*>(pc));
      if (t.in01) a01 = __ldg(reinterpret_cast<const float4*>(pc + sx));
      if (t.in10) a10 = __ldg(reinterpret_cast<const float4*>(pc + sy));
      if (t.in11) a11 = __ldg(reinterpret_cast<const float4*>(pc + sy + sx));
    }
    float4 r;
    r.x = a00.x * t.w00 + a01.x * t.w01 + a10.x * t.w10 + a11.x * t.w11;
    r.y = a00.y * t.w00 + a01.y * t.w01 + a10.y * t.w10 + a11.y * t.w11;
    r.z = a00.z * t.w00 + a01.z * t.w01 + a10.z * t.w10 + a11.z * t.w11;
    r.w = a00.w * t.w00 + a01.w * t.w01 + a10.w * t.w10 + a11.w * t.w11;
    v[j] = r;
    ss += r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
  }
  float inv = 1.f;
  if (do_norm) {
    ss = warp_sum(ss);
    inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  }
  float* o = out + ((int64_t)b * n + p) * D;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c = 4 * (lane + 32 * j);
    if (c < D) {
      float4 r = v[j];
      if (do_norm) { r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv; }
      *reinterpret_cast<float4*>(o + c) = r;
      if (kSink) prep_sink_row(sink, b >> 1, b & 1, p, lane, r);     // D == 128: every lane holds one float4
      if (out_bf16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned*>(&lo);
        pk.y = *reinterpret_cast<unsigned*>(&hi);
        *reinterpret_cast<uint2*>(out_bf16 + ((int64_t)b * n + p) * D + c) = pk;
      }
    }
  }
}

// ---- sparse host->device staging of exactly the pixels the sampler will read --------------------
// Host-buffer callers (PairPipeline.run_host): the dense descriptor map stays in pinned host memory;
// mark_taps_kernel sets one bit per map pixel that some keypoint's 2x2 tap block covers (same
// make_taps as the sampler, so the cover is exact), fetch_pixels_kernel then moves each marked pixel
// (D contiguous floats) once over the host link into the same position of a device-resident map, and
// the ordinary sampler runs on that map.  At 8192 keypoints on a 224x300 map 39 % of the pixels are
// marked: 13.3 MB cross the link instead of 34.4 MB (dense copy) or 16.8 MB (gathering per keypoint).
__global__ void __launch_bounds__(256)
mark_taps_kernel(const float* __restrict__ coord, int n, int h, int w, int words_per_image,
                 unsigned* __restrict__ bitmap) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float2 g = *reinterpret_cast<const float2*>(coord + ((int64_t)b * n + p) * 2);
  const Taps t = make_taps(g.x, g.y, h, w);
  unsigned* bm = bitmap + (int64_t)b * words_per_image;
  const int i00 = t.y0 * w + t.x0;
  if (t.in00) atomicOr(bm + (i00 >> 5), 1u << (i00 & 31));
  if (t.in01) atomicOr(bm + ((i00 + 1) >> 5), 1u << ((i00 + 1) & 31));
  if (t.in10) atomicOr(bm + ((i00 + w) >> 5), 1u << ((i00 + w) & 31));
  if (t.in11) atomicOr(bm + ((i00 + w + 1) >> 5), 1u << ((i00 + w + 1) & 31));
}

template <int VPL>  // float4 per lane and pixel (D <= 128*VPL)
__global__ void __launch_bounds__(256)
fetch_pixels_kernel(const float* __restrict__ src, float* __restrict__ dst, int D, int w, int64_t sb,
                    int64_t sy, int64_t sx, int words_per_image, const unsigned* __restrict__ bitmap,
                    unsigned long long* __restrict__ n_fetched) {
  constexpr int kInFlight = 4;        // pixels whose loads are issued before the first store
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int wd = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wd >= words_per_image) return;
  unsigned m = bitmap[(int64_t)b * words_per_image + wd];
  if (!m) return;
  if (n_fetched && lane == 0) atomicAdd(n_fetched, (unsigned long long)__popc(m));
  const int64_t ib = b * sb;
  while (m) {
    int64_t off[kInFlight];
    float4 v[kInFlight][VPL];
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < kInFlight; ++u) {
      off[u] = -1;
      if (m) {
        const int pix = wd * 32 + __ffs(m) - 1;
        m &= m - 1;
        const int y = pix / w, x = pix - y * w;
        off[u] = ib + (int64_t)y * sy + (int64_t)x * sx;
        ++cnt;
      }
    }
#pragma unroll
    for (int u = 0; u < kInFlight; ++u)
      if (u < cnt) {
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const int c = 4 * (lane + 32 * j);
          if (c < D) v[u][j] = __ldcs(reinterpret_cast<const float4*>(src + off[u] + c));
        }
      }
#pragma unroll
    for (int u = 0; u < kInFlight; ++u)
      if (u < cnt) {
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const int c = 4 * (lane + 32 * j);
          if (c < D) *reinterpret_cast<float4*>(dst + off[u] + c) = v[u][j];
        }
      }
  }
}

// backward of the bilinear gather: g_fmap[taps] += w_tap * g_out (float atomics, as
// ATen's grid_sampler_2d_backward does); gradients w.r.t. the coordinates are not
// needed on the reference's paths (keypoint coordinates are detached).
__global__ void __launch_bounds__(256)
sample_bwd_kernel(const float* __restrict__ g_out, int D, int h, int w, int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                  const float* __restrict__ coord, int n, float* __restrict__ g_fmap) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= n) return;
  const float2 g = *reinterpret_cast<const float2*>(coord + ((int64_t)b * n + p) * 2);
  const Taps t = make_taps(g.x, g.y, h, w);
  float* base = g_fmap + b * sb + (int64_t)t.y0 * sy + (int64_t)t.x0 * sx;
  const float* go = g_out + ((int64_t)b * n + p) * D;
  for (int c = lane; c < D; c += 32) {
    const float gv = __ldg(go + c);
    float* pc = base + (int64_t)c * sc;
    if (t.in00) atomicAdd(pc, gv * t.w00);
    if (t.in01) atomicAdd(pc + sx, gv * t.w01);
    if (t.in10) atomicAdd(pc + sy, gv * t.w10);
    if (t.in11) atomicAdd(pc + sy + sx, gv * t.w11);
  }
}

}  // namespace posfeat

using namespace posfeat;

extern "C" int posfeat_sample_l2norm_f32(const float* fmap, int B, int D, int h, int w, int64_t sb, int64_t sc,
                                         int64_t sy, int64_t sx, const float* coord_n, int n,
                                         const int32_t* n_valid, int do_norm, float* out, void* out_bf16,
                                         void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PF_CHECK_ARG(fmap && coord_n && out, "NULL pointer");
  PF_CHECK_ARG(B >= 1 && D >= 1 && h >= 1 && w >= 1 && n >= 0, "bad shape B=%d D=%d h=%d w=%d n=%d", B, D, h, w, n);
  PF_CHECK_ARG(D <= 512, "D=%d > 512 not supported", D);
  PF_CHECK_ARG(B <= 65535, "B=%d exceeds the grid limit 65535", B);
  if (n == 0) return POSFEAT_OK;
  const int warps = 8;
  dim3 grid((n + warps - 1) / warps, B), block(32 * warps);
  __nv_bfloat16* ob = (__nv_bfloat16*)out_bf16;
  ProfScope prof(PROF_SAMPLE, stream);
  const bool vec = sc == 1 && D % 4 == 0 && sx % 4 == 0 && sy % 4 == 0 && sb % 4 == 0 &&
                   ((uintptr_t)fmap % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                   (!ob || (uintptr_t)ob % 8 == 0);
  if (vec) {
    if (D <= 128)
      sample_nhwc_kernel<1, false><<<grid, block, 0, stream>>>(fmap, D, h, w, sb, sy, sx, coord_n, n, n_valid, do_norm, out, ob, PrepSink{});
    else if (D <= 256)
      sample_nhwc_kernel<2, false><<<grid, block, 0, stream>>>(fmap, D, h, w, sb, sy, sx, coord_n, n, n_valid, do_norm, out, ob, PrepSink{});
    else
      sample_nhwc_kernel<4, false><<<grid, block, 0, stream>>>(fmap, D, h, w, sb, sy, sx, coord_n, n, n_valid, do_norm, out, ob, PrepSink{});
  } else {
    if (D <= 32)
      sample_strided_kernel<1><<<grid, block, 0, stream>>>(fmap, D, h, w, sb, sc, sy, sx, coord_n, n, n_valid, do_norm, out, ob);
    else if (D <= 64)
      sample_strided_kernel<2><<<grid, block, 0, stream>>>(fmap, D, h, w, sb, sc, sy, sx, coord_n, n, n_valid, do_norm, out, ob);
    else if (D <= 128)
      sample_strided_kernel<4><<<grid, block, 0, stream>>>(fmap, D, h, w, sb, sc, sy, sx, coord_n, n, n_valid, do_norm, out, ob);
    else if (D <= 256)
      sample_strided_kernel<8><<<grid, block, 0, stream>>>(fmap, D, h, w, sb, sc, sy, sx, coord_n, n, n_valid, do_norm, out, ob);
    else
      sample_strided_kernel<16><<<grid, block, 0, stream>>>(fmap, D, h, w, sb, sc, sy, sx, coord_n, n, n_valid, do_norm, out, ob);
  }
  PF_LAUNCH_CHECK("sample kernel");
  return POSFEAT_OK;
}

// Sampler + matcher operand preparation in one pass (pair pipeline): images (2p, 2p+1) are the two sides of
// pair p; besides `out` the kernel leaves the tensor-core matcher's bf16 operand rows, row norms, rounding
// error norms and their maxima in the matcher workspace, so posfeat_mnn_batched_f32 can be called with
// POSFEAT_MNN_TC | POSFEAT_MNN_PREPARED and skips its own pass over the descriptors.
extern "C" int posfeat_sample_pairs_f32(const float* fmap, int B, int D, int h, int w, int64_t sb, int64_t sc,
                                        int64_t sy, int64_t sx, const float* coord_n, int n, int do_norm, float* out,
                                        void* mnn_workspace, size_t mnn_ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PF_CHECK_ARG(fmap && coord_n && out && mnn_workspace, "NULL pointer");
  PF_CHECK_ARG(B >= 2 && B % 2 == 0 && B <= 65534, "B=%d: need an even number of images (pairs)", B);
  PF_CHECK_ARG(D == 128 && n >= 1 && h >= 1 && w >= 1, "fused sampling needs D == 128 (got D=%d) and n >= 1", D);
  PF_CHECK_ARG(sc == 1 && sx % 4 == 0 && sy % 4 == 0 && sb % 4 == 0 && ((uintptr_t)fmap % 16 == 0) &&
                   ((uintptr_t)out % 16 == 0),
               "fused sampling needs a channels-last, 16-byte aligned descriptor map");
  PrepSink sink;
  if (int e = tc_prep_sink(mnn_workspace, mnn_ws_bytes, B / 2, n, n, &sink, stream)) return e;
  const int warps = 8;
  dim3 grid((n + warps - 1) / warps, B), block(32 * warps);
  ProfScope prof(PROF_SAMPLE, stream);
  sample_nhwc_kernel<1, true><<<grid, block, 0, stream>>>(fmap, D, h, w, sb, sy, sx, coord_n, n, nullptr, do_norm, out,
                                                           nullptr, sink);
  PF_LAUNCH_CHECK("sample_nhwc_kernel<sink>");
  return POSFEAT_OK;
}

extern "C" size_t posfeat_fetch_taps_workspace_bytes(int B, int h, int w) {
  if (B < 1 || h < 1 || w < 1) return 0;
  return ((size_t)B * (size_t)(((int64_t)h * w + 31) / 32) * 4 + 255) / 256 * 256 + 256;
}

extern "C" int posfeat_fetch_taps_f32(const float* fmap_host, float* fmap_dev, int B, int D, int h, int w, int64_t sb,
                                      int64_t sc, int64_t sy, int64_t sx, const float* coord_n, int n,
                                      void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PF_CHECK_ARG(fmap_host && fmap_dev && workspace && (coord_n || n == 0), "NULL pointer");
  PF_CHECK_ARG(B >= 1 && B <= 65535 && D >= 1 && D <= 512 && h >= 1 && w >= 1 && n >= 0, "bad shape B=%d D=%d h=%d w=%d n=%d",
               B, D, h, w, n);
  PF_CHECK_ARG((int64_t)h * w <= 0x7fffffff - 64, "map of %d x %d pixels is too large", h, w);
  PF_CHECK_ARG(sc == 1 && D % 4 == 0 && sx % 4 == 0 && sy % 4 == 0 && sb % 4 == 0 && ((uintptr_t)fmap_host % 16 == 0) &&
                   ((uintptr_t)fmap_dev % 16 == 0),
               "tap staging needs channels-last, 16-byte aligned descriptor maps with D %% 4 == 0");
  PF_CHECK_ARG(ws_bytes >= posfeat_fetch_taps_workspace_bytes(B, h, w), "workspace too small (%zu bytes)", ws_bytes);
  const int words = (int)(((int64_t)h * w + 31) / 32);
  unsigned* bitmap = (unsigned*)workspace;
  unsigned long long* counter = (unsigned long long*)((char*)workspace + ((size_t)B * words * 4 + 255) / 256 * 256);
  PF_CUDA(cudaMemsetAsync(workspace, 0, posfeat_fetch_taps_workspace_bytes(B, h, w), stream));
  if (n == 0) return POSFEAT_OK;                       // nothing to stage; the pixel counter reads 0
  ProfScope prof(PROF_FETCH, stream);
  mark_taps_kernel<<<dim3((n + 255) / 256, B), 256, 0, stream>>>(coord_n, n, h, w, words, bitmap);
  PF_LAUNCH_CHECK("mark_taps_kernel");
  const int warps = 8;
  dim3 grid((words + warps - 1) / warps, B);
  if (D <= 128)
    fetch_pixels_kernel<1><<<grid, 32 * warps, 0, stream>>>(fmap_host, fmap_dev, D, w, sb, sy, sx, words, bitmap, counter);
  else if (D <= 256)
    fetch_pixels_kernel<2><<<grid, 32 * warps, 0, stream>>>(fmap_host, fmap_dev, D, w, sb, sy, sx, words, bitmap, counter);
  else
    fetch_pixels_kernel<4><<<grid, 32 * warps, 0, stream>>>(fmap_host, fmap_dev, D, w, sb, sy, sx, words, bitmap, counter);
  PF_LAUNCH_CHECK("fetch_pixels_kernel");
  return POSFEAT_OK;
}

extern "C" int posfeat_fetch_taps_count(const void* workspace, int B, int h, int w, unsigned long long* pixels_out,
                                        void* stream_) {
  PF_CHECK_ARG(workspace && pixels_out && B >= 1 && h >= 1 && w >= 1, "bad argument");
  const int words = (int)(((int64_t)h * w + 31) / 32);
  const char* counter = (const char*)workspace + ((size_t)B * words * 4 + 255) / 256 * 256;
  PF_CUDA(cudaMemcpyAsync(pixels_out, counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
  PF_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  return POSFEAT_OK;
}

extern "C" int posfeat_sample_bwd_f32(const float* g_out, int B, int D, int h, int w, int64_t sb, int64_t sc, int64_t sy,
                                      int64_t sx, const float* coord_n, int n, float* g_fmap, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PF_CHECK_ARG(g_out && coord_n && g_fmap, "NULL pointer");
  PF_CHECK_ARG(B >= 1 && B <= 65535 && D >= 1 && h >= 1 && w >= 1 && n >= 0, "bad shape");
  if (n == 0) return POSFEAT_OK;
  dim3 grid((n + 7) / 8, B);
  sample_bwd_kernel<<<grid, 256, 0, stream>>>(g_out, D, h, w, sb, sc, sy, sx, coord_n, n, g_fmap);
  PF_LAUNCH_CHECK("sample_bwd_kernel");
  return POSFEAT_OK;
}
