// T * F.normalize(x, p=2, dim=1) of a descriptor map, written channels-last -- one pass.
//
// Preprocess_Line2Window (losses/preprocess.py:56-57, :84-106) normalises both fine descriptor maps and the line /
// window kernels want them channels-innermost.  In stock tensor operations that is five launches per map forwards
// (norm, clamp, broadcast division, scale, layout copy) and about ten backwards, each a pass over a 79 MB map: 1.4 ms
// of the 5.8 ms training step.  Here: a CTA owns 32 consecutive pixels of one image and all D channels; the NCHW
// reads are 128-byte rows (32 pixels of one channel), the tile is transposed through shared memory and leaves as
// 128-byte rows of the channels-last output.  The backward pass
//     g_x = (T / max(|x|, eps)) * (g - xh <g, xh>),  xh = x / |x|        (|x| >= eps)
//     g_x = (T / eps) * g                                                  (|x| <  eps: the denominator is the constant)
// reads the channels-last gradient and the NCHW input the same way and writes NCHW.
#include "common.cuh"

namespace posfeat {

constexpr int kNzPix = 32;       // pixels per CTA
constexpr int kNzWarps = 8;

// x: [B][D] channels at stride sc, pixels contiguous (NCHW with h*w flattened); out: [B][HW][D]
__global__ void __launch_bounds__(kNzWarps * 32)
normalize_scale_fwd_kernel(const float* __restrict__ x, int D, int HW, int64_t sb, int64_t sc, float scale, float eps,
                           float* __restrict__ out, float* __restrict__ inv_norm) {
  extern __shared__ float tile[];                      // [D][33]
  __shared__ float s_part[kNzWarps][kNzPix];
  __shared__ float s_inv[kNzPix];
  const int b = blockIdx.y, p0 = blockIdx.x * kNzPix;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int p = p0 + lane;
  const float* xb = x + b * sb;
  float ss = 0.f;
  for (int c = w; c < D; c += kNzWarps) {
    const float v = p < HW ? __ldg(xb + c * sc + p) : 0.f;
    tile[c * 33 + lane] = v;
    ss = fmaf(v, v, ss);
  }
  s_part[w][lane] = ss;
  __syncthreads();
  if (w == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kNzWarps; ++k) t += s_part[k][lane];
    const float nrm = sqrtf(t);
    s_inv[lane] = nrm;                                 // the norm itself: the division below is torch's x / max(norm, eps)
    if (p < HW && inv_norm) inv_norm[(int64_t)b * HW + p] = nrm;
  }
  __syncthreads();
  float* ob = out + ((int64_t)b * HW + p0) * D;
  for (int q = w; q < kNzPix && p0 + q < HW; q += kNzWarps) {
    const float den = fmaxf(s_inv[q], eps);
    for (int c = lane; c < D; c += 32) ob[(int64_t)q * D + c] = scale * (tile[c * 33 + q] / den);
  }
}

// g: [B][HW][D] (channels-last gradient of the output); x as above; gx: [B][D][HW] at the strides of x
__global__ void __launch_bounds__(kNzWarps * 32)
normalize_scale_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ norm,
                           int D, int HW, int64_t sb, int64_t sc, float scale, float eps, float* __restrict__ gx) {
  extern __shared__ float tile[];                      // [D][33]: g transposed
  __shared__ float s_dot[kNzPix];
  __shared__ float s_part[kNzWarps][kNzPix];
  const int b = blockIdx.y, p0 = blockIdx.x * kNzPix;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int p = p0 + lane;
  const float* gb = g + ((int64_t)b * HW + p0) * D;
  for (int q = w; q < kNzPix; q += kNzWarps)
    for (int c = lane; c < D; c += 32) tile[c * 33 + q] = p0 + q < HW ? __ldg(gb + (int64_t)q * D + c) : 0.f;
  __syncthreads();
  const float* xb = x + b * sb;
  const float nrm = p < HW ? __ldg(norm + (int64_t)b * HW + p) : 1.f;
  const bool clamped = nrm < eps;
  const float den = fmaxf(nrm, eps);
  // <g, xh> per pixel: this thread's share of the channels
  float dot = 0.f;
  for (int c = w; c < D; c += kNzWarps) {
    const float v = p < HW ? __ldg(xb + c * sc + p) : 0.f;
    dot = fmaf(tile[c * 33 + lane], v / den, dot);
  }
  s_part[w][lane] = dot;
  __syncthreads();
  if (w == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kNzWarps; ++k) t += s_part[k][lane];
    s_dot[lane] = t;
  }
  __syncthreads();
  if (p >= HW) return;
  const float d = clamped ? 0.f : s_dot[lane];
  const float k = scale / den;
  float* gxb = gx + b * sb;
  for (int c = w; c < D; c += kNzWarps) {
    const float v = __ldg(xb + c * sc + p);            // second read: out of L1 / L2
    gxb[c * sc + p] = k * (tile[c * 33 + lane] - (v / den) * d);
  }
}

}  // namespace posfeat

using namespace posfeat;

extern "C" int posfeat_normalize_scale_fwd_f32(const float* x, int B, int D, int HW, int64_t sb, int64_t sc, float scale,
                                               float eps, float* out, float* norm, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PF_CHECK_ARG(x && out, "NULL pointer");
  PF_CHECK_ARG(B >= 1 && B <= 65535 && D >= 1 && D <= 1024 && HW >= 1, "bad shape B=%d D=%d HW=%d", B, D, HW);
  const size_t smem = sizeof(float) * 33 * (size_t)D;
  if (smem > 48 * 1024)
    PF_CUDA(cudaFuncSetAttribute(normalize_scale_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  normalize_scale_fwd_kernel<<<dim3((HW + kNzPix - 1) / kNzPix, B), kNzWarps * 32, smem, stream>>>(x, D, HW, sb, sc, scale, eps,
                                                                                              out, norm);
  PF_LAUNCH_CHECK("normalize_scale_fwd_kernel");
  return POSFEAT_OK;
}

extern "C" int posfeat_normalize_scale_bwd_f32(const float* g, const float* x, const float* norm, int B, int D, int HW,
                                               int64_t sb, int64_t sc, float scale, float eps, float* gx, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PF_CHECK_ARG(g && x && norm && gx, "NULL pointer");
  PF_CHECK_ARG(B >= 1 && B <= 65535 && D >= 1 && D <= 1024 && HW >= 1, "bad shape B=%d D=%d HW=%d", B, D, HW);
  const size_t smem = sizeof(float) * 33 * (size_t)D;
  if (smem > 48 * 1024)
    PF_CUDA(cudaFuncSetAttribute(normalize_scale_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  normalize_scale_bwd_kernel<<<dim3((HW + kNzPix - 1) / kNzPix, B), kNzWarps * 32, smem, stream>>>(g, x, norm, D, HW, sb, sc, scale,
                                                                                              eps, gx);
  PF_LAUNCH_CHECK("normalize_scale_bwd_kernel");
  return POSFEAT_OK;
}
