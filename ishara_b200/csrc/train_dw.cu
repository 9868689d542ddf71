// ishara_b200 — depthwise temporal convolution for the training step (SURVEY.md §8 rows T4/T8/T11 under T15):
// forward with an optional swish on the input (so the PRE-activation tensor is the one kept for backward), the
// data gradient (same kernel: flipped taps, mirrored padding, optional multiply by swish'(pre-activation)) and the
// tap/bias gradient. CausalDWConv1D c5:17-39 (pad_left = k-1), ConvModule.conv2 c5:142 (causal, k = 15),
// ConvolutionModule.depthwise_conv c5:272-279 ('same', pad_left = (k-1)/2, with bias).
//
// HBM-bound stencils: a thread owns two adjacent channels and walks time with a K-deep register window, so a warp
// reads 128 contiguous bytes per time step; CTAs tile (sequence, time chunk, 256-channel slab).
#include <cstdio>

#include "ptx.cuh"
#include "train_kernels.h"

namespace ishara {
namespace {

constexpr int kDwThreads = 128;  // x 2 channels = 256-channel slab
constexpr int kDwChunk = 32;     // output time steps per CTA (forward / data gradient)
constexpr int kDwWgChunk = 64;   // time steps per CTA (tap gradient)

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float swish_f(float x) { return x * sigm(x); }
__device__ __forceinline__ float swish_d(float x) {
  const float s = sigm(x);
  return s * (1.f + x * (1.f - s));
}

constexpr int kDwBatch = 8;  // time steps whose loads are issued together (memory-level parallelism; the walk is latency-bound otherwise)

template <int K>
__global__ void __launch_bounds__(kDwThreads) dw_train_kernel(DwTrainArgs a) {
  const int c = (blockIdx.x * kDwThreads + threadIdx.x) * 2;
  if (c >= a.C) return;
  const int t0 = blockIdx.y * kDwChunk, b = blockIdx.z;
  const int T = a.T, C = a.C;
  const int t_end = t0 + kDwChunk < T ? t0 + kDwChunk : T;  // exclusive
  float w0[K], w1[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const float2 wv = *reinterpret_cast<const float2*>(a.w + static_cast<size_t>(a.flip ? K - 1 - j : j) * C + c);
    w0[j] = wv.x;
    w1[j] = wv.y;
  }
  float b0 = 0.f, b1 = 0.f;
  if (a.bias != nullptr) { b0 = a.bias[c]; b1 = a.bias[c + 1]; }
  const bf16* __restrict__ in = a.in + static_cast<size_t>(b) * T * C + c;
  bf16* __restrict__ out = a.out + static_cast<size_t>(b) * T * C + c;
  const bf16* __restrict__ ref = a.mul_ref != nullptr ? a.mul_ref + static_cast<size_t>(b) * T * C + c : nullptr;
  float x0[K], x1[K];
#pragma unroll
  for (int j = 0; j < K; ++j) { x0[j] = 0.f; x1[j] = 0.f; }
  // step s pushes in[s]; afterwards the window holds in[s-K+1 .. s] and produces out[t], t = s - (K-1) + pad_left
  const int s_begin = t0 - a.pad_left, s_end = t_end - 1 - a.pad_left + (K - 1);
  const int dt = a.pad_left - (K - 1);  // t = s + dt
  for (int sb = s_begin; sb <= s_end; sb += kDwBatch) {
    uint32_t raw[kDwBatch], rr[kDwBatch];
#pragma unroll
    for (int u = 0; u < kDwBatch; ++u) {
      const int s = sb + u;
      raw[u] = (s >= 0 && s < T) ? __ldg(reinterpret_cast<const uint32_t*>(in + static_cast<size_t>(s) * C)) : 0u;  // swish(0) = 0
    }
    if (ref != nullptr) {
#pragma unroll
      for (int u = 0; u < kDwBatch; ++u) {
        const int t = sb + u + dt;
        rr[u] = (t >= t0 && t < t_end) ? __ldg(reinterpret_cast<const uint32_t*>(ref + static_cast<size_t>(t) * C)) : 0u;
      }
    }
#pragma unroll
    for (int u = 0; u < kDwBatch; ++u) {
#pragma unroll
      for (int j = 0; j < K - 1; ++j) { x0[j] = x0[j + 1]; x1[j] = x1[j + 1]; }
      float v0 = bf16_lo(raw[u]), v1 = bf16_hi(raw[u]);
      if (a.pre_act == TACT_SWISH) { v0 = swish_f(v0); v1 = swish_f(v1); }
      x0[K - 1] = v0;
      x1[K - 1] = v1;
      const int t = sb + u + dt;
      if (t >= t0 && t < t_end) {
        float y0 = b0, y1 = b1;
#pragma unroll
        for (int j = 0; j < K; ++j) { y0 = fmaf(w0[j], x0[j], y0); y1 = fmaf(w1[j], x1[j], y1); }
        if (ref != nullptr) {
          y0 *= swish_d(bf16_lo(rr[u]));
          y1 *= swish_d(bf16_hi(rr[u]));
        }
        *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(t) * C) = pack_bf16x2(y0, y1);
      }
    }
  }
}

template <int K>
__global__ void __launch_bounds__(kDwThreads) dw_wgrad_kernel(const bf16* __restrict__ dOut, const bf16* __restrict__ in_all, int pre_act,
                                                               float* __restrict__ dw, float* __restrict__ dbias, int T, int C, int pad_left) {
  const int c = (blockIdx.x * kDwThreads + threadIdx.x) * 2;
  if (c >= C) return;
  const int t0 = blockIdx.y * kDwWgChunk, b = blockIdx.z;
  const bf16* __restrict__ in = in_all + static_cast<size_t>(b) * T * C + c;
  const bf16* __restrict__ dy = dOut + static_cast<size_t>(b) * T * C + c;
  float x0[K], x1[K], g0[K], g1[K];
#pragma unroll
  for (int j = 0; j < K; ++j) { x0[j] = x1[j] = g0[j] = g1[j] = 0.f; }
  float sb0 = 0.f, sb1 = 0.f;
  const int t_end = t0 + kDwWgChunk < T ? t0 + kDwWgChunk : T;  // exclusive
  const int s_begin = t0 - pad_left, s_end = t_end - 1 - pad_left + (K - 1);
  const int dt = pad_left - (K - 1);
  for (int sb = s_begin; sb <= s_end; sb += kDwBatch) {
    uint32_t raw[kDwBatch], dd[kDwBatch];
#pragma unroll
    for (int u = 0; u < kDwBatch; ++u) {
      const int s = sb + u, t = sb + u + dt;
      raw[u] = (s >= 0 && s < T) ? __ldg(reinterpret_cast<const uint32_t*>(in + static_cast<size_t>(s) * C)) : 0u;
      dd[u] = (t >= t0 && t < t_end) ? __ldg(reinterpret_cast<const uint32_t*>(dy + static_cast<size_t>(t) * C)) : 0u;  // 0 gradient outside
    }
#pragma unroll
    for (int u = 0; u < kDwBatch; ++u) {
#pragma unroll
      for (int j = 0; j < K - 1; ++j) { x0[j] = x0[j + 1]; x1[j] = x1[j + 1]; }
      float v0 = bf16_lo(raw[u]), v1 = bf16_hi(raw[u]);
      if (pre_act == TACT_SWISH) { v0 = swish_f(v0); v1 = swish_f(v1); }
      x0[K - 1] = v0;
      x1[K - 1] = v1;
      const float d0 = bf16_lo(dd[u]), d1 = bf16_hi(dd[u]);
      sb0 += d0;
      sb1 += d1;
#pragma unroll
      for (int j = 0; j < K; ++j) { g0[j] = fmaf(d0, x0[j], g0[j]); g1[j] = fmaf(d1, x1[j], g1[j]); }
    }
  }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    atomicAdd(&dw[static_cast<size_t>(j) * C + c], g0[j]);
    atomicAdd(&dw[static_cast<size_t>(j) * C + c + 1], g1[j]);
  }
  if (dbias != nullptr) {
    atomicAdd(&dbias[c], sb0);
    atomicAdd(&dbias[c + 1], sb1);
  }
}

int check(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error(std::string(what) + ": " + cudaGetErrorString(e));
    return 3;
  }
  note_launch();
  return 0;
}

}  // namespace

int dw_train_launch(const DwTrainArgs& a, cudaStream_t s) {
  if (a.C % 2 != 0 || a.k < 1) { set_last_error("dw_train: C must be even"); return 2; }
  const dim3 grid((a.C / 2 + kDwThreads - 1) / kDwThreads, (a.T + kDwChunk - 1) / kDwChunk, a.B);
  switch (a.k) {
#define DW_CASE(KK) case KK: dw_train_kernel<KK><<<grid, kDwThreads, 0, s>>>(a); break;
    DW_CASE(1) DW_CASE(3) DW_CASE(5) DW_CASE(7) DW_CASE(9) DW_CASE(11) DW_CASE(13) DW_CASE(15) DW_CASE(17) DW_CASE(31)
#undef DW_CASE
    default: set_last_error("dw_train: kernel size must be one of 1,3,5,7,9,11,13,15,17,31"); return 2;
  }
  return check("dw_train");
}

int dw_wgrad_launch(const bf16* dOut, const bf16* in, int pre_act, float* dw, float* dbias, int B, int T, int C, int k,
                    int pad_left, cudaStream_t s) {
  if (C % 2 != 0 || k < 1) { set_last_error("dw_wgrad: C must be even"); return 2; }
  const dim3 grid((C / 2 + kDwThreads - 1) / kDwThreads, (T + kDwWgChunk - 1) / kDwWgChunk, B);
  switch (k) {
#define DW_CASE(KK) case KK: dw_wgrad_kernel<KK><<<grid, kDwThreads, 0, s>>>(dOut, in, pre_act, dw, dbias, T, C, pad_left); break;
    DW_CASE(1) DW_CASE(3) DW_CASE(5) DW_CASE(7) DW_CASE(9) DW_CASE(11) DW_CASE(13) DW_CASE(15) DW_CASE(17) DW_CASE(31)
#undef DW_CASE
    default: set_last_error("dw_wgrad: kernel size must be one of 1,3,5,7,9,11,13,15,17,31"); return 2;
  }
  return check("dw_wgrad");
}

}  // namespace ishara
