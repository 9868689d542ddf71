// ishara_b200 — depthwise temporal convolution for the training step (SURVEY.md §8 rows T4/T8/T11 under T15):
// forward with an optional swish on the input (so the PRE-activation tensor is the one kept for backward), the
// data gradient (same kernel: flipped taps, mirrored padding, optional multiply by swish'(pre-activation)) and the
// tap/bias gradient. CausalDWConv1D c5:17-39 (pad_left = k-1), ConvModule.conv2 c5:142 (causal, k = 15),
// ConvolutionModule.depthwise_conv c5:272-279 ('same', pad_left = (k-1)/2, with bias).
//
// HBM-bound stencils in the CTA shape of the inference kernel (dwconv.cu): one CTA = (sequence, 64-channel slab) with the
// whole time axis staged in shared memory (swish applied once while staging), register-blocked packed-fp32x2 taps.
// (The first version walked time per thread straight from global memory: latency-bound at 33-75 us per launch.)
#include <cstdio>

#include "ptx.cuh"
#include "train_kernels.h"

namespace ishara {
namespace {

constexpr int kSlab = 64;     // channels per CTA
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float swish_f(float x) { return x * sigm(x); }
__device__ __forceinline__ float swish_d(float x) {
  const float s = sigm(x);
  return s * (1.f + x * (1.f - s));
}
__device__ __forceinline__ uint32_t swish2(uint32_t u) { return pack_bf16x2(swish_f(bf16_lo(u)), swish_f(bf16_hi(u))); }

template <int K>
__host__ __device__ constexpr int dwt_tb() { return K >= 9 ? 4 : 8; }

// Stage rows [t0 - pad_left, t0 + rows_alloc - pad_left) of one (sequence, 64-channel slab) into smem as bf16x2 words, zero
// outside [0, T); with pre_act the swish is applied HERE, once per element (not once per tap or per halo re-read).
__device__ __forceinline__ void stage_slab(uint32_t* tile, const bf16* src, int T, int C, int t0, int pad_left, int rows_alloc, int pre_act) {
  for (int i = threadIdx.x; i < rows_alloc * 8; i += kThreads) {
    const int r = i >> 3, ch = i & 7;
    const int t = t0 + r - pad_left;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (t >= 0 && t < T) {
      v = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(t) * C) + ch);
      if (pre_act == TACT_SWISH) { v.x = swish2(v.x); v.y = swish2(v.y); v.z = swish2(v.z); v.w = swish2(v.w); }
    }
    reinterpret_cast<uint4*>(tile)[i] = v;
  }
}

// Same CTA shape as the inference kernel (dwconv.cu): one CTA = (sequence, 64-channel slab), the whole time axis of the
// slab in shared memory, a lane owns one bf16x2 channel pair and walks its warp's time range in register blocks with
// packed fp32x2 FMAs.
template <int K>
__global__ void __launch_bounds__(kThreads, (K <= 11 ? 3 : (K <= 17 ? 2 : 1)))
dw_train_kernel(DwTrainArgs a, int rows_alloc, int tchunk) {
  constexpr int kTB = dwt_tb<K>();
  extern __shared__ __align__(16) uint8_t smem_dwt[];
  uint32_t* tile = reinterpret_cast<uint32_t*>(smem_dwt);  // [rows_alloc][32]
  const int b = blockIdx.y, c0 = blockIdx.x * kSlab;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T, C = a.C;
  const int t0 = blockIdx.z * tchunk;  // the time axis is cut into chunks so small batches still fill the GPU
  stage_slab(tile, a.in + static_cast<size_t>(b) * T * C + c0, T, C, t0, a.pad_left, rows_alloc, a.pre_act);
  float2 wt[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wt[j] = __ldg(reinterpret_cast<const float2*>(a.w + static_cast<size_t>(a.flip ? K - 1 - j : j) * C + c0) + lane);
  float2 bs = make_float2(0.f, 0.f);
  if (a.bias != nullptr) bs = __ldg(reinterpret_cast<const float2*>(a.bias + c0) + lane);
  __syncthreads();

  const int rows_per_warp = tchunk / kWarps;
  const int t_begin = t0 + warp * rows_per_warp;
  const int t_end = min(T, t_begin + rows_per_warp);
  const uint32_t* trow = tile + (t_begin - t0) * 32 + lane;  // smem row r holds input row t0 + r - pad_left: tap j of output t reads row t - t0 + j
  const int cw = C >> 1;
  const size_t gofs = (static_cast<size_t>(b) * T + t_begin) * cw + (c0 >> 1) + lane;
  uint32_t* drow = reinterpret_cast<uint32_t*>(a.out) + gofs;
  const uint32_t* rrow = a.mul_ref != nullptr ? reinterpret_cast<const uint32_t*>(a.mul_ref) + gofs : nullptr;
  for (int t = t_begin; t < t_end; t += kTB) {
    float2 x[kTB + K - 1];
#pragma unroll
    for (int i = 0; i < kTB + K - 1; ++i) {
      const uint32_t u = trow[i * 32];
      x[i] = make_float2(bf16_lo(u), bf16_hi(u));
    }
    uint32_t rf[kTB];
    if (rrow != nullptr) {
#pragma unroll
      for (int i = 0; i < kTB; ++i) rf[i] = (t + i < t_end) ? __ldg(rrow + static_cast<size_t>(i) * cw) : 0u;
    }
#pragma unroll
    for (int i = 0; i < kTB; ++i) {
      float2 y = bs;
#pragma unroll
      for (int j = 0; j < K; ++j) ffma2(y.x, y.y, wt[j].x, wt[j].y, x[i + j].x, x[i + j].y, y.x, y.y);
      if (rrow != nullptr) { y.x *= swish_d(bf16_lo(rf[i])); y.y *= swish_d(bf16_hi(rf[i])); }
      if (t + i < t_end) drow[static_cast<size_t>(i) * cw] = pack_bf16x2(y.x, y.y);
    }
    trow += kTB * 32;
    drow += static_cast<size_t>(kTB) * cw;
    if (rrow != nullptr) rrow += static_cast<size_t>(kTB) * cw;
  }
}

// Tap / bias gradient: same staging (input with swish applied once), each warp accumulates K packed partial sums over
// its time range, the eight warps are reduced through shared memory, then 64 x K atomics per CTA.
template <int K>
__global__ void __launch_bounds__(kThreads, (K <= 11 ? 3 : (K <= 17 ? 2 : 1)))
dw_wgrad_kernel(const bf16* __restrict__ dOut, const bf16* __restrict__ in_all, int pre_act, float* __restrict__ dw, float* __restrict__ dbias,
                int T, int C, int pad_left, int rows_alloc, int tchunk) {
  constexpr int kTB = dwt_tb<K>();
  extern __shared__ __align__(16) uint8_t smem_dwt[];
  uint32_t* tile = reinterpret_cast<uint32_t*>(smem_dwt);
  const int b = blockIdx.y, c0 = blockIdx.x * kSlab;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.z * tchunk;
  stage_slab(tile, in_all + static_cast<size_t>(b) * T * C + c0, T, C, t0, pad_left, rows_alloc, pre_act);
  __syncthreads();
  const int rows_per_warp = tchunk / kWarps;
  const int t_begin = t0 + warp * rows_per_warp;
  const int t_end = min(T, t_begin + rows_per_warp);
  const uint32_t* trow = tile + (t_begin - t0) * 32 + lane;
  const int cw = C >> 1;
  const uint32_t* grow = reinterpret_cast<const uint32_t*>(dOut) + (static_cast<size_t>(b) * T + t_begin) * cw + (c0 >> 1) + lane;
  float2 g[K];
#pragma unroll
  for (int j = 0; j < K; ++j) g[j] = make_float2(0.f, 0.f);
  float2 sb = make_float2(0.f, 0.f);
  for (int t = t_begin; t < t_end; t += kTB) {
    float2 x[kTB + K - 1], d[kTB];
#pragma unroll
    for (int i = 0; i < kTB; ++i) {
      const uint32_t u = (t + i < t_end) ? __ldg(grow + static_cast<size_t>(i) * cw) : 0u;
      d[i] = make_float2(bf16_lo(u), bf16_hi(u));
    }
#pragma unroll
    for (int i = 0; i < kTB + K - 1; ++i) {
      const uint32_t u = trow[i * 32];
      x[i] = make_float2(bf16_lo(u), bf16_hi(u));
    }
#pragma unroll
    for (int i = 0; i < kTB; ++i) {
      fadd2(sb.x, sb.y, sb.x, sb.y, d[i].x, d[i].y);
#pragma unroll
      for (int j = 0; j < K; ++j) ffma2(g[j].x, g[j].y, d[i].x, d[i].y, x[i + j].x, x[i + j].y, g[j].x, g[j].y);
    }
    trow += kTB * 32;
    grow += static_cast<size_t>(kTB) * cw;
  }
  __syncthreads();  // every warp is done with the tile: reuse it for the cross-warp reduction
  float* red = reinterpret_cast<float*>(smem_dwt);  // [kWarps][K + 1][64]
#pragma unroll
  for (int j = 0; j < K; ++j) {
    red[(warp * (K + 1) + j) * kSlab + 2 * lane] = g[j].x;
    red[(warp * (K + 1) + j) * kSlab + 2 * lane + 1] = g[j].y;
  }
  red[(warp * (K + 1) + K) * kSlab + 2 * lane] = sb.x;
  red[(warp * (K + 1) + K) * kSlab + 2 * lane + 1] = sb.y;
  __syncthreads();
  for (int i = threadIdx.x; i < (K + 1) * kSlab; i += kThreads) {
    const int j = i / kSlab, c = i % kSlab;
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < kWarps; ++wv) s += red[(wv * (K + 1) + j) * kSlab + c];
    if (j < K) atomicAdd(&dw[static_cast<size_t>(j) * C + c0 + c], s);
    else if (dbias != nullptr) atomicAdd(&dbias[c0 + c], s);
  }
}

template <int K>
int dwt_smem(int T, int* rows_alloc, int* tchunk) {
  constexpr int kTB = dwt_tb<K>();
  const int want = T < 128 ? T : 128;  // outputs per CTA
  const int rows_per_warp = ((want + kWarps - 1) / kWarps + kTB - 1) / kTB * kTB;
  *tchunk = rows_per_warp * kWarps;
  *rows_alloc = *tchunk + K - 1;
  const int tile_bytes = *rows_alloc * 128, red_bytes = kWarps * (K + 1) * kSlab * 4;
  return tile_bytes > red_bytes ? tile_bytes : red_bytes;
}

template <int K>
int launch_fwd(const DwTrainArgs& a, cudaStream_t s) {
  int rows_alloc, tchunk;
  const int smem = dwt_smem<K>(a.T, &rows_alloc, &tchunk);
  static int attr = 0;
  if (smem > attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(dw_train_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = smem;
  }
  dw_train_kernel<K><<<dim3(a.C / kSlab, a.B, (a.T + tchunk - 1) / tchunk), kThreads, smem, s>>>(a, rows_alloc, tchunk);
  return 0;
}
template <int K>
int launch_wgrad(const bf16* dOut, const bf16* in, int pre_act, float* dw, float* dbias, int B, int T, int C, int pad_left, cudaStream_t s) {
  int rows_alloc, tchunk;
  const int smem = dwt_smem<K>(T, &rows_alloc, &tchunk);
  static int attr = 0;
  if (smem > attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(dw_wgrad_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = smem;
  }
  dw_wgrad_kernel<K><<<dim3(C / kSlab, B, (T + tchunk - 1) / tchunk), kThreads, smem, s>>>(dOut, in, pre_act, dw, dbias, T, C, pad_left, rows_alloc,
                                                                                    tchunk);
  return 0;
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
  if (a.C % kSlab != 0 || a.k < 1 || a.T < 1) { set_last_error("dw_train: C must be a multiple of 64"); return 2; }
  int rc = 0;
  switch (a.k) {
#define DW_CASE(KK) case KK: rc = launch_fwd<KK>(a, s); break;
    DW_CASE(1) DW_CASE(3) DW_CASE(5) DW_CASE(7) DW_CASE(9) DW_CASE(11) DW_CASE(13) DW_CASE(15) DW_CASE(17) DW_CASE(31)
#undef DW_CASE
    default: set_last_error("dw_train: kernel size must be one of 1,3,5,7,9,11,13,15,17,31"); return 2;
  }
  if (rc) return rc;
  return check("dw_train");
}

int dw_wgrad_launch(const bf16* dOut, const bf16* in, int pre_act, float* dw, float* dbias, int B, int T, int C, int k,
                    int pad_left, cudaStream_t s) {
  if (C % kSlab != 0 || k < 1 || T < 1) { set_last_error("dw_wgrad: C must be a multiple of 64"); return 2; }
  int rc = 0;
  switch (k) {
#define DW_CASE(KK) case KK: rc = launch_wgrad<KK>(dOut, in, pre_act, dw, dbias, B, T, C, pad_left, s); break;
    DW_CASE(1) DW_CASE(3) DW_CASE(5) DW_CASE(7) DW_CASE(9) DW_CASE(11) DW_CASE(13) DW_CASE(15) DW_CASE(17) DW_CASE(31)
#undef DW_CASE
    default: set_last_error("dw_wgrad: kernel size must be one of 1,3,5,7,9,11,13,15,17,31"); return 2;
  }
  if (rc) return rc;
  return check("dw_wgrad");
}

}  // namespace ishara
