// ishara_b200 — time down/up-sampling kernels of the vendored Squeezeformer (operator-level rows P7/P8 of SURVEY.md §8a).
//
//   time_reduce      TimeReductionLayer (squeezeformer/convolution.py:241-269): the [T, D] activation is treated as a
//                    1-channel image, Conv2d(1->1, k=3, stride 2, no padding, +bias) then Swish
//                    => [(T-3)/2+1, (D-3)/2+1]. Output rows are zero-padded to `ldo` columns so the result is directly the
//                    K-padded A operand of the time_reduction_proj GEMM (encoder.py:80,155).
//   upsample_add     recover step (encoder.py:157-162 + modules.py:137-142): nearest-neighbour x2 repeat along time of
//                    the projected low-rate stream, plus the saved high-rate tensor: out[t] = y[t/2] + rec[t].
//                    (Linear commutes with the row repeat, so time_recover_layer runs on the low-rate rows first.)
//   conv2d_subsample DepthwiseConv2dSubsampling (convolution.py:39-73): Conv2d(1->C,3,s2)+ReLU, depthwise Conv2d(C,3,s2)
//                    +ReLU, output laid out [B, T', C*F'] (channel-major, then frequency) like the reference's permute/view.
// All three are memory-bound gathers: one thread per output element pair, coalesced along the channel axis.
#include "kernels.h"
#include "ptx.cuh"

namespace ishara {
namespace {

__device__ __forceinline__ float ldbf(const bf16* p) { return __bfloat162float(*p); }

__global__ void __launch_bounds__(128)
time_reduce_kernel(TimeReduceArgs a) {
  const int t2 = blockIdx.x, b = blockIdx.y;
  const bf16* x = a.x + (static_cast<size_t>(b) * a.T + 2 * t2) * a.D;
  bf16* o = a.out + (static_cast<size_t>(b) * a.T2 + t2) * a.ldo;
  for (int d2 = threadIdx.x; d2 < a.ldo; d2 += blockDim.x) {
    float v = 0.f;
    if (d2 < a.D2) {
      float acc = a.bias;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc = fmaf(a.w[i * 3 + j], ldbf(x + static_cast<size_t>(i) * a.D + 2 * d2 + j), acc);
      v = acc / (1.f + __expf(-acc));
    }
    o[d2] = __float2bfloat16(v);
  }
}

__global__ void __launch_bounds__(256)
upsample_add_kernel(const bf16* __restrict__ y, const bf16* __restrict__ rec, bf16* __restrict__ out, int B, int T2, int T,
                    int D) {
  const int64_t pairs = static_cast<int64_t>(B) * 2 * T2 * (D / 2);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < pairs;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (D / 2));
    const int64_t row = i / (D / 2);
    const int t = static_cast<int>(row % (2 * T2));
    const int b = static_cast<int>(row / (2 * T2));
    const uint32_t u = reinterpret_cast<const uint32_t*>(y + (static_cast<size_t>(b) * T2 + t / 2) * D)[c];
    const uint32_t r = reinterpret_cast<const uint32_t*>(rec + (static_cast<size_t>(b) * T + t) * D)[c];
    reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * 2 * T2 + t) * D)[c] =
        pack_bf16x2(bf16_lo(u) + bf16_lo(r), bf16_hi(u) + bf16_hi(r));
  }
}

__global__ void __launch_bounds__(128)
conv2d_subsample_kernel(Conv2dSubsampleArgs a) {
  const int t4 = blockIdx.x, b = blockIdx.y;
  const int n = a.C * a.F4;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i / a.F4, f4 = i - c * a.F4;
    const float* w1 = a.w1 + c * 9;
    const float* w2 = a.w2 + c * 9;
    float acc2 = a.b2[c];
    for (int p = 0; p < 3; ++p)
      for (int q = 0; q < 3; ++q) {
        const int t2 = 2 * t4 + p, f2 = 2 * f4 + q;
        float inner = a.b1[c];
        const float* x = a.x + (static_cast<size_t>(b) * a.T + 2 * t2) * a.F + 2 * f2;
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
          for (int v = 0; v < 3; ++v) inner = fmaf(w1[u * 3 + v], x[static_cast<size_t>(u) * a.F + v], inner);
        acc2 = fmaf(w2[p * 3 + q], fmaxf(inner, 0.f), acc2);
      }
    a.out[(static_cast<size_t>(b) * a.T4 + t4) * a.ldo + i] = __float2bfloat16(fmaxf(acc2, 0.f));
  }
  for (int i = n + threadIdx.x; i < a.ldo; i += blockDim.x)
    a.out[(static_cast<size_t>(b) * a.T4 + t4) * a.ldo + i] = __float2bfloat16(0.f);
}

}  // namespace

int time_reduce_launch(const TimeReduceArgs& a, cudaStream_t stream) {
  if (a.T < 3 || a.D < 3 || a.T2 != (a.T - 3) / 2 + 1 || a.D2 != (a.D - 3) / 2 + 1 || a.ldo < a.D2) {
    set_last_error("time_reduce: need T,D >= 3, T2 = (T-3)/2+1, D2 = (D-3)/2+1, ldo >= D2");
    return 2;
  }
  time_reduce_kernel<<<dim3(a.T2, a.B), 128, 0, stream>>>(a);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

int upsample_add_launch(const bf16* y, const bf16* rec, bf16* out, int B, int T2, int T, int D, cudaStream_t stream) {
  if (D % 2 != 0 || 2 * T2 > T || B <= 0) {
    set_last_error("upsample_add: D must be even and 2*T2 <= T");
    return 2;
  }
  const int64_t pairs = static_cast<int64_t>(B) * 2 * T2 * (D / 2);
  int grid = static_cast<int>((pairs + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  upsample_add_kernel<<<grid, 256, 0, stream>>>(y, rec, out, B, T2, T, D);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

int conv2d_subsample_launch(const Conv2dSubsampleArgs& a, cudaStream_t stream) {
  const int T2 = (a.T - 3) / 2 + 1, F2 = (a.F - 3) / 2 + 1;
  if (a.T < 7 || a.F < 7 || a.T4 != (T2 - 3) / 2 + 1 || a.F4 != (F2 - 3) / 2 + 1 || a.ldo < a.C * a.F4) {
    set_last_error("conv2d_subsample: inconsistent output extents");
    return 2;
  }
  conv2d_subsample_kernel<<<dim3(a.T4, a.B), 128, 0, stream>>>(a);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace ishara
