// ishara_b200 — optimiser step of the training path (BASELINE: AdamW lr 4.5e-3, weight decay
// 0.08, global-norm clip 1.0 — integration.py:675-679,750): squared-norm reduction, fused clip + AdamW update on
// the flat fp32 master buffer, and the re-quantisation of every dense kernel into the two bf16 operand layouts the
// tcgen05 GEMMs consume ([out,in] for the forward product, [in,out] for the data gradient).
#include <cstdio>

#include "ptx.cuh"
#include "train_kernels.h"

namespace ishara {
namespace {

__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * 256) {
    const float v = g[i];
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += static_cast<double>(red[w]);
    atomicAdd(out, s);
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                      int64_t n, const double* __restrict__ norm2, AdamWArgs a, float bc1, float bc2) {
  float clip = 1.f;
  {
    // a NaN / inf gradient (e.g. an infeasible CTC alignment: nll = +inf) must not reach theta, m and v: drop the update
    const float total = static_cast<float>(sqrt(*norm2)) * a.grad_scale;
    if (!isfinite(total)) {
      if (blockIdx.x == 0 && threadIdx.x == 0 && a.skipped != nullptr) atomicAdd(a.skipped, 1);
      return;
    }
    if (a.max_norm > 0.f) clip = fminf(1.f, a.max_norm / (total + 1e-6f));
  }
  const float gs = a.grad_scale * clip, decay = 1.f - a.lr * a.weight_decay, step = a.lr / bc1, rbc2 = rsqrtf(bc2);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * 256) {
    const float gi = g[i] * gs;
    const float mi = a.beta1 * m[i] + (1.f - a.beta1) * gi;
    const float vi = a.beta2 * v[i] + (1.f - a.beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    theta[i] = theta[i] * decay - step * mi / (sqrtf(vi) * rbc2 + a.eps);
  }
}

// RectifiedAdam (Liu et al. 2019, as implemented by tensorflow_addons/optimizers/rectified_adam.py, total_steps = 0 so no
// warm-up schedule) followed by the Lookahead slow-weight step (tensorflow_addons/optimizers/lookahead.py):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; mhat = m / (1-b1^t) ; vhat = sqrt(v / (1-b2^t))
//   sma_inf = 2/(1-b2) - 1 ; sma_t = sma_inf - 2 t b2^t / (1-b2^t)
//   r_t = sqrt((sma_t-4)/(sma_inf-4) * (sma_t-2)/(sma_inf-2) * sma_inf/sma_t)
//   upd = sma_t >= sma_threshold ? r_t mhat / (vhat + eps) : mhat ;  upd += wd * theta ;  theta -= lr * upd
//   every sync_period steps: slow += alpha (theta - slow) ; theta = slow     (slow starts as the initial weights)
// rect / use_rect / sync are step-only quantities computed on the host.
__global__ void __launch_bounds__(256) radam_lookahead_kernel(float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ m,
                                                                float* __restrict__ v, float* __restrict__ slow, int64_t n,
                                                                const double* __restrict__ norm2, RAdamArgs a, float bc1, float bc2, float rect,
                                                                int use_rect, int sync) {
  float clip = 1.f;
  {
    const float total = static_cast<float>(sqrt(*norm2)) * a.grad_scale;
    if (!isfinite(total)) {
      if (blockIdx.x == 0 && threadIdx.x == 0 && a.skipped != nullptr) atomicAdd(a.skipped, 1);
      return;
    }
    if (a.max_norm > 0.f) clip = fminf(1.f, a.max_norm / (total + 1e-6f));
  }
  const float gs = a.grad_scale * clip, rbc1 = 1.f / bc1, rbc2 = 1.f / bc2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * 256) {
    const float gi = g[i] * gs;
    const float mi = a.beta1 * m[i] + (1.f - a.beta1) * gi;
    const float vi = a.beta2 * v[i] + (1.f - a.beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float mhat = mi * rbc1;
    float upd = use_rect ? rect * mhat / (sqrtf(vi * rbc2) + a.eps) : mhat;
    float th = theta[i];
    upd += a.weight_decay * th;
    th -= a.lr * upd;
    if (sync) {
      const float sl = slow[i] + a.slow_step * (th - slow[i]);
      slow[i] = sl;
      th = sl;
    }
    theta[i] = th;
  }
}

// one CTA = one 32x32 tile of one dense kernel
__global__ void __launch_bounds__(256) repack_kernel(const RepackEntry* __restrict__ table) {
  __shared__ float tile[32][33];
  const RepackEntry e = table[blockIdx.y];
  const int ti = (e.Ipad + 31) / 32, to = (e.Opad + 31) / 32;
  if (static_cast<int>(blockIdx.x) >= ti * to) return;
  const int i0 = (blockIdx.x / to) * 32, o0 = (blockIdx.x % to) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, o = o0 + tx;
    const float val = (i < e.I && o < e.O) ? e.src[static_cast<size_t>(i) * e.O + o] : 0.f;
    tile[r][tx] = val;
    if (e.bwd != nullptr && i < e.I && o < e.Opad) e.bwd[static_cast<size_t>(i) * e.Opad + o] = __float2bfloat16_rn(val);
  }
  __syncthreads();
  if (e.fwd != nullptr) {
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int o = o0 + r, i = i0 + tx;
      if (o < e.Opad && i < e.Ipad) e.fwd[static_cast<size_t>(o) * e.Ipad + i] = __float2bfloat16_rn(tile[tx][r]);
    }
  }
}

}  // namespace

int sqnorm_launch(const float* g, int64_t n, double* norm2, cudaStream_t s) {
  sqnorm_kernel<<<148 * 4, 256, 0, s>>>(g, n, norm2);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}
int adamw_launch(float* theta, const float* g, float* m, float* v, int64_t n, const double* norm2, const AdamWArgs& a,
                 cudaStream_t s) {
  const float bc1 = 1.f - powf(a.beta1, static_cast<float>(a.step)), bc2 = 1.f - powf(a.beta2, static_cast<float>(a.step));
  adamw_kernel<<<148 * 4, 256, 0, s>>>(theta, g, m, v, n, norm2, a, bc1, bc2);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}
int radam_lookahead_launch(float* theta, const float* g, float* m, float* v, float* slow, int64_t n, const double* norm2,
                           const RAdamArgs& a, cudaStream_t s) {
  const double t = static_cast<double>(a.step), b2 = a.beta2;
  const double b2t = pow(b2, t);
  const double sma_inf = 2.0 / (1.0 - b2) - 1.0;
  const double sma_t = sma_inf - 2.0 * t * b2t / (1.0 - b2t);
  const int use_rect = sma_t >= static_cast<double>(a.sma_threshold) ? 1 : 0;
  double rect = 0.0;
  if (use_rect) rect = sqrt((sma_t - 4.0) / (sma_inf - 4.0) * (sma_t - 2.0) / (sma_inf - 2.0) * sma_inf / sma_t);
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(a.beta1), t)), bc2 = static_cast<float>(1.0 - b2t);
  const int sync = a.sync_period > 0 && (a.step % a.sync_period) == 0 ? 1 : 0;
  radam_lookahead_kernel<<<148 * 4, 256, 0, s>>>(theta, g, m, v, slow, n, norm2, a, bc1, bc2, static_cast<float>(rect), use_rect, sync);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}
int repack_launch(const RepackEntry* table_dev, int n_entries, int max_tiles, cudaStream_t s) {
  if (n_entries <= 0) return 0;
  repack_kernel<<<dim3(max_tiles, n_entries), 256, 0, s>>>(table_dev);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace ishara
