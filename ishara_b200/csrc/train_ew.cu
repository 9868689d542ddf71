// ishara_b200 — memory-bound kernels of the training step: activation forward/backward,
// BatchNorm with batch statistics (Keras training mode: biased variance over (B,T), c5:73, c5:281, c7:17), ECA and
// SqueezeExcite gates and their gradients (c5:1-15, c5:120-133), LayerNorm backward, bias-gradient column sums,
// counter-based dropout. All are HBM-bound streaming passes over bf16 [M, C] tensors (16-byte vector accesses,
// grid sized to a few CTAs per SM); the reductions accumulate per-CTA fp32 partials into fp64 / fp32 atomics.
#include <cstdio>

#include "dropout_hash.cuh"
#include "ptx.cuh"
#include "train_kernels.h"

namespace ishara {
namespace {

constexpr int kEwThreads = 256;

__device__ __forceinline__ void ld8(const bf16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ void st8(bf16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float swish_f(float x) { return x * sigm(x); }
__device__ __forceinline__ float swish_d(float x) {
  const float s = sigm(x);
  return s * (1.f + x * (1.f - s));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline int ew_grid(int64_t nvec) {
  const int64_t b = (nvec + kEwThreads - 1) / kEwThreads;
  return static_cast<int>(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}
#define EW_LOOP(v, nvec) \
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * kEwThreads + threadIdx.x; v < (nvec); v += static_cast<int64_t>(gridDim.x) * kEwThreads)

// ---- activations ------------------------------------------------------------------------------------
// 8 keep factors (inv_keep or 0) of element vector v: the same bits dropout_kernel draws for a tensor of the same shape
__device__ __forceinline__ void drop_factors8(uint64_t key, int64_t v, uint32_t thr16, float inv_keep, float (&k)[8]) {
  const uint64_t r0 = mix64(key + 2ull * static_cast<uint64_t>(v)), r1 = mix64(key + 2ull * static_cast<uint64_t>(v) + 1ull);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t u = static_cast<uint32_t>(((i < 4 ? r0 : r1) >> (16 * (i & 3))) & 0xFFFFu);
    k[i] = u >= thr16 ? inv_keep : 0.f;
  }
}
// out = drop(act(in)): key_ptr == nullptr -> no dropout (one pass instead of activation + in-place dropout)
__global__ void __launch_bounds__(kEwThreads) act_fwd_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int64_t nvec, int act,
                                                               uint32_t thr16, float inv_keep, const uint64_t* __restrict__ key_ptr) {
  const uint64_t key = key_ptr != nullptr ? *key_ptr : 0ull;
  EW_LOOP(v, nvec) {
    float f[8];
    ld8(in + v * 8, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = act == TACT_SWISH ? swish_f(f[i]) : (act == TACT_RELU ? fmaxf(f[i], 0.f) : f[i]);
    if (key_ptr != nullptr) {
      float k[8];
      drop_factors8(key, v, thr16, inv_keep, k);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] *= k[i];
    }
    st8(out + v * 8, f);
  }
}
// dU = act'(ref) * drop'(dH): key_ptr == nullptr -> no dropout
__global__ void __launch_bounds__(kEwThreads) act_bwd_kernel(const bf16* __restrict__ dH, const bf16* __restrict__ ref, bf16* __restrict__ dU,
                                                               int64_t nvec, int act, uint32_t thr16, float inv_keep,
                                                               const uint64_t* __restrict__ key_ptr) {
  const uint64_t key = key_ptr != nullptr ? *key_ptr : 0ull;
  EW_LOOP(v, nvec) {
    float d[8], r[8];
    ld8(dH + v * 8, d);
    ld8(ref + v * 8, r);
    if (key_ptr != nullptr) {
      float k[8];
      drop_factors8(key, v, thr16, inv_keep, k);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] *= k[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = act == TACT_SWISH ? d[i] * swish_d(r[i]) : (act == TACT_RELU ? (r[i] > 0.f ? d[i] : 0.f) : d[i]);
    st8(dU + v * 8, d);
  }
}
__global__ void __launch_bounds__(kEwThreads) glu_fwd_kernel(const bf16* __restrict__ P, bf16* __restrict__ out, int64_t M, int C) {
  const int c8 = C / 8;
  EW_LOOP(v, M * c8) {
    const int64_t row = v / c8;
    const int c = static_cast<int>(v % c8) * 8;
    float a[8], b[8];
    ld8(P + row * 2 * C + c, a);
    ld8(P + row * 2 * C + C + c, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] *= sigm(b[i]);
    st8(out + row * C + c, a);
  }
}
__global__ void __launch_bounds__(kEwThreads) glu_bwd_kernel(const bf16* __restrict__ dOut, const bf16* __restrict__ P, bf16* __restrict__ dP,
                                                               int64_t M, int C) {
  const int c8 = C / 8;
  EW_LOOP(v, M * c8) {
    const int64_t row = v / c8;
    const int c = static_cast<int>(v % c8) * 8;
    float a[8], b[8], d[8], da[8], db[8];
    ld8(P + row * 2 * C + c, a);
    ld8(P + row * 2 * C + C + c, b);
    ld8(dOut + row * C + c, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float s = sigm(b[i]);
      da[i] = d[i] * s;
      db[i] = d[i] * a[i] * s * (1.f - s);
    }
    st8(dP + row * 2 * C + c, da);
    st8(dP + row * 2 * C + C + c, db);
  }
}
__global__ void __launch_bounds__(kEwThreads) affine_gate_add_kernel(const bf16* __restrict__ x, const float* __restrict__ scale,
                                                                       const float* __restrict__ shift, const float* __restrict__ gate,
                                                                       const bf16* __restrict__ resid, bf16* __restrict__ y, int64_t M, int C, int T) {
  const int c8 = C / 8;
  EW_LOOP(v, M * c8) {
    const int64_t row = v / c8;
    const int c = static_cast<int>(v % c8) * 8;
    float f[8];
    ld8(x + row * C + c, f);
    if (scale != nullptr) {
      float sc[8], sh[8];
      ld8f(scale + c, sc);
      ld8f(shift + c, sh);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], sc[i], sh[i]);
    }
    if (gate != nullptr) {
      float g[8];
      ld8f(gate + (row / T) * C + c, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] *= g[i];
    }
    if (resid != nullptr) {
      float r[8];
      ld8(resid + row * C + c, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += r[i];
    }
    st8(y + row * C + c, f);
  }
}
__global__ void __launch_bounds__(kEwThreads) gate_bias_kernel(const bf16* __restrict__ x, const float* __restrict__ g, const float* __restrict__ a,
                                                                 float alpha, bf16* __restrict__ y, int64_t M, int C, int T) {
  const int c8 = C / 8;
  EW_LOOP(v, M * c8) {
    const int64_t row = v / c8;
    const int c = static_cast<int>(v % c8) * 8;
    float f[8], gv[8], av[8];
    ld8(x + row * C + c, f);
    ld8f(g + (row / T) * C + c, gv);
    ld8f(a + (row / T) * C + c, av);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], gv[i], av[i] * alpha);
    st8(y + row * C + c, f);
  }
}
__global__ void __launch_bounds__(kEwThreads) scale_cast_pad_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t M, int V, int Vpad,
                                                                      float alpha) {
  const int c8 = Vpad / 8;
  EW_LOOP(v, M * c8) {
    const int64_t row = v / c8;
    const int c = static_cast<int>(v % c8) * 8;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = (c + i) < V ? in[row * V + c + i] * alpha : 0.f;
    st8(out + row * Vpad + c, f);
  }
}

// ---- dropout ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads) dropout_kernel(const bf16* __restrict__ x, const bf16* __restrict__ resid, bf16* __restrict__ y, int64_t M,
                                                               int C, int T, uint32_t thr16, float inv_keep, uint64_t key_val,
                                                               const uint64_t* __restrict__ key_ptr, int per_sample) {
  const uint64_t key = key_ptr != nullptr ? *key_ptr : key_val;  // key table on the device: the launch is graph-replayable
  const int c8 = C / 8;
  EW_LOOP(v, M * c8) {
    const int64_t row = v / c8;
    const int c = static_cast<int>(v % c8) * 8;
    float f[8];
    ld8(x + row * C + c, f);
    if (per_sample) {
      const uint64_t r = mix64(key + static_cast<uint64_t>(row / T));
      const float k = (static_cast<uint32_t>(r & 0xFFFFu) >= thr16) ? inv_keep : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] *= k;
    } else {
      const uint64_t r0 = mix64(key + 2ull * static_cast<uint64_t>(v)), r1 = mix64(key + 2ull * static_cast<uint64_t>(v) + 1ull);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t u = static_cast<uint32_t>(((i < 4 ? r0 : r1) >> (16 * (i & 3))) & 0xFFFFu);
        f[i] = u >= thr16 ? f[i] * inv_keep : 0.f;
      }
    }
    if (resid != nullptr) {
      float r[8];
      ld8(resid + row * C + c, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += r[i];
    }
    st8(y + row * C + c, f);
  }
}

// ---- per-sequence column reductions: CTA = (64-channel slab, sequence), 8 column groups x 32 row lanes -----------
template <int NQ>
__device__ __forceinline__ void slab_reduce(float (&acc)[NQ][8], float (*red)[32][65], int cg, int rl) {
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[q][rl][cg * 8 + i] = acc[q][i];
  __syncthreads();
}

__global__ void __launch_bounds__(kEwThreads) colstats_kernel(const bf16* __restrict__ x, float* __restrict__ seqsum, double* __restrict__ sum,
                                                                double* __restrict__ sumsq, int T, int C) {
  __shared__ float red[2][32][65];
  const int b = blockIdx.y, c0 = blockIdx.x * 64, cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  float acc[2][8] = {};
  const bf16* base = x + static_cast<size_t>(b) * T * C + c0 + cg * 8;
  for (int t = rl; t < T; t += 32) {
    float f[8];
    ld8(base + static_cast<size_t>(t) * C, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[0][i] += f[i]; acc[1][i] = fmaf(f[i], f[i], acc[1][i]); }
  }
  slab_reduce<2>(acc, red, cg, rl);
  if (threadIdx.x < 128) {
    const int q = threadIdx.x >> 6, c = threadIdx.x & 63;
    float a = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) a += red[q][r][c];
    if (q == 0) {
      seqsum[static_cast<size_t>(b) * C + c0 + c] = a;
      if (sum != nullptr) atomicAdd(&sum[c0 + c], static_cast<double>(a));
    } else if (sumsq != nullptr) {
      atomicAdd(&sumsq[c0 + c], static_cast<double>(a));
    }
  }
}

__global__ void __launch_bounds__(kEwThreads) seq_dot_kernel(const bf16* __restrict__ dG, const bf16* __restrict__ x, const float* __restrict__ scale,
                                                               const float* __restrict__ shift, float* __restrict__ out, int T, int C) {
  __shared__ float red[1][32][65];
  const int b = blockIdx.y, c0 = blockIdx.x * 64, cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  float acc[1][8] = {};
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = 1.f; sh[i] = 0.f; }
  if (scale != nullptr) { ld8f(scale + c0 + cg * 8, sc); ld8f(shift + c0 + cg * 8, sh); }
  const size_t off = static_cast<size_t>(b) * T * C + c0 + cg * 8;
  for (int t = rl; t < T; t += 32) {
    float f[8], d[8];
    ld8(x + off + static_cast<size_t>(t) * C, f);
    ld8(dG + off + static_cast<size_t>(t) * C, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[0][i] = fmaf(d[i], fmaf(f[i], sc[i], sh[i]), acc[0][i]);
  }
  slab_reduce<1>(acc, red, cg, rl);
  if (threadIdx.x < 64) {
    float a = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) a += red[0][r][threadIdx.x];
    out[static_cast<size_t>(b) * C + c0 + threadIdx.x] = a;
  }
}

__global__ void __launch_bounds__(kEwThreads) bn_bwd_reduce_kernel(const bf16* __restrict__ dG, const bf16* __restrict__ x, const float* __restrict__ sgate,
                                                                     const float* __restrict__ dm, float invT, const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd, double* __restrict__ sum1, double* __restrict__ sum2,
                                                                     int T, int C) {
  __shared__ float red[2][32][65];
  const int b = blockIdx.y, c0 = blockIdx.x * 64, cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  float acc[2][8] = {};
  float sg[8], dmv[8], mu[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sg[i] = 1.f; dmv[i] = 0.f; }
  if (sgate != nullptr) ld8f(sgate + static_cast<size_t>(b) * C + c0 + cg * 8, sg);
  if (dm != nullptr) {
    ld8f(dm + static_cast<size_t>(b) * C + c0 + cg * 8, dmv);
#pragma unroll
    for (int i = 0; i < 8; ++i) dmv[i] *= invT;
  }
  ld8f(mean + c0 + cg * 8, mu);
  ld8f(rstd + c0 + cg * 8, rs);
  const size_t off = static_cast<size_t>(b) * T * C + c0 + cg * 8;
  for (int t = rl; t < T; t += 32) {
    float f[8], d[8];
    ld8(x + off + static_cast<size_t>(t) * C, f);
    ld8(dG + off + static_cast<size_t>(t) * C, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float dbn = fmaf(d[i], sg[i], dmv[i]);
      acc[0][i] += dbn;
      acc[1][i] = fmaf(dbn, (f[i] - mu[i]) * rs[i], acc[1][i]);
    }
  }
  slab_reduce<2>(acc, red, cg, rl);
  if (threadIdx.x < 128) {
    const int q = threadIdx.x >> 6, c = threadIdx.x & 63;
    float a = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) a += red[q][r][c];
    atomicAdd(q == 0 ? &sum1[c0 + c] : &sum2[c0 + c], static_cast<double>(a));
  }
}

// per-channel constants of the BatchNorm backward (fp64 sums -> fp32 once, instead of per element) + dgamma / dbeta
__global__ void bn_bwd_coeff_kernel(const double* __restrict__ sum1, const double* __restrict__ sum2, double inv_count, float* __restrict__ coef,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s1 = sum1[c], s2 = sum2[c];
  coef[c] = static_cast<float>(s1 * inv_count);
  coef[C + c] = static_cast<float>(s2 * inv_count);
  dgamma[c] += static_cast<float>(s2);
  dbeta[c] += static_cast<float>(s1);
}

__global__ void __launch_bounds__(kEwThreads) bn_bwd_apply_kernel(const bf16* __restrict__ dG, const bf16* __restrict__ x, const float* __restrict__ sgate,
                                                                    const float* __restrict__ dm, float invT, const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                    const float* __restrict__ coef, bf16* __restrict__ dx, int64_t M, int C, int T) {
  const int c8 = C / 8;
  EW_LOOP(v, M * c8) {
    const int64_t row = v / c8;
    const int c = static_cast<int>(v % c8) * 8;
    float f[8], d[8], mu[8], rs[8], ga[8], s1[8], s2[8];
    ld8(x + row * C + c, f);
    ld8(dG + row * C + c, d);
    ld8f(mean + c, mu);
    ld8f(rstd + c, rs);
    ld8f(gamma + c, ga);
    ld8f(coef + c, s1);
    ld8f(coef + C + c, s2);
    if (sgate != nullptr) {
      float sg[8];
      ld8f(sgate + (row / T) * C + c, sg);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] *= sg[i];
    }
    if (dm != nullptr) {
      float dv[8];
      ld8f(dm + (row / T) * C + c, dv);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = fmaf(dv[i], invT, d[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (f[i] - mu[i]) * rs[i];
      d[i] = ga[i] * rs[i] * (d[i] - s1[i] - xh * s2[i]);
    }
    st8(dx + row * C + c, d);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, double inv_count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* __restrict__ moving_mean,
                                   float* __restrict__ moving_var, float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ scale,
                                   float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = sum[c] * inv_count;
  double var = sumsq[c] * inv_count - mu * mu;
  if (var < 0.0) var = 0.0;
  const float r = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  mean[c] = static_cast<float>(mu);
  rstd[c] = r;
  const float sc = gamma[c] * r;
  scale[c] = sc;
  shift[c] = beta[c] - static_cast<float>(mu) * sc;
  moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * static_cast<float>(mu);
  moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * static_cast<float>(var);
}

// ---- ECA gate over [B, C] ------------------------------------------------------------------------------
__device__ __forceinline__ float eca_mean(const float* seqsum, const float* scale, const float* shift, float invT, int b, int c, int C) {
  return (c < 0 || c >= C) ? 0.f : fmaf(seqsum[static_cast<size_t>(b) * C + c] * invT, scale[c], shift[c]);
}
__global__ void eca_fwd_kernel(const float* __restrict__ seqsum, const float* __restrict__ scale, const float* __restrict__ shift, float invT,
                               const float* __restrict__ w5, float* __restrict__ m, float* __restrict__ sgate, int B, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  float u = 0.f;
#pragma unroll
  for (int j = 0; j < 5; ++j) u = fmaf(w5[j], eca_mean(seqsum, scale, shift, invT, b, c + j - 2, C), u);
  m[i] = eca_mean(seqsum, scale, shift, invT, b, c, C);
  sgate[i] = sigm(u);
}
__global__ void eca_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ sgate, const float* __restrict__ m, const float* __restrict__ w5,
                               float* __restrict__ dm, float* __restrict__ dw5, int B, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = i < B * C;
  const int b = ok ? i / C : 0, c = ok ? i % C : 0;
  auto du_at = [&](int cc) -> float {
    if (cc < 0 || cc >= C) return 0.f;
    const size_t k = static_cast<size_t>(b) * C + cc;
    const float s = sgate[k];
    return ds[k] * s * (1.f - s);
  };
  float acc = 0.f, dw[5];
  const float du = ok ? du_at(c) : 0.f;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    acc = fmaf(w5[j], ok ? du_at(c - j + 2) : 0.f, acc);       // u[c'] = sum_j w[j] m[c'+j-2]  =>  dm[c] = sum_j w[j] du[c-j+2]
    const int cm = c + j - 2;
    dw[j] = (ok && cm >= 0 && cm < C) ? du * m[static_cast<size_t>(b) * C + cm] : 0.f;
  }
  if (ok) dm[i] = acc;
  __shared__ float wred[8][5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const float s = warp_sum(dw[j]);
    if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += wred[w][threadIdx.x];
    atomicAdd(&dw5[threadIdx.x], s);
  }
}

// ---- LayerNorm backward: a warp takes two rows per iteration (their loads are issued together), lane owns the
// 4-column groups (j*32 + lane)*4 ------------------------------------------------------------------------------------
template <int NJ>
__global__ void __launch_bounds__(256, 2) ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma, float eps,
                                                          const bf16* __restrict__ dresid, bf16* __restrict__ dx, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, int64_t M, int D) {
  __shared__ float red[2][8][NJ * 128];
  constexpr int nj = NJ;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float invD = 1.f / static_cast<float>(D);
  float gam[NJ][4], ag[NJ][4] = {}, ab[NJ][4] = {};
#pragma unroll
  for (int j = 0; j < NJ; ++j)
    if (j < nj) {
      const float4 g4 = *reinterpret_cast<const float4*>(gamma + (j * 32 + lane) * 4);
      gam[j][0] = g4.x; gam[j][1] = g4.y; gam[j][2] = g4.z; gam[j][3] = g4.w;
    }
  for (int64_t row0 = (static_cast<int64_t>(blockIdx.x) * 8 + warp) * 2; row0 < M; row0 += static_cast<int64_t>(gridDim.x) * 16) {
    uint2 ux[2][NJ], ud[2][NJ], ur[2][NJ];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (j < nj) {
          const bool ok = row0 + r < M;
          const size_t o = static_cast<size_t>(ok ? row0 + r : row0) * D + (j * 32 + lane) * 4;
          ux[r][j] = *reinterpret_cast<const uint2*>(x + o);
          ud[r][j] = *reinterpret_cast<const uint2*>(dy + o);
          if (dresid != nullptr) ur[r][j] = *reinterpret_cast<const uint2*>(dresid + o);
        }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (row0 + r >= M) break;  // warp-uniform
      float xv[NJ][4], dv[NJ][4];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (j < nj) {
          xv[j][0] = bf16_lo(ux[r][j].x); xv[j][1] = bf16_hi(ux[r][j].x); xv[j][2] = bf16_lo(ux[r][j].y); xv[j][3] = bf16_hi(ux[r][j].y);
          dv[j][0] = bf16_lo(ud[r][j].x); dv[j][1] = bf16_hi(ud[r][j].x); dv[j][2] = bf16_lo(ud[r][j].y); dv[j][3] = bf16_hi(ud[r][j].y);
          s += (xv[j][0] + xv[j][1]) + (xv[j][2] + xv[j][3]);
        }
      const float mean = warp_sum(s) * invD;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (j < nj)
#pragma unroll
          for (int e = 0; e < 4; ++e) { xv[j][e] -= mean; q = fmaf(xv[j][e], xv[j][e], q); }
      const float rstd = rsqrtf(warp_sum(q) * invD + eps);
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (j < nj)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float xh = xv[j][e] * rstd;
            xv[j][e] = xh;
            ag[j][e] = fmaf(dv[j][e], xh, ag[j][e]);
            ab[j][e] += dv[j][e];
            const float dxh = dv[j][e] * gam[j][e];
            dv[j][e] = dxh;
            m1 += dxh;
            m2 = fmaf(dxh, xh, m2);
          }
      m1 = warp_sum(m1) * invD;
      m2 = warp_sum(m2) * invD;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (j < nj) {
          const size_t o = static_cast<size_t>(row0 + r) * D + (j * 32 + lane) * 4;
          float rr[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) rr[e] = rstd * (dv[j][e] - m1 - xv[j][e] * m2);
          if (dresid != nullptr) {
            rr[0] += bf16_lo(ur[r][j].x); rr[1] += bf16_hi(ur[r][j].x); rr[2] += bf16_lo(ur[r][j].y); rr[3] += bf16_hi(ur[r][j].y);
          }
          uint2 uo;
          uo.x = pack_bf16x2(rr[0], rr[1]);
          uo.y = pack_bf16x2(rr[2], rr[3]);
          *reinterpret_cast<uint2*>(dx + o) = uo;
        }
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j)
    if (j < nj)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        red[0][warp][(j * 32 + lane) * 4 + e] = ag[j][e];
        red[1][warp][(j * 32 + lane) * 4 + e] = ab[j][e];
      }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[0][w][c]; b += red[1][w][c]; }
    atomicAdd(&dgamma[c], a);
    atomicAdd(&dbeta[c], b);
  }
}

// ---- bias gradient: column sums of g [M, ld] for the first Cvalid columns -------------------------------------
__global__ void __launch_bounds__(kEwThreads) colsum_kernel(const bf16* __restrict__ g, int ld, float* __restrict__ out, int64_t M, int Cvalid,
                                                              int rows_per_cta) {
  __shared__ float red[1][32][65];
  const int c0 = blockIdx.x * 64, cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  float acc[1][8] = {};
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  if (c0 + cg * 8 < ld) {
    for (int64_t r = r0 + rl; r < r1; r += 32) {
      float f[8];
      ld8(g + r * ld + c0 + cg * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[0][i] += f[i];
    }
  }
  slab_reduce<1>(acc, red, cg, rl);
  if (threadIdx.x < 64 && c0 + threadIdx.x < Cvalid) {
    float a = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) a += red[0][r][threadIdx.x];
    atomicAdd(&out[c0 + threadIdx.x], a);
  }
}

// ---- SqueezeExcite dense layers: one CTA per sequence ------------------------------------------------------
__global__ void __launch_bounds__(256) se_fwd_kernel(SeTrainArgs a) {
  __shared__ float g[512], act[64];
  const int b = blockIdx.x, D = a.D, R = a.R, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < D; c += 256) {
    const float v = a.zsum[static_cast<size_t>(b) * D + c] * a.invT;
    g[c] = v;
    a.g[static_cast<size_t>(b) * D + c] = v;
  }
  __syncthreads();
  for (int r = warp; r < R; r += 8) {
    float p = 0.f;
    for (int c = lane; c < D; c += 32) p = fmaf(g[c], a.fc1_w[static_cast<size_t>(c) * R + r], p);
    p = warp_sum(p) + a.fc1_b[r];
    if (lane == 0) {
      a.a_pre[static_cast<size_t>(b) * R + r] = p;
      act[r] = swish_f(p);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    float p = a.fc2_b[c];
    for (int r = 0; r < R; ++r) p = fmaf(act[r], a.fc2_w[static_cast<size_t>(r) * D + c], p);
    a.gate[static_cast<size_t>(b) * D + c] = sigm(p);
  }
}
__global__ void __launch_bounds__(256) se_bwd_kernel(SeTrainArgs a) {
  __shared__ float dgp[512], act[64], dap[64];
  const int b = blockIdx.x, D = a.D, R = a.R, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = threadIdx.x; r < R; r += 256) act[r] = swish_f(a.a_pre[static_cast<size_t>(b) * R + r]);
  for (int c = threadIdx.x; c < D; c += 256) {
    const float gt = a.gate[static_cast<size_t>(b) * D + c];
    dgp[c] = a.dgate[static_cast<size_t>(b) * D + c] * gt * (1.f - gt);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    const float d = dgp[c];
    atomicAdd(&a.d_fc2_b[c], d);
    for (int r = 0; r < R; ++r) atomicAdd(&a.d_fc2_w[static_cast<size_t>(r) * D + c], act[r] * d);
  }
  for (int r = warp; r < R; r += 8) {
    float p = 0.f;
    for (int c = lane; c < D; c += 32) p = fmaf(a.fc2_w[static_cast<size_t>(r) * D + c], dgp[c], p);
    p = warp_sum(p);
    if (lane == 0) {
      const float d = p * swish_d(a.a_pre[static_cast<size_t>(b) * R + r]);
      dap[r] = d;
      atomicAdd(&a.d_fc1_b[r], d);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    const float gc = a.g[static_cast<size_t>(b) * D + c];
    float p = 0.f;
    for (int r = 0; r < R; ++r) {
      atomicAdd(&a.d_fc1_w[static_cast<size_t>(c) * R + r], gc * dap[r]);
      p = fmaf(a.fc1_w[static_cast<size_t>(c) * R + r], dap[r], p);
    }
    a.dg[static_cast<size_t>(b) * D + c] = p;
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
#define REQUIRE(cond, msg)           \
  do {                               \
    if (!(cond)) {                   \
      set_last_error(msg);           \
      return 2;                      \
    }                                \
  } while (0)

}  // namespace

int act_fwd_launch(const bf16* in, bf16* out, int64_t n, int act, cudaStream_t s) {
  REQUIRE(n % 8 == 0, "act_fwd: n % 8 != 0");
  act_fwd_kernel<<<ew_grid(n / 8), kEwThreads, 0, s>>>(in, out, n / 8, act, 0u, 1.f, nullptr);
  return check("act_fwd");
}
int act_drop_fwd_launch(const bf16* in, bf16* out, int64_t n, int act, float p, const uint64_t* key_dev, cudaStream_t s) {
  REQUIRE(n % 8 == 0 && p >= 0.f && p < 1.f && key_dev != nullptr, "act_drop_fwd: bad arguments");
  act_fwd_kernel<<<ew_grid(n / 8), kEwThreads, 0, s>>>(in, out, n / 8, act, dropout_thr16(p), 1.f / (1.f - p), key_dev);
  return check("act_drop_fwd");
}
int act_bwd_launch(const bf16* dH, const bf16* ref, bf16* dU, int64_t n, int act, cudaStream_t s) {
  REQUIRE(n % 8 == 0, "act_bwd: n % 8 != 0");
  act_bwd_kernel<<<ew_grid(n / 8), kEwThreads, 0, s>>>(dH, ref, dU, n / 8, act, 0u, 1.f, nullptr);
  return check("act_bwd");
}
int act_bwd_drop_launch(const bf16* dH, const bf16* ref, bf16* dU, int64_t n, int act, float p, const uint64_t* key_dev, cudaStream_t s) {
  REQUIRE(n % 8 == 0 && p >= 0.f && p < 1.f && key_dev != nullptr, "act_bwd_drop: bad arguments");
  act_bwd_kernel<<<ew_grid(n / 8), kEwThreads, 0, s>>>(dH, ref, dU, n / 8, act, dropout_thr16(p), 1.f / (1.f - p), key_dev);
  return check("act_bwd_drop");
}
int glu_fwd_launch(const bf16* P, bf16* out, int64_t M, int C, cudaStream_t s) {
  REQUIRE(C % 8 == 0, "glu_fwd: C % 8 != 0");
  glu_fwd_kernel<<<ew_grid(M * C / 8), kEwThreads, 0, s>>>(P, out, M, C);
  return check("glu_fwd");
}
int glu_bwd_launch(const bf16* dOut, const bf16* P, bf16* dP, int64_t M, int C, cudaStream_t s) {
  REQUIRE(C % 8 == 0, "glu_bwd: C % 8 != 0");
  glu_bwd_kernel<<<ew_grid(M * C / 8), kEwThreads, 0, s>>>(dOut, P, dP, M, C);
  return check("glu_bwd");
}
int affine_gate_add_launch(const bf16* x, const float* scale, const float* shift, const float* gate, const bf16* resid,
                           bf16* y, int64_t M, int C, int T, cudaStream_t s) {
  REQUIRE(C % 8 == 0, "affine_gate_add: C % 8 != 0");
  affine_gate_add_kernel<<<ew_grid(M * C / 8), kEwThreads, 0, s>>>(x, scale, shift, gate, resid, y, M, C, T);
  return check("affine_gate_add");
}
int gate_bias_launch(const bf16* x, const float* g, const float* a, float alpha, bf16* y, int64_t M, int C, int T,
                     cudaStream_t s) {
  REQUIRE(C % 8 == 0, "gate_bias: C % 8 != 0");
  gate_bias_kernel<<<ew_grid(M * C / 8), kEwThreads, 0, s>>>(x, g, a, alpha, y, M, C, T);
  return check("gate_bias");
}
int scale_cast_pad_launch(const float* in, bf16* out, int64_t M, int V, int Vpad, float alpha, cudaStream_t s) {
  REQUIRE(Vpad % 8 == 0 && Vpad >= V, "scale_cast_pad: bad padding");
  scale_cast_pad_kernel<<<ew_grid(M * Vpad / 8), kEwThreads, 0, s>>>(in, out, M, V, Vpad, alpha);
  return check("scale_cast_pad");
}
int dropout_launch(const bf16* x, const bf16* resid, bf16* y, int64_t M, int C, int T, float p, uint64_t seed,
                   uint32_t site, int per_sample, cudaStream_t s) {
  REQUIRE(C % 8 == 0 && p >= 0.f && p < 1.f, "dropout: bad arguments");
  const uint32_t thr = dropout_thr16(p);
  const uint64_t key = dropout_key(seed, site);
  dropout_kernel<<<ew_grid(M * C / 8), kEwThreads, 0, s>>>(x, resid, y, M, C, T, thr, 1.f / (1.f - p), key, nullptr, per_sample);
  return check("dropout");
}
// same with the site's key read from device memory at run time (dropout_keys_launch fills the table every step)
int dropout_keyed_launch(const bf16* x, const bf16* resid, bf16* y, int64_t M, int C, int T, float p, const uint64_t* key_dev,
                         int per_sample, cudaStream_t s) {
  REQUIRE(C % 8 == 0 && p >= 0.f && p < 1.f && key_dev != nullptr, "dropout: bad arguments");
  dropout_kernel<<<ew_grid(M * C / 8), kEwThreads, 0, s>>>(x, resid, y, M, C, T, dropout_thr16(p), 1.f / (1.f - p), 0ull, key_dev, per_sample);
  return check("dropout");
}
__global__ void dropout_keys_kernel(uint64_t seed, uint64_t* __restrict__ keys, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = dropout_key(seed, static_cast<uint32_t>(i));
}
// keys[site] = dropout_key(seed, site) for site < n: one tiny launch per step in front of the (replayed) program
int dropout_keys_launch(uint64_t seed, uint64_t* keys_dev, int n, cudaStream_t s) {
  dropout_keys_kernel<<<(n + 127) / 128, 128, 0, s>>>(seed, keys_dev, n);
  return check("dropout_keys");
}
int colstats_launch(const bf16* x, float* seqsum, double* sum, double* sumsq, int B, int T, int C, cudaStream_t s) {
  REQUIRE(C % 64 == 0, "colstats: C % 64 != 0");
  colstats_kernel<<<dim3(C / 64, B), kEwThreads, 0, s>>>(x, seqsum, sum, sumsq, T, C);
  return check("colstats");
}
int seq_dot_launch(const bf16* dG, const bf16* x, const float* scale, const float* shift, float* out, int B, int T,
                   int C, cudaStream_t s) {
  REQUIRE(C % 64 == 0, "seq_dot: C % 64 != 0");
  seq_dot_kernel<<<dim3(C / 64, B), kEwThreads, 0, s>>>(dG, x, scale, shift, out, T, C);
  return check("seq_dot");
}
int bn_finalize_launch(const double* sum, const double* sumsq, double count, const float* gamma, const float* beta,
                       float eps, float momentum, float* moving_mean, float* moving_var, float* mean, float* rstd,
                       float* scale, float* shift, int C, cudaStream_t s) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(sum, sumsq, 1.0 / count, gamma, beta, eps, momentum, moving_mean,
                                                     moving_var, mean, rstd, scale, shift, C);
  return check("bn_finalize");
}
int eca_fwd_launch(const float* seqsum, const float* scale, const float* shift, float invT, const float* w5, float* m,
                   float* sgate, int B, int C, cudaStream_t s) {
  eca_fwd_kernel<<<(B * C + 255) / 256, 256, 0, s>>>(seqsum, scale, shift, invT, w5, m, sgate, B, C);
  return check("eca_fwd");
}
int eca_bwd_launch(const float* ds, const float* sgate, const float* m, const float* w5, float* dm, float* dw5, int B,
                   int C, cudaStream_t s) {
  eca_bwd_kernel<<<(B * C + 255) / 256, 256, 0, s>>>(ds, sgate, m, w5, dm, dw5, B, C);
  return check("eca_bwd");
}
int bn_bwd_reduce_launch(const bf16* dG, const bf16* x, const float* sgate, const float* dm, float invT,
                         const float* mean, const float* rstd, double* sum1, double* sum2, int B, int T, int C,
                         cudaStream_t s) {
  REQUIRE(C % 64 == 0, "bn_bwd_reduce: C % 64 != 0");
  bn_bwd_reduce_kernel<<<dim3(C / 64, B), kEwThreads, 0, s>>>(dG, x, sgate, dm, invT, mean, rstd, sum1, sum2, T, C);
  return check("bn_bwd_reduce");
}
int bn_bwd_apply_launch(const bf16* dG, const bf16* x, const float* sgate, const float* dm, float invT,
                        const float* mean, const float* rstd, const float* gamma, const double* sum1,
                        const double* sum2, double count, float* coef, bf16* dx, float* dgamma, float* dbeta, int B, int T, int C,
                        cudaStream_t s) {
  REQUIRE(C % 8 == 0, "bn_bwd_apply: C % 8 != 0");
  const int64_t M = static_cast<int64_t>(B) * T;
  bn_bwd_coeff_kernel<<<(C + 127) / 128, 128, 0, s>>>(sum1, sum2, 1.0 / count, coef, dgamma, dbeta, C);
  int r = check("bn_bwd_coeff");
  if (r) return r;
  bn_bwd_apply_kernel<<<ew_grid(M * C / 8), kEwThreads, 0, s>>>(dG, x, sgate, dm, invT, mean, rstd, gamma, coef, dx, M, C, T);
  return check("bn_bwd_apply");
}
int ln_bwd_launch(const bf16* dy, const bf16* x, const float* gamma, float eps, const bf16* dresid, bf16* dx,
                  float* dgamma, float* dbeta, int64_t M, int D, cudaStream_t s) {
  REQUIRE(D % 128 == 0 && D <= 512, "ln_bwd: D must be a multiple of 128, <= 512");
  const int64_t want = (M + 15) / 16;
  const int grid = static_cast<int>(want < 148 * 2 ? want : 148 * 2);  // 2 resident CTAs per SM: fewer, fatter CTAs = fewer gamma/beta atomics
  switch (D / 128) {
    case 1: ln_bwd_kernel<1><<<grid, 256, 0, s>>>(dy, x, gamma, eps, dresid, dx, dgamma, dbeta, M, D); break;
    case 2: ln_bwd_kernel<2><<<grid, 256, 0, s>>>(dy, x, gamma, eps, dresid, dx, dgamma, dbeta, M, D); break;
    case 3: ln_bwd_kernel<3><<<grid, 256, 0, s>>>(dy, x, gamma, eps, dresid, dx, dgamma, dbeta, M, D); break;
    default: ln_bwd_kernel<4><<<grid, 256, 0, s>>>(dy, x, gamma, eps, dresid, dx, dgamma, dbeta, M, D); break;
  }
  return check("ln_bwd");
}
int colsum_launch(const bf16* g, int ld, float* out, int64_t M, int Cvalid, cudaStream_t s) {
  REQUIRE(ld % 8 == 0 && Cvalid <= ld, "colsum: bad shape");
  const int rows_per_cta = 1024;
  colsum_kernel<<<dim3((Cvalid + 63) / 64, static_cast<unsigned>((M + rows_per_cta - 1) / rows_per_cta)), kEwThreads, 0, s>>>(
      g, ld, out, M, Cvalid, rows_per_cta);
  return check("colsum");
}
int se_fwd_launch(const SeTrainArgs& a, cudaStream_t s) {
  REQUIRE(a.D <= 512 && a.R <= 64, "se_fwd: D <= 512, R <= 64");
  se_fwd_kernel<<<a.B, 256, 0, s>>>(a);
  return check("se_fwd");
}
int se_bwd_launch(const SeTrainArgs& a, cudaStream_t s) {
  REQUIRE(a.D <= 512 && a.R <= 64, "se_bwd: D <= 512, R <= 64");
  se_bwd_kernel<<<a.B, 256, 0, s>>>(a);
  return check("se_bwd");
}

}  // namespace ishara
