// ishara_b200 — backward of the multi-head self-attention core (MultiHeadSelfAttention.call, nb:conv-hybrid-model
// c5:102-118): given qkv (per-head interleaved), the forward output o and dO, produce
// dqkv in the same interleaved layout so the QKV projection's gradient GEMMs consume it directly.
//
// Flash-style recomputation, two kernels per (sequence, head), no atomics and no [T,T] tensor in HBM:
//   1. attn_bwd_dq : warp = 16 query rows. Sweep 1 recomputes the row log-sum-exp, sweep 2 forms P, dP = dO V^T,
//      dS = P o (dP - rowsum(dO o O)) and accumulates dQ = scale * dS K. Also writes lse and the row sums.
//   2. attn_bwd_dkv: warp = 16 key rows, everything transposed (S^T = K Q^T ...) so P^T and dS^T come out of the MMA
//      in the layout the next MMA wants: dV = P^T dO, dK = scale * dS^T Q.
// Tensor work on mma.sync m16n8k16 (K = dh = 32 contractions are too small for tcgen05 tiles); K/V (kernel 1) and
// Q/dO (kernel 2) are staged once per CTA in shared memory, row-major and transposed.
#include <cstdio>

#include "dropout_hash.cuh"
#include "mma.cuh"
#include "ptx.cuh"
#include "train_kernels.h"

namespace ishara {
namespace {

constexpr int kAbThreads = 256;
constexpr int kAbChunk = 64;
constexpr float kLog2e = 1.4426950408889634f;

// stage rows [0,Tp) of a [T, ld]-pitched bf16 matrix slice (DH columns at `src`) row-major into R [Tp][DH+8] and, when
// Tt != null, transposed into Tt [DH][Tp+8]; rows >= T are zero.
template <int DH>
__device__ __forceinline__ void stage_rows(const bf16* src, size_t ld, int T, int Tp, bf16* R, bf16* Tt) {
  constexpr int KS = DH + 8;
  const int VS = Tp + 8;
  for (int t = threadIdx.x; t < Tp; t += kAbThreads) {
    uint4 v[DH / 8];
    if (t < T) {
      const uint4* p = reinterpret_cast<const uint4*>(src + static_cast<size_t>(t) * ld);
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) v[i] = __ldg(p + i);
    } else {
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) v[i] = make_uint4(0, 0, 0, 0);
    }
    if (R != nullptr) {
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) *reinterpret_cast<uint4*>(R + static_cast<size_t>(t) * KS + 8 * i) = v[i];
    }
    if (Tt != nullptr) {
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) {
        const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          Tt[static_cast<size_t>(8 * i + 2 * e) * VS + t] = __ushort_as_bfloat16(static_cast<unsigned short>(w[e] & 0xFFFFu));
          Tt[static_cast<size_t>(8 * i + 2 * e + 1) * VS + t] = __ushort_as_bfloat16(static_cast<unsigned short>(w[e] >> 16));
        }
      }
    }
  }
}

// A-operand fragments (16 rows x DH) straight from global memory: rows r0 / r1 of a matrix with pitch ld
template <int DH>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[DH / 16][4], const bf16* base, size_t ld, int r0, int r1, int T, int tg) {
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) {
    const int c = kk * 16 + tg * 2;
    a[kk][0] = r0 < T ? __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(r0) * ld + c)) : 0u;
    a[kk][1] = r1 < T ? __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(r1) * ld + c)) : 0u;
    a[kk][2] = r0 < T ? __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(r0) * ld + c + 8)) : 0u;
    a[kk][3] = r1 < T ? __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(r1) * ld + c + 8)) : 0u;
  }
}

// acc[n] (16 x 64 strip, n = 8-column tile) = A(16 x DH) . R[rows c0 + ..][DH]^T  with R row-major [*, DH+8]
template <int DH>
__device__ __forceinline__ void strip_mma(float (&acc)[kAbChunk / 8][4], const uint32_t (&a)[DH / 16][4], const uint32_t* R32, int c0, int g,
                                          int tg) {
  constexpr int KS = DH + 8;
#pragma unroll
  for (int n = 0; n < kAbChunk / 8; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk)
#pragma unroll
    for (int n = 0; n < kAbChunk / 8; ++n) {
      const int row = c0 + n * 8 + g;
      mma16816(acc[n], a[kk], R32[(row * KS + kk * 16 + tg * 2) >> 1], R32[(row * KS + kk * 16 + 8 + tg * 2) >> 1]);
    }
}

// out[n] (16 x DH) += Pfrag(16 x 64, C-fragment layout values in p) . Tt^T, Tt = [DH][VS] transposed operand, rows c0..c0+63
template <int DH>
__device__ __forceinline__ void strip_acc(float (&out)[DH / 8][4], const float (&p)[kAbChunk / 8][4], const uint32_t* Tt32, int VS, int c0, int g,
                                          int tg) {
#pragma unroll
  for (int kk = 0; kk < kAbChunk / 16; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    pa[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
    pa[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    pa[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      const int d = n * 8 + g;
      const int r = c0 + kk * 16 + tg * 2;
      mma16816(out[n], pa, Tt32[(d * VS + r) >> 1], Tt32[(d * VS + r + 8) >> 1]);
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(kAbThreads)
attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ dO, bf16* __restrict__ dqkv,
                   float* __restrict__ lse_out, float* __restrict__ dsum_out, int T, int H, int Tp, float scale, uint32_t drop_thr16,
                   float drop_inv_keep, uint64_t drop_key_val, const uint64_t* __restrict__ drop_key_ptr, int have_lse) {
  const uint64_t drop_key = drop_key_ptr != nullptr ? *drop_key_ptr : drop_key_val;
  constexpr int KS = DH + 8;
  extern __shared__ __align__(16) uint8_t smem_ab[];
  const int VS = Tp + 8;
  bf16* Ks = reinterpret_cast<bf16*>(smem_ab);               // [Tp][KS]
  bf16* Vs = Ks + static_cast<size_t>(Tp) * KS;              // [Tp][KS]
  bf16* Kt = Vs + static_cast<size_t>(Tp) * KS;              // [DH][VS]
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const size_t ld = static_cast<size_t>(3) * DH * H, ldo = static_cast<size_t>(DH) * H;
  const bf16* qbase = qkv + static_cast<size_t>(b) * T * ld + static_cast<size_t>(h) * 3 * DH;
  const bf16* obase = o + static_cast<size_t>(b) * T * ldo + static_cast<size_t>(h) * DH;
  const bf16* dobase = dO + static_cast<size_t>(b) * T * ldo + static_cast<size_t>(h) * DH;
  stage_rows<DH>(qbase + DH, ld, T, Tp, Ks, Kt);
  stage_rows<DH>(qbase + 2 * DH, ld, T, Tp, Vs, nullptr);
  __syncthreads();
  const uint32_t* Ks32 = reinterpret_cast<const uint32_t*>(Ks);
  const uint32_t* Vs32 = reinterpret_cast<const uint32_t*>(Vs);
  const uint32_t* Kt32 = reinterpret_cast<const uint32_t*>(Kt);
  const float scale_log2 = scale * kLog2e;

  for (int q0 = warp * 16; q0 < T; q0 += (kAbThreads / 32) * 16) {
    const int r0 = q0 + g, r1 = q0 + g + 8;
    uint32_t qa[DH / 16][4], da[DH / 16][4], oa[DH / 16][4];
    load_a_frags<DH>(qa, qbase, ld, r0, r1, T, tg);
    load_a_frags<DH>(da, dobase, ldo, r0, r1, T, tg);
    load_a_frags<DH>(oa, obase, ldo, r0, r1, T, tg);
    float d0 = 0.f, d1 = 0.f;  // rowsum(dO o O)
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      d0 += bf16_lo(da[kk][0]) * bf16_lo(oa[kk][0]) + bf16_hi(da[kk][0]) * bf16_hi(oa[kk][0]) +
            bf16_lo(da[kk][2]) * bf16_lo(oa[kk][2]) + bf16_hi(da[kk][2]) * bf16_hi(oa[kk][2]);
      d1 += bf16_lo(da[kk][1]) * bf16_lo(oa[kk][1]) + bf16_hi(da[kk][1]) * bf16_hi(oa[kk][1]) +
            bf16_lo(da[kk][3]) * bf16_lo(oa[kk][3]) + bf16_hi(da[kk][3]) * bf16_hi(oa[kk][3]);
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);

    float lse0, lse1;
    if (have_lse) {
      // the forward kernel saved the row log-sum-exp (log2 domain): no recompute sweep
      const size_t rowbase = (static_cast<size_t>(b) * H + h) * T;
      lse0 = r0 < T ? lse_out[rowbase + r0] : 0.f;
      lse1 = r1 < T ? lse_out[rowbase + r1] : 0.f;
    } else {
      // sweep 1: row log-sum-exp in the log2 domain
      float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
      for (int kc = 0; kc < Tp; kc += kAbChunk) {
        float s[kAbChunk / 8][4];
        strip_mma<DH>(s, qa, Ks32, kc, g, tg);
        float cm0 = -INFINITY, cm1 = -INFINITY;
  #pragma unroll
        for (int n = 0; n < kAbChunk / 8; ++n) {
          const int key = kc + n * 8 + tg * 2;
          const float k0 = key < T ? 0.f : -INFINITY, k1 = key + 1 < T ? 0.f : -INFINITY;
          s[n][0] = fmaf(s[n][0], scale_log2, k0); s[n][1] = fmaf(s[n][1], scale_log2, k1);
          s[n][2] = fmaf(s[n][2], scale_log2, k0); s[n][3] = fmaf(s[n][3], scale_log2, k1);
          cm0 = fmaxf(cm0, fmaxf(s[n][0], s[n][1]));
          cm1 = fmaxf(cm1, fmaxf(s[n][2], s[n][3]));
        }
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
        const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
        const float ms0 = mn0 == -INFINITY ? 0.f : mn0, ms1 = mn1 == -INFINITY ? 0.f : mn1;
        float ps0 = 0.f, ps1 = 0.f;
  #pragma unroll
        for (int n = 0; n < kAbChunk / 8; ++n) {
          ps0 += ex2f(s[n][0] - ms0) + ex2f(s[n][1] - ms0);
          ps1 += ex2f(s[n][2] - ms1) + ex2f(s[n][3] - ms1);
        }
        l0 = l0 * ex2f(m0 - ms0) + ps0;
        l1 = l1 * ex2f(m1 - ms1) + ps1;
        m0 = mn0; m1 = mn1;
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      lse0 = m0 + log2f(l0);
      lse1 = m1 + log2f(l1);
    }
    if (tg == 0) {
      const size_t rowbase = (static_cast<size_t>(b) * H + h) * T;
      if (r0 < T) { if (!have_lse) lse_out[rowbase + r0] = lse0; dsum_out[rowbase + r0] = d0; }
      if (r1 < T) { if (!have_lse) lse_out[rowbase + r1] = lse1; dsum_out[rowbase + r1] = d1; }
    }

    // sweep 2: dQ = scale * (P o (dP - D)) K
    float dq[DH / 8][4];
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) { dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f; }
    for (int kc = 0; kc < Tp; kc += kAbChunk) {
      float s[kAbChunk / 8][4], dp[kAbChunk / 8][4];
      strip_mma<DH>(s, qa, Ks32, kc, g, tg);
      strip_mma<DH>(dp, da, Vs32, kc, g, tg);
#pragma unroll
      for (int n = 0; n < kAbChunk / 8; ++n) {
        const int key = kc + n * 8 + tg * 2;
        const float k0 = key < T ? 0.f : -INFINITY, k1 = key + 1 < T ? 0.f : -INFINITY;
        const float p0 = ex2f(fmaf(s[n][0], scale_log2, k0) - lse0), p1 = ex2f(fmaf(s[n][1], scale_log2, k1) - lse0);
        const float p2 = ex2f(fmaf(s[n][2], scale_log2, k0) - lse1), p3 = ex2f(fmaf(s[n][3], scale_log2, k1) - lse1);
        if (drop_thr16 != 0) {  // O = (P o M) V  =>  dP = M o (dO V^T); rowsum(dO o O) already includes the mask
          const uint64_t Tpair = static_cast<uint64_t>((T + 1) & ~1);
          const uint64_t rb0 = ((static_cast<uint64_t>(b) * H + h) * T + r0) * Tpair;
          float m0, m1;
          attn_keep2(drop_key, rb0, key, drop_thr16, drop_inv_keep, m0, m1);
          dp[n][0] *= m0; dp[n][1] *= m1;
          attn_keep2(drop_key, rb0 + 8 * Tpair, key, drop_thr16, drop_inv_keep, m0, m1);
          dp[n][2] *= m0; dp[n][3] *= m1;
        }
        s[n][0] = p0 * (dp[n][0] - d0); s[n][1] = p1 * (dp[n][1] - d0);
        s[n][2] = p2 * (dp[n][2] - d1); s[n][3] = p3 * (dp[n][3] - d1);
      }
      strip_acc<DH>(dq, s, Kt32, VS, kc, g, tg);
    }
    bf16* dqb = dqkv + static_cast<size_t>(b) * T * ld + static_cast<size_t>(h) * 3 * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      const int c = n * 8 + tg * 2;
      if (r0 < T) *reinterpret_cast<uint32_t*>(dqb + static_cast<size_t>(r0) * ld + c) = pack_bf16x2(dq[n][0] * scale, dq[n][1] * scale);
      if (r1 < T) *reinterpret_cast<uint32_t*>(dqb + static_cast<size_t>(r1) * ld + c) = pack_bf16x2(dq[n][2] * scale, dq[n][3] * scale);
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(kAbThreads)
attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dO, bf16* __restrict__ dqkv, const float* __restrict__ lse_in,
                    const float* __restrict__ dsum_in, int T, int H, int Tp, float scale, uint32_t drop_thr16, float drop_inv_keep,
                    uint64_t drop_key_val, const uint64_t* __restrict__ drop_key_ptr) {
  const uint64_t drop_key = drop_key_ptr != nullptr ? *drop_key_ptr : drop_key_val;
  constexpr int KS = DH + 8;
  extern __shared__ __align__(16) uint8_t smem_ab[];
  const int VS = Tp + 8;
  bf16* Qs = reinterpret_cast<bf16*>(smem_ab);               // [Tp][KS]
  bf16* Ds = Qs + static_cast<size_t>(Tp) * KS;              // [Tp][KS]  dO
  bf16* Qt = Ds + static_cast<size_t>(Tp) * KS;              // [DH][VS]
  bf16* Dt = Qt + static_cast<size_t>(DH) * VS;              // [DH][VS]
  float* lse = reinterpret_cast<float*>(Dt + static_cast<size_t>(DH) * VS);  // [Tp]
  float* dsum = lse + Tp;                                    // [Tp]
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const size_t ld = static_cast<size_t>(3) * DH * H, ldo = static_cast<size_t>(DH) * H;
  const bf16* qbase = qkv + static_cast<size_t>(b) * T * ld + static_cast<size_t>(h) * 3 * DH;
  const bf16* dobase = dO + static_cast<size_t>(b) * T * ldo + static_cast<size_t>(h) * DH;
  stage_rows<DH>(qbase, ld, T, Tp, Qs, Qt);
  stage_rows<DH>(dobase, ldo, T, Tp, Ds, Dt);
  {
    const size_t rowbase = (static_cast<size_t>(b) * H + h) * T;
    for (int t = threadIdx.x; t < Tp; t += kAbThreads) {
      lse[t] = t < T ? lse_in[rowbase + t] : INFINITY;   // padded queries: P = exp2(-inf) = 0
      dsum[t] = t < T ? dsum_in[rowbase + t] : 0.f;
    }
  }
  __syncthreads();
  const uint32_t* Qs32 = reinterpret_cast<const uint32_t*>(Qs);
  const uint32_t* Ds32 = reinterpret_cast<const uint32_t*>(Ds);
  const uint32_t* Qt32 = reinterpret_cast<const uint32_t*>(Qt);
  const uint32_t* Dt32 = reinterpret_cast<const uint32_t*>(Dt);
  const float scale_log2 = scale * kLog2e;

  for (int k0 = warp * 16; k0 < T; k0 += (kAbThreads / 32) * 16) {
    const int r0 = k0 + g, r1 = k0 + g + 8;
    uint32_t ka[DH / 16][4], va[DH / 16][4];
    load_a_frags<DH>(ka, qbase + DH, ld, r0, r1, T, tg);
    load_a_frags<DH>(va, qbase + 2 * DH, ld, r0, r1, T, tg);
    float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) { dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f; dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f; }
    for (int qc = 0; qc < Tp; qc += kAbChunk) {
      float st[kAbChunk / 8][4], dpt[kAbChunk / 8][4];
      strip_mma<DH>(st, ka, Qs32, qc, g, tg);     // S^T  [16 keys x 64 queries]
      strip_mma<DH>(dpt, va, Ds32, qc, g, tg);    // dP^T = V dO^T
#pragma unroll
      for (int n = 0; n < kAbChunk / 8; ++n) {
        const int q = qc + n * 8 + tg * 2;
        const float2 ls = *reinterpret_cast<const float2*>(lse + q), dd = *reinterpret_cast<const float2*>(dsum + q);
        const float p0 = ex2f(fmaf(st[n][0], scale_log2, -ls.x)), p1 = ex2f(fmaf(st[n][1], scale_log2, -ls.y));
        const float p2 = ex2f(fmaf(st[n][2], scale_log2, -ls.x)), p3 = ex2f(fmaf(st[n][3], scale_log2, -ls.y));
        float m00 = 1.f, m01 = 1.f, m10 = 1.f, m11 = 1.f;  // mask[query q / q+1][key r0 / r1]
        if (drop_thr16 != 0) {
          const uint64_t Tpair = static_cast<uint64_t>((T + 1) & ~1);
          const uint64_t rbq = ((static_cast<uint64_t>(b) * H + h) * T + q) * Tpair;
          m00 = attn_keep(drop_key, rbq, r0, drop_thr16, drop_inv_keep);
          m01 = attn_keep(drop_key, rbq + Tpair, r0, drop_thr16, drop_inv_keep);
          m10 = attn_keep(drop_key, rbq, r1, drop_thr16, drop_inv_keep);
          m11 = attn_keep(drop_key, rbq + Tpair, r1, drop_thr16, drop_inv_keep);
        }
        st[n][0] = p0 * m00; st[n][1] = p1 * m01; st[n][2] = p2 * m10; st[n][3] = p3 * m11;  // (P o M)^T feeds dV
        dpt[n][0] = p0 * (dpt[n][0] * m00 - dd.x); dpt[n][1] = p1 * (dpt[n][1] * m01 - dd.y);
        dpt[n][2] = p2 * (dpt[n][2] * m10 - dd.x); dpt[n][3] = p3 * (dpt[n][3] * m11 - dd.y);
      }
      strip_acc<DH>(dv, st, Dt32, VS, qc, g, tg);   // dV += P^T dO
      strip_acc<DH>(dk, dpt, Qt32, VS, qc, g, tg);  // dK += dS^T Q
    }
    bf16* db = dqkv + static_cast<size_t>(b) * T * ld + static_cast<size_t>(h) * 3 * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      const int c = n * 8 + tg * 2;
      if (r0 < T) {
        *reinterpret_cast<uint32_t*>(db + static_cast<size_t>(r0) * ld + DH + c) = pack_bf16x2(dk[n][0] * scale, dk[n][1] * scale);
        *reinterpret_cast<uint32_t*>(db + static_cast<size_t>(r0) * ld + 2 * DH + c) = pack_bf16x2(dv[n][0], dv[n][1]);
      }
      if (r1 < T) {
        *reinterpret_cast<uint32_t*>(db + static_cast<size_t>(r1) * ld + DH + c) = pack_bf16x2(dk[n][2] * scale, dk[n][3] * scale);
        *reinterpret_cast<uint32_t*>(db + static_cast<size_t>(r1) * ld + 2 * DH + c) = pack_bf16x2(dv[n][2], dv[n][3]);
      }
    }
  }
}

template <int DH>
int launch_dh(const AttnBwdArgs& a, cudaStream_t s) {
  const int Tp = (a.T + kAbChunk - 1) / kAbChunk * kAbChunk;
  const size_t smem1 = (static_cast<size_t>(2) * Tp * (DH + 8) + static_cast<size_t>(DH) * (Tp + 8)) * sizeof(bf16);
  const size_t smem2 = (static_cast<size_t>(2) * Tp * (DH + 8) + static_cast<size_t>(2) * DH * (Tp + 8)) * sizeof(bf16) +
                       static_cast<size_t>(2) * Tp * sizeof(float);
  if (smem2 > 227 * 1024) {
    set_last_error("attention_bwd: sequence too long for the shared-memory resident kernel (T*dh too large)");
    return 2;
  }
  static bool attr_done = false;
  if (!attr_done) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ISHARA_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  const dim3 grid(a.H, a.B);
  attn_bwd_dq_kernel<DH><<<grid, kAbThreads, smem1, s>>>(a.qkv, a.o, a.dO, a.dqkv, a.lse2, a.dsum, a.T, a.H, Tp, a.scale, a.drop_thr16,
                                                           a.drop_inv_keep, a.drop_key, a.drop_key_ptr, a.have_lse);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  attn_bwd_dkv_kernel<DH><<<grid, kAbThreads, smem2, s>>>(a.qkv, a.dO, a.dqkv, a.lse2, a.dsum, a.T, a.H, Tp, a.scale, a.drop_thr16, a.drop_inv_keep,
                                                            a.drop_key, a.drop_key_ptr);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace

int attention_bwd_launch(const AttnBwdArgs& a, cudaStream_t s) {
  if (a.qkv == nullptr || a.o == nullptr || a.dO == nullptr || a.dqkv == nullptr || a.lse2 == nullptr || a.dsum == nullptr) {
    set_last_error("attention_bwd: null argument");
    return 1;
  }
  switch (a.dh) {
    case 16: return launch_dh<16>(a, s);
    case 32: return launch_dh<32>(a, s);
    case 48: return launch_dh<48>(a, s);
    case 64: return launch_dh<64>(a, s);
    default: set_last_error("attention_bwd: head dim must be 16/32/48/64"); return 2;
  }
}

}  // namespace ishara
