// ishara_b200 — fused multi-head self-attention core: softmax(q k^T * scale [+ key mask]) v.
//
// Reference: MultiHeadSelfAttention.call (nb:conv-hybrid-model c5:102-118). The QKV
// projection and the output projection are tcgen05 GEMMs (gemm_tc.cu); this kernel consumes the
// per-head-interleaved qkv tensor [B*T, H*3*dh] (head h owns columns [3*dh*h, 3*dh*(h+1)) = q|k|v,
// exactly the Reshape/Permute/split of c5:104-105) and writes the merged-head output [B*T, H*dh].
// The [H,T,T] score tensor never leaves the SM.
//
// One CTA per (sequence, head): K (row-major) and V (transposed) for the whole sequence are staged
// once in shared memory; 8 warps each take 16 query rows per pass and run a flash-style online
// softmax over 64-key chunks with bf16 tensor-core MMAs (fp32 accumulate) and exp2 in fp32.
// Round-1 note: the two small-K contractions here (K=dh=32 and K=T) use warp-level mma.sync; the
// tcgen05 port of this kernel is listed in DESIGN.md §"next".
#include <cstdlib>

#include "dropout_hash.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace ishara {
namespace {

constexpr int kAttnThreads = 256;
constexpr int kKeyChunk = 64;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DH>
__global__ void __launch_bounds__(kAttnThreads)
attn_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, const uint8_t* __restrict__ key_mask, int T, int H,
            int Tp, float scale_log2, uint32_t drop_thr16, float drop_inv_keep, uint64_t drop_key_val,
            const uint64_t* __restrict__ drop_key_ptr, float* __restrict__ lse_out) {
  const uint64_t drop_key = drop_key_ptr != nullptr ? *drop_key_ptr : drop_key_val;
  constexpr int KS = DH + 8;  // K row stride (elements): conflict-free fragment reads
  extern __shared__ __align__(16) uint8_t smem_at[];
  bf16* Ks = reinterpret_cast<bf16*>(smem_at);                 // [Tp][KS]
  bf16* Vt = Ks + static_cast<size_t>(Tp) * KS;                // [DH][Tp + 8]
  float* mb = reinterpret_cast<float*>(Vt + static_cast<size_t>(DH) * (Tp + 8));  // [Tp] additive key bias (log2 domain)
  const int VS = Tp + 8;

  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tg = lane & 3;
  const int ld = 3 * DH * H;  // qkv row pitch (elements)
  const bf16* base = qkv + static_cast<size_t>(b) * T * ld + h * 3 * DH;

  // ---- stage K (row-major) and V^T; zero the padded keys ----
  for (int t = tid; t < Tp; t += kAttnThreads) {
    uint4 kv[DH / 8], vv[DH / 8];
    if (t < T) {
      const uint4* kp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(t) * ld + DH);
      const uint4* vp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(t) * ld + 2 * DH);
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) { kv[i] = __ldg(kp + i); vv[i] = __ldg(vp + i); }
    } else {
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) { kv[i] = make_uint4(0, 0, 0, 0); vv[i] = make_uint4(0, 0, 0, 0); }
    }
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) *reinterpret_cast<uint4*>(Ks + static_cast<size_t>(t) * KS + 8 * i) = kv[i];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      const uint32_t wv[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        Vt[static_cast<size_t>(8 * i + 2 * e) * VS + t] = __ushort_as_bfloat16(static_cast<unsigned short>(wv[e] & 0xFFFFu));
        Vt[static_cast<size_t>(8 * i + 2 * e + 1) * VS + t] = __ushort_as_bfloat16(static_cast<unsigned short>(wv[e] >> 16));
      }
    }
    float bias = 0.f;
    if (t >= T) bias = -INFINITY;
    else if (key_mask != nullptr && key_mask[static_cast<size_t>(b) * T + t] == 0)
      bias = -1.0e9f * 1.4426950408889634f;  // Keras Softmax(mask): logits += (1-mask)*-1e9
    mb[t] = bias;
  }
  __syncthreads();

  const uint32_t* Ks32 = reinterpret_cast<const uint32_t*>(Ks);
  const uint32_t* Vt32 = reinterpret_cast<const uint32_t*>(Vt);

  for (int q0 = warp * 16; q0 < T; q0 += (kAttnThreads / 32) * 16) {
    const int r0 = q0 + g, r1 = q0 + g + 8;
    // Q fragments (A operand), straight from global
    uint32_t qa[DH / 16][4];
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      const int c = kk * 16 + tg * 2;
      qa[kk][0] = r0 < T ? __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(r0) * ld + c)) : 0u;
      qa[kk][1] = r1 < T ? __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(r1) * ld + c)) : 0u;
      qa[kk][2] = r0 < T ? __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(r0) * ld + c + 8)) : 0u;
      qa[kk][3] = r1 < T ? __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(r1) * ld + c + 8)) : 0u;
    }
    float o[DH / 8][4];
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kc = 0; kc < Tp; kc += kKeyChunk) {
      float s[kKeyChunk / 8][4];
#pragma unroll
      for (int n = 0; n < kKeyChunk / 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
        for (int n = 0; n < kKeyChunk / 8; ++n) {
          const int key = kc + n * 8 + g;
          const uint32_t b0 = Ks32[(key * KS + kk * 16 + tg * 2) >> 1];
          const uint32_t b1 = Ks32[(key * KS + kk * 16 + 8 + tg * 2) >> 1];
          mma_bf16_16816(s[n], qa[kk], b0, b1);
        }
      }
      // scale to the log2 domain, add key bias, chunk row max
      float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
      for (int n = 0; n < kKeyChunk / 8; ++n) {
        const float2 kb = *reinterpret_cast<const float2*>(mb + kc + n * 8 + tg * 2);
        s[n][0] = fmaf(s[n][0], scale_log2, kb.x);
        s[n][1] = fmaf(s[n][1], scale_log2, kb.y);
        s[n][2] = fmaf(s[n][2], scale_log2, kb.x);
        s[n][3] = fmaf(s[n][3], scale_log2, kb.y);
        cm0 = fmaxf(cm0, fmaxf(s[n][0], s[n][1]));
        cm1 = fmaxf(cm1, fmaxf(s[n][2], s[n][3]));
      }
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
      const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
      // a fully padded chunk keeps mn = -inf only if everything so far was padded; guard the NaN
      const float ms0 = mn0 == -INFINITY ? 0.f : mn0, ms1 = mn1 == -INFINITY ? 0.f : mn1;
      const float a0 = ex2(m0 - ms0), a1 = ex2(m1 - ms1);
      m0 = mn0; m1 = mn1;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int n = 0; n < kKeyChunk / 8; ++n) {
        s[n][0] = ex2(s[n][0] - ms0); s[n][1] = ex2(s[n][1] - ms0);
        s[n][2] = ex2(s[n][2] - ms1); s[n][3] = ex2(s[n][3] - ms1);
        ps0 += s[n][0] + s[n][1];
        ps1 += s[n][2] + s[n][3];
      }
      l0 = l0 * a0 + ps0;
      l1 = l1 * a1 + ps1;
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) { o[n][0] *= a0; o[n][1] *= a0; o[n][2] *= a1; o[n][3] *= a1; }
      if (drop_thr16 != 0) {
        // training: dropout on the normalised probabilities = mask the numerators that feed P V, keep the row sums
        const uint64_t Tpair = static_cast<uint64_t>((T + 1) & ~1);
        const uint64_t rb0 = ((static_cast<uint64_t>(b) * H + h) * T + r0) * Tpair, rb1 = rb0 + 8 * Tpair;
#pragma unroll
        for (int n = 0; n < kKeyChunk / 8; ++n) {
          const int j = kc + n * 8 + tg * 2;
          float k0, k1;
          attn_keep2(drop_key, rb0, j, drop_thr16, drop_inv_keep, k0, k1);
          s[n][0] *= k0; s[n][1] *= k1;
          attn_keep2(drop_key, rb1, j, drop_thr16, drop_inv_keep, k0, k1);
          s[n][2] *= k0; s[n][3] *= k1;
        }
      }
      // O += P V : P (C-fragment layout) re-packed as the A operand of the next MMA
#pragma unroll
      for (int kk = 0; kk < kKeyChunk / 16; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int n = 0; n < DH / 8; ++n) {
          const int d = n * 8 + g;
          const int key = kc + kk * 16 + tg * 2;
          const uint32_t b0 = Vt32[(d * VS + key) >> 1];
          const uint32_t b1 = Vt32[(d * VS + key + 8) >> 1];
          mma_bf16_16816(o[n], pa, b0, b1);
        }
      }
    }
    // row sums across the quad, normalise, store
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    if (lse_out != nullptr && tg == 0) {  // training: row log-sum-exp (log2 domain) for the backward pass
      const size_t rowbase = (static_cast<size_t>(b) * H + h) * T;
      if (r0 < T) lse_out[rowbase + r0] = m0 + log2f(l0);
      if (r1 < T) lse_out[rowbase + r1] = m1 + log2f(l1);
    }
    bf16* ob = out + static_cast<size_t>(b) * T * (DH * H) + h * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      const int c = n * 8 + tg * 2;
      if (r0 < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0) * (DH * H) + c) = pack_bf16x2(o[n][0] * i0, o[n][1] * i0);
      if (r1 < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r1) * (DH * H) + c) = pack_bf16x2(o[n][2] * i1, o[n][3] * i1);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Transformer-XL relative-position variant (vendored Squeezeformer, squeezeformer/attention.py:25-110;
// SURVEY.md §8a P1): score[i,j] = ((q_i + u)·k_j + (q_i + v)·p[j - i + T - 1]) * scale, where p = pos_proj(pos_emb)
// has 2T-1 rows and row r encodes relative position T-1-r. The reference materialises (q+v)·p^T as [T, 2T-1] and
// applies `_relative_shift` (:102-110); here the shift is index arithmetic: for a 16-query block and a 64-key chunk only
// 80 consecutive rows of p can be hit, so one extra 16x80 MMA strip is computed, parked in a per-warp skew buffer and
// read back along the diagonal. Nothing of size T x 2T ever exists.
// ------------------------------------------------------------------------------------------------
constexpr int kPosLead = 16;   // zero rows in front of p (query rows past T index before row 0)
constexpr int kSkewCols = 88;  // 80 used + padding against bank conflicts

template <int DH>
__global__ void __launch_bounds__(kAttnThreads)
attn_relpos_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ pos, const float* __restrict__ u_bias,
                   const float* __restrict__ v_bias, bf16* __restrict__ out, const uint8_t* __restrict__ key_mask, int T,
                   int H, int Tp, int Rp, float scale_log2) {
  constexpr int KS = DH + 8;
  extern __shared__ __align__(16) uint8_t smem_at[];
  bf16* Ks = reinterpret_cast<bf16*>(smem_at);                                     // [Tp][KS]
  bf16* Vt = Ks + static_cast<size_t>(Tp) * KS;                                    // [DH][Tp + 8]
  bf16* Ps = Vt + static_cast<size_t>(DH) * (Tp + 8);                              // [Rp][KS], row r of p at r + kPosLead
  float* mb = reinterpret_cast<float*>(Ps + static_cast<size_t>(Rp) * KS);         // [Tp]
  float* skew_all = mb + Tp;                                                       // [warps][16][kSkewCols]
  const int VS = Tp + 8;

  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tg = lane & 3;
  const int ld = 3 * DH * H;
  const bf16* base = qkv + static_cast<size_t>(b) * T * ld + h * 3 * DH;
  float* skew = skew_all + warp * 16 * kSkewCols;

  for (int t = tid; t < Tp; t += kAttnThreads) {
    uint4 kv[DH / 8], vv[DH / 8];
    if (t < T) {
      const uint4* kp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(t) * ld + DH);
      const uint4* vp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(t) * ld + 2 * DH);
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) { kv[i] = __ldg(kp + i); vv[i] = __ldg(vp + i); }
    } else {
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) { kv[i] = make_uint4(0, 0, 0, 0); vv[i] = make_uint4(0, 0, 0, 0); }
    }
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) *reinterpret_cast<uint4*>(Ks + static_cast<size_t>(t) * KS + 8 * i) = kv[i];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      const uint32_t wv[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        Vt[static_cast<size_t>(8 * i + 2 * e) * VS + t] = __ushort_as_bfloat16(static_cast<unsigned short>(wv[e] & 0xFFFFu));
        Vt[static_cast<size_t>(8 * i + 2 * e + 1) * VS + t] = __ushort_as_bfloat16(static_cast<unsigned short>(wv[e] >> 16));
      }
    }
    float bias = 0.f;
    if (t >= T) bias = -INFINITY;
    else if (key_mask != nullptr && key_mask[static_cast<size_t>(b) * T + t] == 0) bias = -1.0e9f * 1.4426950408889634f;
    mb[t] = bias;
  }
  for (int r = tid; r < Rp; r += kAttnThreads) {
    const int src = r - kPosLead;
    uint4 pv[DH / 8];
    if (src >= 0 && src < 2 * T - 1) {
      const uint4* pp = reinterpret_cast<const uint4*>(pos + static_cast<size_t>(src) * (DH * H) + h * DH);
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) pv[i] = __ldg(pp + i);
    } else {
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) pv[i] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) *reinterpret_cast<uint4*>(Ps + static_cast<size_t>(r) * KS + 8 * i) = pv[i];
  }
  __syncthreads();

  const uint32_t* Ks32 = reinterpret_cast<const uint32_t*>(Ks);
  const uint32_t* Vt32 = reinterpret_cast<const uint32_t*>(Vt);
  const uint32_t* Ps32 = reinterpret_cast<const uint32_t*>(Ps);

  for (int q0 = warp * 16; q0 < T; q0 += (kAttnThreads / 32) * 16) {
    const int r0 = q0 + g, r1 = q0 + g + 8;
    uint32_t qu[DH / 16][4], qv[DH / 16][4];
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = (e & 1) ? r1 : r0;
        const int c = kk * 16 + tg * 2 + ((e & 2) ? 8 : 0);
        float q0f = 0.f, q1f = 0.f;
        if (row < T) {
          const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(row) * ld + c));
          q0f = bf16_lo(w); q1f = bf16_hi(w);
        }
        const float2 uu = __ldg(reinterpret_cast<const float2*>(u_bias + h * DH + c));
        const float2 vv2 = __ldg(reinterpret_cast<const float2*>(v_bias + h * DH + c));
        qu[kk][e] = pack_bf16x2(q0f + uu.x, q1f + uu.y);
        qv[kk][e] = pack_bf16x2(q0f + vv2.x, q1f + vv2.y);
      }
    }
    float o[DH / 8][4];
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kc = 0; kc < Tp; kc += kKeyChunk) {
      // positional strip: p rows [rbase, rbase + 80), smem row = r + kPosLead
      const int rbase = kc - (q0 + 15) + T - 1 + kPosLead;  // >= 1 because q0 + 15 <= T - 1 + 15
      __syncwarp();
#pragma unroll
      for (int n = 0; n < 10; ++n) {
        float c4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          const int prow = rbase + n * 8 + g;
          const uint32_t b0 = Ps32[(prow * KS + kk * 16 + tg * 2) >> 1];
          const uint32_t b1 = Ps32[(prow * KS + kk * 16 + 8 + tg * 2) >> 1];
          mma_bf16_16816(c4, qv[kk], b0, b1);
        }
        *reinterpret_cast<float2*>(skew + g * kSkewCols + n * 8 + tg * 2) = make_float2(c4[0], c4[1]);
        *reinterpret_cast<float2*>(skew + (g + 8) * kSkewCols + n * 8 + tg * 2) = make_float2(c4[2], c4[3]);
      }
      __syncwarp();
      float s[kKeyChunk / 8][4];
#pragma unroll
      for (int n = 0; n < kKeyChunk / 8; ++n) {
        // key jj = n*8 + tg*2 + e of this chunk, query row ri: strip column jj + 15 - ri
        const int jj = n * 8 + tg * 2;
        s[n][0] = skew[g * kSkewCols + jj + 15 - g];
        s[n][1] = skew[g * kSkewCols + jj + 16 - g];
        s[n][2] = skew[(g + 8) * kSkewCols + jj + 7 - g];
        s[n][3] = skew[(g + 8) * kSkewCols + jj + 8 - g];
      }
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
        for (int n = 0; n < kKeyChunk / 8; ++n) {
          const int key = kc + n * 8 + g;
          const uint32_t b0 = Ks32[(key * KS + kk * 16 + tg * 2) >> 1];
          const uint32_t b1 = Ks32[(key * KS + kk * 16 + 8 + tg * 2) >> 1];
          mma_bf16_16816(s[n], qu[kk], b0, b1);
        }
      }
      float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
      for (int n = 0; n < kKeyChunk / 8; ++n) {
        const float2 kb = *reinterpret_cast<const float2*>(mb + kc + n * 8 + tg * 2);
        s[n][0] = fmaf(s[n][0], scale_log2, kb.x);
        s[n][1] = fmaf(s[n][1], scale_log2, kb.y);
        s[n][2] = fmaf(s[n][2], scale_log2, kb.x);
        s[n][3] = fmaf(s[n][3], scale_log2, kb.y);
        cm0 = fmaxf(cm0, fmaxf(s[n][0], s[n][1]));
        cm1 = fmaxf(cm1, fmaxf(s[n][2], s[n][3]));
      }
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
      const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
      const float ms0 = mn0 == -INFINITY ? 0.f : mn0, ms1 = mn1 == -INFINITY ? 0.f : mn1;
      const float a0 = ex2(m0 - ms0), a1 = ex2(m1 - ms1);
      m0 = mn0; m1 = mn1;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int n = 0; n < kKeyChunk / 8; ++n) {
        s[n][0] = ex2(s[n][0] - ms0); s[n][1] = ex2(s[n][1] - ms0);
        s[n][2] = ex2(s[n][2] - ms1); s[n][3] = ex2(s[n][3] - ms1);
        ps0 += s[n][0] + s[n][1];
        ps1 += s[n][2] + s[n][3];
      }
      l0 = l0 * a0 + ps0;
      l1 = l1 * a1 + ps1;
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) { o[n][0] *= a0; o[n][1] *= a0; o[n][2] *= a1; o[n][3] *= a1; }
#pragma unroll
      for (int kk = 0; kk < kKeyChunk / 16; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int n = 0; n < DH / 8; ++n) {
          const int d = n * 8 + g;
          const int key = kc + kk * 16 + tg * 2;
          const uint32_t b0 = Vt32[(d * VS + key) >> 1];
          const uint32_t b1 = Vt32[(d * VS + key + 8) >> 1];
          mma_bf16_16816(o[n], pa, b0, b1);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    bf16* ob = out + static_cast<size_t>(b) * T * (DH * H) + h * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      const int c = n * 8 + tg * 2;
      if (r0 < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0) * (DH * H) + c) = pack_bf16x2(o[n][0] * i0, o[n][1] * i0);
      if (r1 < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r1) * (DH * H) + c) = pack_bf16x2(o[n][2] * i1, o[n][3] * i1);
    }
  }
}

template <int DH>
int launch_relpos(const AttnArgs& a, cudaStream_t stream) {
  const int Tp = (a.T + kKeyChunk - 1) / kKeyChunk * kKeyChunk;
  const int Rp = Tp + a.T + kPosLead + 16;  // last strip row: (Tp - 64) - 0 - 15 + T - 1 + kPosLead + 79 < Rp
  const size_t smem = static_cast<size_t>(Tp) * (DH + 8) * 2 + static_cast<size_t>(DH) * (Tp + 8) * 2 +
                      static_cast<size_t>(Rp) * (DH + 8) * 2 + Tp * sizeof(float) +
                      static_cast<size_t>(kAttnThreads / 32) * 16 * kSkewCols * sizeof(float);
  if (smem > 227 * 1024) {
    set_last_error("relpos attention: sequence too long for the single-pass K/V/P staging");
    return 2;
  }
  auto kern = attn_relpos_kernel<DH>;
  static size_t smem_attr = 0;
  if (smem > smem_attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    smem_attr = smem;
  }
  kern<<<dim3(a.H, a.B), kAttnThreads, smem, stream>>>(a.qkv, a.pos, a.u_bias, a.v_bias, a.out, a.key_mask, a.T, a.H, Tp,
                                                      Rp, a.scale * 1.4426950408889634f);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

template <int DH>
int launch_inst(const AttnArgs& a, cudaStream_t stream) {
  const int Tp = (a.T + kKeyChunk - 1) / kKeyChunk * kKeyChunk;
  const size_t smem = static_cast<size_t>(Tp) * (DH + 8) * 2 + static_cast<size_t>(DH) * (Tp + 8) * 2 + Tp * sizeof(float);
  if (smem > 227 * 1024) {
    set_last_error("attention: sequence too long for the single-pass K/V staging");
    return 2;
  }
  auto kern = attn_kernel<DH>;
  static size_t smem_attr = 0;
  if (smem > smem_attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    smem_attr = smem;
  }
  kern<<<dim3(a.H, a.B), kAttnThreads, smem, stream>>>(a.qkv, a.out, a.key_mask, a.T, a.H, Tp,
                                                      a.scale * 1.4426950408889634f, a.drop_thr16, a.drop_inv_keep, a.drop_key, a.drop_key_ptr, a.lse_out);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace

bool attention_tc_applicable(const AttnArgs& a);
int attention_tc_launch(const AttnArgs& a, cudaStream_t stream);

int attention_launch(const AttnArgs& a, cudaStream_t stream) {
  static const int use_tc = getenv("ISHARA_ATTN_TC") ? atoi(getenv("ISHARA_ATTN_TC")) : 1;
  if (use_tc && a.drop_thr16 == 0 && attention_tc_applicable(a)) return attention_tc_launch(a, stream);
  if (a.drop_thr16 != 0 && a.pos != nullptr) {
    set_last_error("attention: dropout is not supported by the relative-position kernel");
    return 2;
  }
  if (a.pos != nullptr) {
    if (a.u_bias == nullptr || a.v_bias == nullptr) {
      set_last_error("relpos attention needs u_bias and v_bias");
      return 2;
    }
    switch (a.dh) {
      case 16: return launch_relpos<16>(a, stream);
      case 32: return launch_relpos<32>(a, stream);
      case 64: return launch_relpos<64>(a, stream);
    }
    set_last_error("relpos attention: head dim must be 16, 32 or 64");
    return 2;
  }
  switch (a.dh) {
    case 16: return launch_inst<16>(a, stream);
    case 32: return launch_inst<32>(a, stream);
    case 48: return launch_inst<48>(a, stream);
    case 64: return launch_inst<64>(a, stream);
  }
  set_last_error("attention: head dim must be 16, 32, 48 or 64");
  return 2;
}

}  // namespace ishara
