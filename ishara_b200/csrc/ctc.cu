// ishara_b200 — warp-level CTC loss (forward + gradient) and greedy CTC decode (sm_100a).
//
// CTC: CTCLoss (nb:conv-hybrid-model c6:1-13) = tf.nn.ctc_loss(labels, logits, label_length =
// #(labels != blank), logit_length = T, blank_index = 59, logits_time_major=False); per-sequence
// negative log-likelihood, log-softmax inside, no zero_infinity (infeasible => +inf). SURVEY.md §8a T13.
// One warp per sequence; the S = 2L+1 extended states are blocked SPL-per-lane in registers, the
// alpha/beta recursions run in fp32 log space, neighbours are exchanged with two shuffles per step.
// The gradient (d nll / d logits = softmax - state occupancy) is produced in the beta sweep from
// alphas parked in a global workspace.
//
// Decode: decode_phrase (c8:4-12): argmax over classes (first index on ties) -> keep position t only
// if t < T-1 and ids[t] != ids[t+1] -> drop blanks. NOTE the reference quirk, reproduced on purpose:
// the last run is never emitted because index T-1 is never selected.
#include <algorithm>

#include "kernels.h"
#include "ptx.cuh"

namespace ishara {
namespace {

// labels outside [0, V) would index the staged logits row and the occupancy table out of bounds: clamp for memory safety
// (host entry points that see the labels reject them with ISHARA_ERR_INVALID before any launch)
__device__ __forceinline__ int ctc_label(const int32_t* lab, int i, int V) { return min(max(lab[i], 0), V - 1); }

constexpr float kLogZero = -1.0e30f;
constexpr unsigned kCtcRedoMark = 0x7fc0deadu;  // NaN payload written to nll[b]: "redo this sequence in the log domain"
constexpr float kCtcLinearMaxRange = 100.f;     // largest per-frame (max - min logit) * log2 e the linear sweep accepts
constexpr int kCtcBlk = 32;  // frames staged per shared-memory block

// All recursions run in the log2 domain with raw MUFU ex2/lg2 (no range fix-ups, no branches): the sweep is a serial
// chain of ~T dependent steps executed by ONE warp, so its speed is instructions-per-step times the dependent-issue
// latency (ncu: 390 instructions and ~2.3k cycles per step with __expf/__logf and early-out branches). kLogZero is a
// FINITE sentinel: (-1e30) - (-1e30) = 0 and -1e30 + x = -1e30 in fp32, so "log 0" states stay at the sentinel without
// NaNs or special cases.
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2a(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  return m + lg2a(ex2a(a - m) + ex2a(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  return m + lg2a(ex2a(a - m) + ex2a(b - m) + ex2a(c - m));
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One warp (= one CTA) per sequence. The recursion over T is strictly sequential, so what matters is the latency of one
// step: the logits are staged in shared memory in blocks of 32 frames with cp.async, one block ahead of the sweep
// (the first version read each frame's emissions straight from global one step ahead and spent a DRAM round trip,
// ~2.5k cycles, on every one of the 384 steps). dynamic smem: lse[T] + occ[Vpad] + 2 x [32 x V] logits blocks.
template <int SPL>
__global__ void __launch_bounds__(32)
ctc_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, int B, int T, int V, int L, int blank,
           float* __restrict__ nll, float* __restrict__ grad, float* __restrict__ alpha_ws, float* __restrict__ beta_ws,
           int only_flagged) {
  extern __shared__ __align__(16) float smem_ctc[];
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  // second pass behind ctc_alpha_linear_kernel: only the sequences it marked as outside its dynamic range
  if (only_flagged && __float_as_uint(nll[b]) != kCtcRedoMark) return;
  const int Vpad = (V + 31) & ~31;
  float* lse = smem_ctc;
  float* occ = lse + ((T + 3) & ~3);
  float* blk = occ + Vpad;  // [2][kCtcBlk * V]
  const int blk_floats = kCtcBlk * V;
  const float* lg = logits + static_cast<size_t>(b) * T * V;
  const int32_t* lab = labels + static_cast<size_t>(b) * L;
  const int nblk = (T + kCtcBlk - 1) / kCtcBlk;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(lg) & 15) == 0) && (V % 4 == 0);

  auto stage = [&](int tb) {
    float* dst = blk + (tb & 1) * blk_floats;
    const float* src = lg + static_cast<size_t>(tb) * kCtcBlk * V;
    const int n = min(kCtcBlk, T - tb * kCtcBlk) * V;
    if (vec_ok) {
      for (int i = lane * 4; i < n; i += 128) cp_async_16(dst + i, src + i);
    } else {
      for (int i = lane; i < n; i += 32) cp_async_4(dst + i, src + i);
    }
    cp_async_commit();
  };
  stage(0);

  // label length = number of non-blank entries (reference: reduce_sum(labels != pad))
  int cnt = 0;
  for (int i = lane; i < L; i += 32) cnt += (ctc_label(lab, i, V) != blank) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  const int Lb = cnt;
  const int S = 2 * Lb + 1;

  // per-lane extended states s = lane*SPL + i
  int ext[SPL];
  bool skip_fwd[SPL];  // alpha: transition from s-2 allowed
  bool skip_bwd[SPL];  // beta: transition to s+2 allowed
#pragma unroll
  for (int i = 0; i < SPL; ++i) {
    const int s = lane * SPL + i;
    int e = blank;
    if (s < S && (s & 1)) e = ctc_label(lab, s >> 1, V);
    ext[i] = e;
    skip_fwd[i] = (s < S) && (s & 1) && (s >= 3) && (ctc_label(lab, s >> 1, V) != ctc_label(lab, (s >> 1) - 1, V));
    skip_bwd[i] = (s + 2 < S) && (s & 1) && (ctc_label(lab, (s >> 1) + 1, V) != ctc_label(lab, s >> 1, V));
  }

  // log-sum-exp of frame (tb*32 + lane) from the staged block; lanes walk the classes in rotated order (bank spread)
  auto frame_lse = [&](const float* cur, int nf) -> float {
    float l = 0.f;
    if (lane < nf) {
      const float* row = cur + lane * V;
      float m = -INFINITY;
      int v = lane % V;
      for (int k = 0; k < V; ++k) { m = fmaxf(m, row[v]); v = (v + 1 == V) ? 0 : v + 1; }
      float z = 0.f;
      for (int k = 0; k < V; ++k) { z += __expf(row[v] - m); v = (v + 1 == V) ? 0 : v + 1; }
      l = m + __logf(z);
    }
    return l;
  };

  float* aw = alpha_ws != nullptr ? alpha_ws + static_cast<size_t>(b) * T * (32 * SPL) : nullptr;

  // ---------------- alpha sweep ----------------
  float a[SPL];
  for (int tb = 0; tb < nblk; ++tb) {
    if (tb + 1 < nblk) { stage(tb + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncwarp();
    const float* cur = blk + (tb & 1) * blk_floats;
    const int nf = min(kCtcBlk, T - tb * kCtcBlk);
    const float lse_l = frame_lse(cur, nf);
    if (lane < nf) lse[tb * kCtcBlk + lane] = lse_l;
    for (int tt = 0; tt < nf; ++tt) {
      const int t = tb * kCtcBlk + tt;
      const float lt = __shfl_sync(0xffffffffu, lse_l, tt);
      const float* row = cur + tt * V;
      float em[SPL];
#pragma unroll
      for (int i = 0; i < SPL; ++i) em[i] = (row[ext[i]] - lt) * kLog2e;
      if (t == 0) {
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          const int s = lane * SPL + i;
          a[i] = (s < 2 && s < S) ? em[i] : kLogZero;
        }
      } else {
        float up1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1);
        float up2 = __shfl_up_sync(0xffffffffu, a[SPL - 2], 1);
        if (lane == 0) { up1 = kLogZero; up2 = kLogZero; }
        float na[SPL];
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          const float q1 = (i == 0) ? up1 : a[i - 1];
          const float q2 = (i == 0) ? up2 : (i == 1 ? up1 : a[i - 2]);
          const float acc = lse3(a[i], q1, skip_fwd[i] ? q2 : kLogZero);
          const int s = lane * SPL + i;
          na[i] = (s < S) ? acc + em[i] : kLogZero;
        }
#pragma unroll
        for (int i = 0; i < SPL; ++i) a[i] = na[i];
      }
      if (aw != nullptr) {
        float* dst = aw + static_cast<size_t>(t) * (32 * SPL) + lane * SPL;
#pragma unroll
        for (int i = 0; i < SPL; ++i) dst[i] = a[i];
      }
    }
    __syncwarp();  // every lane is done with `cur` before the block after next is staged over it
  }
  // log p(l|x) = lse(alpha_T-1(S-1), alpha_T-1(S-2))
  float fin = kLogZero;
#pragma unroll
  for (int i = 0; i < SPL; ++i) {
    const int s = lane * SPL + i;
    if (s == S - 1 || (s == S - 2 && S >= 2)) fin = lse2(fin, a[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fin = lse2(fin, __shfl_xor_sync(0xffffffffu, fin, o));
  const float logp = fin;  // log2 p(l|x)
  const bool feasible = logp > 0.5f * kLogZero;
  if (lane == 0) nll[b] = feasible ? -logp * kLn2 : INFINITY;
  if (grad == nullptr) return;

  // ---------------- beta sweep + gradient (blocks in reverse) ----------------
  float* bw = beta_ws + static_cast<size_t>(b) * T * (32 * SPL);
  float bt[SPL];
  stage(nblk - 1);
  for (int tb = nblk - 1; tb >= 0; --tb) {
    // note: stage() selects the buffer by block parity, so the block staged next (tb-1) never aliases `cur`
    if (tb > 0) { stage(tb - 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncwarp();
    const float* cur = blk + (tb & 1) * blk_floats;
    const int nf = min(kCtcBlk, T - tb * kCtcBlk);
    for (int tt = nf - 1; tt >= 0; --tt) {
      const int t = tb * kCtcBlk + tt;
      const float lt = lse[t];
      const float* row = cur + tt * V;
      float em[SPL];
#pragma unroll
      for (int i = 0; i < SPL; ++i) em[i] = (row[ext[i]] - lt) * kLog2e;
      if (t == T - 1) {
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          const int s = lane * SPL + i;
          bt[i] = (s == S - 1 || (s == S - 2 && S >= 2)) ? em[i] : kLogZero;
        }
      } else {
        float dn1 = __shfl_down_sync(0xffffffffu, bt[0], 1);
        float dn2 = __shfl_down_sync(0xffffffffu, bt[SPL > 1 ? 1 : 0], 1);
        if (lane == 31) { dn1 = kLogZero; dn2 = kLogZero; }
        float nb[SPL];
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          const float q1 = (i + 1 < SPL) ? bt[i + 1] : dn1;
          const float q2 = (i + 2 < SPL) ? bt[i + 2] : (i + 1 < SPL ? dn1 : dn2);
          const float acc = lse3(bt[i], q1, skip_bwd[i] ? q2 : kLogZero);
          const int s = lane * SPL + i;
          nb[i] = (s < S) ? acc + em[i] : kLogZero;
        }
#pragma unroll
        for (int i = 0; i < SPL; ++i) bt[i] = nb[i];
      }
      // park beta_t next to alpha_t: the gradient of every frame is formed by ctc_grad_kernel, one warp per frame,
      // instead of ~2k dependent cycles per frame on this single warp
      {
        float* bdst = bw + static_cast<size_t>(t) * (32 * SPL) + lane * SPL;
#pragma unroll
        for (int i = 0; i < SPL; ++i) bdst[i] = bt[i];
      }
    }
    __syncwarp();
  }
}

// Loss-only forward sweep in the LINEAR domain (no gradient requested: inference step, validation). The log-domain
// recursion above pays 3 x ex2 + lg2 per state per step ON the serial chain (~720 cycles per frame, 144 us for T = 384);
// here a step is alpha_t(s) = (alpha_{t-1}(s) + alpha_{t-1}(s-1) + [skip] alpha_{t-1}(s-2)) * y_t(ext(s)) - two adds and a
// multiply behind the neighbour shuffle - and the emission probabilities y_t = 2^((logit - lse_t) log2 e) are off the chain.
// Range: block floating point PER LANE. Lane l keeps its SPL states as a_l * 2^K_l; every step it folds the exponent of its
// own largest value of the previous step into y (one step late, so off the chain) and adds it to K_l; values arriving
// from lane l-1 are converted with 2^(K_{l-1} - K_l), and a lane that holds no mass yet adopts its neighbour's exponent so
// that the first mass to arrive is representable. (One exponent for the whole warp is NOT enough: the states a likely
// path passes through can sit more than 2^-126 below the currently largest state - measured: a 33-label sequence lost
// 18 nats that way.) log2 p(l|x) = lse2 over the last two states of log2(a) + K; no mass there <=> infeasible labelling.
// GRAD: the training variant. The alpha sweep also parks a_t and K_t of every frame, then a mirrored beta sweep
// (beta_t(s) = y_t(s) (beta_{t+1}(s) + beta_{t+1}(s+1) + [skip] beta_{t+1}(s+2)), started from a virtual beta_T(S-1) = 1) parks
// b_t and its exponents; ctc_grad_kernel turns them into log2 alpha + log2 beta per state (mode[b] = 0). Sequences outside
// the linear range get mode[b] = 1 and are redone by the log-domain kernel into the same workspace.
template <int SPL, bool GRAD>
__global__ void __launch_bounds__(32)
ctc_alpha_linear_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, int B, int T, int V, int L, int blank,
                        float* __restrict__ nll, float* __restrict__ alpha_ws, float* __restrict__ beta_ws, int* __restrict__ ka_ws,
                        int* __restrict__ kb_ws, int* __restrict__ mode) {
  extern __shared__ __align__(16) float smem_ctc[];
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  float* lse = smem_ctc;                                   // [T] log2-domain frame log-sum-exp (GRAD only)
  float* blk = smem_ctc + (GRAD ? ((T + 3) & ~3) : 0);     // [2][kCtcBlk * V]
  const int blk_floats = kCtcBlk * V;
  const float* lg = logits + static_cast<size_t>(b) * T * V;
  const int32_t* lab = labels + static_cast<size_t>(b) * L;
  const int nblk = (T + kCtcBlk - 1) / kCtcBlk;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(lg) & 15) == 0) && (V % 4 == 0);
  auto stage = [&](int tb) {
    float* dst = blk + (tb & 1) * blk_floats;
    const float* src = lg + static_cast<size_t>(tb) * kCtcBlk * V;
    const int n = min(kCtcBlk, T - tb * kCtcBlk) * V;
    if (vec_ok) {
      for (int i = lane * 4; i < n; i += 128) cp_async_16(dst + i, src + i);
    } else {
      for (int i = lane; i < n; i += 32) cp_async_4(dst + i, src + i);
    }
    cp_async_commit();
  };
  stage(0);
  int cnt = 0;
  for (int i = lane; i < L; i += 32) cnt += (ctc_label(lab, i, V) != blank) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  const int S = 2 * cnt + 1;
  // everything the serial loop touches is branch-free arithmetic: ONE warp runs it alone on its scheduler, so a step costs
  // (instructions per step) x (dependent-issue latency) - the first version of this loop compiled to 270 instructions
  // with five divergent regions per step and was slower (197 us) than the log-domain kernel
  int extoff[SPL];            // byte offset of the state's class inside a staged logits row
  float skipf[SPL], livef[SPL];
  float skipb[SPL];           // beta: transition to s + 2 allowed (GRAD only)
#pragma unroll
  for (int i = 0; i < SPL; ++i) {
    const int s = lane * SPL + i;
    int e = blank;
    if (s < S && (s & 1)) e = ctc_label(lab, s >> 1, V);
    extoff[i] = e * 4;
    livef[i] = s < S ? 1.f : 0.f;
    skipf[i] = ((s < S) && (s & 1) && (s >= 3) && (ctc_label(lab, s >> 1, V) != ctc_label(lab, (s >> 1) - 1, V))) ? 1.f : 0.f;
    skipb[i] = (GRAD && (s + 2 < S) && (s & 1) && (ctc_label(lab, (s >> 1) + 1, V) != ctc_label(lab, s >> 1, V))) ? 1.f : 0.f;
  }
  float* aw = GRAD ? alpha_ws + static_cast<size_t>(b) * T * (32 * SPL) + lane * SPL : nullptr;
  int* kaw = GRAD ? ka_ws + static_cast<size_t>(b) * T * 32 + lane : nullptr;
  // virtual step -1: alpha(0) = 1 makes the recurrence itself produce alpha_0(0) = y_0(blank), alpha_0(1) = y_0(l_1)
  float a[SPL];
#pragma unroll
  for (int i = 0; i < SPL; ++i) a[i] = (lane == 0 && i == 0) ? 1.f : 0.f;
  bool unsafe = false;  // some frame's logit spread is outside what the linear sweep can represent
  int K = 0;            // this lane's states are a[] * 2^K
  int kpend = 0;        // exponent of this lane's largest value after the previous step: folded into this step's y
  bool empty = lane != 0;  // no mass in this lane yet
  for (int tb = 0; tb < nblk; ++tb) {
    if (tb + 1 < nblk) { stage(tb + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncwarp();
    const float* cur = blk + (tb & 1) * blk_floats;
    const int nf = min(kCtcBlk, T - tb * kCtcBlk);
    // log2-domain log-sum-exp of frame (tb*32 + lane); lanes walk the classes in rotated order (bank spread)
    float lse_l = 0.f;
    if (lane < nf) {
      const float* row = cur + lane * V;
      float m = -INFINITY, mn = INFINITY;
      int v = lane % V;
      for (int k = 0; k < V; ++k) { m = fmaxf(m, row[v]); mn = fminf(mn, row[v]); v = (v + 1 == V) ? 0 : v + 1; }
      float z = 0.f;
      for (int k = 0; k < V; ++k) { z += ex2a((row[v] - m) * kLog2e); v = (v + 1 == V) ? 0 : v + 1; }
      lse_l = m * kLog2e + lg2a(z);
      // a single emission probability below ~2^-106 cannot be held next to this lane's other states in fp32 (the exponent
      // is folded in one step late): such frames - logit spreads beyond ~69 nats, NaN / inf - go to the log-domain kernel
      if (!((m - mn) * kLog2e <= kCtcLinearMaxRange)) unsafe = true;
      if (GRAD) lse[tb * kCtcBlk + lane] = lse_l;
    }
    const char* rowb = reinterpret_cast<const char*>(cur);
#pragma unroll 1
    for (int tt = 0; tt < nf; ++tt, rowb += V * 4) {
      const float lt2 = __shfl_sync(0xffffffffu, lse_l, tt) + static_cast<float>(kpend);
      float y[SPL];
#pragma unroll
      for (int i = 0; i < SPL; ++i)
        y[i] = ex2a(fmaf(*reinterpret_cast<const float*>(rowb + extoff[i]), kLog2e, -lt2)) * livef[i];
      const int Kup = __shfl_up_sync(0xffffffffu, K, 1);
      float up1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1);
      float up2 = __shfl_up_sync(0xffffffffu, a[SPL - 2], 1);
      K = empty ? Kup : K;                                     // nothing here yet: take over the neighbour's scale
      const int d = max(-126, min(126, Kup - K));
      const float f = lane == 0 ? 0.f : __uint_as_float(static_cast<unsigned>(d + 127) << 23);  // 2^(Kup - K)
      up1 *= f;
      up2 *= f;
      float na[SPL];
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        const float q1 = (i == 0) ? up1 : a[i - 1];
        const float q2 = (i == 0) ? up2 : (i == 1 ? up1 : a[i - 2]);
        na[i] = fmaf(skipf[i], q2, a[i] + q1) * y[i];
      }
      K += kpend;  // the exponent folded into y has been applied
      float mx = na[0];
#pragma unroll
      for (int i = 0; i < SPL; ++i) { a[i] = na[i]; mx = fmaxf(mx, na[i]); }
      // exponent for the NEXT step from this step's largest value of this lane (off the chain); 0 (no mass / denormal) and
      // >= 254 (inf / NaN) leave the scale alone
      const unsigned ebits = __float_as_uint(mx) >> 23;
      kpend = (ebits - 1u) < 253u ? static_cast<int>(ebits) - 127 : 0;
      empty = mx == 0.f;
      if (GRAD) {
        const size_t t = static_cast<size_t>(tb) * kCtcBlk + tt;
#pragma unroll
        for (int i = 0; i < SPL; ++i) aw[t * (32 * SPL) + i] = a[i];
        kaw[t * 32] = K;
      }
    }
    __syncwarp();
  }
  // log2 p(l|x) = lse2 over states S-1 and S-2 of log2(a) + K (they may live in neighbouring lanes)
  float fin = kLogZero;
#pragma unroll
  for (int i = 0; i < SPL; ++i) {
    const int s = lane * SPL + i;
    if ((s == S - 1 || (s == S - 2 && S >= 2)) && a[i] > 0.f) fin = lse2(fin, lg2a(a[i]) + static_cast<float>(K));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fin = lse2(fin, __shfl_xor_sync(0xffffffffu, fin, o));
  const bool redo = __any_sync(0xffffffffu, unsafe);
  if (lane == 0) nll[b] = redo ? __uint_as_float(kCtcRedoMark) : (fin > 0.5f * kLogZero ? -fin * kLn2 : INFINITY);
  if constexpr (GRAD) {
    if (lane == 0) mode[b] = redo ? 1 : 0;
    if (redo || !(fin > 0.5f * kLogZero)) return;  // log-domain redo, or infeasible: the gradient kernel writes NaNs without reading
    // ---------------- beta sweep (frames in reverse) ----------------
    float* bw = beta_ws + static_cast<size_t>(b) * T * (32 * SPL) + lane * SPL;
    int* kbw = kb_ws + static_cast<size_t>(b) * T * 32 + lane;
    float bt[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) bt[i] = (lane * SPL + i == S - 1) ? 1.f : 0.f;  // virtual step T
    const float nl31 = lane == 31 ? 0.f : 1.f;
    K = 0;
    kpend = 0;
    empty = (S - 1) / SPL != lane;
    __syncwarp();
    stage(nblk - 1);
    for (int tb = nblk - 1; tb >= 0; --tb) {
      if (tb > 0) { stage(tb - 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
      __syncwarp();
      const float* cur = blk + (tb & 1) * blk_floats;
      const int nf = min(kCtcBlk, T - tb * kCtcBlk);
      const char* rowb = reinterpret_cast<const char*>(cur) + static_cast<size_t>(nf - 1) * V * 4;
#pragma unroll 1
      for (int tt = nf - 1; tt >= 0; --tt, rowb -= V * 4) {
        const size_t t = static_cast<size_t>(tb) * kCtcBlk + tt;
        const float lt2 = lse[t] + static_cast<float>(kpend);
        float y[SPL];
#pragma unroll
        for (int i = 0; i < SPL; ++i)
          y[i] = ex2a(fmaf(*reinterpret_cast<const float*>(rowb + extoff[i]), kLog2e, -lt2)) * livef[i];
        const int Kdn = __shfl_down_sync(0xffffffffu, K, 1);
        float dn1 = __shfl_down_sync(0xffffffffu, bt[0], 1);
        float dn2 = __shfl_down_sync(0xffffffffu, bt[SPL > 1 ? 1 : 0], 1);
        K = (empty && lane != 31) ? Kdn : K;
        const int d = max(-126, min(126, Kdn - K));
        const float f = nl31 * __uint_as_float(static_cast<unsigned>(d + 127) << 23);  // 2^(Kdn - K), nothing beyond lane 31
        dn1 *= f;
        dn2 *= f;
        float nb[SPL];
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          const float q1 = (i + 1 < SPL) ? bt[i + 1] : dn1;
          const float q2 = (i + 2 < SPL) ? bt[i + 2] : (i + 1 < SPL ? dn1 : dn2);
          nb[i] = fmaf(skipb[i], q2, bt[i] + q1) * y[i];
        }
        K += kpend;
        float mx = nb[0];
#pragma unroll
        for (int i = 0; i < SPL; ++i) { bt[i] = nb[i]; mx = fmaxf(mx, nb[i]); }
        const unsigned ebits = __float_as_uint(mx) >> 23;
        kpend = (ebits - 1u) < 253u ? static_cast<int>(ebits) - 127 : 0;
        empty = mx == 0.f;
#pragma unroll
        for (int i = 0; i < SPL; ++i) bw[t * (32 * SPL) + i] = bt[i];
        kbw[t * 32] = K;
      }
      __syncwarp();
    }
  }
}

// d nll_b / d logits[b,t,:] = softmax(logits[b,t,:]) - occupancy_t, occupancy_t(v) = sum_{s: ext(s)=v} alpha_t(s) beta_t(s) / y_t(v)
// normalised per frame (in exact arithmetic the sum over s is p(l|x) at every t; normalising by the per-frame sum keeps
// fp32 drift out of the gradient: each row sums to zero to rounding). One warp per frame, alphas/betas from the sweep's
// workspace (log2 domain, state s = lane*SPL + i).
constexpr int kGradWarps = 8;
template <int SPL>
__global__ void __launch_bounds__(kGradWarps * 32)
ctc_grad_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, int B, int T, int V, int L, int blank,
                const float* __restrict__ nll, const float* __restrict__ alpha_ws, const float* __restrict__ beta_ws, float* __restrict__ grad,
                const int* __restrict__ ka_ws, const int* __restrict__ kb_ws, const int* __restrict__ mode) {
  extern __shared__ __align__(16) float smem_cg[];  // [kGradWarps][Vpad]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Vpad = (V + 31) & ~31;
  float* occ = smem_cg + warp * Vpad;
  const int64_t frame = static_cast<int64_t>(blockIdx.x) * kGradWarps + warp;
  if (frame >= static_cast<int64_t>(B) * T) return;
  const int b = static_cast<int>(frame / T), t = static_cast<int>(frame % T);
  const float* row = logits + frame * V;
  float* gr = grad + frame * V;
  const int32_t* lab = labels + static_cast<size_t>(b) * L;
  const float nl = nll[b];
  const bool feasible = nl < 1.0e29f;  // +inf = no valid alignment (label longer than the frames allow)
  // frame log-sum-exp
  float m = -INFINITY;
  for (int v = lane; v < V; v += 32) m = fmaxf(m, row[v]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float z = 0.f;
  for (int v = lane; v < V; v += 32) z += __expf(row[v] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
  const float lt = m + __logf(z);
  for (int v = lane; v < Vpad; v += 32) occ[v] = 0.f;
  __syncwarp();
  if (feasible) {
    int cnt = 0;
    for (int i = lane; i < L; i += 32) cnt += (ctc_label(lab, i, V) != blank) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    const int S = 2 * cnt + 1;
    const size_t off = (static_cast<size_t>(b) * T + t) * (32 * SPL) + lane * SPL;
    // workspace of this sequence: log2 alpha / beta (mode 1, log-domain sweep) or scaled linear values + per-lane exponents
    const bool linear = mode != nullptr && mode[b] == 0;
    const float kab = linear ? static_cast<float>(ka_ws[(static_cast<size_t>(b) * T + t) * 32 + lane] + kb_ws[(static_cast<size_t>(b) * T + t) * 32 + lane]) : 0.f;
    float e[SPL];
    int ext[SPL];
    float mx = kLogZero;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
      const int s = lane * SPL + i;
      int ex = blank;
      if (s < S && (s & 1)) ex = ctc_label(lab, s >> 1, V);
      ext[i] = ex;
      const float em = (row[ex] - lt) * kLog2e;
      float ab;
      if (linear) {
        const float pa = alpha_ws[off + i], pb = beta_ws[off + i];
        ab = (pa > 0.f && pb > 0.f) ? lg2a(pa) + lg2a(pb) + kab : kLogZero;
      } else {
        ab = alpha_ws[off + i] + beta_ws[off + i];
      }
      e[i] = (s < S) ? ab - em : kLogZero;
      mx = fmaxf(mx, e[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
      e[i] = (e[i] - mx > -100.f) ? ex2a(e[i] - mx) : 0.f;
      sum += e[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < SPL; ++i)
      if (e[i] != 0.f) atomicAdd(&occ[ext[i]], e[i] * inv);
  }
  __syncwarp();
  for (int v = lane; v < V; v += 32) {
    const float y = ex2a((row[v] - lt) * kLog2e);
    gr[v] = feasible ? (y - occ[v]) : __int_as_float(0x7fc00000);
  }
}

// ------------------------------------------------------------------------------------------------
// greedy decode: one warp per sequence
// ------------------------------------------------------------------------------------------------
// One CTA per sequence. Frames are staged 384 at a time in shared memory with coalesced loads (row pitch V + 1 floats;
// thread = frame then walks its row without bank conflicts); the first version let every lane read its own 240-byte row
// from global memory - 32 lines per load instruction - and took 45 us for 23.6 MB.
constexpr int kDecThreads = 384;  // frames per staged chunk = threads per CTA (T = 384: one pass)
__global__ void __launch_bounds__(kDecThreads)
greedy_decode_kernel(const float* __restrict__ logits, int B, int T, int V, int blank, int32_t* __restrict__ ids_out,
                     int32_t* __restrict__ lens, int chunk) {
  extern __shared__ int32_t smem_dec[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int b = blockIdx.x;
  int32_t* ids = smem_dec;                                              // [T + 1]
  float* rows = reinterpret_cast<float*>(smem_dec + ((T + 1 + 3) & ~3));  // [kDecThreads][V + 1]
  const int pitch = V + 1;
  const float* lg = logits + static_cast<size_t>(b) * T * V;
  for (int t0 = 0; t0 < T; t0 += chunk) {  // chunk <= kDecThreads frames fit the shared-memory staging (large vocabularies: fewer)
    const int nf = min(chunk, T - t0);
    const float* src = lg + static_cast<size_t>(t0) * V;
    const int n = nf * V;
    if (((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (V % 4 == 0)) {
      // 16-byte loads, eight in flight per thread before anything is stored (one load per loop trip left every trip
      // exposed to a full memory round trip: 28 us)
      const float4* src4 = reinterpret_cast<const float4*>(src);
      const int n4 = n / 4;
      for (int base = 0; base < n4; base += kDecThreads * 8) {
        float4 buf[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = base + j * kDecThreads + tid;
          if (idx < n4) buf[j] = __ldg(src4 + idx);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = base + j * kDecThreads + tid;
          if (idx < n4) {
            const int e = idx * 4, r = e / V, c = e - r * V;  // V % 4 == 0: the four values stay inside one row
            float* d = rows + r * pitch + c;
            d[0] = buf[j].x; d[1] = buf[j].y; d[2] = buf[j].z; d[3] = buf[j].w;
          }
        }
      }
    } else {
      const int dr = kDecThreads / V, dc = kDecThreads % V;
      int r = tid / V, c = tid - r * V;
      for (int i = tid; i < n; i += kDecThreads) {
        rows[r * pitch + c] = __ldg(src + i);
        r += dr; c += dc;
        if (c >= V) { c -= V; ++r; }
      }
    }
    __syncthreads();
    if (tid < nf) {
      const float* row = rows + tid * pitch;
      float best = row[0];
      int bi = 0;
      for (int v = 1; v < V; ++v) {
        const float x = row[v];
        if (x > best) { best = x; bi = v; }  // first index on ties
      }
      ids[t0 + tid] = bi;
    }
    __syncthreads();
  }
  if (tid >= 32) return;
  int32_t* dst = ids_out + static_cast<size_t>(b) * T;
  int total = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    bool keep = false;
    int id = 0;
    if (t < T - 1) {
      id = ids[t];
      keep = (id != ids[t + 1]) && (id != blank);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) dst[total + __popc(m & ((1u << lane) - 1u))] = id;
    total += __popc(m);
  }
  // pad the tail with -1 so the buffer is deterministic
  for (int i = total + lane; i < T; i += 32) dst[i] = -1;
  if (lane == 0) lens[b] = total;
}

// One states-per-lane instantiation of the whole CTC call:
//   linear sweep(s) -> log-domain kernel for the sequences the sweep marked as out of range (returns at once for the
//   others; with linear == false it does all of them) -> [gradient of every frame, one warp per frame]
template <int SPL>
int ctc_launch_spl(const float* logits, const int32_t* labels, int B, int T, int V, int L, int blank, float* nll, float* grad,
                   float* ws, size_t smem_log, bool linear, cudaStream_t stream) {
  const int Vpad = (V + 31) & ~31;
  const size_t per = static_cast<size_t>(B) * T * 32;
  float* aws = ws;
  float* bws = ws != nullptr ? ws + per * SPL : nullptr;
  int* kaw = ws != nullptr ? reinterpret_cast<int*>(ws + 2 * per * SPL) : nullptr;
  int* kbw = ws != nullptr ? kaw + per : nullptr;
  int* mode = ws != nullptr ? kbw + per : nullptr;
  if (smem_log > 48 * 1024)
    ISHARA_CUDA_OK(cudaFuncSetAttribute(ctc_kernel<SPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_log)));
  if (linear) {
    const size_t smem_lin = (static_cast<size_t>(2 * kCtcBlk * V) + (grad != nullptr ? ((T + 3) & ~3) : 0)) * sizeof(float);
    if (grad != nullptr) {
      if (smem_lin > 48 * 1024)
        ISHARA_CUDA_OK(cudaFuncSetAttribute(ctc_alpha_linear_kernel<SPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_lin)));
      ctc_alpha_linear_kernel<SPL, true><<<B, 32, smem_lin, stream>>>(logits, labels, B, T, V, L, blank, nll, aws, bws, kaw, kbw, mode);
    } else {
      if (smem_lin > 48 * 1024)
        ISHARA_CUDA_OK(cudaFuncSetAttribute(ctc_alpha_linear_kernel<SPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_lin)));
      ctc_alpha_linear_kernel<SPL, false><<<B, 32, smem_lin, stream>>>(logits, labels, B, T, V, L, blank, nll, nullptr, nullptr, nullptr, nullptr, nullptr);
    }
    ISHARA_CUDA_OK(cudaGetLastError());
    note_launch();
  }
  ctc_kernel<SPL><<<B, 32, smem_log, stream>>>(logits, labels, B, T, V, L, blank, nll, grad, aws, bws, linear ? 1 : 0);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  if (grad != nullptr) {
    const int64_t frames = static_cast<int64_t>(B) * T;
    ctc_grad_kernel<SPL><<<static_cast<unsigned>((frames + kGradWarps - 1) / kGradWarps), kGradWarps * 32, kGradWarps * Vpad * sizeof(float), stream>>>(
        logits, labels, B, T, V, L, blank, nll, aws, bws, grad, kaw, kbw, linear ? mode : nullptr);
    ISHARA_CUDA_OK(cudaGetLastError());
    note_launch();
  }
  return 0;
}

}  // namespace

// alphas | betas of the gradient pass: caller-owned (per handle / per call), so two handles, devices or streams never
// share it and nothing is allocated inside a launcher that may run under stream capture
size_t ctc_workspace_bytes(int B, int T, int L) {
  const int spl = (2 * L + 1 + 31) / 32;
  const int SPL = spl <= 5 ? 5 : 9;
  // alphas | betas (floats) | per-lane exponents of the linear sweeps (2 x [B, T, 32] ints) | mode [B] ints
  return 2 * static_cast<size_t>(B) * T * 32 * SPL * sizeof(float) + 2 * static_cast<size_t>(B) * T * 32 * sizeof(int) +
         static_cast<size_t>(B) * sizeof(int);
}

int ctc_loss_launch(const float* logits, const int32_t* labels, int B, int T, int V, int L, int blank, float* nll,
                    float* grad, float* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (B <= 0 || T <= 0 || V <= 0 || L < 0 || blank < 0 || blank >= V) {
    set_last_error("ctc_loss: bad shape");
    return 2;
  }
  const int spl = (2 * L + 1 + 31) / 32;
  const int Vpad = (V + 31) & ~31;
  const size_t smem = (static_cast<size_t>((T + 3) & ~3) + Vpad + 2 * kCtcBlk * V) * sizeof(float);
  if (smem > 200 * 1024) {
    set_last_error("ctc_loss: T / num_classes too large for the per-sequence shared-memory tables");
    return 2;
  }
  const int SPL = spl <= 5 ? 5 : 9;
  if (spl > 9) {
    set_last_error("ctc_loss: label length above 143 is not supported");
    return 2;
  }
  static const int linear_fwd = getenv("ISHARA_CTC_LINEAR") ? atoi(getenv("ISHARA_CTC_LINEAR")) : 1;
  float* ws = nullptr;
  if (grad != nullptr) {
    const size_t need = ctc_workspace_bytes(B, T, L);  // alphas | betas | exponents | mode
    if (workspace == nullptr || workspace_bytes < need) {
      set_last_error("ctc_loss: the gradient pass needs a caller-owned workspace of ctc_workspace_bytes(B, T, L) bytes");
      return 2;
    }
    ws = workspace;
  }
  return SPL == 5 ? ctc_launch_spl<5>(logits, labels, B, T, V, L, blank, nll, grad, ws, smem, linear_fwd != 0, stream)
                  : ctc_launch_spl<9>(logits, labels, B, T, V, L, blank, nll, grad, ws, smem, linear_fwd != 0, stream);
}

int greedy_decode_launch(const float* logits, int B, int T, int V, int blank, int32_t* ids_out, int32_t* lens,
                         cudaStream_t stream) {
  if (B <= 0 || T <= 0 || V <= 0) {
    set_last_error("greedy_decode: bad shape");
    return 2;
  }
  const size_t ids_words = static_cast<size_t>((T + 1 + 3) & ~3);
  const size_t budget = 96 * 1024 / sizeof(int32_t);  // two CTAs per SM
  if (ids_words + static_cast<size_t>(V + 1) > 200 * 1024 / sizeof(int32_t)) {
    set_last_error("greedy_decode: T / num_classes too large for the shared-memory staging");
    return 2;
  }
  int chunk = kDecThreads;
  if (ids_words + static_cast<size_t>(chunk) * (V + 1) > budget) {
    const size_t room = ids_words + static_cast<size_t>(V + 1) > budget ? static_cast<size_t>(V + 1) : budget - ids_words;
    chunk = static_cast<int>(std::max<size_t>(1, std::min<size_t>(kDecThreads, room / (V + 1))));
  }
  const size_t smem = (ids_words + static_cast<size_t>(chunk) * (V + 1)) * sizeof(int32_t);
  static size_t smem_attr = 48 * 1024;
  if (smem > smem_attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(greedy_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    smem_attr = smem;
  }
  greedy_decode_kernel<<<B, kDecThreads, smem, stream>>>(logits, B, T, V, blank, ids_out, lens, chunk);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace ishara
