// ishara_b200 — tcgen05/TMEM multi-head self-attention core for the get_model shape (dh = 32, T <= 384; v4 takes any T,
// v3 whole 128-row blocks).
//
// Reference: MultiHeadSelfAttention.call (nb:conv-hybrid-model c5:102-118): softmax(q k^T * dim^-0.5) v
// on the per-head-interleaved qkv tensor [B*T, H*3*dh] (head h owns columns [96h, 96h+96) = q|k|v). One CTA per
// (sequence, head):
//   * ONE TMA tensor map with a [128 rows x 64 cols] box reads q_h|k_h of 128 tokens as a 128-byte-swizzled tile; the
//     three boxes of a sequence sit contiguously in smem, so the same bytes serve as the A operand (Q rows of a
//     128-query block, K-offsets 0/32 B inside the swizzle row) and as the B operand (all 384 K rows, offsets 64/96 B);
//   * S = Q K^T [128 x 384] fp32 lives in TMEM (384 columns, two tcgen05.mma of N = 192 per K-step);
//   * four softmax warps own one query row per thread (tcgen05.ld 32x32b: lane == row): exact row max, exp2 in fp32,
//     P written as bf16 into a K-major swizzled smem tile (the A operand of the second GEMM), row sum kept in registers;
//   * O = P V [128 x 32] accumulates in TMEM (32 more columns) against V^T, transposed once per CTA into the canonical
//     K-major layout; the epilogue scales by 1/rowsum and writes bf16 straight to the merged-head output.
// The [H,T,T] score tensor never leaves the SM. The softmax exponentials (T^2 per head) bound this kernel on the MUFU
// pipe; the legacy mma.sync kernel (attention.cu) stays for other head sizes / longer sequences / relative positions.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace ishara {
namespace {

constexpr int kDH = 32;
constexpr int kQB = 128;           // query rows per block
constexpr int kMaxT = 384;
constexpr int kPolyPairs = 0;       // column pairs (of 16 per chunk) whose exp2 runs on the FMA pipe; measured no gain at 9 => off
constexpr int kThreadsTc = 288;    // warp 0: TMA + MMA issue + TMEM alloc; warps 1-8: softmax / epilogue (2 per lane quarter)
constexpr int kQKBytes = kMaxT * 128;            // q|k tile: T rows x 128 B
constexpr int kVtBytes = (kMaxT / 64) * 32 * 128;  // V^T: T/64 k-blocks of [32 x 128 B]
constexpr int kPBytes = (kMaxT / 64) * kQB * 128;  // P: T/64 k-blocks of [128 x 128 B]

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= 0 on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5, degree-4 polynomial for
// 2^f (relative error < 5e-5, far below bf16's 4e-3), exponent added with integer arithmetic. Two lanes at a time with
// the packed fp32 instructions. The exponentials and the fp32->bf16 packs of the softmax share the quarter-rate XU pipe;
// moving part of the exponentials here is what balances it (same idea as the published FlashAttention-4 softmax).
__device__ __forceinline__ void exp2_poly2(float& r0, float& r1, float x0, float x1) {
  constexpr float kMagic = 12582912.f;  // 1.5 * 2^23: adding it rounds to the nearest integer in the low mantissa bits
  x0 = fmaxf(x0, -125.f);
  x1 = fmaxf(x1, -125.f);
  float t0, t1, n0, n1, f0, f1, p0, p1;
  fadd2(t0, t1, x0, x1, kMagic, kMagic);
  fadd2(n0, n1, t0, t1, -kMagic, -kMagic);
  fadd2(f0, f1, x0, x1, -n0, -n1);
  ffma2(p0, p1, f0, f1, 9.618129107628477e-3f, 9.618129107628477e-3f, 5.550410866482158e-2f, 5.550410866482158e-2f);
  ffma2(p0, p1, p0, p1, f0, f1, 2.402265069591007e-1f, 2.402265069591007e-1f);
  ffma2(p0, p1, p0, p1, f0, f1, 6.931471805599453e-1f, 6.931471805599453e-1f);
  ffma2(p0, p1, p0, p1, f0, f1, 1.f, 1.f);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// read 32 bf16 (one 64-byte half of a 128-byte swizzled row) and return their squared L2 norm
__device__ __forceinline__ float row_half_norm2(const uint8_t* tile, int row, int first_chunk) {
  const uint8_t* rp = tile + static_cast<size_t>(row) * 128;
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 v = *reinterpret_cast<const uint4*>(rp + (((first_chunk + j) ^ (row & 7)) << 4));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float lo = bf16_lo(w[e]), hi = bf16_hi(w[e]);
      acc = fmaf(lo, lo, acc);
      acc = fmaf(hi, hi, acc);
    }
  }
  return acc;
}

// Pipeline (v3). The softmax is computed against a per-row UPPER BOUND of the scores instead of the exact row maximum:
// s_ij * scale <= scale * |q_i| * max_j |k_j| (Cauchy-Schwarz). Softmax is invariant to the subtracted constant and bf16
// has fp32's exponent range, so P keeps its relative precision; a warp whose bound is so loose that 2^(s-bound) could
// underflow falls back to an exact-maximum pre-pass. Knowing the bound BEFORE the scores exist removes the max pass and
// lets the key axis be processed in two independent halves with no online rescaling:
//   MMA thread:   QK(0,0) QK(0,1) | PV(b,0) QK(b+1,0) PV(b,1) QK(b+1,1) ...      (S halves and O double-buffered in TMEM)
//   softmax warps: sm(0,0) sm(0,1) | sm(b,0) epilogue(b-1) sm(b,1) ...            (never wait for a PV they just enabled)
__global__ void __launch_bounds__(kThreadsTc, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const bf16* __restrict__ qkv, bf16* __restrict__ out,
               const uint8_t* __restrict__ key_mask, int T, int H, float scale_log2, float* __restrict__ lse_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* qk = smem;                          // [T][128 B] swizzled: cols 0-31 q, 32-63 k
  uint8_t* vt = qk + kQKBytes;                 // [T/64][32][128 B] swizzled V^T
  uint8_t* pp = vt + kVtBytes;                 // [T/64][128][128 B] swizzled P (bf16)
  float* mb = reinterpret_cast<float*>(pp + kPBytes);  // [T] additive key bias (log2 domain), only with a mask
  float* xsum = mb + kMaxT;                    // [2 column halves][128 rows] row-sum exchange (+ spare)
  uint64_t* bars = reinterpret_cast<uint64_t*>(xsum + 512);
  uint64_t* qk_full = bars;        // TMA landed
  uint64_t* s_full = bars + 1;     // [2] S half = Q K_half^T complete
  uint64_t* p_ready = bars + 3;    // [2] softmax warps wrote P half (and are done reading S half)
  uint64_t* pv_done = bars + 5;    // [2] PV MMAs of that half retired: P half may be overwritten
  uint64_t* o_full = bars + 7;     // [2] O buffer complete
  uint64_t* o_empty = bars + 9;    // [2] epilogue read O buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  uint32_t* kmax_bits = tmem_slot + 1;  // max_j |k_j|^2 as float bits (non-negative floats order like unsigned ints)

  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nqb = T / kQB, nkb = T / 64, nkb2 = nkb / 2;
  const int ld = 3 * kDH * H;

  if (tid == 0) {
    tma_prefetch_desc(&tmQK);
    mbar_init(qk_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 256);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 256);
    }
    *kmax_bits = 0u;
    mbar_fence_init();
    mbar_arrive_expect_tx(qk_full, static_cast<uint32_t>(T) * 128u);
    for (int r = 0; r < nqb; ++r) tma_load_2d(qk + r * kQB * 128, &tmQK, qk_full, h * 3 * kDH, b * T + r * kQB);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);

  // ---- V^T into the canonical K-major, 128B-swizzled B-operand layout: element (d, key) ----
  {
    const bf16* vbase = qkv + static_cast<size_t>(b) * T * ld + h * 3 * kDH + 2 * kDH;
    for (int key = tid; key < T; key += kThreadsTc) {
      const uint4* vp = reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(key) * ld);
      uint4 vv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) vv[i] = __ldg(vp + i);
      const int kb = key >> 6, kin = key & 63;
      uint8_t* blk = vt + kb * 4096;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t w[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
          for (int hl = 0; hl < 2; ++hl) {
            const int d = 8 * i + 2 * e + hl;
            const uint32_t off = static_cast<uint32_t>(d) * 128u + ((static_cast<uint32_t>(kin >> 3) ^ (d & 7)) << 4) + (kin & 7) * 2;
            *reinterpret_cast<unsigned short*>(blk + off) = static_cast<unsigned short>(hl ? (w[e] >> 16) : (w[e] & 0xFFFFu));
          }
        }
      }
      if (key_mask != nullptr) mb[key] = key_mask[static_cast<size_t>(b) * T + key] ? 0.f : -1.0e9f * 1.4426950408889634f;
    }
  }
  fence_proxy_async_smem();  // V^T (generic-proxy writes) must be visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int half_cols = T / 2;                 // S is produced as two N = T/2 halves (N <= 192)
  const uint32_t tmem_O = tmem_base + 384;     // two 32-column O buffers
  const uint32_t IDESC_S = umma_idesc(128, half_cols, 1);
  constexpr uint32_t IDESC_O = umma_idesc(128, 32, 1);

  if (warp == 0) {
    if (lane == 0) {
      mbar_wait(qk_full, 0);
      const uint32_t qk_addr = smem_base;
      const uint32_t vt_addr = qk_addr + kQKBytes;
      const uint32_t p_addr = vt_addr + kVtBytes;
      auto issue_qk = [&](int blk, int half) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_bf16(tmem_base + half * half_cols, umma_desc_sw128(qk_addr + blk * kQB * 128 + k * 32),
                    umma_desc_sw128(qk_addr + half * half_cols * 128 + 64 + k * 32), IDESC_S, k);
        umma_commit(&s_full[half]);
      };
      issue_qk(0, 0);
      issue_qk(0, 1);
      for (int blk = 0; blk < nqb; ++blk) {
        const uint32_t obuf = blk & 1;
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&p_ready[half], blk & 1);
          if (half == 0 && blk >= 2) mbar_wait(&o_empty[obuf], ((blk >> 1) & 1) ^ 1u);  // block blk-2's epilogue read this O buffer
          tc_fence_after();
          for (int j = 0; j < nkb2; ++j) {
            const int kb = half * nkb2 + j;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_O + obuf * 32, umma_desc_sw128(p_addr + kb * (kQB * 128) + k * 32),
                        umma_desc_sw128(vt_addr + kb * 4096 + k * 32), IDESC_O, (half | j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&pv_done[half]);
          if (half == 1) umma_commit(&o_full[obuf]);
          if (blk + 1 < nqb) issue_qk(blk + 1, half);  // this S half was consumed (p_ready): next block's scores can go in
        }
      }
    }
  } else {
    // ===================== softmax + epilogue =====================
    const int q = warp & 3;                       // TMEM lane quarter of this warp
    const int hh = (warp - 1) >> 2;               // which part of each key half / which 16 output columns
    const int r = q * 32 + lane;                  // row within the 128-query block
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const bool masked = key_mask != nullptr;
    const int nch = half_cols / 32;               // 32-column chunks per key half (6, 4, 2)
    const int c_lo = hh * (nch / 2), c_hi = c_lo + nch / 2;

    // max_j |k_j| of this head (for the score bound)
    mbar_wait(qk_full, 0);
    {
      float km = 0.f;
      for (int key = tid - 32; key < T; key += 256) km = fmaxf(km, row_half_norm2(qk, key, 4));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) km = fmaxf(km, __shfl_xor_sync(0xffffffffu, km, o));
      if (lane == 0) atomicMax(kmax_bits, __float_as_uint(km));
    }
    named_bar_sync(5, 256);
    const float kmax = sqrtf(__uint_as_float(*kmax_bits));

    float sum_prev = 0.f;
    auto epilogue = [&](int blk, float sum) {
      const uint32_t obuf = blk & 1;
      mbar_wait(&o_full[obuf], (blk >> 1) & 1);
      tc_fence_after();
      uint32_t raw[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(raw[0]), "=r"(raw[1]), "=r"(raw[2]), "=r"(raw[3]), "=r"(raw[4]), "=r"(raw[5]), "=r"(raw[6]), "=r"(raw[7]),
            "=r"(raw[8]), "=r"(raw[9]), "=r"(raw[10]), "=r"(raw[11]), "=r"(raw[12]), "=r"(raw[13]), "=r"(raw[14]), "=r"(raw[15])
          : "r"(tmem_O + obuf * 32 + lane_addr + hh * 16)
          : "memory");
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_empty[obuf]);
      const float inv = 1.f / sum;
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(__uint_as_float(raw[2 * j]) * inv, __uint_as_float(raw[2 * j + 1]) * inv);
      uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * T + blk * kQB + r) * (kDH * H) + h * kDH + hh * 16);
      dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    };

    for (int blk = 0; blk < nqb; ++blk) {
      // score bound of this query row in the log2 domain
      float bound = scale_log2 * sqrtf(row_half_norm2(qk, blk * kQB + r, 0)) * kmax;
      const bool exact = __any_sync(0xffffffffu, bound > 100.f);  // loose bound: 2^(s - bound) could underflow
      if (exact) {
        mbar_wait(&s_full[0], blk & 1);
        mbar_wait(&s_full[1], blk & 1);
        tc_fence_after();
        float mx = -INFINITY;
        for (int c = 0; c < 2 * nch; ++c) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + lane_addr + c * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            mx = fmaxf(mx, fmaf(__uint_as_float(raw[j]), scale_log2, masked ? mb[c * 32 + j] : 0.f));
        }
        bound = mx;
      }
      const float nb = -bound;
      float s0 = 0.f, s1 = 0.f;
      for (int half = 0; half < 2; ++half) {
        mbar_wait(&s_full[half], blk & 1);
        if (blk > 0) mbar_wait(&pv_done[half], (blk - 1) & 1);  // previous block's PV MMAs have finished reading this P half
        tc_fence_after();
        for (int c = c_lo; c < c_hi; ++c) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + half * half_cols + lane_addr + c * 32, raw);
          tmem_ld_wait();
          const int key0 = half * half_cols + c * 32;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a0, a1;
            if (masked) {
              a0 = fmaf(__uint_as_float(raw[2 * j]), scale_log2, mb[key0 + 2 * j] + nb);
              a1 = fmaf(__uint_as_float(raw[2 * j + 1]), scale_log2, mb[key0 + 2 * j + 1] + nb);
            } else {
              ffma2(a0, a1, __uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]), scale_log2, scale_log2, nb, nb);
            }
            float p0, p1;
            if (j < kPolyPairs) {
              exp2_poly2(p0, p1, a0, a1);
            } else {
              p0 = ex2f(a0);
              p1 = ex2f(a1);
            }
            fadd2(s0, s1, s0, s1, p0, p1);
            pk[j] = pack_bf16x2(p0, p1);
          }
          const int kc = key0 >> 5;  // global 32-key chunk index
          const uint32_t rowbase = smem_base + kQKBytes + kVtBytes + (kc >> 1) * (kQB * 128) + static_cast<uint32_t>(r) * 128u;
          const uint32_t x = static_cast<uint32_t>(r & 7);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(rowbase + ((static_cast<uint32_t>((kc & 1) * 4 + j) ^ x) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2],
                         pk[4 * j + 3]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_ready[half]);
        if (half == 0 && blk > 0) epilogue(blk - 1, sum_prev);  // overlaps the PV MMAs just enabled
      }
      // row sum over both column parts (the partner warp of this lane quarter holds the other chunks)
      float sum = s0 + s1;
      xsum[((blk & 1) * 2 + hh) * 128 + r] = sum;
      named_bar_sync(1 + q, 64);
      sum += xsum[((blk & 1) * 2 + (hh ^ 1)) * 128 + r];
      sum_prev = sum;
      // training: P = 2^(s*scale_log2 - bound) / sum  =>  row log-sum-exp (log2 domain) = bound + log2(sum)
      if (lse_out != nullptr && hh == 0) lse_out[(static_cast<size_t>(b) * H + h) * T + blk * kQB + r] = bound + log2f(sum);
    }
    epilogue(nqb - 1, sum_prev);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}


// ---------------------------------------------------------------------------------------------------------------------
// v4 ("tc2"): two CTAs per SM. The v3 kernel above keeps one (sequence, head) per SM at a time (170 KB of shared memory,
// 448 TMEM columns), so its per-head fixed costs - TMA latency, the V^T transpose, waiting for the first scores, the last
// PV and the epilogue - are never hidden: measured 207 us per launch against a 65 us MUFU floor. Here the key axis is
// streamed in 64-key blocks through a 2-slot ring (S: 2 x 64 TMEM columns, P: 2 x 16 KB), which brings a CTA down to
// 109 KB / 256 TMEM columns: two heads are resident per SM and one's exponentials run under the other's latencies.
// The exponent offset is v3's per-row Cauchy-Schwarz bound scale * |q_i| * max_j |k_j| (softmax is invariant to it); when
// the bound of ANY row of the head is so loose that 2^(s - bound) could underflow (decided once per CTA in the prologue,
// so the MMA thread knows too), the head runs an extra sweep of Q K^T over all key blocks that yields the exact maxima:
//   MMA thread, per 128-query block:  [QK(kb) for kb < nkb   - sweep 0, exact maxima, only for such heads]
//                                      QK(kb), PV(kb-1) interleaved, PV(last)   [sweep 1: P = 2^(s c - offset), O += P V]
//   8 softmax warps (lane == row):    [sweep 0: tcgen05.ld -> running max]  sweep 1: tcgen05.ld -> ex2 -> bf16 P tile -> PV
// S ring: three 64-column slots (the MMA thread runs two key blocks ahead), P ring: two 16 KB slots.
//   epilogue per block: O / rowsum -> bf16 -> global (merged heads); O is double buffered across query blocks.
// timeline tracing of one CTA (profiling build only, make TRACE=1)
#ifdef ISHARA_TRACE_BUILD
__device__ long long* g_attn_trace = nullptr;
#define ATT_TRACE(ev_) do { if (g_attn_trace != nullptr && blockIdx.y == gridDim.y / 2 && blockIdx.x == 3) g_attn_trace[(ev_)] = clock64(); } while (0)
#else
#define ATT_TRACE(ev_) do { } while (0)
#endif
constexpr int kP2Bytes = 2 * kQB * 128;  // two [128 x 64-key] P tiles
constexpr int kThreadsTc2 = 320;         // warp 0: TMA + QK issue + TMEM alloc; warps 1-8: softmax / epilogue; warp 9: PV issue
__global__ void __launch_bounds__(kThreadsTc2, 2)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQK, const bf16* __restrict__ qkv, bf16* __restrict__ out,
                const uint8_t* __restrict__ key_mask, int T, int H, float scale_log2, float* __restrict__ lse_out, int stagger_cycles) {
  // The two CTAs that share an SM start in lockstep in the first wave, so their prologues and epilogues (no MUFU work)
  // coincide instead of hiding under each other's exponentials: the second resident CTA of every SM waits half a period.
  if (stagger_cycles > 0) {
    const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x;
    unsigned nsm;
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsm));
    if (lin >= nsm && lin < 2 * nsm) {
      if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < stagger_cycles) { }
      }
      __syncthreads();
    }
  }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* qk = smem;                          // [T][128 B] swizzled: cols 0-31 q, 32-63 k
  uint8_t* vt = qk + kQKBytes;                 // [T/64][32][128 B] swizzled V^T
  float* mb = reinterpret_cast<float*>(vt + kVtBytes + kP2Bytes);  // [T] additive key bias (log2 domain), only with a mask
  float* qn = mb + kMaxT;                      // [T] |q_i| of every query row (score bound), written once in the prologue
  float* xch = qn + kMaxT;                     // [2 (max | sum)][2 column halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 512);
  uint64_t* qk_full = bars;        // TMA landed
  uint64_t* s_full = bars + 1;     // [2] S slot holds Q K_kb^T
  uint64_t* s_free = bars + 3;     // [2] softmax warps are done reading the S slot
  uint64_t* p_ready = bars + 5;    // [2] softmax warps wrote the P slot
  uint64_t* pv_done = bars + 7;    // [2] PV MMAs of that P slot retired
  uint64_t* o_full = bars + 9;     // [2] O buffer complete
  uint64_t* o_free = bars + 11;    // [2] epilogue read the O buffer
  uint64_t* s_full3 = bars + 13;   // third S slot
  uint64_t* s_free3 = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
  uint32_t* kmax_bits = tmem_slot + 1;  // max_j |k_j|^2 and max_i |q_i|^2 as float bits (non-negative floats order like unsigned ints)
  uint32_t* qmax_bits = tmem_slot + 2;
  auto sfull = [&](uint32_t slot) { return slot == 2u ? s_full3 : &s_full[slot]; };
  auto sfree = [&](uint32_t slot) { return slot == 2u ? s_free3 : &s_free[slot]; };

  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // T need not be a multiple of the block sizes (the notebook's INPUT_SHAPE has T = 176): the last query block / key block
  // is partial. Rows past T inside the loaded tile are the next sequence's (or TMA zero fill at the end of the tensor):
  // finite garbage that is never stored as a query and is masked out as a key.
  const int nqb = (T + kQB - 1) / kQB, nkb = (T + 63) / 64;
  const bool ragged = (T & 63) != 0;
  const int ld = 3 * kDH * H;

  if (tid == 0) {
    tma_prefetch_desc(&tmQK);
    mbar_init(qk_full, 1);
    mbar_init(s_full3, 1);
    mbar_init(s_free3, 8);      // one arrive per softmax warp
    *kmax_bits = 0u;
    *qmax_bits = 0u;
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 8);
      mbar_init(&p_ready[i], 8);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 8);
    }
    mbar_fence_init();
    mbar_arrive_expect_tx(qk_full, static_cast<uint32_t>(nqb) * kQB * 128u);
    for (int r = 0; r < nqb; ++r) tma_load_2d(qk + r * kQB * 128, &tmQK, qk_full, h * 3 * kDH, b * T + r * kQB);
  }
  if (tid == 0) ATT_TRACE(0);
  if (warp == 0) tmem_alloc(tmem_slot, 256);

  // ---- V^T into the canonical K-major, 128B-swizzled B-operand layout: element (d, key) ----
  {
    const bf16* vbase = qkv + static_cast<size_t>(b) * T * ld + h * 3 * kDH + 2 * kDH;
    for (int key = tid; key < nkb * 64; key += kThreadsTc2) {
      const uint4* vp = reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(key) * ld);
      uint4 vv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) vv[i] = key < T ? __ldg(vp + i) : make_uint4(0u, 0u, 0u, 0u);  // padded keys: V = 0 (their P is 0 too)
      const int kb = key >> 6, kin = key & 63;
      uint8_t* blk = vt + kb * 4096;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t w[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
          for (int hl = 0; hl < 2; ++hl) {
            const int d = 8 * i + 2 * e + hl;
            const uint32_t off = static_cast<uint32_t>(d) * 128u + ((static_cast<uint32_t>(kin >> 3) ^ (d & 7)) << 4) + (kin & 7) * 2;
            *reinterpret_cast<unsigned short*>(blk + off) = static_cast<unsigned short>(hl ? (w[e] >> 16) : (w[e] & 0xFFFFu));
          }
        }
      }
      if (key_mask != nullptr || ragged) {
        const bool valid = key < T && (key_mask == nullptr || key_mask[static_cast<size_t>(b) * T + key] != 0);
        mb[key] = valid ? 0.f : -1.0e9f * 1.4426950408889634f;
      }
    }
  }
  fence_proxy_async_smem();  // V^T (generic-proxy writes) must be visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 32) ATT_TRACE(1);
  // largest |k_j| and |q_i| of this head: the score bound, and whether it is tight enough for every row
  mbar_wait(qk_full, 0);
  if (tid == 32) ATT_TRACE(2);
  // the first three score tiles of query block 0 are the same whether or not the exact-maximum sweep is needed: issue
  // them now, so that the tensor core works while every warp computes the row norms below
  const int npre = nkb < 3 ? nkb : 3;
  if (warp == 0 && lane == 0) {
    const uint32_t tb = *tmem_slot;
    constexpr uint32_t IDESC_S0 = umma_idesc(128, 64, 1);
    tc_fence_after();
    for (int kb = 0; kb < npre; ++kb) {
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2)
        umma_bf16(tb + kb * 64, umma_desc_sw128(smem_base + k2 * 32), umma_desc_sw128(smem_base + kb * 64 * 128 + 64 + k2 * 32), IDESC_S0, k2);
      umma_commit(sfull(static_cast<uint32_t>(kb)));
      if (kb < 8) ATT_TRACE(40 + kb);
    }
  }
  __syncwarp();
  {
    float km = 0.f, qm = 0.f;
    for (int t = tid; t < nqb * kQB; t += kThreadsTc2) {
      const float q2 = row_half_norm2(qk, t, 0);
      qn[t] = sqrtf(q2);
      if (t < T) {
        km = fmaxf(km, row_half_norm2(qk, t, 4));
        qm = fmaxf(qm, q2);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      km = fmaxf(km, __shfl_xor_sync(0xffffffffu, km, o));
      qm = fmaxf(qm, __shfl_xor_sync(0xffffffffu, qm, o));
    }
    if (lane == 0) {
      atomicMax(kmax_bits, __float_as_uint(km));
      atomicMax(qmax_bits, __float_as_uint(qm));
    }
  }
  __syncthreads();
  if (tid == 32) ATT_TRACE(3);
  const float kmax = sqrtf(__uint_as_float(*kmax_bits));
  const bool exact = scale_log2 * sqrtf(__uint_as_float(*qmax_bits)) * kmax > 100.f;  // loose bound: 2^(s - bound) could underflow
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 192;     // two 32-column O buffers behind the three 64-column S slots
  constexpr uint32_t IDESC_S = umma_idesc(128, 64, 1);
  constexpr uint32_t IDESC_O = umma_idesc(128, 32, 1);
  const uint32_t qk_addr = smem_base;
  const uint32_t vt_addr = qk_addr + kQKBytes;
  const uint32_t p_addr = vt_addr + kVtBytes;

  if (warp == 0) {
    // ===================== QK issuer: runs up to three key blocks ahead of the softmax warps =====================
    if (lane == 0) {
      uint32_t g = 0;  // running use counter of the S slots
      for (int blk = 0; blk < nqb; ++blk) {
        for (int sweep = exact ? 0 : 1; sweep < 2; ++sweep) {
          for (int kb = 0; kb < nkb; ++kb, ++g) {
            if (g < static_cast<uint32_t>(npre)) continue;  // issued before the norm pass
            const uint32_t slot = g % 3u;
            if (g >= 3) mbar_wait(sfree(slot), ((g / 3u) - 1u) & 1u);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(tmem_base + slot * 64, umma_desc_sw128(qk_addr + blk * kQB * 128 + k * 32),
                        umma_desc_sw128(qk_addr + kb * 64 * 128 + 64 + k * 32), IDESC_S, k);
            umma_commit(sfull(slot));
            if (g < 8) ATT_TRACE(40 + g);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== PV issuer: O += P V as soon as a P slot is written =====================
    if (lane == 0) {
      uint32_t pc = 0;  // running use counter of the P slots
      for (int blk = 0; blk < nqb; ++blk) {
        const uint32_t obuf = blk & 1;
        for (int kb = 0; kb < nkb; ++kb, ++pc) {
          const uint32_t ps = pc & 1u;
          mbar_wait(&p_ready[ps], (pc >> 1) & 1u);
          if (kb == 0 && blk >= 2) mbar_wait(&o_free[obuf], static_cast<uint32_t>(((blk >> 1) - 1) & 1));  // block blk-2's epilogue read this O buffer
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_O + obuf * 32, umma_desc_sw128(p_addr + ps * (kQB * 128) + k * 32), umma_desc_sw128(vt_addr + kb * 4096 + k * 32),
                      IDESC_O, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&pv_done[ps]);
          if (pc < 8) ATT_TRACE(50 + pc);
        }
        umma_commit(&o_full[obuf]);
        ATT_TRACE(60 + blk);
      }
    }
  } else {
    // ===================== softmax + epilogue =====================
    const int q = warp & 3;                       // TMEM lane quarter of this warp
    const int hh = (warp - 1) >> 2;               // which 32 of the 64 keys of a block / which 16 output columns
    const int r = q * 32 + lane;                  // row within the 128-query block
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const bool masked = key_mask != nullptr || ragged;
    float* xmax = xch;                            // [2][128]
    float* xsum = xch + 256;                      // [2][128]
    uint32_t g = 0, pc = 0;
    // O / sum -> bf16 -> merged-head output for query block eb. Called one key block INTO the next query block (and after
    // the last one), so the wait for the block's last P V MMA hides under the next block's first exponentials.
    auto epilogue = [&](int eb, float esum) {
      const uint32_t eo = eb & 1;
      mbar_wait(&o_full[eo], (eb >> 1) & 1u);
      tc_fence_after();
      if (tid == 32) ATT_TRACE(64 + eb);
      uint32_t raw16[16];
      tmem_ld16(tmem_O + eo * 32 + lane_addr + hh * 16, raw16);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[eo]);
      const float inv = 1.f / esum;
      uint32_t po[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) po[j] = pack_bf16x2(__uint_as_float(raw16[2 * j]) * inv, __uint_as_float(raw16[2 * j + 1]) * inv);
      if (eb * kQB + r < T) {
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * T + eb * kQB + r) * (kDH * H) + h * kDH + hh * 16);
        dst[0] = make_uint4(po[0], po[1], po[2], po[3]);
        dst[1] = make_uint4(po[4], po[5], po[6], po[7]);
      }
      if (tid == 32) ATT_TRACE(68 + eb);
    };
    float sum_prev = 1.f;
    for (int blk = 0; blk < nqb; ++blk) {
      // ---- exponent offset: the Cauchy-Schwarz bound of this row, or (sweep 0) the exact maximum of s * scale_log2 + mask bias ----
      float mx = scale_log2 * qn[blk * kQB + r] * kmax;
      if (exact) {
      mx = -INFINITY;
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        const uint32_t slot = g % 3u;
        mbar_wait(sfull(slot), (g / 3u) & 1u);
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld32(tmem_base + slot * 64 + hh * 32 + lane_addr, raw);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sfree(slot));
        const int key0 = kb * 64 + hh * 32;
        if (masked) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaf(__uint_as_float(raw[j]), scale_log2, mb[key0 + j]));
        } else {
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; j += 2) { m0 = fmaxf(m0, __uint_as_float(raw[j])); m1 = fmaxf(m1, __uint_as_float(raw[j + 1])); }
          mx = fmaxf(mx, fmaxf(m0, m1) * scale_log2);  // scale_log2 > 0
        }
      }
      xmax[hh * 128 + r] = mx;
      named_bar_sync(1 + q, 64);
      mx = fmaxf(mx, xmax[(hh ^ 1) * 128 + r]);
      }
      const float nb = -mx;
      // ---- sweep 1: P = 2^(s c - max) as bf16 into the P ring, row sums in fp32 ----
      float s0 = 0.f, s1 = 0.f;
      for (int kb = 0; kb < nkb; ++kb, ++g, ++pc) {
        const uint32_t slot = g % 3u, ps = pc & 1u;
        mbar_wait(sfull(slot), (g / 3u) & 1u);
        if (tid == 32 && pc < 18) ATT_TRACE(4 + 2 * pc);
        if (pc >= 2) mbar_wait(&pv_done[ps], ((pc >> 1) - 1u) & 1u);  // the PV MMAs that last read this P slot have retired
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld32(tmem_base + slot * 64 + hh * 32 + lane_addr, raw);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sfree(slot));
        const int key0 = kb * 64 + hh * 32;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a0, a1;
          if (masked) {
            a0 = fmaf(__uint_as_float(raw[2 * j]), scale_log2, mb[key0 + 2 * j] + nb);
            a1 = fmaf(__uint_as_float(raw[2 * j + 1]), scale_log2, mb[key0 + 2 * j + 1] + nb);
          } else {
            ffma2(a0, a1, __uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]), scale_log2, scale_log2, nb, nb);
          }
          const float p0 = ex2f(a0), p1 = ex2f(a1);
          fadd2(s0, s1, s0, s1, p0, p1);
          pk[j] = pack_bf16x2(p0, p1);
        }
        const uint32_t rowbase = p_addr + ps * (kQB * 128) + static_cast<uint32_t>(r) * 128u;
        const uint32_t x = static_cast<uint32_t>(r & 7);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(rowbase + ((static_cast<uint32_t>(hh * 4 + j) ^ x) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[ps]);
        if (tid == 32 && pc < 18) ATT_TRACE(5 + 2 * pc);
        if (kb == 0 && blk > 0) epilogue(blk - 1, sum_prev);
      }
      // row sum over both key halves of every block
      float sum = s0 + s1;
      xsum[hh * 128 + r] = sum;
      named_bar_sync(1 + q, 64);
      sum += xsum[(hh ^ 1) * 128 + r];
      const bool row_ok = blk * kQB + r < T;
      if (lse_out != nullptr && hh == 0 && row_ok) lse_out[(static_cast<size_t>(b) * H + h) * T + blk * kQB + r] = mx + log2f(sum);
      sum_prev = sum;
    }
    epilogue(nqb - 1, sum_prev);
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) ATT_TRACE(72);
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace

bool attention_tc_applicable(const AttnArgs& a) {
  static const int tc2 = getenv("ISHARA_ATTN_TC2") ? atoi(getenv("ISHARA_ATTN_TC2")) : 1;
  // v4 (attn_tc2_kernel) takes any T <= 384 (partial last blocks); v3 needs whole 128-row blocks
  return a.dh == kDH && a.pos == nullptr && a.T <= kMaxT && a.T >= 1 && (tc2 || a.T % kQB == 0);
}

int attention_tc_launch(const AttnArgs& a, cudaStream_t stream) {
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, a.qkv, TM_BF16, static_cast<uint64_t>(a.B) * a.T, static_cast<uint64_t>(3) * kDH * a.H,
                        static_cast<uint64_t>(3) * kDH * a.H, kQB, 64);
  if (rc) return rc;
  static const int tc2 = getenv("ISHARA_ATTN_TC2") ? atoi(getenv("ISHARA_ATTN_TC2")) : 1;
  if (tc2) {
    // v4: 64-key streaming, two CTAs per SM
    const int smem2 = kQKBytes + kVtBytes + kP2Bytes + 2 * kMaxT * 4 + 2048 + 256 + 1024;
    static bool attr2 = false;
    if (!attr2) {
      ISHARA_CUDA_OK(cudaFuncSetAttribute(attn_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      ISHARA_CUDA_OK(cudaFuncSetAttribute(attn_tc2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      attr2 = true;
    }
#ifdef ISHARA_TRACE_BUILD
    static long long* tbuf = nullptr;
    static int printed = 0;
    const bool tracing = printed < 2 && a.B >= 8;
    if (tracing) {
      if (tbuf == nullptr) ISHARA_CUDA_OK(cudaMalloc(&tbuf, 128 * sizeof(long long)));
      ISHARA_CUDA_OK(cudaMemsetAsync(tbuf, 0, 128 * sizeof(long long), stream));
      ISHARA_CUDA_OK(cudaMemcpyToSymbolAsync(g_attn_trace, &tbuf, sizeof(tbuf), 0, cudaMemcpyHostToDevice, stream));
    }
#endif
    static const int stagger = getenv("ISHARA_ATTN_STAGGER") ? atoi(getenv("ISHARA_ATTN_STAGGER")) : 0;
    attn_tc2_kernel<<<dim3(a.H, a.B), kThreadsTc2, smem2, stream>>>(tm, a.qkv, a.out, a.key_mask, a.T, a.H, a.scale * 1.4426950408889634f,
                                                               a.lse_out, stagger);
    ISHARA_CUDA_OK(cudaGetLastError());
    note_launch();
#ifdef ISHARA_TRACE_BUILD
    if (tracing) {
      ISHARA_CUDA_OK(cudaStreamSynchronize(stream));
      long long hbuf[128];
      ISHARA_CUDA_OK(cudaMemcpy(hbuf, tbuf, sizeof(hbuf), cudaMemcpyDeviceToHost));
      ++printed;
      fprintf(stderr, "attn trace B=%d T=%d (cycles since CTA start): 1 V^T done | 2 qk landed | 3 norms | 4+2i / 5+2i softmax block i: scores seen / P written | "
                      "40+ QK issue | 50+ PV issue | 60+ o_full commit | 64+ o_full seen | 68+ epilogue done | 72 end\n ", a.B, a.T);
      for (int e = 1; e < 80; ++e) if (hbuf[e]) fprintf(stderr, " %d:%lld", e, hbuf[e] - hbuf[0]);
      fprintf(stderr, "\n");
    }
#endif
    return 0;
  }
  const int smem = kQKBytes + kVtBytes + kPBytes + kMaxT * 4 + 2048 + 128 + 1024;
  static bool attr = false;
  if (!attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  attn_tc_kernel<<<dim3(a.H, a.B), kThreadsTc, smem, stream>>>(tm, a.qkv, a.out, a.key_mask, a.T, a.H,
                                                             a.scale * 1.4426950408889634f, a.lse_out);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace ishara
