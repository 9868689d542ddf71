// ishara_b200 — tcgen05/TMEM multi-head self-attention core for the get_model shape (dh = 32, T <= 384, T % 128 == 0).
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
#include "kernels.h"
#include "ptx.cuh"

namespace ishara {
namespace {

constexpr int kDH = 32;
constexpr int kQB = 128;           // query rows per block
constexpr int kMaxT = 384;
constexpr int kThreadsTc = 288;    // warp 0: TMA + MMA issue + TMEM alloc; warps 1-8: softmax / epilogue (2 per lane quarter)
constexpr int kQKBytes = kMaxT * 128;            // q|k tile: T rows x 128 B
constexpr int kVtBytes = (kMaxT / 64) * 32 * 128;  // V^T: T/64 k-blocks of [32 x 128 B]
constexpr int kPBytes = (kMaxT / 64) * kQB * 128;  // P: T/64 k-blocks of [128 x 128 B]

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreadsTc, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const bf16* __restrict__ qkv, bf16* __restrict__ out,
               const uint8_t* __restrict__ key_mask, int T, int H, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* qk = smem;                          // [T][128 B] swizzled: cols 0-31 q, 32-63 k
  uint8_t* vt = qk + kQKBytes;                 // [T/64][32][128 B] swizzled V^T
  uint8_t* pp = vt + kVtBytes;                 // [T/64][128][128 B] swizzled P (bf16)
  float* mb = reinterpret_cast<float*>(pp + kPBytes);  // [T] additive key bias (log2 domain), only with a mask
  float* xch = mb + kMaxT;                     // [2 kinds][2 column halves][128 rows] row max / row sum exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 512);
  uint64_t* qk_full = bars;      // TMA landed
  uint64_t* s_full = bars + 1;   // S = Q K^T of the current block is complete
  uint64_t* p_ready = bars + 2;  // softmax warps wrote P (and are done reading S)
  uint64_t* o_full = bars + 3;   // O = P V complete
  uint64_t* o_empty = bars + 4;  // epilogue read O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nqb = T / kQB, nkb = T / 64;
  const int ld = 3 * kDH * H;

  if (tid == 0) {
    tma_prefetch_desc(&tmQK);
    mbar_init(qk_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 256);
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
  const uint32_t tmem_S = tmem_base;          // 384 columns
  const uint32_t tmem_O = tmem_base + 384;    // 32 columns
  const int nhalf_cols = T / 2;                              // S is produced as two N = T/2 halves (N <= 256)
  const uint32_t IDESC_S = umma_idesc(128, nhalf_cols, 1);
  constexpr uint32_t IDESC_O = umma_idesc(128, 32, 1);

  if (warp == 0) {
    if (lane == 0) {
      mbar_wait(qk_full, 0);
      const uint32_t qk_addr = smem_base;
      const uint32_t vt_addr = qk_addr + kQKBytes;
      const uint32_t p_addr = vt_addr + kVtBytes;
      for (int blk = 0; blk < nqb; ++blk) {
        // S region is free: block blk-1's softmax has signalled p_ready (waited below before its PV MMAs)
        for (int nh = 0; nh < 2; ++nh)
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_bf16(tmem_S + nh * nhalf_cols, umma_desc_sw128(qk_addr + blk * kQB * 128 + k * 32),
                      umma_desc_sw128(qk_addr + nh * nhalf_cols * 128 + 64 + k * 32), IDESC_S, k);
        umma_commit(s_full);
        mbar_wait(p_ready, blk & 1);
        if (blk > 0) mbar_wait(o_empty, (blk - 1) & 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_O, umma_desc_sw128(p_addr + kb * (kQB * 128) + k * 32),
                      umma_desc_sw128(vt_addr + kb * 4096 + k * 32), IDESC_O, (kb | k) != 0 ? 1u : 0u);
        umma_commit(o_full);
      }
    }
  } else {
    // ===================== softmax + epilogue =====================
    // lane == query row (tcgen05.ld 32x32b); the two warps that share a TMEM lane quarter split the key columns in
    // halves and exchange row max / row sum through smem (64-thread named barrier), so each SM sub-partition has two
    // warps to overlap the MUFU exponentials of one with the FP32/LSU work of the other.
    const int q = warp & 3;                       // TMEM lane quarter of this warp
    const int hh = (warp - 1) >> 2;               // which half of the key columns
    const int r = q * 32 + lane;                  // row within the 128-query block
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const bool masked = key_mask != nullptr;
    const int nc = T / 32, c_lo = hh * (nc / 2), c_hi = c_lo + nc / 2;
    float* xmax = xch;          // [2][128]
    float* xsum = xch + 256;    // [2][128]
    for (int blk = 0; blk < nqb; ++blk) {
      mbar_wait(s_full, blk & 1);
      tc_fence_after();
      // pass 1: exact row maximum of (s * scale + bias) in the log2 domain
      float mx = -INFINITY;
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t raw[32];
        tmem_ld32(tmem_S + lane_addr + c * 32, raw);
        tmem_ld_wait();
        if (masked) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaf(__uint_as_float(raw[j]), scale_log2, mb[c * 32 + j]));
        } else {
          float m0 = __uint_as_float(raw[0]), m1 = __uint_as_float(raw[1]);
#pragma unroll
          for (int j = 2; j < 32; j += 2) {
            m0 = fmaxf(m0, __uint_as_float(raw[j]));
            m1 = fmaxf(m1, __uint_as_float(raw[j + 1]));
          }
          mx = fmaxf(mx, fmaxf(m0, m1) * scale_log2);
        }
      }
      xmax[hh * 128 + r] = mx;
      named_bar_sync(1 + q, 64);
      mx = fmaxf(mx, xmax[(hh ^ 1) * 128 + r]);
      // pass 2: p = 2^(s*scale + bias - max), row sum, bf16 P into the swizzled A-operand tile
      if (blk > 0) mbar_wait(o_full, (blk - 1) & 1);  // P was being read by the previous block's PV MMAs
      float s0 = 0.f, s1 = 0.f;
      const float nmx = -mx;
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t raw[32];
        tmem_ld32(tmem_S + lane_addr + c * 32, raw);
        tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a0, a1;
          if (masked) {
            a0 = fmaf(__uint_as_float(raw[2 * j]), scale_log2, mb[c * 32 + 2 * j] + nmx);
            a1 = fmaf(__uint_as_float(raw[2 * j + 1]), scale_log2, mb[c * 32 + 2 * j + 1] + nmx);
          } else {
            ffma2(a0, a1, __uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]), scale_log2, scale_log2, nmx, nmx);
          }
          p[2 * j] = ex2f(a0);
          p[2 * j + 1] = ex2f(a1);
          fadd2(s0, s1, s0, s1, p[2 * j], p[2 * j + 1]);
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(p[2 * j], p[2 * j + 1]);
        const uint32_t rowbase = smem_base + kQKBytes + kVtBytes + (c >> 1) * (kQB * 128) + static_cast<uint32_t>(r) * 128u;
        const uint32_t x = static_cast<uint32_t>(r & 7);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(rowbase + ((static_cast<uint32_t>((c & 1) * 4 + j) ^ x) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2],
                       pk[4 * j + 3]);
      }
      float sum = s0 + s1;
      xsum[hh * 128 + r] = sum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
      named_bar_sync(1 + q, 64);
      sum += xsum[(hh ^ 1) * 128 + r];
      // epilogue of this block: O / rowsum -> bf16 -> global (each warp of the pair writes 16 of the 32 head columns)
      mbar_wait(o_full, blk & 1);
      tc_fence_after();
      {
        uint32_t raw[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(raw[0]), "=r"(raw[1]), "=r"(raw[2]), "=r"(raw[3]), "=r"(raw[4]), "=r"(raw[5]), "=r"(raw[6]), "=r"(raw[7]),
              "=r"(raw[8]), "=r"(raw[9]), "=r"(raw[10]), "=r"(raw[11]), "=r"(raw[12]), "=r"(raw[13]), "=r"(raw[14]), "=r"(raw[15])
            : "r"(tmem_O + lane_addr + hh * 16)
            : "memory");
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(o_empty);
        const float inv = 1.f / sum;
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(__uint_as_float(raw[2 * j]) * inv, __uint_as_float(raw[2 * j + 1]) * inv);
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * T + blk * kQB + r) * (kDH * H) + h * kDH + hh * 16);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool attention_tc_applicable(const AttnArgs& a) {
  return a.dh == kDH && a.pos == nullptr && a.T <= kMaxT && a.T % kQB == 0;  // T in {128, 256, 384}
}

int attention_tc_launch(const AttnArgs& a, cudaStream_t stream) {
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, a.qkv, TM_BF16, static_cast<uint64_t>(a.B) * a.T, static_cast<uint64_t>(3) * kDH * a.H,
                        static_cast<uint64_t>(3) * kDH * a.H, kQB, 64);
  if (rc) return rc;
  const int smem = kQKBytes + kVtBytes + kPBytes + kMaxT * 4 + 2048 + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  attn_tc_kernel<<<dim3(a.H, a.B), kThreadsTc, smem, stream>>>(tm, a.qkv, a.out, a.key_mask, a.T, a.H,
                                                             a.scale * 1.4426950408889634f);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace ishara
