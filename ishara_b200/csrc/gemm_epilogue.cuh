// ishara_b200 — fused GEMM epilogue building blocks shared by the single-CTA (gemm_tc.cu) and CTA-pair (gemm_tc2.cu)
// tcgen05 kernels: per-thread row view of the TMEM accumulator, packed-fp32 epilogue math, per-warp staging + TMA store.
#pragma once
#include "kernels.h"
#include "ptx.cuh"

namespace ishara {
namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;  // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kAStageBytes = kBM * kBK * 2;
constexpr int kStgBytes = 128 * 128;  // staging per 4 epilogue warps: 4 boxes of 32 rows x 128 bytes
constexpr int kWarpStgBytes = 32 * 128;
constexpr int kMaxSmem = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100

struct EpiThread {
  int row;        // global row
  bool valid;     // row < M
  int seq;        // row / rows_per_seq
  int t;          // row % rows_per_seq
  uint32_t taddr; // TMEM address of (lane quarter, buffer col 0)
};

// All epilogue math runs on packed fp32 pairs (FFMA2 / FADD2 / FMUL2): the epilogue, not the MMA, paces these
// small-K GEMMs (ncu: tensor pipe 22 % active while the 8 epilogue warps are 85 % busy), so instruction count matters.

// v = (acc + bias) * gate + rowtab for 32 consecutive columns starting at global column `col0`
__device__ __forceinline__ void epi_affine(float (&v)[32], const GemmEpi& ep, const EpiThread& th, int col0, int ldn) {
  if (ep.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      fadd2(v[4 * j + 0], v[4 * j + 1], v[4 * j + 0], v[4 * j + 1], b.x, b.y);
      fadd2(v[4 * j + 2], v[4 * j + 3], v[4 * j + 2], v[4 * j + 3], b.z, b.w);
    }
  }
  if (ep.gate != nullptr && th.valid) {
    const float4* g4 = reinterpret_cast<const float4*>(ep.gate + static_cast<size_t>(th.seq) * ldn + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 g = __ldg(g4 + j);
      fmul2(v[4 * j + 0], v[4 * j + 1], v[4 * j + 0], v[4 * j + 1], g.x, g.y);
      fmul2(v[4 * j + 2], v[4 * j + 3], v[4 * j + 2], v[4 * j + 3], g.z, g.w);
    }
  }
  if (ep.rowtab != nullptr && th.valid) {
    const float4* t4 = reinterpret_cast<const float4*>(ep.rowtab + static_cast<size_t>(th.t) * ldn + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 g = __ldg(t4 + j);
      fadd2(v[4 * j + 0], v[4 * j + 1], v[4 * j + 0], v[4 * j + 1], g.x, g.y);
      fadd2(v[4 * j + 2], v[4 * j + 3], v[4 * j + 2], v[4 * j + 3], g.z, g.w);
    }
  }
}

// swish(x) = x * sigmoid(x) = h + h * tanh(h), h = x / 2: FMUL2, 2 x MUFU.TANH, FFMA2 per pair
__device__ __forceinline__ void epi_swish(float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float h0, h1;
    fmul2(h0, h1, v[2 * j], v[2 * j + 1], 0.5f, 0.5f);
    const float t0 = fast_tanh(h0), t1 = fast_tanh(h1);
    ffma2(v[2 * j], v[2 * j + 1], h0, h1, t0, t1, h0, h1);
  }
}

__device__ __forceinline__ void epi_resid(float (&v)[32], const GemmEpi& ep, const EpiThread& th, int col0) {
  if (ep.resid != nullptr && th.valid) {
    const uint4* r4 = reinterpret_cast<const uint4*>(ep.resid + static_cast<size_t>(th.row) * ep.ld_resid + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 q = __ldg(r4 + j);
      fadd2(v[8 * j + 0], v[8 * j + 1], v[8 * j + 0], v[8 * j + 1], bf16_lo(q.x), bf16_hi(q.x));
      fadd2(v[8 * j + 2], v[8 * j + 3], v[8 * j + 2], v[8 * j + 3], bf16_lo(q.y), bf16_hi(q.y));
      fadd2(v[8 * j + 4], v[8 * j + 5], v[8 * j + 4], v[8 * j + 5], bf16_lo(q.z), bf16_hi(q.z));
      fadd2(v[8 * j + 6], v[8 * j + 7], v[8 * j + 6], v[8 * j + 7], bf16_lo(q.w), bf16_hi(q.w));
    }
  }
}

// running row statistics in two packed lanes (even / odd columns)
struct RowStats {
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  __device__ __forceinline__ void add(const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      fadd2(s0, s1, s0, s1, v[2 * j], v[2 * j + 1]);
      ffma2(q0, q1, v[2 * j], v[2 * j + 1], v[2 * j], v[2 * j + 1], q0, q1);
    }
  }
  __device__ __forceinline__ void finish(int n, float eps, float* mean, float* rstd) const {
    const float m = (s0 + s1) * (1.f / n);
    const float var = fmaxf((q0 + q1) * (1.f / n) - m * m, 0.f);
    *mean = m;
    *rstd = rsqrtf(var + eps);
  }
};

// v = ((v - mean) * rstd) * gamma + beta: two FFMA2 per pair
__device__ __forceinline__ void epi_layernorm(float (&v)[32], const float* g, const float* b, float mean, float rstd,
                                              int col0) {
  const float4* g4 = reinterpret_cast<const float4*>(g + col0);
  const float4* b4 = reinterpret_cast<const float4*>(b + col0);
  const float nm = -mean * rstd;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 gg = __ldg(g4 + j), bb = __ldg(b4 + j);
    float x0, x1, x2, x3;
    ffma2(x0, x1, v[4 * j + 0], v[4 * j + 1], rstd, rstd, nm, nm);
    ffma2(x2, x3, v[4 * j + 2], v[4 * j + 3], rstd, rstd, nm, nm);
    ffma2(v[4 * j + 0], v[4 * j + 1], x0, x1, gg.x, gg.y, bb.x, bb.y);
    ffma2(v[4 * j + 2], v[4 * j + 3], x2, x3, gg.z, gg.w, bb.z, bb.w);
  }
}

__device__ __forceinline__ void to_float(float (&v)[32], const uint32_t (&raw)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
}
__device__ __forceinline__ void to_raw(uint32_t (&raw)[32], const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(v[j]);
}

// write 32 values of this thread's row into the warp's swizzled staging box (32 rows x 128 bytes, 16-byte chunks
// XOR-ed with row%8 — identical to CU_TENSOR_MAP_SWIZZLE_128B, so the TMA store un-swizzles it).
template <bool F32>
__device__ __forceinline__ void stage_write(uint32_t stg, int r, int sub, const float (&v)[32]) {
  const uint32_t rowbase = stg + static_cast<uint32_t>(r) * 128u;
  const uint32_t x = static_cast<uint32_t>(r & 7);
  if constexpr (F32) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      st_shared_v4(rowbase + ((static_cast<uint32_t>(j) ^ x) << 4), __float_as_uint(v[4 * j + 0]),
                   __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      st_shared_v4(rowbase + ((static_cast<uint32_t>(sub * 4 + j) ^ x) << 4), pack_bf16x2(v[8 * j + 0], v[8 * j + 1]),
                   pack_bf16x2(v[8 * j + 2], v[8 * j + 3]), pack_bf16x2(v[8 * j + 4], v[8 * j + 5]),
                   pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
  }
}

// Per-WARP staging + TMA store: every epilogue warp owns the 32 rows of its TMEM lane quarter, stages them in its own
// 4 KB box(es) and issues its own bulk tensor store, so the epilogue needs no cross-warp barrier at all.
struct WarpStore {
  uint32_t base;   // smem address of this warp's staging boxes
  uint32_t iter;   // running box counter (selects the buffer)
  int lane;
  bool single;     // one box per warp instead of two
  bool skip_store = false;
  bool skip_fence = false;
  __device__ __forceinline__ uint32_t acquire() {
    if (lane == 0) {  // the store that last used this box has finished reading it
      if (single) tma_store_wait_read<0>();
      else tma_store_wait_read<1>();
    }
    __syncwarp();
    return base + (single ? 0u : (iter & 1u) * kWarpStgBytes);
  }
  __device__ __forceinline__ void release(const CUtensorMap* tm, uint32_t buf, int c0, int r0) {
    if (!skip_fence) fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && !skip_store) {
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                       reinterpret_cast<uint64_t>(tm)),
                   "r"(buf), "r"(c0), "r"(r0)
                   : "memory");
      tma_store_commit();
    }
    ++iter;
  }
};

// Walk NCH 32-column chunks of this thread's accumulator row with the NEXT chunk's tcgen05.ld already in flight while
// the current one is processed (f(raw, chunk)); the load latency disappears behind the epilogue math.
template <int NCH, class F>
__device__ __forceinline__ void chunk_loop(uint32_t taddr, F&& f) {
  static_assert(NCH % 2 == 0, "chunk_loop handles chunk pairs");
  uint32_t ra[32], rb[32];
  tmem_ld32(taddr, ra);
#pragma unroll 1
  for (int c = 0; c < NCH; c += 2) {
    tmem_ld_fence(ra);
    tmem_ld32(taddr + (c + 1) * 32, rb);
    f(ra, c);
    tmem_ld_fence(rb);
    if (c + 2 < NCH) tmem_ld32(taddr + (c + 2) * 32, ra);
    f(rb, c + 1);
  }
}

}  // namespace
}  // namespace ishara
