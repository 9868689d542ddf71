// ishara_b200 — fused GEMM epilogue building blocks shared by the single-CTA (gemm_tc.cu) and CTA-pair (gemm_tc2.cu)
// tcgen05 kernels: per-thread row view of the TMEM accumulator, packed-fp32 epilogue math, per-warp staging + TMA store.
#pragma once
#include "kernels.h"
#include "ptx.cuh"

namespace ishara {
namespace {

// Timing-bisect bits (GemmEpi::dbg, ISHARA_GEMM_DBG) and timeline tracing exist only in the profiling build
// (make TRACE=1 -> libishara_b200_trace.so); the shipped kernels contain neither the branches nor the clock reads.
#ifdef ISHARA_TRACE_BUILD
#define ISHARA_DBG_BIT(ep_, bit_) (((ep_).dbg & (bit_)) != 0)
#else
#define ISHARA_DBG_BIT(ep_, bit_) false
#endif
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

// Residual prefetch: the 32 bf16 of the NEXT chunk's residual row segment are requested while the current chunk is
// processed (and, for the first chunk of a tile, before the accumulator is even ready), so the L2/HBM latency of the
// residual stream never sits on the epilogue's critical path (timeline trace: it used to cost ~1.8k cycles per chunk).
struct ResidRegs {
  uint4 q[4];
};
__device__ __forceinline__ void resid_load(ResidRegs& r, const GemmEpi& ep, const EpiThread& th, int col0) {
  if (ep.resid != nullptr && th.valid) {
    const uint4* r4 = reinterpret_cast<const uint4*>(ep.resid + static_cast<size_t>(th.row) * ep.ld_resid + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) r.q[j] = __ldg(r4 + j);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) r.q[j] = make_uint4(0u, 0u, 0u, 0u);
  }
}
// residual segment of this thread's row from a staged [32 rows x 128 B] swizzled box (filled by TMA): 32 bf16 = 64 B
__device__ __forceinline__ void resid_load_smem(ResidRegs& r, uint32_t box, int lane, int sub) {
  const uint32_t rb = box + static_cast<uint32_t>(lane) * 128u, x = static_cast<uint32_t>(lane & 7);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.q[j].x), "=r"(r.q[j].y), "=r"(r.q[j].z), "=r"(r.q[j].w)
                 : "r"(rb + ((static_cast<uint32_t>(sub * 4 + j) ^ x) << 4))
                 : "memory");
}
__device__ __forceinline__ void resid_add(float (&v)[32], const ResidRegs& r) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 q = r.q[j];
    fadd2(v[8 * j + 0], v[8 * j + 1], v[8 * j + 0], v[8 * j + 1], bf16_lo(q.x), bf16_hi(q.x));
    fadd2(v[8 * j + 2], v[8 * j + 3], v[8 * j + 2], v[8 * j + 3], bf16_lo(q.y), bf16_hi(q.y));
    fadd2(v[8 * j + 4], v[8 * j + 5], v[8 * j + 4], v[8 * j + 5], bf16_lo(q.z), bf16_hi(q.z));
    fadd2(v[8 * j + 6], v[8 * j + 7], v[8 * j + 6], v[8 * j + 7], bf16_lo(q.w), bf16_hi(q.w));
  }
}

// Row statistics of the two warps that share a TMEM lane quarter (same 32 rows, alternate 64-column boxes) are
// combined through a small smem slot + a 64-thread named barrier.
constexpr int kXchBytes = 2 * 2 * 2 * 128 * 16;  // [tile parity][exchange A|B][column half][row] x float4 = 16 KB
__device__ __forceinline__ void exchange_stats(RowStats& rs, float4* slot /*[2][128]*/, int h, int r, uint32_t bar_id) {
  slot[h * 128 + r] = make_float4(rs.s0, rs.s1, rs.q0, rs.q1);
  named_bar_sync(bar_id, 64);
  const float4 o = slot[(h ^ 1) * 128 + r];
  rs.s0 += o.x; rs.s1 += o.y; rs.q0 += o.z; rs.q1 += o.w;
}

// The whole epilogue of one output tile for ONE warp. All 8 epilogue warps work on the same tile (timeline trace: with
// two 4-warp groups on alternate tiles one tile's epilogue lasted 4.5k cycles and gated the MMA two tiles later; eight
// warps halve that): warp (q, h) owns rows [32q, 32q+32) of the tile and the 64-column boxes b = h, h+2, ...
//   tile_col0: first output column of this tile in the output tensor; acc_col0: first accumulator column (n_tile * BN)
template <int BN, bool ROW, bool OUT_F32>
__device__ __forceinline__ void epilogue_tile(const GemmEpi& ep, const EpiThread& th, int n_tile, int N, int row0, int q,
                                              int h, int lane, WarpStore& st, const CUtensorMap* tmO0,
                                              const CUtensorMap* tmO1, float4* xch, int tile_parity, ResidRegs rr,
                                              uint32_t resid_smem = 0u) {
  // resid_smem != 0 (row mode): the residual boxes of this warp were fetched by TMA into its two staging boxes - box k of
  // the tile sits in staging slot (st.iter + k) & 1, the slot the k-th acquire of this tile hands out. Reading the
  // residual per thread from global memory touches 32 different lines per load instruction (~8k L1 wavefronts per tile).
  constexpr int CH = OUT_F32 ? 32 : 64;  // output columns per staging box
  constexpr int SUBS = CH / 32;
  if constexpr (!ROW) {
    const bool glu = ep.act == ACT_GLU;
    const int cols_out = glu ? BN / 2 : BN;
    const int nbox = cols_out / CH;
    for (int b = h; b < nbox; b += 2) {
      const uint32_t buf = st.acquire();
#pragma unroll 1
      for (int sub = 0; sub < SUBS; ++sub) {
        const int tc = b * CH + sub * 32;
        uint32_t raw[32];
        float v[32];
        if (!ISHARA_DBG_BIT(ep, 32)) tmem_ld32(th.taddr + tc, raw);
        ResidRegs rn;
        {  // next chunk of this warp (if any)
          int nb = b, nsub = sub + 1;
          if (nsub == SUBS) { nsub = 0; nb = b + 2; }
          if (nb < nbox) {
            resid_load(rn, ep, th, n_tile * cols_out + nb * CH + nsub * 32);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) rn.q[j] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        tmem_ld_wait();
        to_float(v, raw);
        epi_affine(v, ep, th, n_tile * BN + tc, N);
        if (glu) {
          float u[32];
          tmem_ld32(th.taddr + cols_out + tc, raw);
          tmem_ld_wait();
          to_float(u, raw);
          epi_affine(u, ep, th, n_tile * BN + cols_out + tc, N);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= fast_sigmoid(u[j]);
        } else if (ep.act == ACT_SWISH) {
          epi_swish(v);
        } else if (ep.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (ep.resid != nullptr) resid_add(v, rr);
        rr = rn;
        if (!ISHARA_DBG_BIT(ep, 8)) stage_write<OUT_F32>(buf, lane, sub, v);
      }
      st.release(tmO0, buf, n_tile * cols_out + b * CH, row0);
    }
  } else {
    // ---- full-row epilogue: (resid add) -> [LN0] -> out0 -> [LN1 -> out1]; lane == row; the partner warp (same q,
    //      other h) holds the other boxes of the same rows, so row statistics are exchanged once per LayerNorm; values
    //      are parked in TMEM (tcgen05.st) between passes ----
    static_assert(!OUT_F32 || !ROW, "row epilogue writes bf16");
    constexpr int nbox = BN / 64;
    const bool ln0 = ep.ln0_g != nullptr, ln1 = ep.ln1_g != nullptr;
    const int r = q * 32 + lane;
    float4* slotA = xch + (tile_parity * 2 + 0) * 256;
    float4* slotB = xch + (tile_parity * 2 + 1) * 256;
    RowStats rs;
    const uint32_t iter0 = st.iter;
    int kbox = 0;
    // pass A
    for (int b = h; b < nbox; b += 2, ++kbox) {
      uint32_t buf = 0;
      if (!ln0) buf = st.acquire();
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        const int tc = b * 64 + sub * 32;
        uint32_t raw[32];
        float v[32];
        tmem_ld32(th.taddr + tc, raw);
        ResidRegs rn;
        if (resid_smem != 0u) {
          resid_load_smem(rr, resid_smem + ((iter0 + static_cast<uint32_t>(kbox)) & 1u) * kWarpStgBytes, lane, sub);
          rn = rr;
        } else {
          int nb = b, nsub = sub + 1;
          if (nsub == 2) { nsub = 0; nb = b + 2; }
          if (nb < nbox) {
            resid_load(rn, ep, th, nb * 64 + nsub * 32);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) rn.q[j] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        tmem_ld_wait();
        to_float(v, raw);
        epi_affine(v, ep, th, tc, N);
        if (ep.act == ACT_SWISH) epi_swish(v);
        if (ep.resid != nullptr) resid_add(v, rr);
        rr = rn;
        if (ln0 || ln1) {
          rs.add(v);
          to_raw(raw, v);
          tmem_st32(th.taddr + tc, raw);
        }
        if (!ln0) stage_write<false>(buf, lane, sub, v);
      }
      if (!ln0) st.release(tmO0, buf, b * 64, row0);
    }
    if (ln0 || ln1) {
      tmem_st_wait();
      exchange_stats(rs, slotA, h, r, 1 + q);
    }
    if (ln0) {
      // pass B: u = LN0(v) -> out0; stats of u for LN1
      float mean, rstd;
      rs.finish(BN, ep.ln0_eps, &mean, &rstd);
      rs = RowStats();
      for (int b = h; b < nbox; b += 2) {
        const uint32_t buf = st.acquire();
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int tc = b * 64 + sub * 32;
          uint32_t raw[32];
          float v[32];
          tmem_ld32(th.taddr + tc, raw);
          tmem_ld_wait();
          to_float(v, raw);
          epi_layernorm(v, ep.ln0_g, ep.ln0_b, mean, rstd, tc);
          if (ln1) {
            rs.add(v);
            to_raw(raw, v);
            tmem_st32(th.taddr + tc, raw);
          }
          stage_write<false>(buf, lane, sub, v);
        }
        st.release(tmO0, buf, b * 64, row0);
      }
      if (ln1) {
        tmem_st_wait();
        exchange_stats(rs, slotB, h, r, 1 + q);
      }
    }
    if (ln1) {
      // pass C: out1 = bf16(LN1(stream))
      float mean, rstd;
      rs.finish(BN, ep.ln1_eps, &mean, &rstd);
      for (int b = h; b < nbox; b += 2) {
        const uint32_t buf = st.acquire();
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int tc = b * 64 + sub * 32;
          uint32_t raw[32];
          float v[32];
          tmem_ld32(th.taddr + tc, raw);
          tmem_ld_wait();
          to_float(v, raw);
          epi_layernorm(v, ep.ln1_g, ep.ln1_b, mean, rstd, tc);
          stage_write<false>(buf, lane, sub, v);
        }
        st.release(tmO1, buf, b * 64, row0);
      }
    }
  }
}

// 16-epilogue-warp variant for the plain BN = 256 GEMMs: warp (q, c) owns rows [32q, +32) and ONE 64-column output box
// c of the tile (4 boxes; GLU tiles have 2, warps c >= 2 idle). Slim on purpose: 640 threads leave 96 registers each.
template <int BN>
__device__ __forceinline__ void epilogue_box16(const GemmEpi& ep, const EpiThread& th, int n_tile, int N, int row0, int c,
                                               int lane, WarpStore& st, const CUtensorMap* tmO0) {
  const bool glu = ep.act == ACT_GLU;
  const int cols_out = glu ? BN / 2 : BN;
  if (c * 64 >= cols_out) return;
  const uint32_t buf = st.acquire();
#pragma unroll 1
  for (int sub = 0; sub < 2; ++sub) {
    const int tc = c * 64 + sub * 32;
    uint32_t raw[32];
    float v[32];
    tmem_ld32(th.taddr + tc, raw);
    tmem_ld_wait();
    to_float(v, raw);
    epi_affine(v, ep, th, n_tile * BN + tc, N);
    if (glu) {
      tmem_ld32(th.taddr + cols_out + tc, raw);
      tmem_ld_wait();
      // gate half: (acc + bias) through the same affine, one column at a time to stay within the register budget
      const float* gb = ep.bias != nullptr ? ep.bias + n_tile * BN + cols_out + tc : nullptr;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float g = __uint_as_float(raw[j]);
        if (gb != nullptr) g += __ldg(gb + j);
        v[j] *= fast_sigmoid(g);
      }
    } else if (ep.act == ACT_SWISH) {
      epi_swish(v);
    } else if (ep.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    epi_resid(v, ep, th, n_tile * cols_out + tc);
    stage_write<false>(buf, lane, sub, v);
  }
  st.release(tmO0, buf, n_tile * cols_out + c * 64, row0);
}

// 16-epilogue-warp FULL-ROW variant (BN = 256 = the whole model dimension): warp (q, c) owns rows [32q, +32) and the ONE
// 64-column box c, lane == row. Same dataflow as epilogue_tile's row mode - (bias, gate, rowtab, residual) -> [LN0] -> out0
// -> [LN1 -> out1], values parked in TMEM between passes - but each pass is half as long per warp and four warps per
// scheduler hide the tcgen05.ld / shared-memory latencies that two could not (the 8-warp epilogue paced these GEMMs at
// 15-24k cycles per tile against 2-4k cycles of MMA). One staging box per warp: the residual box lands in it by TMA, is
// read into registers, and the same bytes then stage out0 and, after that store has been read, out1.
//   cvec: [5][256] floats in shared memory: bias | ln0 gamma | ln0 beta | ln1 gamma | ln1 beta (zeros where absent)
//   xch:  [2 slots][4 q][4 c][32] float2 row-statistics exchange; slots alternate per exchange, so a slot is rewritten
//         only after a later named barrier that every reader of its previous contents has already passed
struct Row16State {
  uint32_t xc = 0;  // exchanges done so far by this warp (selects the slot)
};
__device__ __forceinline__ void row16_exchange(RowStats& rs, float2* xch, Row16State& rst, int q, int c, int lane, float eps,
                                               float* mean, float* rstd) {
  float2* slot = xch + (rst.xc & 1u) * 512;
  ++rst.xc;
  slot[(q * 4 + c) * 32 + lane] = make_float2(rs.s0 + rs.s1, rs.q0 + rs.q1);
  named_bar_sync(1 + q, 128);
  float s = 0.f, qq = 0.f;
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const float2 o = slot[(q * 4 + cc) * 32 + lane];
    s += o.x;
    qq += o.y;
  }
  const float m = s * (1.f / 256.f);
  const float var = fmaxf(qq * (1.f / 256.f) - m * m, 0.f);
  *mean = m;
  *rstd = rsqrtf(var + eps);
}
__device__ __forceinline__ void row16_layernorm(float (&v)[32], const float* g_s, const float* b_s, float mean, float rstd) {
  const float nm = -mean * rstd;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 gg = *reinterpret_cast<const float4*>(g_s + 4 * j);
    const float4 bb = *reinterpret_cast<const float4*>(b_s + 4 * j);
    float x0, x1, x2, x3;
    ffma2(x0, x1, v[4 * j + 0], v[4 * j + 1], rstd, rstd, nm, nm);
    ffma2(x2, x3, v[4 * j + 2], v[4 * j + 3], rstd, rstd, nm, nm);
    ffma2(v[4 * j + 0], v[4 * j + 1], x0, x1, gg.x, gg.y, bb.x, bb.y);
    ffma2(v[4 * j + 2], v[4 * j + 3], x2, x3, gg.z, gg.w, bb.z, bb.w);
  }
}
// 32 values of this thread's row into a HALF staging box (32 rows x 64 bytes, CU_TENSOR_MAP_SWIZZLE_64B: 16-byte chunk
// index XOR-ed with (row / 2) % 4)
__device__ __forceinline__ void stage_write_half(uint32_t stg, int r, const float (&v)[32]) {
  const uint32_t rowbase = stg + static_cast<uint32_t>(r) * 64u;
  const uint32_t x = static_cast<uint32_t>(r >> 1) & 3u;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    st_shared_v4(rowbase + ((static_cast<uint32_t>(j) ^ x) << 4), pack_bf16x2(v[8 * j + 0], v[8 * j + 1]),
                 pack_bf16x2(v[8 * j + 2], v[8 * j + 3]), pack_bf16x2(v[8 * j + 4], v[8 * j + 5]),
                 pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
}
// this thread's residual row segment (64 columns of box c) straight from global memory, to be issued well before use
__device__ __forceinline__ void row16_resid_ldg(const GemmEpi& ep, const EpiThread& th, int c, uint4 (&rq)[8]) {
  if (ep.resid != nullptr && th.valid) {
    const uint4* r4 = reinterpret_cast<const uint4*>(ep.resid + static_cast<size_t>(th.row) * ep.ld_resid + c * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) rq[j] = __ldg(r4 + j);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) rq[j] = make_uint4(0u, 0u, 0u, 0u);
  }
}
// resid_bar != nullptr: the residual box of this warp was requested by TMA into its staging box (phase = tile parity)
// HALF: the warp's staging box is 2 KB (32 rows x 32 columns; tmO0 / tmO1 are the matching [32 x 32] maps): every
// 32-column chunk is staged and stored on its own (ffn_tc.cu, where 16 full boxes do not fit beside the weight ring)
template <bool HALF = false>
__device__ __forceinline__ void epilogue_row16(const GemmEpi& ep, const EpiThread& th, int row0, int q, int c, int lane,
                                               WarpStore& st, const CUtensorMap* tmO0, const CUtensorMap* tmO1, float2* xch,
                                               const float* cvec, Row16State& rst, uint64_t* resid_bar, uint32_t resid_phase,
                                               const uint4* rq_pre = nullptr) {
  const bool ln0 = ep.ln0_g != nullptr, ln1 = ep.ln1_g != nullptr;
  const uint32_t stg = st.base;
  uint4 rq[8];
  if (rq_pre != nullptr) {  // the caller requested the residual row segment earlier (row16_resid_ldg)
#pragma unroll
    for (int j = 0; j < 8; ++j) rq[j] = rq_pre[j];
  } else if (resid_bar != nullptr) {
    mbar_wait(resid_bar, resid_phase);
    const uint32_t rb = stg + static_cast<uint32_t>(lane) * 128u, xr = static_cast<uint32_t>(lane & 7);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(rq[j].x), "=r"(rq[j].y), "=r"(rq[j].z), "=r"(rq[j].w)
                   : "r"(rb + ((static_cast<uint32_t>(j) ^ xr) << 4))
                   : "memory");
  } else if (ep.resid != nullptr && th.valid) {
    const uint4* r4 = reinterpret_cast<const uint4*>(ep.resid + static_cast<size_t>(th.row) * ep.ld_resid + c * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) rq[j] = __ldg(r4 + j);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) rq[j] = make_uint4(0u, 0u, 0u, 0u);
  }
  RowStats rs;
  if (!HALF && !ln0 && resid_bar == nullptr) st.acquire();  // the box was last read by the previous tile's out1 / out0 store
  // ---- pass A ----
#pragma unroll
  for (int sub = 0; sub < 2; ++sub) {
    const int col = c * 64 + sub * 32;
    uint32_t raw[32];
    float v[32];
    tmem_ld32(th.taddr + col, raw);
    tmem_ld_wait();
    to_float(v, raw);
    if (ep.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = *reinterpret_cast<const float4*>(cvec + col + 4 * j);
        fadd2(v[4 * j + 0], v[4 * j + 1], v[4 * j + 0], v[4 * j + 1], bb.x, bb.y);
        fadd2(v[4 * j + 2], v[4 * j + 3], v[4 * j + 2], v[4 * j + 3], bb.z, bb.w);
      }
    }
    if (ep.gate != nullptr && th.valid) {
      const float4* g4 = reinterpret_cast<const float4*>(ep.gate + static_cast<size_t>(th.seq) * 256 + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 g = __ldg(g4 + j);
        fmul2(v[4 * j + 0], v[4 * j + 1], v[4 * j + 0], v[4 * j + 1], g.x, g.y);
        fmul2(v[4 * j + 2], v[4 * j + 3], v[4 * j + 2], v[4 * j + 3], g.z, g.w);
      }
    }
    if (ep.rowtab != nullptr && th.valid) {
      const float4* t4 = reinterpret_cast<const float4*>(ep.rowtab + static_cast<size_t>(th.t) * 256 + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 g = __ldg(t4 + j);
        fadd2(v[4 * j + 0], v[4 * j + 1], v[4 * j + 0], v[4 * j + 1], g.x, g.y);
        fadd2(v[4 * j + 2], v[4 * j + 3], v[4 * j + 2], v[4 * j + 3], g.z, g.w);
      }
    }
    if (ep.act == ACT_SWISH) epi_swish(v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 r4 = rq[4 * sub + j];
      fadd2(v[8 * j + 0], v[8 * j + 1], v[8 * j + 0], v[8 * j + 1], bf16_lo(r4.x), bf16_hi(r4.x));
      fadd2(v[8 * j + 2], v[8 * j + 3], v[8 * j + 2], v[8 * j + 3], bf16_lo(r4.y), bf16_hi(r4.y));
      fadd2(v[8 * j + 4], v[8 * j + 5], v[8 * j + 4], v[8 * j + 5], bf16_lo(r4.z), bf16_hi(r4.z));
      fadd2(v[8 * j + 6], v[8 * j + 7], v[8 * j + 6], v[8 * j + 7], bf16_lo(r4.w), bf16_hi(r4.w));
    }
    if (ln0 || ln1) {
      rs.add(v);
      to_raw(raw, v);
      tmem_st32(th.taddr + col, raw);
    }
    if (!ln0) {
      if constexpr (HALF) {
        st.acquire();
        stage_write_half(stg, lane, v);
        st.release(tmO0, stg, col, row0);
      } else {
        stage_write<false>(stg, lane, sub, v);
      }
    }
  }
  if (!HALF && !ln0) st.release(tmO0, stg, c * 64, row0);
  if (ln0 || ln1) tmem_st_wait();
  if (ln0) {
    float mean, rstd;
    row16_exchange(rs, xch, rst, q, c, lane, ep.ln0_eps, &mean, &rstd);
    rs = RowStats();
    if (!HALF && resid_bar == nullptr) st.acquire();
#pragma unroll 1
    for (int sub = 0; sub < 2; ++sub) {
      const int col = c * 64 + sub * 32;
      uint32_t raw[32];
      float v[32];
      tmem_ld32(th.taddr + col, raw);
      tmem_ld_wait();
      to_float(v, raw);
      row16_layernorm(v, cvec + 256 + col, cvec + 512 + col, mean, rstd);
      if (ln1) {
        rs.add(v);
        to_raw(raw, v);
        tmem_st32(th.taddr + col, raw);
      }
      if constexpr (HALF) {
        st.acquire();
        stage_write_half(stg, lane, v);
        st.release(tmO0, stg, col, row0);
      } else {
        stage_write<false>(stg, lane, sub, v);
      }
    }
    if (!HALF) st.release(tmO0, stg, c * 64, row0);
    if (ln1) tmem_st_wait();
  }
  if (ln1) {
    float mean, rstd;
    row16_exchange(rs, xch, rst, q, c, lane, ep.ln1_eps, &mean, &rstd);
    if (!HALF) st.acquire();  // the out0 store has finished reading the box
#pragma unroll 1
    for (int sub = 0; sub < 2; ++sub) {
      const int col = c * 64 + sub * 32;
      uint32_t raw[32];
      float v[32];
      tmem_ld32(th.taddr + col, raw);
      tmem_ld_wait();
      to_float(v, raw);
      row16_layernorm(v, cvec + 768 + col, cvec + 1024 + col, mean, rstd);
      if constexpr (HALF) {
        st.acquire();
        stage_write_half(stg, lane, v);
        st.release(tmO1, stg, col, row0);
      } else {
        stage_write<false>(stg, lane, sub, v);
      }
    }
    if (!HALF) st.release(tmO1, stg, c * 64, row0);
  }
}

// first residual column a warp needs for a tile (matches the first chunk epilogue_tile processes)
template <int BN, bool ROW, bool OUT_F32>
__device__ __forceinline__ int epilogue_first_col(const GemmEpi& ep, int n_tile, int h) {
  constexpr int CH = OUT_F32 ? 32 : 64;
  if constexpr (ROW) return h * 64;
  const int cols_out = ep.act == ACT_GLU ? BN / 2 : BN;
  return n_tile * cols_out + h * CH;
}

}  // namespace
}  // namespace ishara
