// ishara_b200 — dense weight gradient on tcgen05: dW[I,O] += X[M,I]^T @ G[M,O]  (SURVEY.md §8 row T15).
//
// Both operands are contracted over their ROW index, so for the MMA they are "MN-major": the 64-column (128-byte)
// TMA boxes of the row-major activations ARE the canonical MN-major SWIZZLE_128B operand tiles — box row = one k
// index, 8-row groups = the 1024-byte swizzle atoms (SBO), consecutive 64-column boxes = consecutive MN atoms (LBO).
// No transposed copy of any activation is written; instruction-descriptor bits 15/16 select the MN-major read.
// One CTA = one 128 x BN tile of dW for one slice of the row axis: warp 0 drives TMA through a 4-stage ring of 64-row
// stages, warp 1 issues 4 x (M128 x BN x K16) MMAs per stage into a TMEM accumulator, warps 2-5 drain it (tcgen05.ld ->
// per-warp smem transpose -> coalesced fp32 red.global.add). Reads X and G once per tile row/column (L2-resident
// re-reads), so the launch is HBM-bound for I, O <= 768.
#include <cstdio>
#include <cstdlib>

#include "ptx.cuh"
#include "train_kernels.h"

#ifdef ISHARA_TRACE_BUILD
#define ISHARA_WGRAD_SKIP(dbg_) (((dbg_) & 1) != 0)
#else
#define ISHARA_WGRAD_SKIP(dbg_) false
#endif

namespace ishara {
namespace {

constexpr int kTcBK = 64;       // rows (k) per stage
constexpr int kTcStages = 4;
constexpr int kTcBoxBytes = kTcBK * 128;                  // one [64 rows x 64 cols] bf16 box
constexpr int kTcStageBytes = (2 + 4) * kTcBoxBytes;      // A: 2 boxes (128 i), B: up to 4 boxes (256 o)
constexpr int kTcThreads = 192;
constexpr int kTcEpiFloats = 32 * 33;

// MN-major SWIZZLE_128B shared-memory descriptor (cute::UMMA::make_umma_desc<Major::MN>, canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): LBO = distance between 64-element MN atoms, SBO = distance
// between 8-row k groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kTcThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, float* __restrict__ dW, int ldw,
                float* __restrict__ dbias, int M, int Ivalid, int Ovalid, int BN, int rows_per_split, int dbg) {
  extern __shared__ uint8_t smem_wg_raw[];
  uint8_t* smem_wg = smem_wg_raw + (((smem_u32(smem_wg_raw) + 1023u) & ~1023u) - smem_u32(smem_wg_raw));  // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* stages = smem_wg;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_wg + kTcStages * kTcStageBytes);
  uint64_t* empty = full + kTcStages;
  uint64_t* tfull = empty + kTcStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  float* epi = reinterpret_cast<float*>(smem_wg + kTcStages * kTcStageBytes + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o0 = blockIdx.x * BN, i0 = blockIdx.y * 128;
  const int m_begin = blockIdx.z * rows_per_split;
  const int m_end = m_begin + rows_per_split < M ? m_begin + rows_per_split : M;
  const int nst = (m_end - m_begin + kTcBK - 1) / kTcBK;  // >= 1 by construction of the grid
  const int nbB = BN / 64;
  // Dense bias gradient = column sums of G: the i-tile-0 CTAs see every row of their G columns, and their four
  // epilogue warps are idle during the main loop, so they add the staged tiles up straight from shared memory.
  const bool do_bias = dbias != nullptr && blockIdx.y == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], do_bias ? 5 : 1); }
    mbar_init(tfull, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t bytes = static_cast<uint32_t>(2 + nbB) * kTcBoxBytes;
      for (int st = 0; st < nst; ++st) {
        const int slot = st % kTcStages;
        if (st >= kTcStages) mbar_wait(&empty[slot], ((st / kTcStages) - 1) & 1);
        mbar_arrive_expect_tx(&full[slot], bytes);
        uint8_t* base = stages + slot * kTcStageBytes;
        const int m0 = m_begin + st * kTcBK;
        tma_load_2d(base, &tmX, &full[slot], i0, m0);
        tma_load_2d(base + kTcBoxBytes, &tmX, &full[slot], i0 + 64, m0);
        for (int b = 0; b < nbB; ++b) tma_load_2d(base + (2 + b) * kTcBoxBytes, &tmG, &full[slot], o0 + b * 64, m0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(128, BN, 1) | (1u << 15) | (1u << 16);  // both operands MN-major
      for (int st = 0; st < nst; ++st) {
        const int slot = st % kTcStages;
        mbar_wait(&full[slot], (st / kTcStages) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(stages + slot * kTcStageBytes), b0 = a0 + 2 * kTcBoxBytes;
#pragma unroll
        for (int kk = 0; kk < kTcBK / 16; ++kk) {
          const uint64_t ad = umma_desc_mn_sw128(a0 + kk * 2048, kTcBoxBytes, 1024);
          const uint64_t bd = umma_desc_mn_sw128(b0 + kk * 2048, kTcBoxBytes, 1024);
          umma_bf16(tmem, ad, bd, idesc, (st | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[slot]);  // the stage may be overwritten once these MMAs have read it
      }
      umma_commit(tfull);
    }
  } else {
    // epilogue: warp q = warp % 4 owns TMEM lanes [32q, 32q+32) = rows i0 + 32q + lane
    const int q = warp & 3;
    float* buf = epi + (warp - 2) * kTcEpiFloats;
    if (do_bias) {
      const int c = ((warp - 2) * 32 + lane) * 2;  // this thread's two columns of the BN-wide tile
      float s0 = 0.f, s1 = 0.f;
      for (int st = 0; st < nst; ++st) {
        const int slot = st % kTcStages;
        mbar_wait(&full[slot], (st / kTcStages) & 1);
        if (c < BN) {
          const uint8_t* box = stages + slot * kTcStageBytes + (2 + (c >> 6)) * kTcBoxBytes;
          const int cc = c & 63;
#pragma unroll 8
          for (int r = 0; r < kTcBK; ++r) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(box + r * 128 + ((((cc >> 3) ^ (r & 7)) << 4) | ((cc & 7) << 1)));
            s0 += bf16_lo(v);
            s1 += bf16_hi(v);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
      if (c < BN) {
        if (o0 + c < Ovalid) atomicAdd(&dbias[o0 + c], s0);
        if (o0 + c + 1 < Ovalid) atomicAdd(&dbias[o0 + c + 1], s1);
      }
    }
    mbar_wait(tfull, 0);
    tc_fence_after();
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_ld_fence(r);
#pragma unroll
      for (int j = 0; j < 32; ++j) buf[lane * 33 + j] = __uint_as_float(r[j]);
      __syncwarp();
      const int o = o0 + c * 32 + lane;
      if (o < Ovalid) {
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
          const int i = i0 + q * 32 + rr;
          if (i < Ivalid && !ISHARA_WGRAD_SKIP(dbg)) atomicAdd(dW + static_cast<size_t>(i) * ldw + o, buf[rr * 33 + lane]);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace

struct WgradTcPlanImpl {
  CUtensorMap tmX, tmG;
  int M, I, O, BN, splits, rps, ti, to;
};

int wgrad_tc_plan_init(WgradTcPlan* p, const bf16* X, int ldx, const bf16* G, int ldg, int64_t M, int I, int O, int num_sms) {
  static_assert(sizeof(WgradTcPlanImpl) <= sizeof(WgradTcPlan), "WgradTcPlan storage too small");
  WgradTcPlanImpl* q = reinterpret_cast<WgradTcPlanImpl*>(p);
  if (I % 8 != 0 || O % 64 != 0 || M <= 0 || M > 0x7fffffff) {
    set_last_error("wgrad_tc: I must be a multiple of 8, O a multiple of 64");
    return 2;
  }
  int rc;
  if ((rc = make_tmap_2d(&q->tmX, X, TM_BF16, static_cast<uint64_t>(M), static_cast<uint64_t>(I), ldx, kTcBK, 64))) return rc;
  if ((rc = make_tmap_2d(&q->tmG, G, TM_BF16, static_cast<uint64_t>(M), static_cast<uint64_t>(O), ldg, kTcBK, 64))) return rc;
  q->M = static_cast<int>(M); q->I = I; q->O = O;
  q->BN = O % 256 == 0 ? 256 : (O % 128 == 0 ? 128 : 64);
  q->ti = (I + 127) / 128;
  q->to = O / q->BN;
  int splits = num_sms / (q->ti * q->to);
  if (splits < 1) splits = 1;
  int64_t rps = (M + splits - 1) / splits;
  rps = (rps + kTcBK - 1) / kTcBK * kTcBK;  // stage-aligned slices: only the last one has an (OOB, zero-filled) tail
  q->rps = static_cast<int>(rps);
  q->splits = static_cast<int>((M + rps - 1) / rps);
  return 0;
}

int wgrad_tc_launch(const WgradTcPlan* p, float* dW, int ldw, float* dbias, int Ivalid, int Ovalid, cudaStream_t s) {
  const WgradTcPlanImpl* q = reinterpret_cast<const WgradTcPlanImpl*>(p);
  const size_t smem = static_cast<size_t>(kTcStages) * kTcStageBytes + 128 + 4 * kTcEpiFloats * sizeof(float) + 1024;
#ifdef ISHARA_TRACE_BUILD
  static const int dbg = getenv("ISHARA_WGRAD_DBG") ? atoi(getenv("ISHARA_WGRAD_DBG")) : 0;  // bit 0: skip the atomics (timing bisect only)
#else
  static const int dbg = 0;
#endif
  static bool attr_done = false;
  if (!attr_done) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_done = true;
  }
  wgrad_tc_kernel<<<dim3(q->to, q->ti, q->splits), kTcThreads, smem, s>>>(q->tmX, q->tmG, dW, ldw, dbias, q->M, Ivalid, Ovalid, q->BN, q->rps, dbg);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace ishara
