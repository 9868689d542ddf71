// ishara_b200 — fused front half of Conv1DBlock: G = ECA(BN(CausalDWConv1D(swish(x @ We + be))))  (sm_100a only).
//
// Reference: Conv1DBlock (nb:conv-hybrid-model c5:41-89) = Dense(2D, swish) c5:61-65 -> CausalDWConv1D(k) c5:17-39,68-71
// -> BatchNorm c5:73 -> ECA c5:1-15,75 -> Dense(D) -> +skip. As separate launches the 2D-wide tensor crosses HBM three
// times (written by the expand GEMM, read + written by the depthwise kernel). Here it exists only in shared memory:
//
//   cluster = the T/128 CTAs of ONE sequence; CTA r owns frames [128 r, 128 r + 128)
//   1. GEMM: acc[128 x 512] fp32 in TMEM (all 512 columns, two N = 256 halves) = x_tile[128 x 256] @ We^T, TMA ring
//   2. epilogue (8 warps, lane = row): + be, swish, bf16 -> H[128][512] in shared memory (row pitch padded by 16 B),
//      aliased over the dead TMA ring
//   3. ECA needs the mean over ALL T frames of the BatchNorm'd conv output. The conv is linear in time, so that mean is
//      wsum * colsum(H) minus a tail correction from the last k-1 frames of the sequence: every CTA publishes its
//      column sums, the last CTA the correction, one cluster barrier, then every CTA derives all 512 channel scales
//      (5-tap conv across channels + sigmoid) from DSMEM reads
//   4. stencil: thread = channel pair, register sliding window down the 128 frames; the k-1 halo frames come straight
//      from the previous CTA's shared memory (DSMEM); BatchNorm is folded into the taps/bias; result * scale -> bf16 ->
//      global, 128 B per warp per frame
// Only D == 256 (2D = 512 TMEM columns), T % 128 == 0, T <= 1024; other shapes use the three-kernel path.
#include <cooperative_groups.h>

#include "gemm_epilogue.cuh"

namespace cg = cooperative_groups;

namespace ishara {
namespace {

constexpr int kCD = 256;                 // model dim (K of the GEMM)
constexpr int kCE = 512;                 // expanded channels = TMEM columns
constexpr int kCStage = kAStageBytes + kCD * kBK * 2;   // A 16 KB + B 32 KB
constexpr int kCStages = 2;
constexpr int kHPitch = kCE * 2 + 16;    // bytes per H row (16 B pad: conflict-free 16-byte row-wise stores)
constexpr int kHBytes = kBM * kHPitch;   // 133,120 B (aliases the ring)
constexpr int kHaloRows = 16;            // room for the k-1 <= 14 frames in front of the tile
constexpr int kHaloBytes = kHaloRows * kHPitch;

__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.f / (1.f + __expf(-x)); }

template <int K>
__global__ void __launch_bounds__(384, 1)
conv1d_front_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const float* __restrict__ bias_e, const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                    const float* __restrict__ eca_w, bf16* __restrict__ out, int T) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* ring = smem;                       // phase 1 (dead before anything below is written)
  uint8_t* Hs = smem + kHaloBytes;            // phase 2+: [128][kHPitch]; rows -16..-1 in front of it hold the halo frames
  float* part = reinterpret_cast<float*>(Hs + kHBytes);     // [512] column sums of this CTA's H tile
  float* corr = part + kCE;                                 // [512] tail correction (last CTA only)
  float* mean = corr + kCE;                                 // [512] channel means of the BatchNorm'd conv output
  uint64_t* bars = reinterpret_cast<uint64_t*>(mean + kCE);
  uint64_t* full = bars;                // [2]
  uint64_t* empty = bars + 2;           // [2]
  uint64_t* acc_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = blockIdx.x;                    // tile index inside the sequence == rank in the cluster
  const int nrank = gridDim.x;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_g0 = b * T + rank * kBM;          // first global row of this tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kCStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t IDESC = umma_idesc(kBM, 256, 1);
  constexpr int NKB = kCD / kBK;   // 4
  constexpr int NLOAD = 2 * NKB;   // two N halves

  // ===================== phase 1: GEMM into TMEM =====================
  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int i = 0; i < NLOAD; ++i) {
        const int nh = i / NKB, kb = i % NKB;
        mbar_wait(&empty[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full[stage], kCStage);
        uint8_t* sa = ring + stage * kCStage;
        tma_load_2d(sa, &tmA, &full[stage], kb * kBK, row_g0);
        tma_load_2d(sa + kAStageBytes, &tmB, &full[stage], kb * kBK, nh * 256);
        if (++stage == kCStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int i = 0; i < NLOAD; ++i) {
        const int nh = i / NKB, kb = i % NKB;
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * kCStage;
        const uint32_t sb = sa + kAStageBytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          umma_bf16(tmem_base + nh * 256, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), IDESC,
                    (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty[stage]);
        if (++stage == kCStages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(acc_full);
    }
  }

  // ===================== phase 2: TMEM -> +be -> swish -> bf16 H tile in shared memory =====================
  if (warp >= 4) {
    const int q = warp & 3, h = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    mbar_wait(acc_full, 0);   // every MMA has retired, so every ring slot has been consumed: the ring may be overwritten
    tc_fence_after();
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t rowbase = smem_base + kHaloBytes + static_cast<uint32_t>(r) * kHPitch;
#pragma unroll 1
    for (int c = h * 8; c < h * 8 + 8; ++c) {
      uint32_t raw[32];
      float v[32];
      tmem_ld32(tmem_base + lane_addr + c * 32, raw);
      tmem_ld_wait();
      to_float(v, raw);
      const float4* b4 = reinterpret_cast<const float4*>(bias_e + c * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = __ldg(b4 + j);
        fadd2(v[4 * j + 0], v[4 * j + 1], v[4 * j + 0], v[4 * j + 1], bb.x, bb.y);
        fadd2(v[4 * j + 2], v[4 * j + 3], v[4 * j + 2], v[4 * j + 3], bb.z, bb.w);
      }
      epi_swish(v);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_shared_v4(rowbase + c * 64 + j * 16, pack_bf16x2(v[8 * j + 0], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                     pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
    }
  }
  tc_fence_before();
  __syncthreads();

  // ===================== phase 3: column sums, tail correction, ECA scales =====================
  const int p = threadIdx.x - 128;  // channel pair owned by this thread in phases 3-4 (warps 4-11)
  float2 wt[K];
  float2 bdw = make_float2(0.f, 0.f);
  if (p >= 0) {
#pragma unroll
    for (int j = 0; j < K; ++j) wt[j] = __ldg(reinterpret_cast<const float2*>(dw_w + static_cast<size_t>(j) * kCE) + p);
    bdw = __ldg(reinterpret_cast<const float2*>(dw_b) + p);
    const uint32_t* col = reinterpret_cast<const uint32_t*>(Hs) + p;
    float2 s = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int t = 0; t < kBM; ++t) {
      const uint32_t u = col[t * (kHPitch / 4)];
      fadd2(s.x, s.y, s.x, s.y, bf16_lo(u), bf16_hi(u));
    }
    part[2 * p] = s.x;
    part[2 * p + 1] = s.y;
    if (rank == nrank - 1) {
      // sum_t y[t] = T*b + sum_j w_j * (S - tail_{K-1-j}),  tail_m = sum of the LAST m frames of the sequence
      float2 tail = make_float2(0.f, 0.f), acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int m = 1; m <= K - 1; ++m) {  // tap j = K-1-m
        const uint32_t u = col[(kBM - m) * (kHPitch / 4)];
        tail.x += bf16_lo(u); tail.y += bf16_hi(u);
        acc.x = fmaf(wt[K - 1 - m].x, tail.x, acc.x);
        acc.y = fmaf(wt[K - 1 - m].y, tail.y, acc.y);
      }
      corr[2 * p] = acc.x;
      corr[2 * p + 1] = acc.y;
    }
  }
  cluster.sync();
  if (p >= 0) {
    float2 S = make_float2(0.f, 0.f);
    for (int rr = 0; rr < nrank; ++rr) {
      const float2 v = *reinterpret_cast<const float2*>(cluster.map_shared_rank(part, rr) + 2 * p);
      S.x += v.x; S.y += v.y;
    }
    const float2 cr = *reinterpret_cast<const float2*>(cluster.map_shared_rank(corr, nrank - 1) + 2 * p);
    float2 wsum = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < K; ++j) { wsum.x += wt[j].x; wsum.y += wt[j].y; }
    const float invT = 1.f / static_cast<float>(T);
    mean[2 * p] = fmaf(fmaf(wsum.x, S.x, -cr.x), invT, bdw.x);
    mean[2 * p + 1] = fmaf(fmaf(wsum.y, S.y, -cr.y), invT, bdw.y);
  }
  __syncthreads();
  float2 scale = make_float2(1.f, 1.f);
  if (p >= 0) {
    float z0 = 0.f, z1 = 0.f;
#pragma unroll
    for (int d = -2; d <= 2; ++d) {
      const float e = __ldg(eca_w + d + 2);
      const int c0 = 2 * p + d, c1 = 2 * p + 1 + d;
      if (c0 >= 0 && c0 < kCE) z0 = fmaf(e, mean[c0], z0);
      if (c1 >= 0 && c1 < kCE) z1 = fmaf(e, mean[c1], z1);
    }
    scale = make_float2(sigmoidf_exact(z0), sigmoidf_exact(z1));
  }

  // ===================== phase 4: causal depthwise stencil down the frames, * scale, -> global =====================
  if (p >= 0) {
    uint32_t* col = reinterpret_cast<uint32_t*>(Hs) + p;
    constexpr int PW = kHPitch / 4;  // row pitch in 32-bit words
    // halo: the previous CTA's last K-1 frames (zeros in front of the sequence) -> rows -(K-1)..-1 of this thread's column
#pragma unroll
    for (int j = 1; j <= K - 1; ++j) {
      uint32_t u = 0u;
      if (rank > 0) u = (reinterpret_cast<const uint32_t*>(cluster.map_shared_rank(Hs, rank - 1)) + p)[(kBM - j) * PW];
      (col - j * PW)[0] = u;
    }
    constexpr int TB = K >= 9 ? 4 : 8;  // frames per register block
    const uint32_t* trow = col - (K - 1) * PW;  // window start of output frame 0
    uint32_t* drow = reinterpret_cast<uint32_t*>(out + static_cast<size_t>(row_g0) * kCE) + p;
#pragma unroll 1
    for (int t0 = 0; t0 < kBM; t0 += TB) {
      float2 x[TB + K - 1];
#pragma unroll
      for (int i = 0; i < TB + K - 1; ++i) {
        const uint32_t u = trow[i * PW];
        x[i] = make_float2(bf16_lo(u), bf16_hi(u));
      }
#pragma unroll
      for (int i = 0; i < TB; ++i) {
        float2 a = bdw;
#pragma unroll
        for (int j = 0; j < K; ++j) ffma2(a.x, a.y, wt[j].x, wt[j].y, x[i + j].x, x[i + j].y, a.x, a.y);
        fmul2(a.x, a.y, a.x, a.y, scale.x, scale.y);
        *drow = pack_bf16x2(a.x, a.y);
        drow += kCE / 2;
      }
      trow += TB * PW;
    }
  }
  // neighbours may still be reading this CTA's H tile / partial sums
  cluster.sync();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int K>
int launch_k(const Conv1dFrontPlan& p, cudaStream_t stream) {
  auto kern = conv1d_front_kernel<K>;
  const int smem = kHaloBytes + kHBytes + 3 * kCE * 4 + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.T / kBM, p.B, 1);
  cfg.blockDim = dim3(384, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.T / kBM;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  ISHARA_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p.tmA, p.tmB, p.bias_e, p.dw_w, p.dw_b, p.eca_w, p.out, p.T));
  note_launch();
  return 0;
}

}  // namespace

bool conv1d_front_applicable(int D, int T, int k) {
  return D == kCD && T % kBM == 0 && T / kBM >= 1 && T / kBM <= 8 && (k == 3 || k == 5 || k == 7 || k == 9 || k == 11 || k == 15);
}

int conv1d_front_plan_init(Conv1dFrontPlan* p, const bf16* x, const bf16* wet) {
  static_assert(kCStages * kCStage <= kHaloBytes + kHBytes, "the TMA ring must fit under the H tile it is aliased with");
  int rc;
  if ((rc = make_tmap_2d(&p->tmA, x, TM_BF16, static_cast<uint64_t>(p->B) * p->T, kCD, kCD, kBM, kBK))) return rc;
  return make_tmap_2d(&p->tmB, wet, TM_BF16, kCE, kCD, kCD, 256, kBK);
}

int conv1d_front_launch(const Conv1dFrontPlan& p, cudaStream_t stream) {
  switch (p.k) {
    case 3: return launch_k<3>(p, stream);
    case 5: return launch_k<5>(p, stream);
    case 7: return launch_k<7>(p, stream);
    case 9: return launch_k<9>(p, stream);
    case 11: return launch_k<11>(p, stream);
    case 15: return launch_k<15>(p, stream);
  }
  set_last_error("conv1d_front: unsupported kernel size");
  return 2;
}

}  // namespace ishara
