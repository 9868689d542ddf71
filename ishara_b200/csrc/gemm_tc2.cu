// ishara_b200 — CTA-pair (cta_group::2) weight-stationary tcgen05 GEMM (sm_100a only).
//
// Why this variant exists (measured on B200, tools/gpu_probe.py + ISHARA_GEMM_DBG bisect): in the
// single-CTA kernel the phases of a tile ADD instead of overlapping — TMA latency under load is ~2.1-2.5k cycles and the
// smem ring only holds one tile of operands, so loads (21 us) + MMA (13 us) + epilogue (15 us) ~= the 49 us measured for
// the QKV GEMM. The fix is bytes in flight, and shared memory is what limits them. Pairing two CTAs with cta_group::2
// splits the weight tile across the pair (each CTA supplies N/2 rows of B), so each CTA can keep its half of the weights
// RESIDENT for the whole kernel (64 KB at K=256, 128 KB at K=512) and spend the rest on a deep A ring (8 stages at
// K=256 = two full tiles of lookahead). Per tile only A crosses L2->SM.
//
//   cluster (2,1,1): CTA rank r owns rows [256*mp + 128*r, +128) of pair-tile mp and N-rows [256*nt + 128*r, +128) of B
//   warp 0 lane 0   TMA producer (both CTAs): own half of B once, then A k-blocks into the ring; completion bytes of BOTH
//                   CTAs' loads are credited to the LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2)
//   warp 1 lane 0   MMA issuer (leader only): tcgen05.mma.cta_group::2 M=256 N=256 K=16, fp32 accumulators in the TMEM
//                   of both CTAs (each CTA: its own 128 rows x 256 columns, double buffered); tcgen05.commit multicast
//                   frees the ring slot / publishes the accumulator in both CTAs at once
//   warp 2          TMEM allocator (cta_group::2, same warp id in both CTAs)
//   warps 4-11      epilogue, identical to gemm_tc.cu (gemm_epilogue.cuh); "accumulator drained" is one arrive per
//                   warp on the leader's barrier (peer CTA: mbarrier.arrive.shared::cluster)
#include <cstdio>
#include <cstdlib>

#include "gemm_epilogue.cuh"

namespace ishara {
namespace {

#ifdef ISHARA_TRACE_BUILD
#define ISHARA_TRACE(it_, ev_) do { if (ep.trace != nullptr && blockIdx.x == 0) ep.trace[(it_) * 8 + (ev_)] = clock64(); } while (0)
#else
#define ISHARA_TRACE(it_, ev_) do { } while (0)
#endif

constexpr int kBN2 = 256;
constexpr int kBHalfKbBytes = (kBN2 / 2) * kBK * 2;  // one k-block of this CTA's half of the weight tile: 16 KB

template <bool ROW, int STG_TILES>  // STG_TILES: 4 = two staging boxes per epilogue warp, 2 = one
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1, const GemmEpi ep, int M,
             int N, int K, int num_m_pairs, int num_n_tiles, int num_stages) {
  constexpr int BN = kBN2;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr uint32_t IDESC = umma_idesc(2 * kBM, BN, /*bf16*/ 1);
  const int STAGES = num_stages;
  const int num_kb = K / kBK;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t resb_bytes = static_cast<uint32_t>(num_kb) * kBHalfKbBytes;
  uint8_t* resb_ptr = smem;                         // [num_kb][128 x 64] resident half of B
  uint8_t* stage_ptr = smem + resb_bytes;           // A ring
  uint8_t* stg_ptr = stage_ptr + STAGES * kAStageBytes;
  float4* xch = reinterpret_cast<float4*>(stg_ptr + STG_TILES * kStgBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_ptr + STG_TILES * kStgBytes + kXchBytes);
  uint64_t* full = bars;                            // leader's copy is the live one
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;         // leader's copy is the live one
  uint64_t* bfull = bars + 2 * STAGES + 4;
  uint64_t* bpeer = bars + 2 * STAGES + 5;          // leader: the peer's half of B has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;             // a multiple of num_n_tiles: every pair owns one n-tile
  const int num_tiles = num_m_pairs * num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO0);
    if (ep.ln1_g != nullptr) tma_prefetch_desc(&tmO1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 16);  // 8 epilogue warps in each of the two CTAs
    mbar_init(&tempty[1], 16);
    mbar_init(bfull, 1);
    mbar_init(bpeer, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();  // barrier inits + TMEM allocation visible in both CTAs before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0 && pair < num_tiles) {
      const int n_tile = pair % num_n_tiles;
      // resident half of B: the leader's half is credited to its bfull, the peer's half to the LEADER's bpeer, so the
      // MMA issuer learns about both without any thread of the peer having to wait
      if (leader) {
        mbar_arrive_expect_tx(bfull, resb_bytes);
        mbar_arrive_expect_tx(bpeer, resb_bytes);
      }
      const uint32_t bbar = leader ? smem_u32(bfull) : mapa_shared(smem_u32(bpeer), 0);
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d_2cta(resb_ptr + kb * kBHalfKbBytes, &tmB, bbar, kb * kBK, n_tile * BN + static_cast<int>(rank) * (BN / 2));
      uint32_t stage = 0, phase = 0;
      int pit = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++pit) {
        const int m_pair = tile / num_n_tiles;
        const int row0 = m_pair * 2 * kBM + static_cast<int>(rank) * kBM;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          if (kb == 0) ISHARA_TRACE(pit, 0);
          if (kb == num_kb - 1) ISHARA_TRACE(pit, 1);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * kAStageBytes);  // bytes of both CTAs' A sub-tiles
          tma_load_2d_2cta(stage_ptr + stage * kAStageBytes, &tmA, mapa_shared(smem_u32(&full[stage]), 0), kb * kBK, row0);
          if (++stage == static_cast<uint32_t>(STAGES)) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && leader && pair < num_tiles) {
      mbar_wait(bfull, 0);
      mbar_wait(bpeer, 0);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        ISHARA_TRACE(it, 2);
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (kb == 0) ISHARA_TRACE(it, 3);
          if (kb == num_kb - 1) ISHARA_TRACE(it, 4);
          const uint32_t sa = smem_base + resb_bytes + stage * kAStageBytes;
          const uint32_t sb = smem_base + kb * kBHalfKbBytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            umma_bf16_2cta(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), IDESC,
                           (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2cta(&empty[stage], 0b11);  // frees this ring slot in both CTAs when these MMAs retire
          if (kb == num_kb - 1) { umma_commit_2cta(&tfull[buf], 0b11); ISHARA_TRACE(it, 5); }
          if (++stage == static_cast<uint32_t>(STAGES)) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): all 8 warps on every tile =====================
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    WarpStore st;
    st.single = STG_TILES == 2;
    st.base = smem_u32(stg_ptr) + static_cast<uint32_t>((warp - 4) * (st.single ? 1 : 2)) * kWarpStgBytes;
    st.iter = 0;
    st.lane = lane;
    st.skip_store = ISHARA_DBG_BIT(ep, 1);
    st.skip_fence = ISHARA_DBG_BIT(ep, 16);

    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const uint32_t buf = it & 1;
      const int m_pair = tile / num_n_tiles, n_tile = tile % num_n_tiles;
      const int row0 = m_pair * 2 * kBM + static_cast<int>(rank) * kBM + q * 32;
      EpiThread th;
      th.row = row0 + lane;
      th.valid = th.row < M;
      th.seq = th.valid ? th.row / ep.rows_per_seq : 0;
      th.t = th.valid ? th.row - th.seq * ep.rows_per_seq : 0;
      th.taddr = tmem_base + buf * BN + (static_cast<uint32_t>(q * 32) << 16);
      ResidRegs rr;
      resid_load(rr, ep, th, epilogue_first_col<BN, ROW, false>(ep, n_tile, h));

      mbar_wait(&tfull[buf], (it >> 1) & 1);
      tc_fence_after();
      if (q == 0 && h == 0 && lane == 0) ISHARA_TRACE(it, 6);
      if (!ISHARA_DBG_BIT(ep, 2)) epilogue_tile<BN, ROW, false>(ep, th, n_tile, N, row0, q, h, lane, st, &tmO0, &tmO1, xch, it & 1, rr);
      // this warp is done with its CTA's half of accumulator buffer `buf`: one arrive on the LEADER's barrier
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[buf], 0);
      if (q == 0 && h == 0 && lane == 0) ISHARA_TRACE(it, 7);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  // no CTA may exit (or free TMEM) while its peer can still multicast into its barriers / read its smem
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

template <bool ROW, int STG_TILES>
int launch2(const GemmPlan& p, int num_sms, int stages, int smem, cudaStream_t stream) {
  auto kern = gemm2_kernel<ROW, STG_TILES>;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  const int mp = (p.M + 2 * kBM - 1) / (2 * kBM);
  const int nt = p.N / kBN2;
  int pairs = mp * nt;
  if (pairs > num_sms / 2) pairs = num_sms / 2;
  pairs = pairs / nt * nt;
  kern<<<2 * pairs, 384, smem, stream>>>(p.tmA, p.tmB, p.tmO0, p.tmO1, p.epi, p.M, p.N, p.K, mp, nt, stages);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace

int gemm2_fixed_smem(int K, int stg_tiles) {
  return (K / kBK) * kBHalfKbBytes + stg_tiles * kStgBytes + kXchBytes + 256 + 1024;
}

// The B tensor map of a pair plan has a 128-row box (each CTA loads its half of the 256-wide weight tile).
bool gemm2_applicable(const GemmPlan& p, int num_sms) {
  if (p.block_n != kBN2 || p.out_f32 || p.N % kBN2 != 0 || p.K % kBK != 0) return false;
  const int mp = (p.M + 2 * kBM - 1) / (2 * kBM);
  if (mp * (p.N / kBN2) < num_sms / 2) return false;  // not enough pair-tiles to fill the machine
  return gemm2_fixed_smem(p.K, 2) + 3 * kAStageBytes <= kMaxSmem;
}

int gemm2_launch(const GemmPlan& p, int num_sms, cudaStream_t stream) {
  // two staging boxes per warp when that still leaves >= 6 ring stages, else one
  static const int force_stg = getenv("ISHARA_GEMM_STG") ? atoi(getenv("ISHARA_GEMM_STG")) : 0;
  int stg = 4;
  if ((kMaxSmem - gemm2_fixed_smem(p.K, 4)) / kAStageBytes < 6) stg = 2;
  if (force_stg == 2 || force_stg == 4) stg = force_stg;
  if ((kMaxSmem - gemm2_fixed_smem(p.K, stg)) / kAStageBytes < 3) stg = 2;
  const int fixed = gemm2_fixed_smem(p.K, stg);
  int stages = (kMaxSmem - fixed) / kAStageBytes;
  if (stages > 10) stages = 10;
  static const int cap = getenv("ISHARA_GEMM_STAGES") ? atoi(getenv("ISHARA_GEMM_STAGES")) : 0;
  if (cap >= 2 && stages > cap) stages = cap;
  const int smem = fixed + stages * kAStageBytes;
  if (stg == 4)
    return p.row_mode ? launch2<true, 4>(p, num_sms, stages, smem, stream) : launch2<false, 4>(p, num_sms, stages, smem, stream);
  return p.row_mode ? launch2<true, 2>(p, num_sms, stages, smem, stream) : launch2<false, 2>(p, num_sms, stages, smem, stream);
}

}  // namespace ishara
