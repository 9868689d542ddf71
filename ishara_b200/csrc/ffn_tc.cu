// ishara_b200 — fused feed-forward module: S = [LN](S + swish(XN @ W1 + b1) @ W2 + b2), XN' = LN'(S)  (sm_100a only).
//
// Reference: the FFN sub-modules of SqueezeformerBlock / ConformerBlock (nb:conv-hybrid-model c5:162-166,187-190,
// 202-205,237-247,323-326,338-341): Dense(D -> E, swish) -> Dropout -> Dense(E -> D), residual add. As two separate
// GEMM launches the E-wide intermediate makes a 200 MB round trip through HBM per module and pays two epilogue passes;
// here it never leaves the SM:
//
//   per 128-row tile, E is walked in quarters of 128 columns:
//     G1(q): acc1[q&1] (TMEM, 128 cols, double buffered) = XN_tile[128 x 256] @ W1[q]^T         tcgen05.mma N=128
//     epi1(q): +b1, swish, bf16 -> H[128 x 128] in shared memory (K-major, 128B swizzle)         8 epilogue warps
//     G2(q): acc2 (TMEM, 256 cols) += H @ W2[:, q]^T                                             tcgen05.mma N=256
//   epi2: +b2, +residual, [LayerNorm], TMA store of the stream and of its normalised copy        gemm_epilogue.cuh
//
//   MMA issue order G1(0) G1(1) G2(0) G1(2) G2(1) G1(3) G2(2) G2(3): the tensor pipe always has the next GEMM queued
//   while the epilogue warps turn acc1 into H. Weights stream from L2 through a ring of 32 KB slots in exactly that
//   order (W1 quarter = two slots of two [128 x 64] k-blocks, W2 K-slice = two slots of one [256 x 64] k-block).
//   smem: XN tile 64 KB + H 32 KB (doubles as the epilogue's TMA-store staging: every warp owns the same 4 KB region in
//   both roles) + 3 weight slots 96 KB + row-statistics exchange.
//   warp 0 TMA producer | warp 1 MMA issuer | warp 2 TMEM alloc | warps 4-11 epilogue.
// Only D == 256 (full-row LayerNorm epilogue) and E % 128 == 0; other shapes use the two-GEMM path.
#include "gemm_epilogue.cuh"

namespace ishara {
namespace {

constexpr int kFD = 256;               // model dim: K of GEMM1, N of GEMM2
constexpr int kFQ = 128;               // quarter width of the hidden dimension
constexpr int kFSlot = 32 * 1024;      // weight ring slot
constexpr int kFStages = 3;
constexpr int kFA1Bytes = (kFD / kBK) * kAStageBytes;  // 64 KB
constexpr int kFHBytes = 2 * kAStageBytes;             // 32 KB: H quarter = two k-blocks of [128 x 64]

struct FfnBars {
  uint64_t a1_full, a1_empty;
  uint64_t w_full[kFStages], w_empty[kFStages];
  uint64_t acc1_full[2], acc1_empty[2];
  uint64_t h_full, h_empty;
  uint64_t acc2_full, acc2_empty;
  uint32_t tmem_slot;
};

// Timeline tracing (profiling builds only: make TRACE=1): clock64 of CTA 60 for its first three tiles, 32 events per tile
#ifdef ISHARA_TRACE_BUILD
__device__ long long* g_ffn_trace = nullptr;
#define FFN_TRACE(it_, ev_) do { if (g_ffn_trace != nullptr && blockIdx.x == 60 && (it_) < 3) g_ffn_trace[(it_) * 32 + (ev_)] = clock64(); } while (0)
#else
#define FFN_TRACE(it_, ev_) do { } while (0)
#endif

// EW = epilogue warps: 8 (two per TMEM lane quarter) or 16 (four per quarter). With 8 the epilogue warps were the
// critical resource - 4 x swish quarter + the full-row residual / LayerNorm pass took ~2/3 of the 31k-cycle tile period
// against 8k cycles of MMA; 16 warps halve every pass and give each scheduler four warps to hide tcgen05.ld latency.
//   EW = 16: warp (q4, c) owns rows [32 q4, +32) and, of a 128-column hidden quarter, the 32 columns [32c, +32); in the
//   final epilogue the 64-column box c of the 256-wide output, staged through a 2 KB half box (the H bytes).
template <int EW>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
ffn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO0,
              const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmO0h,
              const __grid_constant__ CUtensorMap tmO1h, const GemmEpi ep, const float* __restrict__ bias1, int M, int E,
              int num_m_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* a1_ptr = smem;                                  // [4][128 x 64] XN tile
  uint8_t* h_ptr = a1_ptr + kFA1Bytes;                     // [2][128 x 64] H quarter / epilogue staging
  uint8_t* w_ptr = h_ptr + kFHBytes;                       // [3][32 KB] weight slots
  float4* xch = reinterpret_cast<float4*>(w_ptr + kFStages * kFSlot);
  FfnBars* bars = reinterpret_cast<FfnBars*>(reinterpret_cast<uint8_t*>(xch) + kXchBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = E / kFQ;  // quarters of the hidden dimension (4 for E = 512)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmO0);
    if (ep.ln1_g != nullptr) tma_prefetch_desc(&tmO1);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bars->a1_full, 1);
    mbar_init(&bars->a1_empty, 1);
    for (int s = 0; s < kFStages; ++s) {
      mbar_init(&bars->w_full[s], 1);
      mbar_init(&bars->w_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->acc1_full[i], 1);
      mbar_init(&bars->acc1_empty[i], EW);
    }
    mbar_init(&bars->h_full, EW);
    mbar_init(&bars->h_empty, 1);
    mbar_init(&bars->acc2_full, 1);
    mbar_init(&bars->acc2_empty, EW);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  const uint32_t tmem_acc2 = tmem_base + 256;
  constexpr uint32_t IDESC1 = umma_idesc(kBM, kFQ, 1);
  constexpr uint32_t IDESC2 = umma_idesc(kBM, kFD, 1);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t ws = 0, wphase = 0;
      auto slot_acquire = [&]() -> uint8_t* {
        mbar_wait(&bars->w_empty[ws], wphase ^ 1u);
        mbar_arrive_expect_tx(&bars->w_full[ws], kFSlot);
        return w_ptr + ws * kFSlot;
      };
      auto slot_advance = [&]() { if (++ws == kFStages) { ws = 0; wphase ^= 1u; } };
      auto load_g1 = [&](int q) {  // W1 rows [128q, +128), all K = 256: two slots of two k-blocks
        for (int j = 0; j < 2; ++j) {
          uint8_t* dst = slot_acquire();
          tma_load_2d(dst, &tmW1, &bars->w_full[ws], (2 * j) * kBK, q * kFQ);
          tma_load_2d(dst + kAStageBytes, &tmW1, &bars->w_full[ws], (2 * j + 1) * kBK, q * kFQ);
          slot_advance();
        }
      };
      auto load_g2 = [&](int q) {  // W2 all 256 rows, K slice [128q, +128): two slots of one k-block
        for (int j = 0; j < 2; ++j) {
          uint8_t* dst = slot_acquire();
          tma_load_2d(dst, &tmW2, &bars->w_full[ws], q * kFQ + j * kBK, 0);
          slot_advance();
        }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++it) {
        mbar_wait(&bars->a1_empty, (it & 1) ^ 1u);
        FFN_TRACE(it, 0);
        mbar_arrive_expect_tx(&bars->a1_full, kFA1Bytes);
        for (int kb = 0; kb < kFD / kBK; ++kb) tma_load_2d(a1_ptr + kb * kAStageBytes, &tmA, &bars->a1_full, kb * kBK, tile * kBM);
        load_g1(0);
        if (nq > 1) load_g1(1);
        for (int q = 0; q < nq; ++q) {
          load_g2(q);
          if (q + 2 < nq) load_g1(q + 2);
        }
        FFN_TRACE(it, 1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t ws = 0, wphase = 0;
      uint32_t n_g1 = 0, n_g2 = 0;  // running counts of issued GEMM1 quarters / GEMM2 quarters
      const uint32_t a1_addr = smem_base, h_addr = a1_addr + kFA1Bytes, w_addr = h_addr + kFHBytes;
      auto issue_g1 = [&](int q) {
        const uint32_t buf = n_g1 & 1;
        mbar_wait(&bars->acc1_empty[buf], ((n_g1 >> 1) & 1) ^ 1u);  // epilogue drained this accumulator buffer
        tc_fence_after();
        for (int j = 0; j < 2; ++j) {
          mbar_wait(&bars->w_full[ws], wphase);
          tc_fence_after();
          const uint32_t sb = w_addr + ws * kFSlot;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16(tmem_base + buf * kFQ, umma_desc_sw128(a1_addr + (2 * j + kk) * kAStageBytes + k * 32),
                        umma_desc_sw128(sb + kk * kAStageBytes + k * 32), IDESC1, (j | kk | k) != 0 ? 1u : 0u);
          umma_commit(&bars->w_empty[ws]);
          if (++ws == kFStages) { ws = 0; wphase ^= 1u; }
        }
        umma_commit(&bars->acc1_full[buf]);
        ++n_g1;
        (void)q;
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++it) {
        mbar_wait(&bars->a1_full, it & 1);
        tc_fence_after();
        FFN_TRACE(it, 2);
        issue_g1(0);
        FFN_TRACE(it, 3);
        if (nq > 1) issue_g1(1);
        FFN_TRACE(it, 4);
        for (int q = 0; q < nq; ++q) {
          // G2(q): needs H(q) written by the epilogue warps; the first one of a tile also needs acc2 drained
          if (q == 0) mbar_wait(&bars->acc2_empty, (it & 1) ^ 1u);
          if (q == 0) FFN_TRACE(it, 5);
          mbar_wait(&bars->h_full, n_g2 & 1);
          tc_fence_after();
          if (q < 4) FFN_TRACE(it, 6 + q);   // H(q) seen
          for (int j = 0; j < 2; ++j) {
            mbar_wait(&bars->w_full[ws], wphase);
            tc_fence_after();
            const uint32_t sb = w_addr + ws * kFSlot;
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16(tmem_acc2, umma_desc_sw128(h_addr + j * kAStageBytes + k * 32), umma_desc_sw128(sb + k * 32), IDESC2,
                        (q | j | k) != 0 ? 1u : 0u);
            umma_commit(&bars->w_empty[ws]);
            if (++ws == kFStages) { ws = 0; wphase ^= 1u; }
          }
          umma_commit(&bars->h_empty);  // H may be overwritten once these MMAs retire
          ++n_g2;
          if (q < 4) FFN_TRACE(it, 10 + q);  // G2(q) issued
          if (q == nq - 1) umma_commit(&bars->acc2_full);
          if (q + 2 < nq) issue_g1(q + 2);
          if (q + 2 == nq - 1 || (nq <= 2 && q == 0)) umma_commit(&bars->a1_empty);  // last GEMM1 of the tile issued
        }
      }
    }
  } else if (warp >= 4 && EW == 16) {
    // ===================== 16 epilogue warps =====================
    const int q4 = warp & 3;         // TMEM lane quarter
    const int c = (warp - 4) >> 2;   // 32-column slice of a hidden quarter / 64-column box of the output tile
    float2* xch2 = reinterpret_cast<float2*>(xch);
    float* cvec = reinterpret_cast<float*>(xch) + 2048;   // [5][256]: b2 | ln0 g | ln0 b | ln1 g | ln1 b
    float* b1s = cvec + 5 * 256;                          // [E] first-layer bias (E <= 768, ffn_applicable)
    {
      const int t16 = static_cast<int>(threadIdx.x) - 128;
      for (int i = t16; i < 5 * 256; i += 512) {
        const int which = i >> 8, col = i & 255;
        const float* src = which == 0 ? ep.bias : which == 1 ? ep.ln0_g : which == 2 ? ep.ln0_b : which == 3 ? ep.ln1_g : ep.ln1_b;
        cvec[i] = src != nullptr ? __ldg(src + col) : 0.f;
      }
      for (int i = t16; i < E; i += 512) b1s[i] = __ldg(bias1 + i);
      named_bar_sync(5, 512);
    }
    WarpStore st;
    st.single = true;
    st.base = smem_u32(h_ptr) + static_cast<uint32_t>(warp - 4) * (kWarpStgBytes / 2);  // 2 KB half box inside the H bytes
    st.iter = 0;
    st.lane = lane;
    Row16State rst;
    // this warp's slice of an H quarter: k-block c / 2 of H, rows [32 q4, +32), 64-byte half c % 2 of every 128-byte row
    const uint32_t hbox = smem_u32(h_ptr) + static_cast<uint32_t>((c >> 1) * 4 + q4) * kWarpStgBytes;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    uint32_t n_q = 0;  // running count of processed quarters
    int it = 0;
    for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++it) {
      const int row0 = tile * kBM + q4 * 32;
      EpiThread th;
      th.row = row0 + lane;
      th.valid = th.row < M;
      th.seq = th.valid ? th.row / ep.rows_per_seq : 0;
      th.t = th.valid ? th.row - th.seq * ep.rows_per_seq : 0;
      th.taddr = tmem_acc2 + lane_addr;
      for (int q = 0; q < nq; ++q, ++n_q) {
        const uint32_t buf = n_q & 1;
        mbar_wait(&bars->acc1_full[buf], (n_q >> 1) & 1);
        tc_fence_after();
        if (warp == 4 && lane == 0 && q < 4) FFN_TRACE(it, 14 + q);  // acc1(q) seen
        if (n_q > 0) mbar_wait(&bars->h_empty, (n_q - 1) & 1);
        if (warp == 4 && lane == 0 && q < 4) FFN_TRACE(it, 18 + q);  // H free
        if (q == 0 && it > 0) {
          // the previous tile's output stores staged through the H bytes - any warp's half box may overlap the bytes this
          // warp is about to write, so every warp's stores must have been read before anyone continues
          if (lane == 0) tma_store_wait_read<0>();
          named_bar_sync(6, 512);
        }
        {
          uint32_t raw[32];
          float v[32];
          const int tc = c * 32;
          tmem_ld32(tmem_base + buf * kFQ + lane_addr + tc, raw);
          tmem_ld_wait();
          to_float(v, raw);
          const float* bq = b1s + q * kFQ + tc;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = *reinterpret_cast<const float4*>(bq + 4 * j);
            fadd2(v[4 * j + 0], v[4 * j + 1], v[4 * j + 0], v[4 * j + 1], bb.x, bb.y);
            fadd2(v[4 * j + 2], v[4 * j + 3], v[4 * j + 2], v[4 * j + 3], bb.z, bb.w);
          }
          epi_swish(v);
          stage_write<false>(hbox, lane, c & 1, v);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->acc1_empty[buf]);
          mbar_arrive(&bars->h_full);
        }
        if (warp == 4 && lane == 0 && q < 4) FFN_TRACE(it, 22 + q);  // epi1(q) done
      }
      // ---- final epilogue of the tile: acc2 -> +b2 -> +residual -> [LN] -> S (and LN'(S)) ----
      uint4 rq[8];
      row16_resid_ldg(ep, th, c, rq);       // residual row segment: in flight while the last GEMM2 finishes
      mbar_wait(&bars->acc2_full, it & 1);  // every GEMM2 of the tile has retired: the H bytes are free for staging
      tc_fence_after();
      if (warp == 4 && lane == 0) FFN_TRACE(it, 26);
      epilogue_row16<true>(ep, th, row0, q4, c, lane, st, &tmO0h, &tmO1h, xch2, cvec, rst, nullptr, 0u, rq);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc2_empty);
      if (warp == 4 && lane == 0) FFN_TRACE(it, 27);
    }
    if (lane == 0) tma_store_wait_all<0>();
  } else if (warp >= 4) {
    // ===================== 8 epilogue warps =====================
    const int q4 = warp & 3;         // TMEM lane quarter
    const int h = (warp - 4) >> 2;   // which 64-column box of a 128-column quarter / alternate boxes of the final tile
    WarpStore st;
    st.single = true;
    st.base = smem_u32(h_ptr) + static_cast<uint32_t>(h * 4 + q4) * kWarpStgBytes;  // = H k-block h, rows [32 q4, +32)
    st.iter = 0;
    st.lane = lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    uint32_t n_q = 0;  // running count of processed quarters
    int it = 0;
    for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++it) {
      const int row0 = tile * kBM + q4 * 32;
      EpiThread th;
      th.row = row0 + lane;
      th.valid = th.row < M;
      th.seq = th.valid ? th.row / ep.rows_per_seq : 0;
      th.t = th.valid ? th.row - th.seq * ep.rows_per_seq : 0;
      th.taddr = tmem_acc2 + lane_addr;
      for (int q = 0; q < nq; ++q, ++n_q) {
        const uint32_t buf = n_q & 1;
        mbar_wait(&bars->acc1_full[buf], (n_q >> 1) & 1);
        tc_fence_after();
        // H region free? the previous quarter's GEMM2 has retired; at a tile boundary the previous tile's TMA stores
        // (which used the same 4 KB region as staging) must have finished reading it as well
        if (n_q > 0) mbar_wait(&bars->h_empty, (n_q - 1) & 1);
        if (q == 0) {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t raw[32];
          float v[32];
          const int tc = h * 64 + sub * 32;
          tmem_ld32(tmem_base + buf * kFQ + lane_addr + tc, raw);
          tmem_ld_wait();
          to_float(v, raw);
          const float4* b4 = reinterpret_cast<const float4*>(bias1 + q * kFQ + tc);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b4 + j);
            fadd2(v[4 * j + 0], v[4 * j + 1], v[4 * j + 0], v[4 * j + 1], bb.x, bb.y);
            fadd2(v[4 * j + 2], v[4 * j + 3], v[4 * j + 2], v[4 * j + 3], bb.z, bb.w);
          }
          epi_swish(v);
          stage_write<false>(st.base, lane, sub, v);
        }
        // accumulator buffer drained, H quarter written (generic proxy -> async proxy fence before the MMA reads it)
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->acc1_empty[buf]);
          mbar_arrive(&bars->h_full);
        }
      }
      // ---- final epilogue of the tile: acc2 -> +b2 -> +residual -> [LN] -> S (and LN'(S)) ----
      ResidRegs rr;
      resid_load(rr, ep, th, epilogue_first_col<kFD, true, false>(ep, 0, h));
      mbar_wait(&bars->acc2_full, it & 1);
      tc_fence_after();
      // staging == this warp's H region: every GEMM2 of the tile has retired (acc2_full), so it is free
      st.iter = 0;
      epilogue_tile<kFD, true, false>(ep, th, 0, kFD, row0, q4, h, lane, st, &tmO0, &tmO1, xch, it & 1, rr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc2_empty);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------------
// v2: hidden dimension in steps of 64 columns with the H tile DOUBLE-buffered (2 x 16 KB in the same 32 KB).
// The timeline of the quarter kernel (FFN_TRACE, profiles/r02 notes in DESIGN.md section 6) showed one serial chain per
// quarter: epi1(q) 1.6k cycles (MUFU-bound: 16k tanh at 16/clk) -> G2(q) waits for its weights and runs (1.5k) -> only
// then is the single H buffer free for epi1(q+1); 4 x 3.1k = 12.5k cycles per tile against 8.2k of MMA, plus the final
// epilogue. With two H buffers epi1(e+1) runs under G2(e):
//   G1(e): acc1[e&1] (TMEM, 64 columns) = XN_tile[128 x 256] @ W1[64e..+64]^T      tcgen05.mma N=64, one 32 KB weight slot
//   epi1(e): +b1, swish, bf16 -> H[e&1] [128 x 64] (one K-major SW128 k-block)      16 warps: warp (q4, c) = 32 rows x 16 columns
//   G2(e): acc2 (256 columns) += H[e&1] @ W2[:, 64e..+64]^T                          tcgen05.mma N=256, one 32 KB weight slot
//   issue order G1(0) G1(1) G2(0) G1(2) G2(1) ... ; final epilogue as in ffn_tc_kernel<16> (half staging boxes in the H bytes),
//   with the residual row segments requested before the wait for the last GEMM.
struct Ffn2Bars {
  uint64_t a1_full, a1_empty;
  uint64_t w_full[kFStages], w_empty[kFStages];
  uint64_t acc1_full[2], acc1_empty[2];
  uint64_t h_full[2], h_empty[2];
  uint64_t acc2_full, acc2_empty;
  uint32_t tmem_slot;
};
constexpr int kFE = 64;  // hidden columns per step

__global__ void __launch_bounds__(640, 1)
ffn_tc_v2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1e,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO0h,
                 const __grid_constant__ CUtensorMap tmO1h, const GemmEpi ep, const float* __restrict__ bias1, int M, int E,
                 int num_m_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* a1_ptr = smem;                                  // [4][128 x 64] XN tile
  uint8_t* h_ptr = a1_ptr + kFA1Bytes;                     // [2][128 x 64] H buffers / epilogue staging
  uint8_t* w_ptr = h_ptr + kFHBytes;                       // [3][32 KB] weight slots
  float4* xch = reinterpret_cast<float4*>(w_ptr + kFStages * kFSlot);
  Ffn2Bars* bars = reinterpret_cast<Ffn2Bars*>(reinterpret_cast<uint8_t*>(xch) + kXchBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ne = E / kFE;  // steps over the hidden dimension (8 for E = 512)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1e);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmO0h);
    if (ep.ln1_g != nullptr) tma_prefetch_desc(&tmO1h);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bars->a1_full, 1);
    mbar_init(&bars->a1_empty, 1);
    for (int s = 0; s < kFStages; ++s) {
      mbar_init(&bars->w_full[s], 1);
      mbar_init(&bars->w_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->acc1_full[i], 1);
      mbar_init(&bars->acc1_empty[i], 16);
      mbar_init(&bars->h_full[i], 16);
      mbar_init(&bars->h_empty[i], 1);
    }
    mbar_init(&bars->acc2_full, 1);
    mbar_init(&bars->acc2_empty, 16);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  const uint32_t tmem_acc2 = tmem_base + 256;
  constexpr uint32_t IDESC1 = umma_idesc(kBM, kFE, 1);
  constexpr uint32_t IDESC2 = umma_idesc(kBM, kFD, 1);
  constexpr int kW1KbBytes = kFE * kBK * 2;  // one [64 x 64] k-block of W1: 8 KB

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t ws = 0, wphase = 0;
      auto slot_acquire = [&]() -> uint8_t* {
        mbar_wait(&bars->w_empty[ws], wphase ^ 1u);
        mbar_arrive_expect_tx(&bars->w_full[ws], kFSlot);
        return w_ptr + ws * kFSlot;
      };
      auto slot_advance = [&]() { if (++ws == kFStages) { ws = 0; wphase ^= 1u; } };
      auto load_g1 = [&](int e) {  // W1 rows [64e, +64), all K = 256: four [64 x 64] k-blocks in one slot
        uint8_t* dst = slot_acquire();
        for (int kb = 0; kb < kFD / kBK; ++kb) tma_load_2d(dst + kb * kW1KbBytes, &tmW1e, &bars->w_full[ws], kb * kBK, e * kFE);
        slot_advance();
      };
      auto load_g2 = [&](int e) {  // W2 all 256 rows, K slice [64e, +64): one [256 x 64] k-block
        uint8_t* dst = slot_acquire();
        tma_load_2d(dst, &tmW2, &bars->w_full[ws], e * kFE, 0);
        slot_advance();
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++it) {
        mbar_wait(&bars->a1_empty, (it & 1) ^ 1u);
        FFN_TRACE(it, 0);
        mbar_arrive_expect_tx(&bars->a1_full, kFA1Bytes);
        for (int kb = 0; kb < kFD / kBK; ++kb) tma_load_2d(a1_ptr + kb * kAStageBytes, &tmA, &bars->a1_full, kb * kBK, tile * kBM);
        load_g1(0);
        if (ne > 1) load_g1(1);
        for (int e = 0; e < ne; ++e) {
          load_g2(e);
          if (e + 2 < ne) load_g1(e + 2);
        }
        FFN_TRACE(it, 1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t ws = 0, wphase = 0;
      uint32_t n_g1 = 0, n_g2 = 0;
      const uint32_t a1_addr = smem_base, h_addr = a1_addr + kFA1Bytes, w_addr = h_addr + kFHBytes;
      auto issue_g1 = [&]() {
        const uint32_t buf = n_g1 & 1;
        mbar_wait(&bars->acc1_empty[buf], ((n_g1 >> 1) & 1) ^ 1u);
        mbar_wait(&bars->w_full[ws], wphase);
        tc_fence_after();
        const uint32_t sb = w_addr + ws * kFSlot;
#pragma unroll
        for (int kb = 0; kb < kFD / kBK; ++kb)
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16(tmem_base + buf * kFE, umma_desc_sw128(a1_addr + kb * kAStageBytes + k * 32),
                      umma_desc_sw128(sb + kb * kW1KbBytes + k * 32), IDESC1, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&bars->w_empty[ws]);
        if (++ws == kFStages) { ws = 0; wphase ^= 1u; }
        umma_commit(&bars->acc1_full[buf]);
        ++n_g1;
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++it) {
        mbar_wait(&bars->a1_full, it & 1);
        tc_fence_after();
        FFN_TRACE(it, 2);
        issue_g1();
        FFN_TRACE(it, 3);
        if (ne > 1) issue_g1();
        FFN_TRACE(it, 4);
        for (int e = 0; e < ne; ++e) {
          if (e == 0) mbar_wait(&bars->acc2_empty, (it & 1) ^ 1u);
          if (e == 0) FFN_TRACE(it, 5);
          const uint32_t hb = n_g2 & 1;
          mbar_wait(&bars->h_full[hb], (n_g2 >> 1) & 1);
          mbar_wait(&bars->w_full[ws], wphase);
          tc_fence_after();
          if (e < 4) FFN_TRACE(it, 6 + e);
          const uint32_t sb = w_addr + ws * kFSlot;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16(tmem_acc2, umma_desc_sw128(h_addr + hb * kAStageBytes + k * 32), umma_desc_sw128(sb + k * 32), IDESC2,
                      (e | k) != 0 ? 1u : 0u);
          umma_commit(&bars->w_empty[ws]);
          if (++ws == kFStages) { ws = 0; wphase ^= 1u; }
          umma_commit(&bars->h_empty[hb]);
          ++n_g2;
          if (e < 4) FFN_TRACE(it, 10 + e);
          if (e == ne - 1) umma_commit(&bars->acc2_full);
          if (e + 2 < ne) issue_g1();
          if (e + 2 == ne - 1 || (ne <= 2 && e == 0)) umma_commit(&bars->a1_empty);  // last GEMM1 of the tile issued
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== 16 epilogue warps =====================
    const int q4 = warp & 3;         // TMEM lane quarter
    const int c = (warp - 4) >> 2;   // 16-column slice of a hidden step / 64-column box of the output tile
    float2* xch2 = reinterpret_cast<float2*>(xch);
    float* cvec = reinterpret_cast<float*>(xch) + 2048;   // [5][256]: b2 | ln0 g | ln0 b | ln1 g | ln1 b
    float* b1s = cvec + 5 * 256;                          // [E] first-layer bias
    {
      const int t16 = static_cast<int>(threadIdx.x) - 128;
      for (int i = t16; i < 5 * 256; i += 512) {
        const int which = i >> 8, col = i & 255;
        const float* src = which == 0 ? ep.bias : which == 1 ? ep.ln0_g : which == 2 ? ep.ln0_b : which == 3 ? ep.ln1_g : ep.ln1_b;
        cvec[i] = src != nullptr ? __ldg(src + col) : 0.f;
      }
      for (int i = t16; i < E; i += 512) b1s[i] = __ldg(bias1 + i);
      named_bar_sync(5, 512);
    }
    WarpStore st;
    st.single = true;
    st.base = smem_u32(h_ptr) + static_cast<uint32_t>(warp - 4) * (kWarpStgBytes / 2);  // 2 KB half box inside the H bytes
    st.iter = 0;
    st.lane = lane;
    Row16State rst;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    // this thread's 32 bytes of an H row: 16-byte chunks 2c and 2c+1 of row r = 32 q4 + lane, 128B-swizzled
    const uint32_t hrow = static_cast<uint32_t>(q4 * 32 + lane) * 128u;
    const uint32_t hx = static_cast<uint32_t>(lane & 7);
    uint32_t n_e = 0;  // running count of processed steps
    int it = 0;
    for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++it) {
      const int row0 = tile * kBM + q4 * 32;
      EpiThread th;
      th.row = row0 + lane;
      th.valid = th.row < M;
      th.seq = th.valid ? th.row / ep.rows_per_seq : 0;
      th.t = th.valid ? th.row - th.seq * ep.rows_per_seq : 0;
      th.taddr = tmem_acc2 + lane_addr;
      uint4 rq[8];
      for (int e = 0; e < ne; ++e, ++n_e) {
        const uint32_t buf = n_e & 1;
        mbar_wait(&bars->acc1_full[buf], (n_e >> 1) & 1);
        tc_fence_after();
        if (warp == 4 && lane == 0 && e < 4) FFN_TRACE(it, 14 + e);
        // H buffer free? the GEMM2 that read it two steps ago has retired; at a tile boundary the previous tile's output
        // stores (staged through the same bytes by ANY warp) must have been read as well
        if (n_e >= 2) mbar_wait(&bars->h_empty[buf], ((n_e >> 1) - 1u) & 1u);
        if (e == 0 && it > 0) {
          if (lane == 0) tma_store_wait_read<0>();
          named_bar_sync(6, 512);
        }
        if (warp == 4 && lane == 0 && e < 4) FFN_TRACE(it, 18 + e);
        {
          uint32_t raw[16];
          tmem_ld16(tmem_base + buf * kFE + lane_addr + c * 16, raw);
          tmem_ld_wait();
          const float* bq = b1s + e * kFE + c * 16;
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 bb = *reinterpret_cast<const float4*>(bq + 4 * j);
            float h0, h1, h2, h3;
            fadd2(h0, h1, __uint_as_float(raw[4 * j + 0]), __uint_as_float(raw[4 * j + 1]), bb.x, bb.y);
            fadd2(h2, h3, __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]), bb.z, bb.w);
            fmul2(h0, h1, h0, h1, 0.5f, 0.5f);  // swish(x) = h + h tanh(h), h = x / 2
            fmul2(h2, h3, h2, h3, 0.5f, 0.5f);
            const float t0 = fast_tanh(h0), t1 = fast_tanh(h1), t2 = fast_tanh(h2), t3 = fast_tanh(h3);
            float y0, y1, y2, y3;
            ffma2(y0, y1, h0, h1, t0, t1, h0, h1);
            ffma2(y2, y3, h2, h3, t2, t3, h2, h3);
            pk[2 * j] = pack_bf16x2(y0, y1);
            pk[2 * j + 1] = pack_bf16x2(y2, y3);
          }
          const uint32_t hb_addr = smem_u32(h_ptr) + buf * kAStageBytes + hrow;
          st_shared_v4(hb_addr + ((static_cast<uint32_t>(2 * c) ^ hx) << 4), pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(hb_addr + ((static_cast<uint32_t>(2 * c + 1) ^ hx) << 4), pk[4], pk[5], pk[6], pk[7]);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->acc1_empty[buf]);
          mbar_arrive(&bars->h_full[buf]);
        }
        if (warp == 4 && lane == 0 && e < 4) FFN_TRACE(it, 22 + e);
        if (e == ne - 2) row16_resid_ldg(ep, th, c, rq);  // in flight under the last step and the last GEMM2
      }
      if (ne < 2) row16_resid_ldg(ep, th, c, rq);
      // ---- final epilogue of the tile: acc2 -> +b2 -> +residual -> [LN] -> S (and LN'(S)) ----
      mbar_wait(&bars->acc2_full, it & 1);  // every GEMM2 of the tile has retired: the H bytes are free for staging
      tc_fence_after();
      if (warp == 4 && lane == 0) FFN_TRACE(it, 26);
      epilogue_row16<true>(ep, th, row0, q4, c, lane, st, &tmO0h, &tmO1h, xch2, cvec, rst, nullptr, 0u, rq);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc2_empty);
      if (warp == 4 && lane == 0) FFN_TRACE(it, 27);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool ffn_applicable(int D, int E, int M, int num_sms) {
  (void)num_sms;
  return D == kFD && E % kFQ == 0 && E >= 2 * kFQ && E <= 768 && M >= 1;  // E <= 768: the first-layer bias is kept in shared memory
}

// plan->tmA: XN [M, 256]; tmW1: W1^T [E, 256] box [128 x 64]; tmW2: W2^T [256, E] box [256 x 64]; tmO0 / tmO1: S / LN'(S)
int ffn_plan_init(FfnPlan* p, const bf16* xn, const bf16* w1t, const bf16* w2t, bf16* out0, bf16* out1) {
  int rc;
  if ((rc = make_tmap_2d(&p->tmA, xn, TM_BF16, p->M, kFD, kFD, kBM, kBK))) return rc;
  if ((rc = make_tmap_2d(&p->tmW1, w1t, TM_BF16, p->E, kFD, kFD, kFQ, kBK))) return rc;
  if ((rc = make_tmap_2d(&p->tmW2, w2t, TM_BF16, kFD, p->E, p->E, kFD, kBK))) return rc;
  if ((rc = make_tmap_2d(&p->tmO0, out0, TM_BF16, p->M, kFD, kFD, 32, 64))) return rc;
  if (out1 != nullptr) {
    if ((rc = make_tmap_2d(&p->tmO1, out1, TM_BF16, p->M, kFD, kFD, 32, 64))) return rc;
  } else {
    p->tmO1 = p->tmO0;
  }
  if ((rc = make_tmap_2d(&p->tmW1e, w1t, TM_BF16, p->E, kFD, kFD, kFE, kBK))) return rc;
  if ((rc = make_tmap_2d(&p->tmO0h, out0, TM_BF16, p->M, kFD, kFD, 32, 32))) return rc;
  if (out1 != nullptr) {
    if ((rc = make_tmap_2d(&p->tmO1h, out1, TM_BF16, p->M, kFD, kFD, 32, 32))) return rc;
  } else {
    p->tmO1h = p->tmO0h;
  }
  return 0;
}

int ffn_launch(const FfnPlan& p, int num_sms, cudaStream_t stream) {
  const int smem = kFA1Bytes + kFHBytes + kFStages * kFSlot + kXchBytes + 256 + 1024;
  static bool attr = false;
  if (!attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(ffn_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ISHARA_CUDA_OK(cudaFuncSetAttribute(ffn_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ISHARA_CUDA_OK(cudaFuncSetAttribute(ffn_tc_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  // v2 (eighths, two H buffers) is bit-identical and measured SLOWER (102.6 vs 93.7 us): both versions are paced by the
  // weight ring - 96 KB in flight against ~2.4k cycles of TMA latency is ~40 B/clk, i.e. >= 13k cycles for the 512 KB of
  // weights every tile needs - and v2 doubles the number of ring turnarounds. Opt-in.
  static const int v2 = getenv("ISHARA_FFN_V2") ? atoi(getenv("ISHARA_FFN_V2")) : 0;
  static const int ew16 = getenv("ISHARA_FFN_EW16") ? atoi(getenv("ISHARA_FFN_EW16")) : 1;
  const int mt = (p.M + kBM - 1) / kBM;
  const int grid = mt < num_sms ? mt : num_sms;
#ifdef ISHARA_TRACE_BUILD
  static long long* tbuf = nullptr;
  static int printed = 0;
  const bool tracing = printed < 2 && grid > 60;
  if (tracing) {
    if (tbuf == nullptr) ISHARA_CUDA_OK(cudaMalloc(&tbuf, 96 * sizeof(long long)));
    ISHARA_CUDA_OK(cudaMemsetAsync(tbuf, 0, 96 * sizeof(long long), stream));
    ISHARA_CUDA_OK(cudaMemcpyToSymbolAsync(g_ffn_trace, &tbuf, sizeof(tbuf), 0, cudaMemcpyHostToDevice, stream));
  }
#endif
  if (v2 && ew16)
    ffn_tc_v2_kernel<<<grid, 640, smem, stream>>>(p.tmA, p.tmW1e, p.tmW2, p.tmO0h, p.tmO1h, p.epi, p.bias1, p.M, p.E, mt);
  else if (ew16)
    ffn_tc_kernel<16><<<grid, 640, smem, stream>>>(p.tmA, p.tmW1, p.tmW2, p.tmO0, p.tmO1, p.tmO0h, p.tmO1h, p.epi, p.bias1, p.M, p.E, mt);
  else
    ffn_tc_kernel<8><<<grid, 384, smem, stream>>>(p.tmA, p.tmW1, p.tmW2, p.tmO0, p.tmO1, p.tmO0h, p.tmO1h, p.epi, p.bias1, p.M, p.E, mt);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
#ifdef ISHARA_TRACE_BUILD
  if (tracing) {
    ISHARA_CUDA_OK(cudaStreamSynchronize(stream));
    long long h[96];
    ISHARA_CUDA_OK(cudaMemcpy(h, tbuf, sizeof(h), cudaMemcpyDeviceToHost));
    ++printed;
    fprintf(stderr, "ffn trace M=%d (CTA 60, cycles since its first event): P0 a1_empty ok, P1 loads issued | M2 a1 landed, M3 G1(0) issued, M4 G1(1) issued, "
                    "M5 acc2 free, M6-9 H(q) seen, M10-13 G2(q) issued | E14-17 acc1(q) seen, E18-21 H free, E22-25 epi1(q) done, E26 acc2 seen, E27 epi2 done\n", p.M);
    long long t0 = 0;
    for (int i = 0; i < 32; ++i) if (h[i] != 0 && (t0 == 0 || h[i] < t0)) t0 = h[i];
    for (int t = 0; t < 3; ++t) {
      fprintf(stderr, "  tile %d:", t);
      for (int e = 0; e < 28; ++e) fprintf(stderr, " %d:%lld", e, h[t * 32 + e] ? h[t * 32 + e] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
#endif
  return 0;
}

}  // namespace ishara
