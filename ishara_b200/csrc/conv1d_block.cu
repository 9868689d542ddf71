// ishara_b200 — the whole Conv1DBlock in ONE launch (sm_100a only):
//   S <- S + (ECA(BN(CausalDWConv1D_k(swish(S @ We + be)))) @ Wp + bp)      [, XN <- LayerNorm(S) for the next module]
//
// Reference: Conv1DBlock (nb:conv-hybrid-model c5:41-89) = Dense(2D, swish) c5:61-65 -> CausalDWConv1D(k) c5:17-39,68-71
// -> BatchNorm c5:73 -> ECA c5:1-15,75 -> Dense(D) c5:77-80 -> (+ skip) c5:85-86; 12 of them per forward (c7:20-29).
// As three launches (expand GEMM, depthwise kernel, project GEMM) the 2D-wide tensor crossed HBM four times
// (571 MB per block at B = 256 against 100.7 MB of compulsory traffic). Here it exists only in shared memory / TMEM:
// the kernel reads the [T, D] stream once and writes it once.
//
//   cluster = the T/128 CTAs of ONE sequence; CTA r owns frames [128 r, 128 r + 128); 640 threads:
//   warp 0 TMA producer | warp 1 MMA issuer | warp 2 TMEM allocator | warps 4-19: 16 worker warps
//
//   1. expand:  acc[128 x 512] fp32 (all 512 TMEM columns, two N = 256 halves) = S_tile @ We^T, operands through three
//      48 KB TMA slots; the first half is drained while the second half's MMAs run
//   2. drain:   + be, swish, bf16 -> H[128 x 512] in shared memory, laid out as the eight [128 x 64] K-major
//      128B-swizzled boxes that tcgen05.mma consumes as the A operand of the project GEMM (16 halo rows in front of
//      each box hold the k-1 frames of the previous CTA)
//   3. ECA:     mean over the VALID frames of the BatchNorm'd conv output. The conv is linear in time, so the mean is
//      wsum * colsum(H) - (a correction from the last k-1 valid frames): every CTA publishes its partial sums,
//      cluster barrier, every CTA derives all 512 channel scales (5-tap conv across channels + sigmoid) from DSMEM
//   4. stencil: IN PLACE, box pair by box pair: warp = (box, 16-frame group), lane = channel pair, register sliding
//      window walking DOWN the frames (so a frame is overwritten only after every tap that needs it has read it);
//      BatchNorm and the ECA scale are folded into the taps. As soon as a box is finished the MMA warp issues its
//      four project MMAs (K = 64) against the matching Wp slice streamed through a two-slot ring
//   5. epilogue (16 warps, lane = row): + bp, + residual, [LayerNorm statistics exchanged between the four warps of a
//      lane quarter, values parked in TMEM], TMA stores of S and of LN(S)
// Only D == 256 (2D = 512 TMEM columns), T % 128 == 0, T <= 1024; other shapes use the three-kernel path.
#include <cstdio>
#include <cstdlib>

#include "gemm_epilogue.cuh"

namespace ishara {
namespace {

constexpr int kKD = 256;                      // model dim
constexpr int kKC = 512;                      // expanded channels
constexpr int kBoxHalo = 2048;                // 16 halo rows x 128 B in front of every box
constexpr int kBoxStride = kBoxHalo + kAStageBytes;  // 18 KB (keeps every box 1024-byte aligned)
constexpr int kHRegion = 8 * kBoxStride;      // 144 KB
constexpr int kWSlot = 32 * 1024;             // one [256 x 64] k-slice of Wp^T
constexpr int kWRing = 2 * kWSlot;
constexpr int kXSlot = 48 * 1024;             // expand slot: S chunk [128 x 64] + We chunk [256 x 64]
constexpr int kFloatBytes = 16 * 1024;        // part | corr | tapbuf (overlaps mean, which first holds the expand bias) | scale | constants; LN exchange over the first 8 KB
constexpr int kCbSmem = kHRegion + kWRing + kFloatBytes + 384 + 1024;
constexpr int kWorkers = 16;
constexpr int kCbThreads = 128 + 32 * kWorkers;

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + __expf(-x)); }

// Timeline tracing (profiling builds only: make TRACE=1 -> libishara_b200_trace.so): clock64 at phase boundaries of one
// sequence's CTAs. Compiled out of the production library.
#ifdef ISHARA_C1B_TRACE
#define CB_TRACE(ev_) do { if (pr.trace != nullptr && blockIdx.y == pr.trace_b) pr.trace[blockIdx.x * 32 + (ev_)] = clock64(); } while (0)
#else
#define CB_TRACE(ev_) do { } while (0)
#endif

struct CbBars {
  uint64_t full[3], empty[3];   // expand slots
  uint64_t accf[2];             // expand accumulator half ready
  uint64_t wfull[2], wempty[2]; // Wp ring
  uint64_t boxr[8];             // stencil finished box b (8 warps each)
  uint64_t acc2f;               // project accumulator ready
  uint64_t taps_full, taps_empty;  // tap staging buffer: filled by warps 2-3, drained by the 16 worker warps
  uint64_t stg_free;            // boxes 0-3 are dead (project MMAs of chunks 0-3 retired): the epilogue staging may be filled
  uint64_t resid_full[kWorkers];  // residual box of each epilogue warp has landed in its staging box
  uint32_t tmem_slot;
};

struct CbParams {
  const float* bias_e;   // [512]
  const float* dw_w;     // [k, 512] taps, BatchNorm folded
  const float* dw_b;     // [512]
  const float* eca_w;    // [5]
  const float* bias_p;   // [256]
  const bf16* resid;     // S [M, 256]
  const float* ln_g;     // LayerNorm of the NEXT module (null: none)
  const float* ln_b;
  float ln_eps;
  const float* dw_wsum;    // [512] sum over the k taps (per channel)
  const uint16_t* wbits;   // [B*T] window validity bits (mask_mode="propagated": ECA averages over valid frames), null = all valid
  const int32_t* valid_cnt;  // [B] valid frames per sequence (with wbits)
  int T;
  long long* trace;        // ISHARA_C1B_TRACE builds only
  int trace_b;
};

template <int K>
__global__ void __launch_bounds__(kCbThreads, 1)
conv1d_block_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
                    const __grid_constant__ CUtensorMap tmWp, const __grid_constant__ CUtensorMap tmO0,
                    const __grid_constant__ CUtensorMap tmO1, const CbParams pr) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  float* part = reinterpret_cast<float*>(smem + kHRegion + kWRing);   // [2 row halves][512] column sums of H over this tile's valid frames
  float* corr = part + 2 * kKC;                                       // [512] this tile's share of the tail correction
  float* tapbuf = part;                                               // [(K+1)][128] taps + offset of the NEXT stencil round (6 KB; part/corr are dead by then)
  float* mean = part + 4 * kKC;                                       // [512] (holds the expand bias until the drain is over)
  float* scale = part + 5 * kKC;                                      // [512]
  float* cvec = part + 6 * kKC;                                          // bias_p[256] | ln_g[256] | ln_b[256] | eca_w[5]
  CbBars* bars = reinterpret_cast<CbBars*>(smem + kHRegion + kWRing + kFloatBytes);
  static_assert(sizeof(CbBars) <= 384, "barrier block");
  uint16_t* wb_s = reinterpret_cast<uint16_t*>(cvec + 800);           // [128] window bits of this tile's frames
  uint32_t* mixed_s = reinterpret_cast<uint32_t*>(cvec + 800 + 64);   // [4] ballots: frames whose k-window is partly valid

  const int rank = blockIdx.x, nrank = gridDim.x;   // tile index inside the sequence == rank in the cluster
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = pr.T;
  const int row_g0 = b * T + rank * kBM;
  const uint32_t w_base = smem_base + kHRegion;

  // Expand operand slots. Loads 0-3 feed accumulator half 0, loads 4-7 half 1; the second half only uses the two slots
  // that do not overlap boxes 0-3, so the drain of half 0 can run under the MMAs of half 1.
  //   slot 0 = [0, 48K) (boxes 0-2), slot 1 = [72K, 120K) (boxes 4-6), slot 2 = [120K, 168K) (boxes 6-7 + Wp slot 0)
  const uint32_t slot_off[3] = {0u, 72u * 1024u, 120u * 1024u};
  const int slot_of[8] = {0, 1, 2, 0, 1, 2, 1, 2};
  const int use_of[8] = {0, 0, 0, 1, 1, 1, 2, 2};

  if (warp == 0 && lane == 0) {
    // the TMA thread initialises the barriers its first loads signal and issues those loads BEFORE the CTA-wide
    // barrier below (timeline: the first expand MMA used to start 2.4k cycles into the CTA, ~1k of it this prologue)
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWe);
    for (int s = 0; s < 3; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int s = 0; s < 2; ++s) mbar_init(&bars->wfull[s], 1);
    mbar_fence_init();
    for (int i = 0; i < 3; ++i) {
      const int s = slot_of[i];
      mbar_arrive_expect_tx(&bars->full[s], kXSlot);
      uint8_t* dst = smem + slot_off[s];
      tma_load_2d(dst, &tmX, &bars->full[s], i * kBK, row_g0);
      tma_load_2d(dst + kAStageBytes, &tmWe, &bars->full[s], i * kBK, 0);
    }
    tma_prefetch_desc(&tmWp);
    // Wp slice 1 lives in [176K, 208K): never touched by the expand slots, prefetch it right away
    mbar_arrive_expect_tx(&bars->wfull[1], kWSlot);
    tma_load_2d(smem + kHRegion + kWSlot, &tmWp, &bars->wfull[1], 1 * kBK, 0);
    tma_prefetch_desc(&tmO0);
    if (pr.ln_g != nullptr) tma_prefetch_desc(&tmO1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&bars->accf[s], 1); mbar_init(&bars->wempty[s], 1); }
    for (int s = 0; s < 8; ++s) mbar_init(&bars->boxr[s], 8);
    mbar_init(&bars->acc2f, 1);
    mbar_init(&bars->taps_full, 2);
    mbar_init(&bars->taps_empty, kWorkers);
    mbar_init(&bars->stg_free, 1);
    for (int s = 0; s < kWorkers; ++s) mbar_init(&bars->resid_full[s], 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  constexpr uint32_t IDESC = umma_idesc(kBM, 256, 1);
  if (threadIdx.x == 0) CB_TRACE(0);

  // ======================================= phase 1: expand GEMM + drain =======================================
  if (warp == 0) {
    if (lane == 0) {
      // (loads 0-2 and Wp slice 1 were issued in the prologue)
#pragma unroll
      for (int i = 3; i < 8; ++i) {
        const int s = slot_of[i], u = use_of[i];
        const int nh = i >> 2, kb = i & 3;
        if (u > 0) mbar_wait(&bars->empty[s], static_cast<uint32_t>((u - 1) & 1));
        mbar_arrive_expect_tx(&bars->full[s], kXSlot);
        uint8_t* dst = smem + slot_off[s];
        tma_load_2d(dst, &tmX, &bars->full[s], kb * kBK, row_g0);
        tma_load_2d(dst + kAStageBytes, &tmWe, &bars->full[s], kb * kBK, nh * 256);
      }
      // every expand MMA has retired -> slot 2 is dead -> Wp slice 0 may land in [144K, 176K)
      mbar_wait(&bars->accf[1], 0);
      mbar_arrive_expect_tx(&bars->wfull[0], kWSlot);
      tma_load_2d(smem + kHRegion, &tmWp, &bars->wfull[0], 0, 0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int s = slot_of[i], u = use_of[i];
        const int nh = i >> 2, kb = i & 3;
        mbar_wait(&bars->full[s], static_cast<uint32_t>(u & 1));
        tc_fence_after();
        const uint32_t sa = smem_base + slot_off[s];
        const uint32_t sb = sa + kAStageBytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          umma_bf16(tmem_base + nh * 256, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), IDESC,
                    (kb | k) != 0 ? 1u : 0u);
        umma_commit(&bars->empty[s]);
        if (kb == 3) umma_commit(&bars->accf[nh]);
        if (i == 0) CB_TRACE(1);
        if (i == 3) CB_TRACE(2);
        if (i == 7) CB_TRACE(3);
      }
    }
  } else if (warp >= 4) {
    // constants the later phases need go to shared memory now, while the expand GEMM runs: with the shared-memory
    // carve-out at its maximum the L1 is ~2 KB, so every __ldg on a critical path costs an L2 round trip
    const int ww = warp - 4;
    {
      const int wt = threadIdx.x - 128;
      if (wt < 128) {
        reinterpret_cast<float4*>(mean)[wt] = __ldg(reinterpret_cast<const float4*>(pr.bias_e) + wt);
      } else if (wt < 192) {
        reinterpret_cast<float4*>(cvec)[wt - 128] = __ldg(reinterpret_cast<const float4*>(pr.bias_p) + (wt - 128));
      } else if (wt < 256) {
        if (pr.ln_g != nullptr) reinterpret_cast<float4*>(cvec)[wt - 128] = __ldg(reinterpret_cast<const float4*>(pr.ln_g) + (wt - 192));
      } else if (wt < 320) {
        if (pr.ln_g != nullptr) reinterpret_cast<float4*>(cvec)[wt - 128] = __ldg(reinterpret_cast<const float4*>(pr.ln_b) + (wt - 256));
      } else if (wt < 325) {
        cvec[768 + (wt - 320)] = __ldg(pr.eca_w + (wt - 320));
      } else if (wt >= 384) {
        // window validity of this tile's 128 frames, cut to the k taps of this block: bit d = frame t+d counts in the ECA mean
        const int r = wt - 384, t = rank * kBM + r;
        constexpr uint32_t KM = (1u << K) - 1u;
        uint32_t bits = pr.wbits != nullptr ? pr.wbits[static_cast<size_t>(b) * T + t] : (T - t >= 16 ? 0xFFFFu : ((1u << (T - t)) - 1u));
        bits &= KM;
        wb_s[r] = static_cast<uint16_t>(bits);
        const uint32_t mixed = __ballot_sync(0xffffffffu, bits != 0u && bits != KM);
        if (lane == 0) mixed_s[r >> 5] = mixed;
      }
      named_bar_sync(3, 32 * kWorkers);
    }
    // drain: warp (q, c) turns rows [32q, +32) x channels [256 nh + 64 c, +64) into box 4 nh + c; 16 columns at a time,
    // the TMEM read of the next 16 in flight under the math of the current ones
    const int q = warp & 3, c = ww >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float* bias_s = mean;
#pragma unroll 1
    for (int nh = 0; nh < 2; ++nh) {
      mbar_wait(&bars->accf[nh], 0);
      tc_fence_after();
      if (warp == 4 && lane == 0) CB_TRACE(4 + 2 * nh);
      const uint32_t rowbase = smem_base + static_cast<uint32_t>(4 * nh + c) * kBoxStride + kBoxHalo + static_cast<uint32_t>(q * 32 + lane) * 128u;
      const uint32_t xr = static_cast<uint32_t>(lane & 7);
      const int col0 = nh * 256 + c * 64;
      const uint32_t t0 = tmem_base + lane_addr + static_cast<uint32_t>(col0);
      uint32_t ra[16], rb[16];
      tmem_ld16(t0, ra);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t (&cur)[16] = (i & 1) ? rb : ra;
        uint32_t (&nxt)[16] = (i & 1) ? ra : rb;
        tmem_ld_fence16(cur);
        if (i < 3) tmem_ld16(t0 + 16u * (i + 1), nxt);
        float v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 bb = *reinterpret_cast<const float4*>(bias_s + col0 + 16 * i + 4 * j);
          fadd2(v[4 * j + 0], v[4 * j + 1], __uint_as_float(cur[4 * j + 0]), __uint_as_float(cur[4 * j + 1]), bb.x, bb.y);
          fadd2(v[4 * j + 2], v[4 * j + 3], __uint_as_float(cur[4 * j + 2]), __uint_as_float(cur[4 * j + 3]), bb.z, bb.w);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // swish(x) = h + h tanh(h), h = x / 2
          float h0, h1;
          fmul2(h0, h1, v[2 * j], v[2 * j + 1], 0.5f, 0.5f);
          const float t0f = fast_tanh(h0), t1f = fast_tanh(h1);
          ffma2(v[2 * j], v[2 * j + 1], h0, h1, t0f, t1f, h0, h1);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
          st_shared_v4(rowbase + ((static_cast<uint32_t>(2 * i + j) ^ xr) << 4), pack_bf16x2(v[8 * j + 0], v[8 * j + 1]),
                       pack_bf16x2(v[8 * j + 2], v[8 * j + 3]), pack_bf16x2(v[8 * j + 4], v[8 * j + 5]),
                       pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
      }
      if (warp == 4 && lane == 0) CB_TRACE(5 + 2 * nh);
    }
    tc_fence_before();
    named_bar_sync(3, 32 * kWorkers);  // the whole H tile is in shared memory
    if (warp == 4 && lane == 0) CB_TRACE(8);

    // ---- ECA mean without a second conv pass. With m = frame validity (all ones unless mask_mode="propagated"):
    //   sum_t m_t y_t = b * sum_t m_t + sum_u h[u] * (sum_d w_{K-1-d} m_{u+d})
    // i.e. wsum * h[u] for every frame whose k-frame window [u, u+K) is entirely valid ("full"), an explicit weight for
    // the few frames next to a validity edge or the end of the sequence ("mixed"), nothing for the others.
    // Full frames: warp = (box, 64-frame half); lane = (16-byte chunk c8 of the 128-byte row, frame phase sub): the 8
    // lanes of a quarter warp read one whole swizzled row (conflict-free LDS.128), 8 channel accumulators per thread,
    // the 4 frame phases are folded with two shuffles ----
    {
      constexpr uint32_t KM = (1u << K) - 1u;
      const int box = ww >> 1, rh = ww & 1;
      const int c8 = lane & 7, sub = lane >> 3;
      const uint32_t bx = smem_base + static_cast<uint32_t>(box) * kBoxStride + kBoxHalo;
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      // which of this thread's 16 rows are "full": read first, so that the row loads below do not wait on a branch each
      // (the branchy version took 2.8k cycles for a pass whose shared-memory floor is 1k)
      uint32_t fullmask = 0u;
#pragma unroll
      for (int i = 0; i < 16; ++i) fullmask |= (wb_s[rh * 64 + 4 * i + sub] == KM ? 1u : 0u) << i;
#pragma unroll 8
      for (int i = 0; i < 16; ++i) {
        const int r = rh * 64 + 4 * i + sub;
        uint4 v;
        const uint32_t addr = bx + static_cast<uint32_t>(r) * 128u + ((static_cast<uint32_t>(c8) ^ static_cast<uint32_t>(r & 7)) << 4);
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
        if (!((fullmask >> i) & 1u)) v = make_uint4(0u, 0u, 0u, 0u);  // bf16 zeros: adds nothing
        fadd2(acc[0], acc[1], acc[0], acc[1], bf16_lo(v.x), bf16_hi(v.x));
        fadd2(acc[2], acc[3], acc[2], acc[3], bf16_lo(v.y), bf16_hi(v.y));
        fadd2(acc[4], acc[5], acc[4], acc[5], bf16_lo(v.z), bf16_hi(v.z));
        fadd2(acc[6], acc[7], acc[6], acc[7], bf16_lo(v.w), bf16_hi(v.w));
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
      }
      if (warp == 4 && lane == 0) CB_TRACE(25);
      if (sub == 0) {
        float4* dst = reinterpret_cast<float4*>(part + rh * kKC + box * 64 + c8 * 8);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    }
    // mixed frames (warp = box, lane = channel pair): the last k-1 frames of the sequence, plus the frames next to every
    // validity edge in propagated mode
    if (ww < 8) {
      const int box = ww;
      const uint32_t bx = smem_base + static_cast<uint32_t>(box) * kBoxStride + kBoxHalo;
      const uint32_t lsw = static_cast<uint32_t>(lane >> 2), lw = static_cast<uint32_t>(lane & 3) << 2;
      float2 cr = make_float2(0.f, 0.f);
      if ((mixed_s[0] | mixed_s[1] | mixed_s[2] | mixed_s[3]) != 0u) {
        const int ch = box * 64 + 2 * lane;
        float2 wj[K];
#pragma unroll
        for (int j = 0; j < K; ++j) wj[j] = __ldg(reinterpret_cast<const float2*>(pr.dw_w + static_cast<size_t>(j) * kKC + ch));
#pragma unroll 1
        for (int g32 = 0; g32 < 4; ++g32) {
          uint32_t mm = mixed_s[g32];
          while (mm != 0u) {
            const int r = g32 * 32 + __ffs(static_cast<int>(mm)) - 1;
            mm &= mm - 1u;
            const uint32_t bits = wb_s[r];
            float2 w = make_float2(0.f, 0.f);
#pragma unroll
            for (int d = 0; d < K; ++d)
              if ((bits >> d) & 1u) { w.x += wj[K - 1 - d].x; w.y += wj[K - 1 - d].y; }
            const uint32_t u = ld_shared_u32(bx + static_cast<uint32_t>(r) * 128u + (((lsw ^ static_cast<uint32_t>(r & 7)) << 4) | lw));
            cr.x = fmaf(w.x, bf16_lo(u), cr.x);
            cr.y = fmaf(w.y, bf16_hi(u), cr.y);
          }
        }
      }
      *reinterpret_cast<float2*>(corr + box * 64 + 2 * lane) = cr;
    }
    if (warp == 4 && lane == 0) CB_TRACE(26);
    named_bar_sync(3, 32 * kWorkers);  // every worker's shared-memory writes are ordered before the one fence below
    if (warp == 4 && lane == 0) CB_TRACE(27);
    if (warp == 4 && lane == 0) asm volatile("fence.acq_rel.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 4 && lane == 0) CB_TRACE(9);
  // #1: every CTA of the sequence has its H tile and its partial sums in shared memory. A release-arrive by all 640
  // threads cost ~2k cycles of MEMBAR per barrier (ncu: ERRBAR / UCGABAR_ARV membar stalls); here the workers order their
  // writes with a CTA barrier, ONE thread issues the cluster-scope fence and everybody arrives relaxed.
  cluster_arrive_relaxed();
  cluster_wait();
  if (warp == 4 && lane == 0) CB_TRACE(10);

  if (warp >= 4) {
    const int ww = warp - 4;
    // ---- halo: the previous CTA's last K-1 frames (zeros in front of the sequence) -> rows -(K-1)..-1 of every box.
    //      (128 - j) & 7 == (-j) & 7, so a physical 128-byte row keeps its swizzle phase and is copied word by word ----
    {
      const int box = ww >> 1;
      const uint32_t bx = smem_base + static_cast<uint32_t>(box) * kBoxStride + kBoxHalo;
      constexpr int NH = K / 2;  // ceil((K - 1) / 2) rows per warp
      uint32_t u[NH];
#pragma unroll
      for (int jj = 0; jj < NH; ++jj) {  // all remote loads first (a DSMEM load is ~200 cycles)
        const int j = 1 + (ww & 1) + 2 * jj;
        u[jj] = 0u;
        if (rank > 0 && j <= K - 1)
          u[jj] = ld_dsmem_u32(mapa_shared(bx + static_cast<uint32_t>(kBM - j) * 128u + static_cast<uint32_t>(lane) * 4u, rank - 1));
      }
#pragma unroll
      for (int jj = 0; jj < NH; ++jj) {
        const int j = 1 + (ww & 1) + 2 * jj;
        if (j <= K - 1) st_shared_u32(bx - static_cast<uint32_t>(j) * 128u + static_cast<uint32_t>(lane) * 4u, u[jj]);
      }
    }
    // ---- channel means of the BatchNorm'd conv output over the valid frames (every CTA computes all 512) ----
    const int wt = threadIdx.x - 128;
    if (wt < kKC / 2) {
      const int ch = 2 * wt;
      float2 S = make_float2(0.f, 0.f), C = make_float2(0.f, 0.f);
      const float2 wsum = __ldg(reinterpret_cast<const float2*>(pr.dw_wsum + ch));
      const float2 bdw = __ldg(reinterpret_cast<const float2*>(pr.dw_b + ch));
      // GlobalAveragePooling1D(mask) divides by the number of valid frames (0 valid frames: 0/0 = NaN, as in Keras)
      const float invL = 1.f / static_cast<float>(pr.valid_cnt != nullptr ? pr.valid_cnt[b] : T);
      for (int r0 = 0; r0 < nrank; r0 += 4) {  // four CTAs' partial sums in flight at a time
        float2 a[4][2], c2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i][0] = a[i][1] = c2[i] = make_float2(0.f, 0.f);
          if (r0 + i < nrank) {
            a[i][0] = ld_dsmem_f32x2(mapa_shared(smem_u32(part + ch), r0 + i));
            a[i][1] = ld_dsmem_f32x2(mapa_shared(smem_u32(part + kKC + ch), r0 + i));
            c2[i] = ld_dsmem_f32x2(mapa_shared(smem_u32(corr + ch), r0 + i));
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          S.x += a[i][0].x + a[i][1].x; S.y += a[i][0].y + a[i][1].y;
          C.x += c2[i].x; C.y += c2[i].y;
        }
      }
      mean[ch] = fmaf(fmaf(wsum.x, S.x, C.x), invL, bdw.x);
      mean[ch + 1] = fmaf(fmaf(wsum.y, S.y, C.y), invL, bdw.y);
    }
  }
  __syncwarp();
  if (warp == 4 && lane == 0) CB_TRACE(11);
  // #2 (arrive): this CTA no longer reads its neighbours' shared memory. Relaxed: the remote loads above have completed
  // (their values were consumed), which is all the write-after-read hazard needs; a release-arrive costs a MEMBAR per warp
  cluster_arrive_relaxed();
  float2 wt[K];                    // taps (BatchNorm folded, then ECA-scaled) of this thread's channel pair in the current round
  float2 bs = make_float2(0.f, 0.f);
  if (warp >= 4) {
    {  // round 0 straight from global memory, requested now so the L2 latency hides under the barriers below
      const int ch0 = ((warp - 4) >> 3) * 64 + 2 * lane;
#pragma unroll
      for (int j = 0; j < K; ++j) wt[j] = __ldg(reinterpret_cast<const float2*>(pr.dw_w + static_cast<size_t>(j) * kKC + ch0));
      bs = __ldg(reinterpret_cast<const float2*>(pr.dw_b + ch0));
    }
    named_bar_sync(3, 32 * kWorkers);  // mean[] complete
    if (warp == 4 && lane == 0) CB_TRACE(28);
    const int wt = threadIdx.x - 128;  // one channel per worker thread
    float z = 0.f;
#pragma unroll
    for (int d = -2; d <= 2; ++d) {
      const int cc = wt + d;
      if (cc >= 0 && cc < kKC) z = fmaf(cvec[768 + d + 2], mean[cc], z);
    }
    scale[wt] = sigmoid_exact(z);
    if (warp == 4 && lane == 0) CB_TRACE(29);
    named_bar_sync(3, 32 * kWorkers);  // scale[] complete, halo rows written
    asm volatile("bar.arrive 8, %0;" ::"r"(32 * kWorkers + 64) : "memory");  // mean[] is dead: warps 2-3 may overwrite it with taps
  }
  if (warp == 4 && lane == 0) CB_TRACE(12);
  cluster_wait();     // #2 (wait): the neighbour has copied its halo, this CTA's H tile may now be overwritten in place
  if (warp == 4 && lane == 0) CB_TRACE(13);

  // ======================================= phase 2: stencil -> project GEMM -> epilogue =======================================
  if (warp == 0) {
    if (lane == 0) {
      for (int c = 2; c < 8; ++c) {
        const int s = c & 1;
        mbar_wait(&bars->wempty[s], static_cast<uint32_t>(((c >> 1) - 1) & 1));
        mbar_arrive_expect_tx(&bars->wfull[s], kWSlot);
        tma_load_2d(smem + kHRegion + s * kWSlot, &tmWp, &bars->wfull[s], c * kBK, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int c = 0; c < 8; ++c) {
        const int s = c & 1;
        mbar_wait(&bars->wfull[s], static_cast<uint32_t>((c >> 1) & 1));
        if (c == 7) CB_TRACE(23);
        mbar_wait(&bars->boxr[c], 0);
        tc_fence_after();
        if (c == 0) CB_TRACE(20);
        if (c == 6) CB_TRACE(30);
        if (c == 7) CB_TRACE(21);
        const uint32_t sa = smem_base + static_cast<uint32_t>(c) * kBoxStride + kBoxHalo;
        const uint32_t sb = w_base + static_cast<uint32_t>(s) * kWSlot;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          umma_bf16(tmem_base, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), IDESC, (c | k) != 0 ? 1u : 0u);
        umma_commit(&bars->wempty[s]);
        if (c == 3) umma_commit(&bars->stg_free);  // boxes 0-3 are no longer read: the epilogue staging (same bytes) may be filled
      }
      umma_commit(&bars->acc2f);
      CB_TRACE(24);
    }
  } else if (warp == 2 || warp == 3) {
    // ---- tap staging: taps + offsets of stencil round r+1 (128 channels) go global -> tapbuf while round r computes ----
    const int t64 = (warp - 2) * 32 + lane;
    named_bar_sync(8, 32 * kWorkers + 64);  // tapbuf overlaps mean[]: wait until the workers have derived the ECA scales
#pragma unroll 1
    for (int round = 1; round < 4; ++round) {
      if (round > 1) mbar_wait(&bars->taps_empty, static_cast<uint32_t>(round & 1));  // completion index round-2
      for (int i = t64; i < (K + 1) * 32; i += 64) {
        const int j = i >> 5, piece = i & 31;  // row j of the tap table (row K = offsets), 16-byte piece of 128 channels
        const float* src = (j < K ? pr.dw_w + static_cast<size_t>(j) * kKC : pr.dw_b) + round * 128 + piece * 4;
        cpa_16(smem_u32(tapbuf + j * 128 + piece * 4), src);
      }
      cpa_commit();
      cpa_wait_all();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->taps_full);
    }
  } else if (warp >= 4) {
    const int ww = warp - 4;
    // ---- stencil, in place: 4 rounds of two boxes; warp = (box of the round, 16-frame group), lane = channel pair ----
    {
      const int bi = ww >> 3, g = ww & 7;
      constexpr int TB = 4;
      const uint32_t lsw = static_cast<uint32_t>(lane >> 2), lw = static_cast<uint32_t>(lane & 3) << 2;
      uint32_t sw[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) sw[e] = ((lsw ^ static_cast<uint32_t>(e)) << 4) | lw;
#pragma unroll 1
      for (int round = 0; round < 4; ++round) {
        const int box = 2 * round + bi;
        const int ch = box * 64 + 2 * lane;
        const uint32_t base = smem_base + static_cast<uint32_t>(box) * kBoxStride + kBoxHalo + static_cast<uint32_t>(g) * 16u * 128u;
        // the K-1 frames in front of this group belong to the previous group (or to the halo rows): read them before
        // anybody overwrites them
        uint32_t hw[K - 1];
#pragma unroll
        for (int j = 0; j < K - 1; ++j) {
          const int r = j - (K - 1);  // -(K-1) .. -1
          hw[j] = ld_shared_u32(base + static_cast<uint32_t>(r * 128) + sw[r & 7]);
        }
        named_bar_sync(1 + bi, 256);  // the 8 warps of this box have read their halo frames
        if (round > 0) {
          mbar_wait(&bars->taps_full, static_cast<uint32_t>((round - 1) & 1));
#pragma unroll
          for (int j = 0; j < K; ++j) wt[j] = *reinterpret_cast<const float2*>(tapbuf + j * 128 + bi * 64 + 2 * lane);
          bs = *reinterpret_cast<const float2*>(tapbuf + K * 128 + bi * 64 + 2 * lane);
          __syncwarp();
          if (lane == 0 && round < 3) mbar_arrive(&bars->taps_empty);  // this warp holds the round's taps in registers
        }
        {
          const float2 sc = *reinterpret_cast<const float2*>(scale + ch);
#pragma unroll
          for (int j = 0; j < K; ++j) { wt[j].x *= sc.x; wt[j].y *= sc.y; }
          bs.x *= sc.x; bs.y *= sc.y;
        }
        float2 x[TB + K - 1];
#pragma unroll
        for (int blk = 16 / TB - 1; blk >= 0; --blk) {
          // window index i <-> frame 4 blk - (K-1) + i of the group
#pragma unroll
          for (int i = 0; i < TB + K - 1; ++i) {
            if (blk == 16 / TB - 1 || i < TB) {
              const int r = TB * blk - (K - 1) + i;
              const uint32_t u = r >= 0 ? ld_shared_u32(base + static_cast<uint32_t>(r * 128) + sw[r & 7]) : hw[r + (K - 1)];
              x[i] = make_float2(bf16_lo_prmt(u), bf16_hi(u));
            }
          }
          uint32_t o[TB];
#pragma unroll
          for (int t = 0; t < TB; ++t) {
            float2 a = bs;
#pragma unroll
            for (int j = 0; j < K; ++j) ffma2(a.x, a.y, wt[j].x, wt[j].y, x[t + j].x, x[t + j].y, a.x, a.y);
            o[t] = pack_bf16x2(a.x, a.y);
          }
#pragma unroll
          for (int t = 0; t < TB; ++t) {
            const int r = TB * blk + t;
            st_shared_u32(base + static_cast<uint32_t>(r * 128) + sw[r & 7], o[t]);
          }
          // keep the lowest K-1 frames of the window for the next (lower) block
#pragma unroll
          for (int i = K - 2; i >= 0; --i) x[i + TB] = x[i];
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->boxr[box]);
        if (warp == 4 && lane == 0) CB_TRACE(14 + round);
      }
    }

    // ---- epilogue: warp (q, c) owns rows [32q, +32) x columns [64c, +64) ----
    const int q = warp & 3, c = ww >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int row = row_g0 + q * 32 + lane;
    const bool ln = pr.ln_g != nullptr;
    const uint32_t stg0 = smem_base + static_cast<uint32_t>(ww) * kWarpStgBytes;             // H region is dead once acc2f fires
    const uint32_t stg1 = smem_base + static_cast<uint32_t>(kWorkers + ww) * kWarpStgBytes;
    // residual: this warp's [32 x 64] box of S goes straight into its staging box by TMA (one bulk copy instead of 8
    // loads per thread that each touch 32 different lines: ~4k cycles of L1 wavefronts per tile, ncu/trace r2g)
    if (lane == 0) {
      mbar_wait(&bars->stg_free, 0);
      mbar_arrive_expect_tx(&bars->resid_full[ww], kWarpStgBytes);
      tma_load_2d(smem + static_cast<uint32_t>(ww) * kWarpStgBytes, &tmO0, &bars->resid_full[ww], c * 64, row_g0 + q * 32);
    }
    (void)row;
    mbar_wait(&bars->acc2f, 0);
    tc_fence_after();
    if (warp == 4 && lane == 0) CB_TRACE(18);
    RowStats rs;
    uint4 rq[8];
    mbar_wait(&bars->resid_full[ww], 0);
    {
      const uint32_t rb = stg0 + static_cast<uint32_t>(lane) * 128u, xr = static_cast<uint32_t>(lane & 7);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(rq[j].x), "=r"(rq[j].y), "=r"(rq[j].z), "=r"(rq[j].w)
                     : "r"(rb + ((static_cast<uint32_t>(j) ^ xr) << 4))
                     : "memory");
    }
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      const int col = c * 64 + sub * 32;
      uint32_t raw[32];
      float v[32];
      tmem_ld32(tmem_base + lane_addr + col, raw);
      tmem_ld_wait();
      to_float(v, raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = *reinterpret_cast<const float4*>(cvec + col + 4 * j);
        fadd2(v[4 * j + 0], v[4 * j + 1], v[4 * j + 0], v[4 * j + 1], bb.x, bb.y);
        fadd2(v[4 * j + 2], v[4 * j + 3], v[4 * j + 2], v[4 * j + 3], bb.z, bb.w);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 r4 = rq[4 * sub + j];
        fadd2(v[8 * j + 0], v[8 * j + 1], v[8 * j + 0], v[8 * j + 1], bf16_lo(r4.x), bf16_hi(r4.x));
        fadd2(v[8 * j + 2], v[8 * j + 3], v[8 * j + 2], v[8 * j + 3], bf16_lo(r4.y), bf16_hi(r4.y));
        fadd2(v[8 * j + 4], v[8 * j + 5], v[8 * j + 4], v[8 * j + 5], bf16_lo(r4.z), bf16_hi(r4.z));
        fadd2(v[8 * j + 6], v[8 * j + 7], v[8 * j + 6], v[8 * j + 7], bf16_lo(r4.w), bf16_hi(r4.w));
      }
      if (ln) {
        rs.add(v);
        to_raw(raw, v);
        tmem_st32(tmem_base + lane_addr + col, raw);  // parked for the LayerNorm pass
      }
      stage_write<false>(stg0, lane, sub, v);
      if (warp == 4 && lane == 0 && sub == 1) CB_TRACE(31);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&tmO0, smem + static_cast<uint32_t>(ww) * kWarpStgBytes, c * 64, row_g0 + q * 32);
      tma_store_commit();
    }
    if (ln) {
      // row statistics of the four warps sharing this lane quarter
      float4* xch = reinterpret_cast<float4*>(part);  // [4 q][4 c][32] = 8 KB (aliases sums / tap staging / the dead mean: all consumed by now)
      tmem_st_wait();
      xch[(q * 4 + c) * 32 + lane] = make_float4(rs.s0, rs.s1, rs.q0, rs.q1);
      named_bar_sync(4 + q, 128);
      float s = 0.f, qq = 0.f;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const float4 o = xch[(q * 4 + cc) * 32 + lane];
        s += o.x + o.y;
        qq += o.z + o.w;
      }
      const float m = s * (1.f / kKD);
      const float var = fmaxf(qq * (1.f / kKD) - m * m, 0.f);
      const float rstd = rsqrtf(var + pr.ln_eps);
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        const int col = c * 64 + sub * 32;
        uint32_t raw[32];
        float v[32];
        tmem_ld32(tmem_base + lane_addr + col, raw);
        tmem_ld_wait();
        to_float(v, raw);
        {  // ((v - mean) * rstd) * gamma + beta, gamma / beta from shared memory
          const float nm = -m * rstd;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 gg = *reinterpret_cast<const float4*>(cvec + 256 + col + 4 * j);
            const float4 bb = *reinterpret_cast<const float4*>(cvec + 512 + col + 4 * j);
            float x0, x1, x2, x3;
            ffma2(x0, x1, v[4 * j + 0], v[4 * j + 1], rstd, rstd, nm, nm);
            ffma2(x2, x3, v[4 * j + 2], v[4 * j + 3], rstd, rstd, nm, nm);
            ffma2(v[4 * j + 0], v[4 * j + 1], x0, x1, gg.x, gg.y, bb.x, bb.y);
            ffma2(v[4 * j + 2], v[4 * j + 3], x2, x3, gg.z, gg.w, bb.z, bb.w);
          }
        }
        stage_write<false>(stg1, lane, sub, v);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmO1, smem + static_cast<uint32_t>(kWorkers + ww) * kWarpStgBytes, c * 64, row_g0 + q * 32);
        tma_store_commit();
      }
    }
    if (warp == 4 && lane == 0) CB_TRACE(19);
    if (lane == 0) tma_store_wait_all<0>();
    if (warp == 4 && lane == 0) CB_TRACE(22);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int K>
int launch_k(const Conv1dBlockPlan& p, cudaStream_t stream) {
  auto kern = conv1d_block_kernel<K>;
  static bool attr = false;
  if (!attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kCbSmem));
    attr = true;
  }
  CbParams pr;
  pr.bias_e = p.bias_e; pr.dw_w = p.dw_w; pr.dw_b = p.dw_b; pr.eca_w = p.eca_w; pr.bias_p = p.bias_p;
  pr.resid = p.resid; pr.ln_g = p.ln_g; pr.ln_b = p.ln_b; pr.ln_eps = p.ln_eps; pr.T = p.T;
  pr.dw_wsum = p.dw_wsum; pr.wbits = p.wbits; pr.valid_cnt = p.valid_cnt;
  pr.trace = nullptr; pr.trace_b = 0;
#ifdef ISHARA_C1B_TRACE
  static long long* tbuf = nullptr;
  static int printed = 0;
  const int want = getenv("ISHARA_C1B_TRACE_N") ? atoi(getenv("ISHARA_C1B_TRACE_N")) : 6;
  const bool tracing = printed < want && p.B >= 8;
  if (tracing) {
    if (tbuf == nullptr) ISHARA_CUDA_OK(cudaMalloc(&tbuf, 8 * 32 * sizeof(long long)));
    ISHARA_CUDA_OK(cudaMemsetAsync(tbuf, 0, 8 * 32 * sizeof(long long), stream));
    pr.trace = tbuf;
    pr.trace_b = p.B * 3 / 5;  // a sequence from the middle of the launch
  }
#endif
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.T / kBM, p.B, 1);
  cfg.blockDim = dim3(kCbThreads, 1, 1);
  cfg.dynamicSmemBytes = kCbSmem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.T / kBM;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  ISHARA_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p.tmX, p.tmWe, p.tmWp, p.tmO0, p.tmO1, pr));
  note_launch();
#ifdef ISHARA_C1B_TRACE
  if (tracing) {
    ISHARA_CUDA_OK(cudaStreamSynchronize(stream));
    long long h[8 * 32];
    ISHARA_CUDA_OK(cudaMemcpy(h, tbuf, sizeof(h), cudaMemcpyDeviceToHost));
    ++printed;
    fprintf(stderr, "c1b trace K=%d B=%d seq=%d ln=%d (cycles since CTA start): 0 start | 1,2,3 expand MMA issued: first, half0, half1 | "
                    "4/5 drain0 begin/end | 6/7 drain1 begin/end | 8 H complete | 9 sums | 10 cluster#1 | 11 halo+mean | 12 scale | 13 cluster#2 | "
                    "14-17 stencil rounds | 20/21 project MMA box0/box7 | 18 acc2 ready | 19 epilogue issued | 22 stores drained\n",
            K, p.B, pr.trace_b, p.ln_g != nullptr);
    for (int r = 0; r < p.T / kBM; ++r) {
      fprintf(stderr, "  rank %d:", r);
      const int order[] = {1, 2, 3, 4, 5, 6, 7, 8, 25, 26, 27, 9, 10, 11, 28, 29, 12, 13, 14, 15, 16, 17, 20, 30, 23, 21, 24, 18, 31, 19, 22};
      for (int e : order) fprintf(stderr, " %d:%lld", e, h[r * 32 + e] ? h[r * 32 + e] - h[r * 32] : -1);
      fprintf(stderr, "\n");
    }
  }
#endif
  return 0;
}

}  // namespace

bool conv1d_block_applicable(int D, int T, int k) {
  return D == kKD && T % kBM == 0 && T / kBM >= 1 && T / kBM <= 8 && (k == 3 || k == 5 || k == 7 || k == 9 || k == 11);
}

// x = S [B*T, 256] (read as the expand operand and as the residual, overwritten in place); wet = We^T [512, 256];
// wpt = Wp^T [256, 512]; xn = LN(S) output or null
int conv1d_block_plan_init(Conv1dBlockPlan* p, bf16* s, const bf16* wet, const bf16* wpt, bf16* xn) {
  static_assert(kCbSmem <= kMaxSmem, "conv1d_block: shared memory budget");
  static_assert(120 * 1024 + kXSlot <= kHRegion + kWSlot, "expand slot 2 must end before Wp slot 1");
  static_assert(kXSlot <= 4 * kBoxStride && 72 * 1024 == 4 * kBoxStride, "expand slots 1-2 must not overlap boxes 0-3");
  const uint64_t M = static_cast<uint64_t>(p->B) * p->T;
  int rc;
  if ((rc = make_tmap_2d(&p->tmX, s, TM_BF16, M, kKD, kKD, kBM, kBK))) return rc;
  if ((rc = make_tmap_2d(&p->tmWe, wet, TM_BF16, kKC, kKD, kKD, 256, kBK))) return rc;
  if ((rc = make_tmap_2d(&p->tmWp, wpt, TM_BF16, kKD, kKC, kKC, 256, kBK))) return rc;
  if ((rc = make_tmap_2d(&p->tmO0, s, TM_BF16, M, kKD, kKD, 32, 64))) return rc;
  if (xn != nullptr) {
    if ((rc = make_tmap_2d(&p->tmO1, xn, TM_BF16, M, kKD, kKD, 32, 64))) return rc;
  } else {
    p->tmO1 = p->tmO0;
  }
  p->resid = s;
  return 0;
}

int conv1d_block_launch(const Conv1dBlockPlan& p, cudaStream_t stream) {
  switch (p.k) {
    case 3: return launch_k<3>(p, stream);
    case 5: return launch_k<5>(p, stream);
    case 7: return launch_k<7>(p, stream);
    case 9: return launch_k<9>(p, stream);
    case 11: return launch_k<11>(p, stream);
  }
  set_last_error("conv1d_block: unsupported kernel size");
  return 2;
}

}  // namespace ishara
