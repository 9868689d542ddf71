// ishara_b200 — tcgen05/TMEM GEMM with fused epilogues (sm_100a only).
//
// Serves every dense contraction on the get_model hot path (SURVEY.md §8a T2,T3,T6-T12): Dense /
// 1x1-conv layers of the reference (nb:conv-hybrid-model c5:61-65,77-80,97-99,140-142,163-165,
// 258-278; c7:14,61-63) become  out = epilogue(A[M,K] @ Wt[N,K]^T)  with M = B*T rows.
//
// Structure (one persistent CTA per SM, 384 threads):
//   warp 0      TMA producer   : A tile 128x64 and W tile BNx64 (bf16, 128B swizzle) -> smem ring
//   warp 1      MMA issuer     : tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, fp32 accum in TMEM
//   warp 2      TMEM allocator
//   warps 4-7   epilogue group 0  (even tiles, TMEM buffer 0)
//   warps 8-11  epilogue group 1  (odd tiles,  TMEM buffer 1)
// The accumulator is double buffered in TMEM, so the epilogue of tile i overlaps the MMAs of tile
// i+1. Each epilogue thread owns one output row (tcgen05.ld 32x32b: lane == row), which makes the
// LayerNorm statistics of a full 256-wide row thread-local: the residual add and up to two chained
// LayerNorms are fused here and the normalised copy for the next GEMM is produced in the same
// pass (values parked in TMEM between passes with tcgen05.st). Output leaves through swizzled
// staging tiles and TMA stores.
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

constexpr int kNumStg = 4;            // 2 per epilogue group

__host__ __device__ constexpr int gemm_stages(int bn) { return bn >= 256 ? 3 : 4; }
__host__ __device__ constexpr int gemm_stage_bytes(int bn) { return kAStageBytes + bn * kBK * 2; }
__host__ __device__ constexpr int gemm_smem_bytes(int bn) {
  return gemm_stages(bn) * gemm_stage_bytes(bn) + kNumStg * kStgBytes + kXchBytes + 256 /*barriers*/ + 1024 /*align slack*/;
}

// RESB ("resident B", weight-stationary): the CTA keeps its whole [BN x K] weight slice in shared memory for the
// lifetime of the kernel and streams only A tiles through the ring, so the L2->SM traffic per tile drops from
// A + B to A alone (ncu on the streaming variant: MMA issue waits on the TMA ring, tensor pipe 22 % active). Each CTA
// is pinned to one n-tile; m-tiles are strided over the CTAs of that n-tile.
// EW = number of epilogue warps: 8 (two per TMEM lane quarter) or 16 (four per quarter, one 64-column box each — the
// plain BN = 256 GEMMs are epilogue-paced, so their epilogue gets twice the warps; 640 threads cap registers at 96).
template <int BN, bool ROW, bool OUT_F32, bool RESB, int EW>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
               const __grid_constant__ CUtensorMap tmR, const GemmEpi ep, int M, int N, int K, int num_m_tiles, int num_n_tiles,
               int num_stages, int resid_tma) {
  constexpr int B_KB_BYTES = BN * kBK * 2;                              // one k-block of the weight slice
  constexpr int STAGE_BYTES = RESB ? kAStageBytes : gemm_stage_bytes(BN);
  constexpr int NUM_STG = RESB ? 2 : kNumStg;                           // staging tiles (per group: NUM_STG / 2)
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  constexpr uint32_t IDESC = umma_idesc(kBM, BN, /*bf16*/ 1);
  const int STAGES = num_stages;
  const int num_kb = K / kBK;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t resb_bytes = RESB ? static_cast<uint32_t>(num_kb) * B_KB_BYTES : 0u;
  uint8_t* resb_ptr = smem;                                             // [num_kb][BN x 64] (RESB only)
  uint8_t* stage_ptr = smem + resb_bytes;
  uint8_t* stg_ptr = stage_ptr + STAGES * STAGE_BYTES;
  float4* xch = reinterpret_cast<float4*>(stg_ptr + NUM_STG * kStgBytes);   // row-statistics exchange slots
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_ptr + NUM_STG * kStgBytes + kXchBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint64_t* bfull = bars + 2 * STAGES + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);
  uint64_t* rbar = bars + 2 * STAGES + 6;  // [EW] per epilogue warp: residual boxes landed (row mode, resid_tma)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = num_m_tiles * num_n_tiles;
  // tile walk: streaming variant strides over all (m, n) tiles; RESB pins the CTA to n-tile (blockIdx.x % num_n_tiles)
  // and strides m-tiles by the CTAs sharing that n-tile (gridDim.x is a multiple of num_n_tiles).
  const int tile_begin = RESB ? (blockIdx.x / num_n_tiles) * num_n_tiles + (blockIdx.x % num_n_tiles) : blockIdx.x;
  const int tile_step = gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO0);
    if (ep.ln1_g != nullptr) tma_prefetch_desc(&tmO1);
    if (resid_tma) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], EW);  // one arrive per epilogue warp
    mbar_init(&tempty[1], EW);
    mbar_init(bfull, 1);
    if (ROW) for (int w = 0; w < EW; ++w) mbar_init(&rbar[w], 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      if constexpr (RESB) {
        if (tile_begin < num_tiles) {
          const int n_tile = tile_begin % num_n_tiles;
          mbar_arrive_expect_tx(bfull, resb_bytes);
          for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(resb_ptr + kb * B_KB_BYTES, &tmB, bfull, kb * kBK, n_tile * BN);
        }
      }
      int pit = 0;
      for (int tile = tile_begin; tile < num_tiles; tile += tile_step, ++pit) {
        const int m_tile = tile / num_n_tiles, n_tile = tile % num_n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          if (kb == 0) ISHARA_TRACE(pit, 0);
          if (kb == num_kb - 1) ISHARA_TRACE(pit, 1);
          mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* sa = stage_ptr + stage * STAGE_BYTES;
          tma_load_2d(sa, &tmA, &full[stage], kb * kBK, m_tile * kBM);
          if constexpr (!RESB) tma_load_2d(sa + kAStageBytes, &tmB, &full[stage], kb * kBK, n_tile * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int it = 0;
      if constexpr (RESB) {
        if (tile_begin < num_tiles) mbar_wait(bfull, 0);
      }
      for (int tile = tile_begin; tile < num_tiles; tile += tile_step, ++it) {
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
          const uint32_t sa = smem_base + resb_bytes + stage * STAGE_BYTES;
          const uint32_t sb = RESB ? smem_base + kb * B_KB_BYTES : sa + kAStageBytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            if (ISHARA_DBG_BIT(ep, 4)) break;
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row
            umma_bf16(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), IDESC,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees the smem slot when these MMAs retire
          if (kb == num_kb - 1) { umma_commit(&tfull[buf]); ISHARA_TRACE(it, 5); }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: all 8 warps on every tile =====================
    const int q = warp & 3;         // TMEM lane quarter this warp may access
    const int h = (warp - 4) >> 2;  // which alternate 64-column boxes of the tile this warp owns
    WarpStore st;
    st.single = NUM_STG == 2 || EW == 16;  // 16 warps: one staging box each (it is reused a whole tile later)
    st.base = smem_u32(stg_ptr) + static_cast<uint32_t>((warp - 4) * (st.single ? 1 : 2)) * kWarpStgBytes;
    st.iter = 0;
    st.lane = lane;
    st.skip_store = ISHARA_DBG_BIT(ep, 1);
    st.skip_fence = ISHARA_DBG_BIT(ep, 16);

    // 16-warp full-row epilogue: per-column vectors live in shared memory (the xch region: 8 KB exchange + 5 KB vectors)
    float2* xch2 = reinterpret_cast<float2*>(xch);
    float* cvec = reinterpret_cast<float*>(xch) + 2048;
    Row16State rst;
    if constexpr (ROW && EW == 16) {
      const int t16 = static_cast<int>(threadIdx.x) - 128;
      for (int i = t16; i < 5 * 256; i += 512) {
        const int which = i >> 8, col = i & 255;
        const float* src = which == 0 ? ep.bias : which == 1 ? ep.ln0_g : which == 2 ? ep.ln0_b : which == 3 ? ep.ln1_g : ep.ln1_b;
        cvec[i] = src != nullptr ? __ldg(src + col) : 0.f;
      }
      named_bar_sync(5, 512);
    }

    int it = 0;
    for (int tile = tile_begin; tile < num_tiles; tile += tile_step, ++it) {
      const uint32_t buf = it & 1;
      const int m_tile = tile / num_n_tiles, n_tile = tile % num_n_tiles;
      const int row0 = m_tile * kBM + q * 32;  // first row of this warp's boxes
      EpiThread th;
      th.row = row0 + lane;
      th.valid = th.row < M;
      th.seq = th.valid ? th.row / ep.rows_per_seq : 0;
      th.t = th.valid ? th.row - th.seq * ep.rows_per_seq : 0;
      th.taddr = tmem_base + buf * BN + (static_cast<uint32_t>(q * 32) << 16);
      ResidRegs rr;
      uint32_t rsm = 0u;
      if constexpr (ROW && EW == 8) {
        if (resid_tma) {
          // residual of this warp's output boxes: TMA into its own staging boxes (slot of the k-th acquire of this tile),
          // once the stores of the previous tile have finished reading them; lands while the MMAs of this tile run
          rsm = st.base;
          if (lane == 0) {
            tma_store_wait_read<0>();
            constexpr int NB = BN / 128;  // boxes per warp per tile
            mbar_arrive_expect_tx(&rbar[warp - 4], NB * kWarpStgBytes);
#pragma unroll
            for (int k = 0; k < NB; ++k)
              tma_load_2d(stg_ptr + static_cast<uint32_t>((warp - 4) * 2) * kWarpStgBytes + ((st.iter + k) & 1u) * kWarpStgBytes, &tmR,
                          &rbar[warp - 4], (h + 2 * k) * 64, row0);
          }
          __syncwarp();
        }
      }
      bool r16_tma = false;
      if constexpr (ROW && EW == 16) {
        if (resid_tma && ep.resid != nullptr) {
          // this warp's residual box: TMA into its staging box once the previous tile's last store has read it
          r16_tma = true;
          if (lane == 0) {
            tma_store_wait_read<0>();
            mbar_arrive_expect_tx(&rbar[warp - 4], kWarpStgBytes);
            tma_load_2d(stg_ptr + static_cast<uint32_t>(warp - 4) * kWarpStgBytes, &tmR, &rbar[warp - 4], h * 64, row0);
          }
          __syncwarp();
        }
      }
      if constexpr (EW != 16) {
        if (rsm == 0u) resid_load(rr, ep, th, epilogue_first_col<BN, ROW, OUT_F32>(ep, n_tile, h));  // in flight while the MMAs finish
      }

      mbar_wait(&tfull[buf], (it >> 1) & 1);
      tc_fence_after();
      if (q == 0 && h == 0 && lane == 0) ISHARA_TRACE(it, 6);
      if (!ISHARA_DBG_BIT(ep, 2)) {
        if constexpr (EW == 16 && ROW) {
          epilogue_row16(ep, th, row0, q, h, lane, st, &tmO0, &tmO1, xch2, cvec, rst, r16_tma ? &rbar[warp - 4] : nullptr,
                         static_cast<uint32_t>(it & 1));
        } else if constexpr (EW == 16) {
          epilogue_box16<BN>(ep, th, n_tile, N, row0, h, lane, st, &tmO0);
        } else {
          if constexpr (ROW && EW == 8) {
            if (rsm != 0u) mbar_wait(&rbar[warp - 4], static_cast<uint32_t>(it & 1));
          }
          epilogue_tile<BN, ROW, OUT_F32>(ep, th, n_tile, N, row0, q, h, lane, st, &tmO0, &tmO1, xch, it & 1, rr, rsm);
        }
      }
      // this warp is done with accumulator buffer `buf`
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
      if (q == 0 && h == 0 && lane == 0) ISHARA_TRACE(it, 7);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int BN, bool ROW, bool OUT_F32, bool RESB, int EW = 8>
int launch_inst(const GemmPlan& p, int num_sms, cudaStream_t stream) {
  auto kern = gemm_tc_kernel<BN, ROW, OUT_F32, RESB, EW>;
  static int attr_smem = 0;
  const int mt = (p.M + kBM - 1) / kBM;
  const int nt = p.N / BN;
  int stages, smem;
  if (RESB) {
    const int fixed = (p.K / kBK) * BN * kBK * 2 + 2 * kStgBytes + kXchBytes + 256 + 1024;
    stages = (kMaxSmem - fixed) / kAStageBytes;
    if (stages > 8) stages = 8;
    smem = fixed + stages * kAStageBytes;
  } else {
    stages = gemm_stages(BN);
    smem = gemm_smem_bytes(BN);
  }
  if (smem > attr_smem) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  int grid = mt * nt;
  if (grid > num_sms) grid = num_sms;
  if (RESB) grid = grid / nt * nt;  // every CTA owns exactly one n-tile
  static const int rtma_env = getenv("ISHARA_GEMM_RESID_TMA") ? atoi(getenv("ISHARA_GEMM_RESID_TMA")) : 1;
  const int rtma = (ROW && !RESB && p.resid_tma && rtma_env) ? 1 : 0;
  kern<<<grid, 128 + 32 * EW, smem, stream>>>(p.tmA, p.tmB, p.tmO0, p.tmO1, p.tmR, p.epi, p.M, p.N, p.K, mt, nt, stages, rtma);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

// weight-stationary variant is used when the [BN x K] slice plus >= 3 A stages fit in shared memory
bool resident_fits(int bn, int K) {
  const int fixed = (K / kBK) * bn * kBK * 2 + 2 * kStgBytes + kXchBytes + 256 + 1024;
  return fixed + 3 * kAStageBytes <= kMaxSmem;
}

}  // namespace

bool gemm2_applicable(const GemmPlan& p, int num_sms);
int gemm_launch_inner(const GemmPlan& p, int num_sms, cudaStream_t stream);
int gemm2_launch(const GemmPlan& p, int num_sms, cudaStream_t stream);

int make_tmap_2d(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled not available from the driver");
    return 3;
  }
  const uint32_t es = dt == TM_BF16 ? 2 : 4;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld_elems * es};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const uint32_t inner = box_cols * es;
  CUtensorMapSwizzle sw = inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : inner == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                        : CU_TENSOR_MAP_SWIZZLE_NONE;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (gstr[0] & 15) != 0) {
    set_last_error("make_tmap_2d: base/pitch not 16-byte aligned");
    return 2;
  }
  CUresult r = fn(out, dt == TM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed, CUresult " + std::to_string(static_cast<int>(r)) + " rows " +
                   std::to_string(rows) + " cols " + std::to_string(cols) + " ld " + std::to_string(ld_elems) +
                   " box " + std::to_string(box_rows) + "x" + std::to_string(box_cols));
    return 3;
  }
  return 0;
}

int gemm_plan_init(GemmPlan* p, const bf16* A, int lda, const bf16* Wt, void* out0, int ldo0, int nout, bf16* out1,
                   int ldo1) {
  if (p->K % kBK != 0 || p->N % p->block_n != 0) {
    set_last_error("gemm_plan_init: K must be a multiple of 64 and N a multiple of block_n");
    return 2;
  }
  if (p->row_mode && p->N != p->block_n) {
    set_last_error("gemm_plan_init: row_mode needs N == block_n");
    return 2;
  }
  int rc;
  if ((rc = make_tmap_2d(&p->tmA, A, TM_BF16, p->M, p->K, lda, kBM, kBK))) return rc;
  if ((rc = make_tmap_2d(&p->tmB, Wt, TM_BF16, p->N, p->K, p->K, p->block_n, kBK))) return rc;
  if (p->block_n == 256)  // CTA-pair variant: each CTA loads its 128-row half of the weight tile
    if ((rc = make_tmap_2d(&p->tmBh, Wt, TM_BF16, p->N, p->K, p->K, 128, kBK))) return rc;
  if (p->out_f32) {
    if ((rc = make_tmap_2d(&p->tmO0, out0, TM_F32, p->M, nout, ldo0, 32, 32))) return rc;
  } else {
    if ((rc = make_tmap_2d(&p->tmO0, out0, TM_BF16, p->M, nout, ldo0, 32, 64))) return rc;
  }
  if (out1 != nullptr) {
    if ((rc = make_tmap_2d(&p->tmO1, out1, TM_BF16, p->M, nout, ldo1, 32, 64))) return rc;
  } else {
    p->tmO1 = p->tmO0;
  }
  p->tmR = p->tmO0;
  p->resid_tma = false;
  if (p->row_mode && !p->out_f32 && p->epi.resid != nullptr && p->epi.ld_resid % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(p->epi.resid) & 15) == 0) {
    // full-row epilogue: the residual tile is fetched by TMA ([32 x 64] boxes) instead of per-thread row reads
    if ((rc = make_tmap_2d(&p->tmR, p->epi.resid, TM_BF16, p->M, nout, p->epi.ld_resid, 32, 64))) return rc;
    p->resid_tma = true;
  }
  return 0;
}

int gemm_launch(const GemmPlan& p_in, int num_sms, cudaStream_t stream) {
#ifdef ISHARA_TRACE_BUILD
  static const int dbg = getenv("ISHARA_GEMM_DBG") ? atoi(getenv("ISHARA_GEMM_DBG")) : 0;
#else
  static const int dbg = 0;  // bisect bits and tracing are compiled out of the shipped library
#endif
  static const int force = getenv("ISHARA_GEMM_RESIDENT") ? atoi(getenv("ISHARA_GEMM_RESIDENT")) : 0;  // 1 = wherever it fits, -1 = never, 0 = qkv only
  static const int pair = getenv("ISHARA_GEMM_PAIR") ? atoi(getenv("ISHARA_GEMM_PAIR")) : 0;  // CTA-pair variant: opt-in (measured slower)
  GemmPlan p = p_in;
  p.epi.dbg = dbg;
  // weight-stationary only where it measured faster: plain GEMMs with three or more n-tiles (qkv, N = 768: 55 -> 51 us;
  // the one- and two-tile GEMMs lose more on the 8-warp epilogue the variant is limited to than they gain on the ring)
  if (force < 0 || (force == 0 && !(!p.row_mode && !p.out_f32 && p.block_n == 256 && p.N >= 3 * 256))) p.no_resident = true;
#ifdef ISHARA_TRACE_BUILD
  static const int trace = getenv("ISHARA_GEMM_TRACE") ? atoi(getenv("ISHARA_GEMM_TRACE")) : 0;
#else
  static const int trace = 0;
#endif
  if (trace > 0) {
    // debugging aid: timeline of CTA 0 (clock64 at 8 events per tile), printed to stderr after a blocking launch
    static long long* dbuf = nullptr;
    const int ntile = 64;
    if (dbuf == nullptr) ISHARA_CUDA_OK(cudaMalloc(&dbuf, ntile * 8 * sizeof(long long)));
    ISHARA_CUDA_OK(cudaMemset(dbuf, 0, ntile * 8 * sizeof(long long)));
    p.epi.trace = dbuf;
    int rc;
    if (pair != 0 && !p.no_pair && gemm2_applicable(p, num_sms)) {
      p.tmB = p.tmBh;
      rc = gemm2_launch(p, num_sms, stream);
    } else {
      p.no_pair = true;
      GemmPlan q = p;
      q.epi.trace = dbuf;
      static const int depth = 0;
      (void)depth;
      rc = gemm_launch_inner(q, num_sms, stream);
    }
    if (rc) return rc;
    ISHARA_CUDA_OK(cudaDeviceSynchronize());
    static int printed = 0;
    if (printed++ < trace) {
      long long h[64 * 8];
      ISHARA_CUDA_OK(cudaMemcpy(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost));
      long long t0 = 0;
      for (int i = 0; i < 8; ++i) if (h[i] != 0 && (t0 == 0 || h[i] < t0)) t0 = h[i];
      fprintf(stderr, "gemm trace M=%d N=%d K=%d row=%d (cycles since first event; P0 first load issue, P1 last load issue, "
                      "M2 tempty ok, M3 first full ok, M4 last full ok, M5 tfull commit, E6 tfull seen, E7 epilogue done)\n",
              p.M, p.N, p.K, (int)p.row_mode);
      for (int t = 0; t < 14; ++t) {
        fprintf(stderr, "  tile %2d:", t);
        for (int e = 0; e < 8; ++e) fprintf(stderr, " %8lld", h[t * 8 + e] ? h[t * 8 + e] - t0 : -1);
        fprintf(stderr, "\n");
      }
    }
    return 0;
  }
  if (pair != 0 && !p.no_pair && gemm2_applicable(p, num_sms)) {
    p.tmB = p.tmBh;
    return gemm2_launch(p, num_sms, stream);
  }
  return gemm_launch_inner(p, num_sms, stream);
}

int gemm_launch_inner(const GemmPlan& p, int num_sms, cudaStream_t stream) {
  const bool res = !p.no_resident && resident_fits(p.block_n, p.K) && (p.M + kBM - 1) / kBM * (p.N / p.block_n) >= num_sms;
  if (p.row_mode) {
    if (p.block_n == 256 && !p.out_f32) {
      static const int row16 = getenv("ISHARA_GEMM_ROW16") ? atoi(getenv("ISHARA_GEMM_ROW16")) : 1;  // 16-warp full-row epilogue
      if (res) return launch_inst<256, true, false, true>(p, num_sms, stream);
      // (the stem's per-row positional table is read per thread: measured 55 us with 16 warps vs 49 us with 8)
      return row16 && p.epi.rowtab == nullptr ? launch_inst<256, true, false, false, 16>(p, num_sms, stream)
                                              : launch_inst<256, true, false, false>(p, num_sms, stream);
    }
    if (p.block_n == 128 && !p.out_f32) return launch_inst<128, true, false, false>(p, num_sms, stream);
  } else {
    if (p.block_n == 256 && !p.out_f32) {
      static const int ew16 = getenv("ISHARA_GEMM_EW16") ? atoi(getenv("ISHARA_GEMM_EW16")) : 1;
      if (res) return launch_inst<256, false, false, true>(p, num_sms, stream);
      return ew16 ? launch_inst<256, false, false, false, 16>(p, num_sms, stream)
                  : launch_inst<256, false, false, false>(p, num_sms, stream);
    }
    if (p.block_n == 128 && !p.out_f32) return launch_inst<128, false, false, false>(p, num_sms, stream);
    if (p.block_n == 64 && !p.out_f32) return launch_inst<64, false, false, false>(p, num_sms, stream);
    if (p.block_n == 64 && p.out_f32) return launch_inst<64, false, true, false>(p, num_sms, stream);
  }
  set_last_error("gemm_launch: unsupported (block_n, row_mode, out_f32) combination");
  return 2;
}

}  // namespace ishara
