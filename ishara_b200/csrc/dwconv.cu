// ishara_b200 — depthwise temporal Conv1D over a whole sequence per CTA (memory-bound, sm_100a).
//
// Covers the three depthwise convolutions of the get_model path (SURVEY.md §8a T4,T5,T8,T11):
//   * Conv1DBlock:   CausalDWConv1D(k=11/5/3) -> BatchNorm -> ECA      (nb:conv-hybrid-model c5:17-39,73,1-15)
//   * Squeezeformer: CausalDWConv1D(k=15) -> swish, + column sums for SqueezeExcite (c5:141,149-150,120-133)
//   * Conformer:     Conv1D(k=15,'same',groups=D,+bias) -> BatchNorm   (c5:265-271,298-301)
// BatchNorm (inference) is folded into the taps/bias by the caller.
//
// Layout: activations are [B, T, C] bf16 channels-last. One CTA owns (sequence b, 64-channel slab):
// the T x 64 tile (128 B per row) is staged once in shared memory with its zero halo rows, every
// lane owns one bf16x2 channel pair (conflict-free 4-byte smem reads, 128-byte coalesced stores) and
// slides a register window down its time range. Because the whole T axis of the slab lives in the
// CTA, the global-average-pool that ECA needs is a CTA-local reduction; the 5-tap conv across the
// CHANNEL axis needs two channels from each neighbouring slab, which are exchanged through
// distributed shared memory: the C/64 CTAs of one sequence form one thread-block cluster.
#include <cooperative_groups.h>

#include "kernels.h"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace ishara {
namespace {

constexpr int kSlab = 64;    // channels per CTA
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTBmax = 8;    // outputs per register block (small kernels); wide kernels use 4 to stay under 64 registers

__device__ __forceinline__ float exact_sigmoid(float x) { return 1.f / (1.f + __expf(-x)); }

template <int K>
__host__ __device__ constexpr int dw_tb() { return K >= 9 ? 4 : kTBmax; }

template <int K, int POST>
__global__ void __launch_bounds__(kThreads, (K <= 11 ? 4 : (K <= 15 ? 3 : 1)))
dwconv_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, const float* __restrict__ w,
              const float* __restrict__ bias, const float* __restrict__ eca_w, float* __restrict__ colsum, int T,
              int C, int pad_left, int rows_alloc, const uint8_t* __restrict__ key_mask, const int32_t* __restrict__ valid_cnt) {
  constexpr int kTB = dw_tb<K>();
  extern __shared__ __align__(16) uint8_t smem_dw[];
  uint32_t* tile = reinterpret_cast<uint32_t*>(smem_dw);           // [rows_alloc][32] bf16x2
  float* red = reinterpret_cast<float*>(tile + rows_alloc * 32);   // [kWarps][64]
  float* mean_s = red + kWarps * kSlab;                            // [64] (+ scale reuse)

  const int b = blockIdx.y;
  const int c0 = blockIdx.x * kSlab;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- stage the slab: smem row r holds input row (r - pad_left); zero outside [0, T) ----
  const bf16* src = in + (static_cast<size_t>(b) * T) * C + c0;
  for (int i = tid; i < rows_alloc * 8; i += kThreads) {
    const int r = i >> 3, ch = i & 7;
    const int t = r - pad_left;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (t >= 0 && t < T) v = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(t) * C) + ch);
    reinterpret_cast<uint4*>(tile)[i] = v;
  }

  // taps for this lane's channel pair
  float2 wt[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wt[j] = __ldg(reinterpret_cast<const float2*>(w + static_cast<size_t>(j) * C + c0) + lane);
  float2 bs = make_float2(0.f, 0.f);
  if (bias != nullptr) bs = __ldg(reinterpret_cast<const float2*>(bias + c0) + lane);
  if constexpr (POST == 1) {
    // swish(y) = h + h tanh(h) with h = y / 2: the halving is folded into the taps, so the activation costs one MUFU op
    // and one FMA per element (the ex2 + IEEE division form was a third of this kernel's issue slots)
#pragma unroll
    for (int j = 0; j < K; ++j) { wt[j].x *= 0.5f; wt[j].y *= 0.5f; }
    bs.x *= 0.5f; bs.y *= 0.5f;
  }
  __syncthreads();

  const int rows_per_warp = ((T + kWarps - 1) / kWarps + kTB - 1) / kTB * kTB;
  const int t_begin = warp * rows_per_warp;
  const int t_end = min(T, t_begin + rows_per_warp);

  float2 scale = make_float2(1.f, 1.f);
  cg::cluster_group cluster = cg::this_cluster();
  const uint8_t* mrow = key_mask != nullptr ? key_mask + static_cast<size_t>(b) * T : nullptr;  // 1 = frame counts (propagated mask)
  if constexpr (POST == 2) {
    if (mrow != nullptr) {
      // ---- ECA with a propagated Keras mask (c5:8-9): GlobalAveragePooling1D(mask) = mean of y over the VALID frames.
      //      Arbitrary masks: one explicit conv pass over the valid frames (this is the non-default mode; the dense
      //      case below needs no conv pass at all) ----
      float2 s = make_float2(0.f, 0.f);
      for (int t = t_begin; t < t_end; ++t) {
        if (mrow[t] == 0) continue;
        float2 a = bs;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const uint32_t u = tile[(t + j) * 32 + lane];
          ffma2(a.x, a.y, wt[j].x, wt[j].y, bf16_lo(u), bf16_hi(u), a.x, a.y);
        }
        s.x += a.x; s.y += a.y;
      }
      red[warp * kSlab + 2 * lane] = s.x;
      red[warp * kSlab + 2 * lane + 1] = s.y;
      __syncthreads();
      if (warp == 0) {
        float2 S = make_float2(0.f, 0.f);
#pragma unroll
        for (int wv = 0; wv < kWarps; ++wv) {
          S.x += red[wv * kSlab + 2 * lane];
          S.y += red[wv * kSlab + 2 * lane + 1];
        }
        const float invn = 1.f / static_cast<float>(valid_cnt[b]);  // 0 valid frames: 0/0 = NaN, as in Keras
        mean_s[2 * lane] = S.x * invn;
        mean_s[2 * lane + 1] = S.y * invn;
      }
    } else {
    // ---- ECA: mean over T of y = conv + bias, then 5-tap conv over channels, sigmoid ----
    // The convolution is linear in time, so its mean needs no convolution pass: with S = sum_t x[t],
    //   sum_t y[t] = T*bias + sum_j w[j] * (S - [rows that tap j shifts out of the window]),
    // i.e. one add per staged element plus a head/tail correction of at most K-1 rows (ncu: the two-pass version was
    // issue-bound, 0.6 IPC per scheduler with DRAM at 15 %).
    float2 s = make_float2(0.f, 0.f);
    for (int t = t_begin; t < t_end; ++t) {
      const uint32_t u = tile[(t + pad_left) * 32 + lane];
      fadd2(s.x, s.y, s.x, s.y, bf16_lo(u), bf16_hi(u));
    }
    red[warp * kSlab + 2 * lane] = s.x;
    red[warp * kSlab + 2 * lane + 1] = s.y;
    __syncthreads();
    if (warp == 0) {
      float2 S = make_float2(0.f, 0.f);
#pragma unroll
      for (int wv = 0; wv < kWarps; ++wv) {
        S.x += red[wv * kSlab + 2 * lane];
        S.y += red[wv * kSlab + 2 * lane + 1];
      }
      float2 acc = make_float2(0.f, 0.f);
      // taps with offset o = j - pad_left < 0 lose the last |o| rows, taps with o > 0 lose the first o rows
      float2 tail = make_float2(0.f, 0.f);
      for (int m = 1; m <= pad_left; ++m) {  // tap j = pad_left - m
        const uint32_t u = tile[(T - m + pad_left) * 32 + lane];
        tail.x += bf16_lo(u); tail.y += bf16_hi(u);
        const float2 wj = __ldg(reinterpret_cast<const float2*>(w + static_cast<size_t>(pad_left - m) * C + c0) + lane);
        acc.x = fmaf(wj.x, S.x - tail.x, acc.x); acc.y = fmaf(wj.y, S.y - tail.y, acc.y);
      }
      {
        const float2 wj = __ldg(reinterpret_cast<const float2*>(w + static_cast<size_t>(pad_left) * C + c0) + lane);
        acc.x = fmaf(wj.x, S.x, acc.x); acc.y = fmaf(wj.y, S.y, acc.y);
      }
      float2 head = make_float2(0.f, 0.f);
      for (int m = 1; m < K - pad_left; ++m) {  // tap j = pad_left + m
        const uint32_t u = tile[(m - 1 + pad_left) * 32 + lane];
        head.x += bf16_lo(u); head.y += bf16_hi(u);
        const float2 wj = __ldg(reinterpret_cast<const float2*>(w + static_cast<size_t>(pad_left + m) * C + c0) + lane);
        acc.x = fmaf(wj.x, S.x - head.x, acc.x); acc.y = fmaf(wj.y, S.y - head.y, acc.y);
      }
      const float invT = 1.f / static_cast<float>(T);
      mean_s[2 * lane] = fmaf(acc.x, invT, bs.x);
      mean_s[2 * lane + 1] = fmaf(acc.y, invT, bs.y);
    }
    }
    // exchange slab-edge means with the neighbouring slabs of the same sequence (DSMEM)
    cluster.sync();
    if (tid < kSlab) {
      const int rank = static_cast<int>(cluster.block_rank());
      const int nrank = static_cast<int>(cluster.num_blocks());
      float acc = 0.f;
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int c = tid + d;  // channel within this slab, may spill into a neighbour
        float mv = 0.f;
        if (c >= 0 && c < kSlab) {
          mv = mean_s[c];
        } else if (c < 0 && rank > 0) {
          mv = cluster.map_shared_rank(mean_s, rank - 1)[c + kSlab];
        } else if (c >= kSlab && rank + 1 < nrank) {
          mv = cluster.map_shared_rank(mean_s, rank + 1)[c - kSlab];
        }
        acc = fmaf(__ldg(eca_w + d + 2), mv, acc);
      }
      red[tid] = exact_sigmoid(acc);  // red[0..63] reused for the per-channel scale
    }
    // neighbours may still be reading mean_s: arrive now, wait only at the very end of the kernel (nothing below
    // touches mean_s, the wait merely keeps this CTA's shared memory alive)
    cluster.barrier_arrive();
    __syncthreads();  // red[] visible to the CTA
    scale = make_float2(red[2 * lane], red[2 * lane + 1]);
  }

  // ---- main pass (packed fp32x2 FMAs: one issue slot per tap for the lane's channel pair) ----
  // Pointers advance incrementally and full kTB-row blocks run without per-row predicates: the first version spent more
  // issue slots on address arithmetic and bounds checks than on the taps (ncu: 72 thread-instructions per output pair).
  float2 cs = make_float2(0.f, 0.f);
  {
    const uint32_t* trow = tile + t_begin * 32 + lane;  // smem row (t - pad_left) lives at index t, so tap j of output t reads row t + j
    const int cw = C >> 1;                              // output row pitch in 32-bit words
    uint32_t* drow = reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * T + t_begin) * C + c0) + lane;
    const int nrows = max(0, t_end - t_begin);
    const int nfull = nrows / kTB;
    for (int blk = 0; blk < nfull; ++blk) {
      float2 x[kTB + K - 1];
#pragma unroll
      for (int i = 0; i < kTB + K - 1; ++i) {
        const uint32_t u = trow[i * 32];
        x[i] = make_float2(bf16_lo(u), bf16_hi(u));
      }
#pragma unroll
      for (int i = 0; i < kTB; ++i) {
        float2 a = bs;
#pragma unroll
        for (int j = 0; j < K; ++j) ffma2(a.x, a.y, wt[j].x, wt[j].y, x[i + j].x, x[i + j].y, a.x, a.y);
        if constexpr (POST == 1) { a.x = fmaf(a.x, fast_tanh(a.x), a.x); a.y = fmaf(a.y, fast_tanh(a.y), a.y); }
        if constexpr (POST == 2) fmul2(a.x, a.y, a.x, a.y, scale.x, scale.y);
        const uint32_t packed = pack_bf16x2(a.x, a.y);
        *drow = packed;
        drow += cw;
        if (colsum != nullptr && (mrow == nullptr || mrow[t_begin + blk * kTB + i] != 0)) { cs.x += bf16_lo(packed); cs.y += bf16_hi(packed); }
      }
      trow += kTB * 32;
    }
    for (int i = nfull * kTB; i < nrows; ++i) {  // ragged tail (T not a multiple of the register block)
      float2 a = bs;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const uint32_t u = trow[j * 32];
        ffma2(a.x, a.y, wt[j].x, wt[j].y, bf16_lo(u), bf16_hi(u), a.x, a.y);
      }
      if constexpr (POST == 1) { a.x = fmaf(a.x, fast_tanh(a.x), a.x); a.y = fmaf(a.y, fast_tanh(a.y), a.y); }
      if constexpr (POST == 2) fmul2(a.x, a.y, a.x, a.y, scale.x, scale.y);
      const uint32_t packed = pack_bf16x2(a.x, a.y);
      *drow = packed;
      drow += cw;
      if (colsum != nullptr && (mrow == nullptr || mrow[t_begin + i] != 0)) { cs.x += bf16_lo(packed); cs.y += bf16_hi(packed); }
      trow += 32;
    }
  }
  if (colsum != nullptr) {
    __syncthreads();  // red[] free again (POST==2 consumers are done: scale already in registers)
    red[warp * kSlab + 2 * lane] = cs.x;
    red[warp * kSlab + 2 * lane + 1] = cs.y;
    __syncthreads();
    if (tid < kSlab) {
      float m = 0.f;
#pragma unroll
      for (int wv = 0; wv < kWarps; ++wv) m += red[wv * kSlab + tid];
      colsum[static_cast<size_t>(b) * C + c0 + tid] = m;
    }
  }
  if constexpr (POST == 2) cluster.barrier_wait();
}

template <int K, int POST>
int launch_inst(const DwConvArgs& a, cudaStream_t stream) {
  constexpr int kTB = dw_tb<K>();
  const int rows_per_warp = ((a.T + kWarps - 1) / kWarps + kTB - 1) / kTB * kTB;
  const int rows_alloc = rows_per_warp * kWarps + K - 1;
  const size_t smem = static_cast<size_t>(rows_alloc) * 128 + (kWarps * kSlab + kSlab) * sizeof(float);
  auto kern = dwconv_kernel<K, POST>;
  static size_t smem_attr = 0;
  if (smem > smem_attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    smem_attr = smem;
  }
  const int slabs = a.C / kSlab;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(slabs, a.B, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int nattr = 0;
  if (POST == 2) {
    if (slabs > 16) {
      set_last_error("dwconv: ECA needs C/64 <= 16 (one cluster per sequence)");
      return 2;
    }
    if (slabs > 8) ISHARA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = slabs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    nattr = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  ISHARA_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a.in, a.out, a.w, a.bias, a.eca_w, a.colsum, a.T, a.C, a.pad_left,
                                    rows_alloc, a.key_mask, a.valid_cnt));
  note_launch();
  return 0;
}

template <int K>
int launch_k(const DwConvArgs& a, cudaStream_t stream) {
  switch (a.post) {
    case 0: return launch_inst<K, 0>(a, stream);
    case 1: return launch_inst<K, 1>(a, stream);
    case 2: return launch_inst<K, 2>(a, stream);
  }
  set_last_error("dwconv: bad post mode");
  return 2;
}

}  // namespace

int dwconv_launch(const DwConvArgs& a, cudaStream_t stream) {
  if (a.C % kSlab != 0 || a.T < a.k || a.B <= 0) {
    set_last_error("dwconv: C must be a multiple of 64 and T >= k");
    return 2;
  }
  if (a.post == 2 && a.eca_w == nullptr) {
    set_last_error("dwconv: ECA post-op needs eca_w");
    return 2;
  }
  switch (a.k) {
    case 3: return launch_k<3>(a, stream);
    case 5: return launch_k<5>(a, stream);
    case 7: return launch_k<7>(a, stream);
    case 9: return launch_k<9>(a, stream);
    case 11: return launch_k<11>(a, stream);
    case 15: return launch_k<15>(a, stream);
    case 31: return launch_k<31>(a, stream);
  }
  set_last_error("dwconv: unsupported kernel size (3,5,7,9,11,15,31)");
  return 2;
}

}  // namespace ishara
