// ishara_b200 — dense weight gradient dW[I,O] += X[M,I]^T @ G[M,O] for every Dense / 1x1 Conv1D on the path
// (SURVEY.md §8 row T15). Both operands are contracted over their ROW index, i.e. they are "MN-major" for the MMA;
// ldmatrix.trans turns the row-major [rows, cols] shared-memory tiles into A (X^T) and B (G) fragments, so no
// transposed copy of any activation is ever written. The row axis (M = B*T, 24k-98k) is split across CTAs; each CTA
// owns a 128x128 tile of dW for its row range and adds it to the fp32 gradient with vector atomics.
// Roofline: reads X and G once per (tile column / tile row) = HBM/L2-bound for I,O <= 512; tensor work on mma.sync.
#include "mma.cuh"
#include <cstdio>

#include "ptx.cuh"
#include "train_kernels.h"

namespace ishara {
namespace {

constexpr int kWgBI = 128, kWgBO = 128, kWgBK = 32, kWgLd = 136, kWgThreads = 256;

__global__ void __launch_bounds__(kWgThreads)
wgrad_kernel(const bf16* __restrict__ X, int ldx, const bf16* __restrict__ G, int ldg, float* __restrict__ dW, int ldw,
             float* __restrict__ dbias, int64_t M, int I, int O, int Ivalid, int Ovalid, int rows_per_split) {
  __shared__ __align__(16) bf16 Xs[2][kWgBK][kWgLd];
  __shared__ __align__(16) bf16 Gs[2][kWgBK][kWgLd];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.y * kWgBI, o0 = blockIdx.x * kWgBO;
  const int64_t m_begin = static_cast<int64_t>(blockIdx.z) * rows_per_split;
  const int64_t m_end = m_begin + rows_per_split < M ? m_begin + rows_per_split : M;
  if (m_begin >= m_end) return;
  const int wi = (warp & 3) * 32, wo = (warp >> 2) * 64;
  const bool do_bias = dbias != nullptr && blockIdx.y == 0;  // the i-tile-0 CTAs see every row of G for their columns
  float bsum = 0.f;
  float acc[2][8][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[a][n][0] = acc[a][n][1] = acc[a][n][2] = acc[a][n][3] = 0.f;

  const int nchunks = static_cast<int>((m_end - m_begin + kWgBK - 1) / kWgBK);
  auto load = [&](int stage, int chunk) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + q * kWgThreads;  // 0..511 : 32 rows x 16 column groups of 8
      const int row = idx >> 4, cg = (idx & 15) * 8;
      const int64_t m = m_begin + static_cast<int64_t>(chunk) * kWgBK + row;
      const bool rv = m < m_end;
      const bool xv = rv && (i0 + cg) < I, gv = rv && (o0 + cg) < O;
      const bf16* xs = xv ? X + m * ldx + i0 + cg : X;
      const bf16* gs = gv ? G + m * ldg + o0 + cg : G;
      cp_async16(smem_u32(&Xs[stage][row][cg]), xs, xv ? 16u : 0u);
      cp_async16(smem_u32(&Gs[stage][row][cg]), gs, gv ? 16u : 0u);
    }
  };
  load(0, 0);
  cp_async_commit();
  for (int ch = 0; ch < nchunks; ++ch) {
    const int st = ch & 1;
    if (ch + 1 < nchunks) {
      load(st ^ 1, ch + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (do_bias) {
      const int col = tid & 127, r0 = (tid >> 7) * 16;
#pragma unroll
      for (int r = 0; r < 16; ++r) bsum += __bfloat162float(Gs[st][r0 + r][col]);
    }
#pragma unroll
    for (int kk = 0; kk < kWgBK / 16; ++kk) {
      uint32_t af[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        ldmatrix_x4_trans(af[mt], smem_u32(&Xs[st][kk * 16 + ((lane >> 4) & 1) * 8 + (lane & 7)][wi + mt * 16 + ((lane >> 3) & 1) * 8]));
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bfr[4];
        ldmatrix_x4_trans(bfr, smem_u32(&Gs[st][kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)][wo + np * 16 + ((lane >> 4) & 1) * 8]));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(acc[mt][2 * np], af[mt], bfr[0], bfr[1]);
          mma16816(acc[mt][2 * np + 1], af[mt], bfr[2], bfr[3]);
        }
      }
    }
    __syncthreads();
  }
  if (do_bias) {
    __shared__ float bred[kWgThreads];
    bred[tid] = bsum;
    __syncthreads();
    if (tid < 128 && o0 + tid < Ovalid) atomicAdd(&dbias[o0 + tid], bred[tid] + bred[tid + 128]);
  }
  const int g = lane >> 2, tg = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int o = o0 + wo + n * 8 + tg * 2;
      if (o >= Ovalid) continue;
      const int ia = i0 + wi + mt * 16 + g, ib = ia + 8;
      if (ia < Ivalid) atomicAdd(reinterpret_cast<float2*>(dW + static_cast<size_t>(ia) * ldw + o), make_float2(acc[mt][n][0], acc[mt][n][1]));
      if (ib < Ivalid) atomicAdd(reinterpret_cast<float2*>(dW + static_cast<size_t>(ib) * ldw + o), make_float2(acc[mt][n][2], acc[mt][n][3]));
    }
}

}  // namespace

int wgrad_launch(const bf16* X, int ldx, const bf16* G, int ldg, float* dW, int ldw, float* dbias, int64_t M, int I, int O,
                 int Ivalid, int Ovalid, int num_sms, cudaStream_t s) {
  if (I % 8 != 0 || O % 8 != 0 || ldx % 8 != 0 || ldg % 8 != 0 || ldw % 2 != 0 || Ovalid % 2 != 0 || Ivalid > I || Ovalid > O) {
    set_last_error("wgrad: I, O, ldx, ldg must be multiples of 8; ldw and Ovalid even");
    return 2;
  }
  if ((reinterpret_cast<uintptr_t>(dW) & 7) != 0) { set_last_error("wgrad: dW must be 8-byte aligned"); return 2; }
  const int ti = (I + kWgBI - 1) / kWgBI, to = (O + kWgBO - 1) / kWgBO;
  int splits = (2 * num_sms + ti * to - 1) / (ti * to);
  const int64_t max_splits = (M + 127) / 128;
  if (splits > max_splits) splits = static_cast<int>(max_splits);
  if (splits < 1) splits = 1;
  int64_t rps = (M + splits - 1) / splits;
  rps = (rps + kWgBK - 1) / kWgBK * kWgBK;
  splits = static_cast<int>((M + rps - 1) / rps);
  wgrad_kernel<<<dim3(to, ti, splits), kWgThreads, 0, s>>>(X, ldx, G, ldg, dW, ldw, dbias, M, I, O, Ivalid, Ovalid, static_cast<int>(rps));
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error(std::string("wgrad: ") + cudaGetErrorString(e)); return 3; }
  note_launch();
  return 0;
}

}  // namespace ishara
