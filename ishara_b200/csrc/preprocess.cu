// ishara_b200 — landmark preprocessing in front of the encoder: raw MediaPipe frames
// [N, 276] in the reference's SEL_COLS order -> the model input [T, 276].
//
// Reference (nb:conv-hybrid-model): TFLiteModel.__call__ c13:9-15 = pre_process00 (group gather + hand-frame filter,
// c3:57-101) -> pre_process1 (resize_pad c3:1-7: NaN-pad to FRAME_LEN or bilinear resize of the time axis with
// tf.image.resize = half-pixel centres, no antialias; (x - mean) / std per group; concat [lip, rhand, lhand, rpose,
// lpose]; reshape; NaN -> 0, c3:103-115). filter = 0 gives the training-path variant (pre_process1 only).
//
// Memory-bound gather. One CTA per (sequence, slice of output rows): every CTA of a sequence recomputes the keep
// flags and their prefix sum (N <= a few thousand frames, cheaper than a second launch), then writes its rows with
// column-fastest (coalesced) stores. The SEL_COLS bookkeeping is closed-form, no index tables in memory.
#include <cstdio>

#include "kernels.h"
#include "ptx.cuh"

namespace ishara {
namespace {

constexpr int kPpThreads = 256;
constexpr int kF = 276;        // len(SEL_COLS) = 3 coordinates x 92 landmarks
constexpr int kPerAxis = 92;   // per coordinate block: rhand 0..20, lhand 21..41, pose 42..51 (LPOSE, RPOSE), lip 52..91

// output column (landmark-major, xyz-minor; groups lip | rhand | lhand | rpose | lpose) -> SEL_COLS column
__device__ __forceinline__ int source_column(int oc) {
  const int l = oc / 3, c = oc - 3 * l;
  int pos;
  if (l < 40) pos = 52 + l;                 // lip
  else if (l < 61) pos = l - 40;            // right hand
  else if (l < 82) pos = 21 + (l - 61);     // left hand
  else if (l < 87) pos = 47 + (l - 82);     // right pose (second half of POSE = LPOSE + RPOSE)
  else pos = 42 + (l - 87);                 // left pose
  return c * kPerAxis + pos;
}

__global__ void __launch_bounds__(kPpThreads)
preprocess_kernel(const float* __restrict__ frames, const int32_t* __restrict__ offsets, const float* __restrict__ mean,
                  const float* __restrict__ stdv, float* __restrict__ out, int T, int filter, int rows_per_cta) {
  extern __shared__ int32_t sel[];  // kept frame indices of this sequence
  __shared__ int32_t warp_tot[kPpThreads / 32];
  __shared__ int32_t n_kept_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t f0 = offsets[b];
  const int N = offsets[b + 1] - offsets[b];
  const float* x = frames + f0 * kF;
  const bool zero_frame = N == 0;           // c13:11: an empty input becomes one all-zero frame
  const int Neff = zero_frame ? 1 : N;

  // ---- keep flags: one warp per frame, coalesced over the 3 x 42 hand columns (cols 0..41 of each coordinate block) ----
  uint8_t* flags = reinterpret_cast<uint8_t*>(sel + Neff);  // [Neff] after the index list
  for (int f = warp; f < Neff; f += kPpThreads / 32) {
    bool keep = true;
    if (filter && !zero_frame) {
      const float* r = x + static_cast<int64_t>(f) * kF;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v0 = r[c * kPerAxis + lane];
        s += (v0 != v0) ? 0.f : v0;
        if (lane < 10) {
          const float v1 = r[c * kPerAxis + 32 + lane];
          s += (v1 != v1) ? 0.f : v1;
        }
      }
      // any non-zero lane partial makes the frame's sum non-zero except under exact cancellation, which the reference's
      // own float32 reduction order does not define either; coordinates are non-negative in practice
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      keep = (s != 0.f) || ((f & 1) == 0);   // cumsum(ones) % 2 == 1 <=> 0-based even index
    }
    if (lane == 0) flags[f] = keep ? 1 : 0;
  }
  __syncthreads();
  // ---- compaction: thread owns a contiguous run of flags ----
  const int per = (Neff + kPpThreads - 1) / kPpThreads;
  const int fb = tid * per, fe = min(Neff, fb + per);
  int cnt = 0;
  for (int f = fb; f < fe; ++f) cnt += flags[f];
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int base = incl - cnt;
  for (int w = 0; w < warp; ++w) base += warp_tot[w];
  if (tid == kPpThreads - 1) n_kept_s = base + cnt;
  for (int f = fb; f < fe; ++f)
    if (flags[f]) sel[base++] = f;
  __syncthreads();
  const int Nk = n_kept_s;

  // ---- output rows of this CTA: one warp per row. The one or two source frames are read as whole rows (coalesced)
  // into a per-warp staging buffer, then permuted into the output column order on the way out ----
  __shared__ float stage[kPpThreads / 32][2][kF + 4];
  const int t_begin = blockIdx.y * rows_per_cta, t_end = min(T, t_begin + rows_per_cta);
  const bool resize = Nk >= T;
  const float scale = static_cast<float>(Nk) / static_cast<float>(T);
  float* ob = out + static_cast<int64_t>(b) * T * kF;
  for (int t = t_begin + warp; t < t_end; t += kPpThreads / 32) {
    int lo = -1, hi = -1;
    float w = 0.f;
    if (resize) {
      // tf.image.resize bilinear, half_pixel_centers: in = (t + 0.5) * scale - 0.5 (each op rounded, no FMA contraction)
      const float in_f = __fsub_rn(__fmul_rn(static_cast<float>(t) + 0.5f, scale), 0.5f);
      const float fl = floorf(in_f);
      lo = sel[max(static_cast<int>(fl), 0)];
      hi = sel[min(static_cast<int>(ceilf(in_f)), Nk - 1)];
      w = __fsub_rn(in_f, fl);
    } else if (t < Nk && !zero_frame) {
      lo = sel[t];
    }
    float* s0 = stage[warp][0];
    float* s1 = stage[warp][1];
    if (lo >= 0) {
      const float* r0 = x + static_cast<int64_t>(lo) * kF;
      for (int c = lane; c < kF; c += 32) s0[c] = r0[c];
      if (hi >= 0 && hi != lo) {
        const float* r1 = x + static_cast<int64_t>(hi) * kF;
        for (int c = lane; c < kF; c += 32) s1[c] = r1[c];
      }
    }
    __syncwarp();
    for (int oc = lane; oc < kF; oc += 32) {
      const int sc = source_column(oc);
      float v;
      if (resize) {
        const float top = s0[sc], bot = (hi != lo) ? s1[sc] : top;
        v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), w));
      } else if (t < Nk) {
        v = zero_frame ? 0.f : s0[sc];
      } else {
        v = __int_as_float(0x7fc00000);     // NaN padding (c3:4) -> 0 below
      }
      v = __fdiv_rn(__fsub_rn(v, mean[oc]), stdv[oc]);
      ob[static_cast<int64_t>(t) * kF + oc] = (v != v) ? 0.f : v;
    }
    __syncwarp();
  }
}

}  // namespace

int preprocess_launch(const float* frames, const int32_t* offsets, int B, int max_frames, const float* mean, const float* stdv, int T,
                      int filter, float* out, cudaStream_t stream) {
  if (B <= 0 || T <= 0 || max_frames < 0) { set_last_error("preprocess: bad shape"); return 2; }
  const size_t nf = static_cast<size_t>(max_frames < 1 ? 1 : max_frames);
  const size_t smem = nf * sizeof(int32_t) + ((nf + 15) & ~static_cast<size_t>(15));  // kept-frame indices + one flag byte per frame
  if (smem > 200 * 1024) { set_last_error("preprocess: more than 40000 frames in one sequence"); return 2; }
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    ISHARA_CUDA_OK(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  int slices = B >= 592 ? 1 : (592 + B - 1) / B;  // ~4 CTAs per SM in flight even for a small batch
  if (slices > 16) slices = 16;
  const int rows_per_cta = (T + slices - 1) / slices;
  slices = (T + rows_per_cta - 1) / rows_per_cta;
  preprocess_kernel<<<dim3(B, slices), kPpThreads, smem, stream>>>(frames, offsets, mean, stdv, out, T, filter, rows_per_cta);
  ISHARA_CUDA_OK(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace ishara
