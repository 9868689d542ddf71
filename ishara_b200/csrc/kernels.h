// ishara_b200 — internal launcher interface between the model runtime (model.cu / capi.cu) and the
// sm_100a kernels. Not part of the public C ABI (that is include/ishara_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

namespace ishara {

typedef __nv_bfloat16 bf16;

// ---- error plumbing -------------------------------------------------------------------------
void set_last_error(const std::string& msg);
const char* get_last_error();
// every kernel launcher bumps this process-wide counter (bench.py reports it as gpu_launches)
void note_launch();
void note_launches(int n);  // a replayed CUDA graph launches n kernels at once
uint64_t launch_count();
#define ISHARA_CUDA_OK(expr)                                                                           \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      ::ishara::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ +  \
                               ":" + std::to_string(__LINE__) + ")");                                  \
      return 3; /* ISHARA_ERR_CUDA */                                                                  \
    }                                                                                                  \
  } while (0)

// ---- TMA tensor maps --------------------------------------------------------------------------
enum TmapDtype { TM_BF16 = 0, TM_F32 = 1 };
// 2-D row-major tensor [rows, cols] with row pitch `ld_elems`; box = [box_rows, box_cols];
// 128-byte swizzle when box_cols*elemsize == 128, 64-byte swizzle when == 64.
int make_tmap_2d(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t rows, uint64_t cols,
                 uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols);

// ---- tcgen05 GEMM ------------------------------------------------------------------------------
enum GemmAct { ACT_NONE = 0, ACT_SWISH = 1, ACT_RELU = 2, ACT_GLU = 3 };

// out0[M, Nout] = f( A[M,K] @ Wt[N,K]^T ), f assembled from the non-null fields, in this order:
//   v = acc + bias[n]; v *= gate[row / rows_per_seq][n]; v += rowtab[row % rows_per_seq][n];
//   v = act(v)  (GLU: v = a * sigmoid(b), a/b = the two halves of the N-tile; Nout = N/2);
//   v += resid[row][n];
//   if ln0: v = LN(v; ln0_g, ln0_b, ln0_eps)        (needs the full row in one tile: N == block_n)
//   out0 = v (bf16 or fp32)
//   if ln1: out1 = bf16( LN(v; ln1_g, ln1_b, ln1_eps) )
struct GemmEpi {
  const float* bias = nullptr;
  const float* gate = nullptr;
  const float* rowtab = nullptr;
  const bf16* resid = nullptr;
  const float* ln0_g = nullptr;
  const float* ln0_b = nullptr;
  const float* ln1_g = nullptr;
  const float* ln1_b = nullptr;
  float ln0_eps = 0.f, ln1_eps = 0.f;
  int rows_per_seq = 1;
  int ld_resid = 0;
  int act = ACT_NONE;
  long long* trace = nullptr;  // ISHARA_GEMM_TRACE: clock64 timeline of CTA 0, [tile][8 events]
  int dbg = 0;  // ISHARA_GEMM_DBG bisect bits: 1 skip TMA stores, 2 skip epilogue work, 4 skip MMA issue, 8 skip staging writes
};

struct GemmPlan {
  CUtensorMap tmA, tmB, tmBh, tmO0, tmO1;  // tmBh: 128-row B box for the CTA-pair kernel
  CUtensorMap tmR;                         // residual [M, Nout] with [32 x 64] boxes (row mode: fetched by TMA into the staging boxes)
  bool resid_tma = false;
  GemmEpi epi;
  int M = 0, N = 0, K = 0;  // N = packed weight rows (MMA N extent), K multiple of 64
  int block_n = 256;        // 64 | 128 | 256
  bool out_f32 = false;
  bool row_mode = false;    // full-row epilogue (LN capable); requires N == block_n
  bool no_resident = false; // force the streaming-B variant (tests / A-B timing)
  bool no_pair = false;     // force a single-CTA variant
};

// Fill tensor maps of a plan. A [M,K] bf16 (ld = lda), Wt [N,K] bf16 (ld = K),
// out0 [M,Nout] (bf16|f32, ld = ldo0), out1 [M,Nout] bf16 (ld = ldo1) or null.
int gemm_plan_init(GemmPlan* p, const bf16* A, int lda, const bf16* Wt, void* out0, int ldo0, int nout,
                   bf16* out1, int ldo1);
int gemm_launch(const GemmPlan& p, int num_sms, cudaStream_t stream);

// ---- fused feed-forward module (ffn_tc.cu): S = [LN0](S + swish(XN @ W1 + b1) @ W2 + b2) [, XN' = LN1(S)] ------
struct FfnPlan {
  CUtensorMap tmA, tmW1, tmW2, tmO0, tmO1;
  CUtensorMap tmO0h, tmO1h;    // the same outputs as [32 x 32] boxes (16-warp epilogue: half staging boxes)
  CUtensorMap tmW1e;           // W1^T as [64 x 64] boxes (ffn_tc_v2_kernel walks the hidden dimension in eighths)
  GemmEpi epi;                 // epilogue of the second GEMM: bias = b2, resid, ln0 / ln1
  const float* bias1 = nullptr;  // [E]
  int M = 0, E = 0;
};
bool ffn_applicable(int D, int E, int M, int num_sms);
int ffn_plan_init(FfnPlan* p, const bf16* xn, const bf16* w1t, const bf16* w2t, bf16* out0, bf16* out1);
int ffn_launch(const FfnPlan& p, int num_sms, cudaStream_t stream);

// ---- fused front half of Conv1DBlock (conv1d_front.cu): G = ECA(BN(CausalDW(swish(x @ We + be)))) ----------
struct Conv1dFrontPlan {
  CUtensorMap tmA, tmB;
  const float* bias_e = nullptr;  // [512] expand bias
  const float* dw_w = nullptr;    // [k, 512] depthwise taps with BatchNorm folded in
  const float* dw_b = nullptr;    // [512] BatchNorm offset
  const float* eca_w = nullptr;   // [5]
  bf16* out = nullptr;            // [B*T, 512]
  int B = 0, T = 0, k = 0;
};
bool conv1d_front_applicable(int D, int T, int k);
int conv1d_front_plan_init(Conv1dFrontPlan* p, const bf16* x, const bf16* wet);
int conv1d_front_launch(const Conv1dFrontPlan& p, cudaStream_t stream);

// ---- the whole Conv1DBlock in one launch (conv1d_block.cu): S <- S + ECA(BN(CausalDW(swish(S We + be)))) Wp + bp ----
struct Conv1dBlockPlan {
  CUtensorMap tmX, tmWe, tmWp, tmO0, tmO1;
  const float* bias_e = nullptr;  // [512] expand bias
  const float* dw_w = nullptr;    // [k, 512] depthwise taps with BatchNorm folded in
  const float* dw_b = nullptr;    // [512] BatchNorm offset
  const float* dw_wsum = nullptr; // [512] sum of the k taps per channel
  const float* eca_w = nullptr;   // [5]
  const float* bias_p = nullptr;  // [256] project bias
  const bf16* resid = nullptr;    // = S (set by plan_init)
  const float* ln_g = nullptr;    // LayerNorm of the next module: XN = LN(S) (null: none)
  const float* ln_b = nullptr;
  float ln_eps = 0.f;
  // mask_mode="propagated": ECA averages over the valid frames only (c5:8-9). Null = every frame counts.
  const uint16_t* wbits = nullptr;     // [B*T] window validity bits (mask_prep_launch)
  const int32_t* valid_cnt = nullptr;  // [B]
  int B = 0, T = 0, k = 0;
};
bool conv1d_block_applicable(int D, int T, int k);
int conv1d_block_plan_init(Conv1dBlockPlan* p, bf16* s, const bf16* wet, const bf16* wpt, bf16* xn);
int conv1d_block_launch(const Conv1dBlockPlan& p, cudaStream_t stream);

// ---- depthwise temporal convolution over a whole sequence per CTA ---------------------------
// in/out [B, T, C] bf16 channels-last. y[t,c] = post( sum_j w[j,c] * in[t - pad_left + j, c] + bias[c] )
// (zeros outside [0,T)); BatchNorm is folded into w/bias by the caller.
//   post: 0 none, 1 swish, 2 ECA: y *= sigmoid( conv1d_k5_same over channels of mean_t(y) )
// colsum (optional): [B, C] fp32 = sum_t of the post-activated output (bf16-rounded values).
struct DwConvArgs {
  const bf16* in = nullptr;
  bf16* out = nullptr;
  const float* w = nullptr;     // [k, C] fp32
  const float* bias = nullptr;  // [C] fp32 or null
  const float* eca_w = nullptr; // [5] fp32 (post == 2)
  float* colsum = nullptr;      // [B, C] or null
  int B = 0, T = 0, C = 0, k = 0, pad_left = 0, post = 0;
  // mask_mode="propagated": frames with key_mask == 0 do not count in the ECA mean (post 2) / the column sums (SE)
  const uint8_t* key_mask = nullptr;   // [B, T], 1 = valid
  const int32_t* valid_cnt = nullptr;  // [B]
};
int dwconv_launch(const DwConvArgs& a, cudaStream_t stream);

// ---- attention -------------------------------------------------------------------------------
// qkv [B*T, 3*D] bf16, per-head interleaved columns [h][q|k|v][dh]; out [B*T, D] bf16 (head-major).
// softmax(q k^T * scale + maskbias) v ; key_mask [B,T] uint8 (1 = keep) or null.
struct AttnArgs {
  const bf16* qkv = nullptr;
  bf16* out = nullptr;
  const uint8_t* key_mask = nullptr;
  // Transformer-XL relative-position variant (squeezeformer/attention.py:25-110) when pos != null:
  // pos [2T-1, H*dh] bf16 = pos_proj(pos_emb) (head-major columns), u_bias / v_bias [H*dh] fp32.
  const bf16* pos = nullptr;
  const float* u_bias = nullptr;
  const float* v_bias = nullptr;
  int B = 0, T = 0, H = 0, dh = 0;
  float scale = 1.f;
  // training only: row log-sum-exp of the scaled scores in the log2 domain, [B*H*T] fp32 (saved for the backward pass)
  float* lse_out = nullptr;
  // training only: dropout on the attention probabilities (c5:113); thr16 = 0 disables. See dropout_hash.cuh.
  uint32_t drop_thr16 = 0;
  float drop_inv_keep = 1.f;
  uint64_t drop_key = 0;
  const uint64_t* drop_key_ptr = nullptr;  // when set, the key is read from device memory at run time
};
int attention_launch(const AttnArgs& a, cudaStream_t stream);

// ---- small memory-bound kernels ---------------------------------------------------------------
// x fp32 [M, F] -> bf16 [M, Fpad] (zero pad)
int cast_pad_launch(const float* x, bf16* out, int64_t M, int F, int Fpad, cudaStream_t stream);
// SE gate from column sums (linearity of the 1x1 conv): gate[b,:] = sigmoid(fc2(swish(fc1(mean @ W3 + b3))))
struct SeGateArgs {
  const float* colsum = nullptr;  // [B, C]
  const bf16* w3kn = nullptr;     // [C, D] bf16: conv3 kernel in its native orientation (row c = input channel)
  const float* b3 = nullptr;      // [D]
  const float* fc1_w = nullptr;   // [D, R] fp32
  const float* fc1_b = nullptr;   // [R]
  const float* fc2_w = nullptr;   // [R, D] fp32
  const float* fc2_b = nullptr;   // [D]
  float* gate = nullptr;          // [B, D]
  int B = 0, C = 0, D = 0, R = 0;
  float inv_T = 1.f;
  const int32_t* valid_cnt = nullptr;  // [B] valid frames per sequence (mask_mode="propagated"): mean = colsum / valid_cnt
};
// Keras Masking(0.0) for mask_mode="propagated": mask_out[B*T] (1 = frame carries data) from x fp32 [B,T,F] (or from
// user_mask when given and *use_user != 0), wbits[B*T] (bit d = frame t+d valid, d < 16), valid_cnt[B]
int mask_prep_launch(const float* x, const uint8_t* user_mask, const int* use_user, int B, int T, int F, uint8_t* mask_out,
                     uint16_t* wbits, int32_t* valid_cnt, cudaStream_t stream);
int se_gate_launch(const SeGateArgs& a, cudaStream_t stream);
// standalone LayerNorm (fallback / operator-level use): x bf16 [M, D] -> bf16
int layernorm_launch(const bf16* x, bf16* out, const float* g, const float* b, float eps, int64_t M, int D,
                     cudaStream_t stream);

// ---- time down/up-sampling (vendored Squeezeformer, rows P7/P8) -------------------------------
struct TimeReduceArgs {
  const bf16* x = nullptr;  // [B, T, D]
  bf16* out = nullptr;      // [B, T2, ldo], columns >= D2 zero
  float w[9] = {0};         // Conv2d(1,1,3,3) weight, row-major (time, channel)
  float bias = 0.f;
  int B = 0, T = 0, D = 0, T2 = 0, D2 = 0, ldo = 0;
};
int time_reduce_launch(const TimeReduceArgs& a, cudaStream_t stream);
// out[b, t, :] = y[b, t/2, :] + rec[b, t, :] for t < 2*T2  (y [B,T2,D], rec [B,T,D], out [B,2*T2,D])
int upsample_add_launch(const bf16* y, const bf16* rec, bf16* out, int B, int T2, int T, int D, cudaStream_t stream);
struct Conv2dSubsampleArgs {
  const float* x = nullptr;   // [B, T, F] fp32 features
  bf16* out = nullptr;        // [B, T4, ldo]: column c*F4 + f
  const float* w1 = nullptr;  // [C, 9]  Conv2d(1->C, 3, stride 2)
  const float* b1 = nullptr;  // [C]
  const float* w2 = nullptr;  // [C, 9]  depthwise Conv2d(C, 3, stride 2)
  const float* b2 = nullptr;  // [C]
  int B = 0, T = 0, F = 0, C = 0, T4 = 0, F4 = 0, ldo = 0;
};
int conv2d_subsample_launch(const Conv2dSubsampleArgs& a, cudaStream_t stream);

// ---- landmark preprocessing (preprocess.cu) --------------------------------
// frames fp32 [total_frames, 276] (SEL_COLS order), offsets int32 [B+1]; mean / stdv [276] in OUTPUT column order;
// out fp32 [B, T, 276]. filter != 0 applies the hand-frame filter of pre_process00.
int preprocess_launch(const float* frames, const int32_t* offsets, int B, int max_frames, const float* mean, const float* stdv, int T,
                      int filter, float* out, cudaStream_t stream);

// ---- CTC + greedy decode ----------------------------------------------------------------------
// logits fp32 [B,T,V]; labels int32 [B,L] padded with `blank`; nll [B]; grad [B,T,V] or null
// (grad = d nll_b / d logits, unreduced).
// The gradient pass (grad != null) needs a caller-owned workspace of ctc_workspace_bytes(B, T, L) bytes.
size_t ctc_workspace_bytes(int B, int T, int L);
int ctc_loss_launch(const float* logits, const int32_t* labels, int B, int T, int V, int L, int blank, float* nll,
                    float* grad, float* workspace, size_t workspace_bytes, cudaStream_t stream);
// ids_out int32 [B, T] (first lens[b] valid), lens int32 [B]; reproduces the reference decode_phrase quirk.
int greedy_decode_launch(const float* logits, int B, int T, int V, int blank, int32_t* ids_out, int32_t* lens,
                         cudaStream_t stream);

}  // namespace ishara
