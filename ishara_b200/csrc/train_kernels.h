// ishara_b200 — launcher interface of the training-step kernels. Internal; the public
// surface is ishara_model_train_* in include/ishara_b200.h. Everything here runs on sm_100a only.
//
// Conventions: activations and activation gradients are bf16 [M, C] row-major (M = B*T rows, channels-last), C a
// multiple of 8; statistics, gates and parameter gradients are fp32 (column sums over all rows accumulate in fp64).
// Every parameter-gradient output ACCUMULATES (+=) into the flat gradient buffer, which the step zeroes once.
#pragma once
#include "kernels.h"

namespace ishara {

enum TrainAct { TACT_NONE = 0, TACT_SWISH = 1, TACT_RELU = 2 };

// ---- elementwise (train_ew.cu) --------------------------------------------------------------------
// out = swish(in)                                                    [n elements, n % 8 == 0]
int act_fwd_launch(const bf16* in, bf16* out, int64_t n, int act, cudaStream_t s);
// act = swish: dU = dH * swish'(U) (ref = pre-activation U). act = relu: dU = dH * (ref > 0) (ref = post-activation).
int act_bwd_launch(const bf16* dH, const bf16* ref, bf16* dU, int64_t n, int act, cudaStream_t s);
// GLU (c5:294-295): out[M,C] = P[:, :C] * sigmoid(P[:, C:]) for P [M, 2C]; and its backward dP [M, 2C]
int glu_fwd_launch(const bf16* P, bf16* out, int64_t M, int C, cudaStream_t s);
int glu_bwd_launch(const bf16* dOut, const bf16* P, bf16* dP, int64_t M, int C, cudaStream_t s);
// y = ((scale ? x*scale[c] + shift[c] : x) * (gate ? gate[row/T][c] : 1)) + (resid ? resid : 0)
int affine_gate_add_launch(const bf16* x, const float* scale, const float* shift, const float* gate, const bf16* resid,
                           bf16* y, int64_t M, int C, int T, cudaStream_t s);
// y = x * g[row/T][c] + a[row/T][c] * alpha      (SqueezeExcite backward: dZ = dOut*gate + dmean/T)
int gate_bias_launch(const bf16* x, const float* g, const float* a, float alpha, bf16* y, int64_t M, int C, int T,
                     cudaStream_t s);
// out = bf16(in * alpha), [M, V] fp32 -> [M, Vpad] bf16 with zero pad columns (CTC gradient -> classifier dgrad operand)
int scale_cast_pad_launch(const float* in, bf16* out, int64_t M, int V, int Vpad, float alpha, cudaStream_t s);
// counter-based dropout: keep(i) = hash(seed, site, i) >= p; y = (resid ? resid : 0) + x * keep/(1-p).
// per_sample != 0: one decision per sequence (noise_shape=(None,1,1), c5:83), i = row / T.
int dropout_launch(const bf16* x, const bf16* resid, bf16* y, int64_t M, int C, int T, float p, uint64_t seed,
                   uint32_t site, int per_sample, cudaStream_t s);
int dropout_keyed_launch(const bf16* x, const bf16* resid, bf16* y, int64_t M, int C, int T, float p, const uint64_t* key_dev,
                         int per_sample, cudaStream_t s);
int dropout_keys_launch(uint64_t seed, uint64_t* keys_dev, int n, cudaStream_t s);
// activation and elementwise dropout of the same [M, C] tensor in one pass (same mask bits as dropout_kernel)
int act_drop_fwd_launch(const bf16* in, bf16* out, int64_t n, int act, float p, const uint64_t* key_dev, cudaStream_t s);
int act_bwd_drop_launch(const bf16* dH, const bf16* ref, bf16* dU, int64_t n, int act, float p, const uint64_t* key_dev, cudaStream_t s);

// ---- reductions (train_ew.cu) ---------------------------------------------------------------------
// per-sequence column sums seqsum[B,C] (fp32, overwritten) and, when sum/sumsq != null, whole-batch sum / sum of
// squares [C] (fp64, ACCUMULATED: zero them first)
int colstats_launch(const bf16* x, float* seqsum, double* sum, double* sumsq, int B, int T, int C, cudaStream_t s);
// out[B,C] = sum_t dG[b,t,c] * (scale ? x*scale[c]+shift[c] : x)
int seq_dot_launch(const bf16* dG, const bf16* x, const float* scale, const float* shift, float* out, int B, int T,
                   int C, cudaStream_t s);
// BatchNorm training statistics (Keras: biased variance over (B,T)); also updates the moving statistics in place:
// moving = momentum*moving + (1-momentum)*batch. Writes mean, rstd, scale = gamma*rstd, shift = beta - mean*scale.
int bn_finalize_launch(const double* sum, const double* sumsq, double count, const float* gamma, const float* beta,
                       float eps, float momentum, float* moving_mean, float* moving_var, float* mean, float* rstd,
                       float* scale, float* shift, int C, cudaStream_t s);
// ECA gate (c5:1-15): m[b,c] = seqsum[b,c]*invT*scale[c] + shift[c]; s = sigmoid(conv1d_k5_same over c of m)
int eca_fwd_launch(const float* seqsum, const float* scale, const float* shift, float invT, const float* w5, float* m,
                   float* sgate, int B, int C, cudaStream_t s);
// given ds[b,c] = dL/ds: dm[b,c] (gradient w.r.t. the per-sequence mean), dw5 += ...
int eca_bwd_launch(const float* ds, const float* sgate, const float* m, const float* w5, float* dm, float* dw5, int B,
                   int C, cudaStream_t s);
// BatchNorm backward. dBn[b,t,c] = dG*(sgate ? sgate[b,c] : 1) + (dm ? dm[b,c]*invT : 0); xhat = (x-mean)*rstd.
//   reduce: sum1[c] += sum dBn, sum2[c] += sum dBn*xhat (fp64, zero first)
//   apply : dx = gamma*rstd*(dBn - sum1/count - xhat*sum2/count); dgamma += sum2, dbeta += sum1
int bn_bwd_reduce_launch(const bf16* dG, const bf16* x, const float* sgate, const float* dm, float invT,
                         const float* mean, const float* rstd, double* sum1, double* sum2, int B, int T, int C,
                         cudaStream_t s);
int bn_bwd_apply_launch(const bf16* dG, const bf16* x, const float* sgate, const float* dm, float invT,
                        const float* mean, const float* rstd, const float* gamma, const double* sum1,
                        const double* sum2, double count, float* coef /* [2C] scratch */, bf16* dx, float* dgamma,
                        float* dbeta, int B, int T, int C, cudaStream_t s);
// LayerNorm backward over rows of x [M,D] (statistics recomputed from x): dx = LN'(dy) (+ dresid if non-null);
// dgamma/dbeta += column sums. D multiple of 128, D <= 512.
int ln_bwd_launch(const bf16* dy, const bf16* x, const float* gamma, float eps, const bf16* dresid, bf16* dx,
                  float* dgamma, float* dbeta, int64_t M, int D, cudaStream_t s);
// out[c] += sum_rows g[row, c]  for c < Cvalid  (bias gradients)
int colsum_launch(const bf16* g, int ld, float* out, int64_t M, int Cvalid, cudaStream_t s);

// ---- SqueezeExcite dense layers over [B, D] (c5:120-133) -----------------------------------------
struct SeTrainArgs {
  const float* zsum = nullptr;   // [B, D] per-sequence column sums of conv3's output
  const float *fc1_w = nullptr, *fc1_b = nullptr, *fc2_w = nullptr, *fc2_b = nullptr;  // [D,R],[R],[R,D],[D] (fp32 masters)
  float *g = nullptr, *a_pre = nullptr, *gate = nullptr;  // saved: [B,D], [B,R], [B,D]
  // backward
  const float* dgate = nullptr;  // [B, D]
  float* dg = nullptr;           // [B, D] gradient w.r.t. the per-sequence mean
  float *d_fc1_w = nullptr, *d_fc1_b = nullptr, *d_fc2_w = nullptr, *d_fc2_b = nullptr;
  int B = 0, D = 0, R = 0;
  float invT = 1.f;
};
int se_fwd_launch(const SeTrainArgs& a, cudaStream_t s);
int se_bwd_launch(const SeTrainArgs& a, cudaStream_t s);

// ---- depthwise temporal convolution, training flavour (train_dw.cu) ------------------------------
// y[t,c] = (sum_j w[flip ? k-1-j : j][c] * pre(in[t - pad_left + j, c]) + bias[c]) * (mul_ref ? swish'(mul_ref[t,c]) : 1)
// pre = swish when pre_act == TACT_SWISH. Backward-data = the same kernel with flip = 1, pad_left' = k-1-pad_left.
struct DwTrainArgs {
  const bf16* in = nullptr;
  bf16* out = nullptr;
  const float* w = nullptr;      // [k, C] fp32 (Keras [k,C,1] or [k,1,C] — same memory order)
  const float* bias = nullptr;   // [C] or null
  const bf16* mul_ref = nullptr; // optional pre-activation tensor whose swish' multiplies the result
  int B = 0, T = 0, C = 0, k = 0, pad_left = 0, flip = 0, pre_act = 0;
};
int dw_train_launch(const DwTrainArgs& a, cudaStream_t s);
// dw[j,c] += sum_{b,t} dOut[b,t,c] * pre(in[b, t - pad_left + j, c]);  dbias[c] += sum dOut (when non-null)
int dw_wgrad_launch(const bf16* dOut, const bf16* in, int pre_act, float* dw, float* dbias, int B, int T, int C, int k,
                    int pad_left, cudaStream_t s);

// ---- dense weight gradient (train_wgrad.cu): dW[I,O] += X[M,I]^T @ G[M,O] --------------------------
// X, G bf16 row-major (ldx, ldg); dW fp32 row-major [Ivalid, ldw] (Keras [in,out]); only i < Ivalid, o < Ovalid stored.
// dbias (optional): [Ovalid] += column sums of G (the Dense bias gradient), taken from the tiles already in smem.
int wgrad_launch(const bf16* X, int ldx, const bf16* G, int ldg, float* dW, int ldw, float* dbias, int64_t M, int I, int O,
                 int Ivalid, int Ovalid, int num_sms, cudaStream_t s);

// tcgen05 version (train_wgrad_tc.cu): operands read MN-major straight from the row-major activations through TMA.
// The plan holds the two tensor maps and the grid; build once, launch every step. O must be a multiple of 64.
struct alignas(64) WgradTcPlan { unsigned char storage[384]; };
int wgrad_tc_plan_init(WgradTcPlan* p, const bf16* X, int ldx, const bf16* G, int ldg, int64_t M, int I, int O, int num_sms);
int wgrad_tc_launch(const WgradTcPlan* p, float* dW, int ldw, float* dbias, int Ivalid, int Ovalid, cudaStream_t s);

// ---- attention backward (train_attn.cu) ----------------------------------------------------------
// qkv / dqkv [B*T, 3*H*dh] per-head interleaved; o / dO [B*T, H*dh]; lse2, dsum scratch [B*H*T] fp32
struct AttnBwdArgs {
  const bf16 *qkv = nullptr, *o = nullptr, *dO = nullptr;
  bf16* dqkv = nullptr;
  float *lse2 = nullptr, *dsum = nullptr;
  int have_lse = 0;  // lse2 already holds the forward's row log-sum-exp (AttnArgs::lse_out): skip the recompute sweep
  int B = 0, T = 0, H = 0, dh = 0;
  float scale = 1.f;
  uint32_t drop_thr16 = 0;  // same dropout mask as the forward (AttnArgs)
  float drop_inv_keep = 1.f;
  uint64_t drop_key = 0;
  const uint64_t* drop_key_ptr = nullptr;  // when set, the key is read from device memory at run time
};
int attention_bwd_launch(const AttnBwdArgs& a, cudaStream_t s);

// ---- optimiser (train_opt.cu) --------------------------------------------------------------------
// norm2[0] += sum g^2 (fp64)
int sqnorm_launch(const float* g, int64_t n, double* norm2, cudaStream_t s);
struct AdamWArgs {
  float lr = 4.5e-3f, weight_decay = 0.08f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, max_norm = 1.0f;
  int step = 1;            // 1-based
  float grad_scale = 1.f;  // applied before clipping (1/world after a sum all-reduce)
  int* skipped = nullptr;  // device counter: bumped (and the update dropped) when the gradient norm is not finite
};
// tfa.optimizers.Lookahead(tfa.optimizers.RectifiedAdam(sma_threshold=4), sync_period=5) -- the optimiser the reference
// compiles the model with (nb:conv-hybrid-model c7:68-69). One fused pass over the flat buffers.
struct RAdamArgs {
  float lr = 1e-3f, weight_decay = 0.f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-7f, max_norm = 0.f;
  float sma_threshold = 4.f;
  int sync_period = 5;         // Lookahead k
  float slow_step = 0.5f;      // Lookahead alpha
  int step = 1;                // 1-based
  float grad_scale = 1.f;
  int* skipped = nullptr;
};
int radam_lookahead_launch(float* theta, const float* g, float* m, float* v, float* slow, int64_t n, const double* norm2,
                           const RAdamArgs& a, cudaStream_t s);
// theta/g/m/v [n]; clip scale derived on the device from norm2[0] (already of the scaled gradient)
int adamw_launch(float* theta, const float* g, float* m, float* v, int64_t n, const double* norm2, const AdamWArgs& a,
                 cudaStream_t s);
// bf16 working copies of a dense kernel W [I, O] fp32: fwd [Opad, Ipad] (= W^T, K-major B operand of y = x W) and
// bwd [I, Opad] (= W, K-major B operand of dx = dy W^T); zero padding; either pointer may be null.
struct RepackEntry {
  const float* src;
  bf16* fwd;
  bf16* bwd;
  int I, O, Ipad, Opad;
};
int repack_launch(const RepackEntry* table_dev, int n_entries, int max_tiles, cudaStream_t s);

}  // namespace ishara
