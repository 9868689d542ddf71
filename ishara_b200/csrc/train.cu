// ishara_b200 — the training step of get_model: model(x, training=True) -> CTCLoss ->
// gradients of all 7.59 M parameters -> global-norm clip + AdamW, on one B200. Reference semantics:
// nb:conv-hybrid-model c5:1-343 / c7:12-65 in Keras training mode (BatchNormalization on batch statistics with the
// moving-average update, dropout sites c5:83,113,162-166,183,190,204 and c7:62), loss c6:1-13, optimiser per
// BASELINE.json (integration.py:675-679,750).
//
// Layout in HBM: fp32 master parameters, gradients and both Adam moments are four flat buffers in the Keras layouts
// (trainable tensors first, BatchNorm moving statistics last); every Dense / 1x1-conv kernel additionally lives as
// two bf16 operand copies ([out,in] for y = xW on tcgen05, [in,out] for dx = dy W^T), refreshed by one repack
// launch after each update. Activations needed by the backward pass are kept as bf16 [B*T, C] buffers (one per
// tensor, ~2 GB at B = 64 — sized for 180 GB, nothing is recomputed except elementwise activations and the
// attention probabilities). The residual-stream gradient ping-pongs between two bf16 buffers.
//
// Program: built once per batch size as two lists of closures (forward order / backward order); the forward and
// data-gradient products run on the tcgen05 GEMM of gemm_tc.cu, weight gradients on train_wgrad.cu, attention
// backward on train_attn.cu, the rest on the streaming kernels of train_ew.cu / train_dw.cu.
#include <cmath>
#include <cstdlib>
#include <functional>
#include <memory>

#include "dropout_hash.cuh"
#include "model_internal.h"
#include "train_kernels.h"

namespace ishara {

int model_finalize(ishara_model* m);

using Step = std::function<int(cudaStream_t)>;

// comm.cu
void comm_bucket_plan(const int64_t* hi, int n, int64_t n_train, int64_t min_elems, int64_t* lo_out, int64_t* up_out);
int comm_allreduce_after(ishara_model* m, float* buf, int64_t count, cudaStream_t after);
int comm_join(ishara_model* m, cudaStream_t stream);

struct NamedBuf {
  const void* ptr = nullptr;
  int64_t rows = 0;
  int cols = 0, ld = 0;
  int dtype = 0;  // 0 bf16, 1 fp32
};

struct TrainState {
  // flat parameter storage
  std::vector<int64_t> off;  // per parameter index: offset in floats
  int64_t n_train = 0, n_total = 0;
  float *theta = nullptr, *grad = nullptr, *adam_m = nullptr, *adam_v = nullptr;
  float* slow = nullptr;  // Lookahead slow weights (allocated with the first RAdam + Lookahead step, = the weights at that time)
  double* norm2 = nullptr;
  int step = 0;
  std::unordered_map<std::string, bf16*> wf, wb;
  std::vector<void*> wallocs;
  RepackEntry* repack_dev = nullptr;
  int repack_n = 0, repack_max_tiles = 0;
  // program
  int batch = 0, labels_len = 0;
  float dropout = 0.f;
  uint64_t seed = 0;       // effective dropout seed of the CURRENT forward/backward = step_seed(seed_base, fb_count)
  uint64_t seed_base = 0;  // as configured
  int64_t fb_count = 0;    // forward/backward passes since train_configure: fresh dropout noise every step (Keras Dropout)
  bool debug = false;
  std::vector<Step> fwd, bwd;
  std::vector<void*> allocs;
  std::map<std::string, NamedBuf> named, named_grad;
  double* stats = nullptr;  // fp64 arena, zeroed at the start of every step
  size_t stats_count = 0;
  float* x_dev = nullptr;
  int32_t* labels_dev = nullptr;
  float *logits = nullptr, *dlogits = nullptr, *nll = nullptr, *loss_dev = nullptr;
  float* ctc_ws = nullptr;    // alphas | betas of the CTC gradient pass: owned by this handle's program (ctc.cu)
  size_t ctc_ws_bytes = 0;
  int* skipped_dev = nullptr; // optimiser steps skipped because the gradient norm was not finite
  // data-parallel exchange: per module (forward order) the end offset of the highest gradient its backward writes, and
  // the range of the flat gradient buffer that is final - and reduced - right after that module's backward
  std::vector<int64_t> mod_hi, bucket_lo, bucket_up;
  float* loss_pinned = nullptr;
  // weight gradients run on a side stream next to the data-gradient GEMM that follows them (both read the same dY and
  // neither fills the 148 SMs alone at 192 row tiles); the main stream rejoins right after that next step
  uint64_t* keys_dev = nullptr;          // [kMaxDropSites] dropout key of every site for the CURRENT step (device)
  // CUDA-graph replay of forward + CTC + backward (single-GPU, non-debug): captured on the second step of a program
  cudaGraphExec_t step_graph = nullptr;
  int step_graph_launches = 0;
  int steps_since_build = 0;
  bool step_graph_broken = false;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_side = nullptr;
  bool side_pending = false, last_was_side = false;
  std::shared_ptr<void> builder;  // closures may refer to builder members: it lives as long as the program
};

constexpr int kMaxDropSites = 256;

// run `fn` on the side stream, ordered after everything enqueued on `s` so far (see TrainState::side)
template <typename F>
int side_run(TrainState* ts, cudaStream_t s, F&& fn) {
  static const int enabled = getenv("ISHARA_WGRAD_SIDE") ? atoi(getenv("ISHARA_WGRAD_SIDE")) : 1;
  if (!enabled) return fn(s);
  if (ts->side == nullptr) {
    ISHARA_CUDA_OK(cudaStreamCreateWithFlags(&ts->side, cudaStreamNonBlocking));
    ISHARA_CUDA_OK(cudaEventCreateWithFlags(&ts->ev_fork, cudaEventDisableTiming));
    ISHARA_CUDA_OK(cudaEventCreateWithFlags(&ts->ev_side, cudaEventDisableTiming));
  }
  ISHARA_CUDA_OK(cudaEventRecord(ts->ev_fork, s));
  ISHARA_CUDA_OK(cudaStreamWaitEvent(ts->side, ts->ev_fork, 0));
  const int rc = fn(ts->side);
  ISHARA_CUDA_OK(cudaEventRecord(ts->ev_side, ts->side));
  ts->side_pending = true;
  ts->last_was_side = true;
  return rc;
}
inline int side_join(TrainState* ts, cudaStream_t s) {
  if (ts->side_pending) {
    ISHARA_CUDA_OK(cudaStreamWaitEvent(s, ts->ev_side, 0));
    ts->side_pending = false;
  }
  return 0;
}

namespace {

__global__ void mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) acc += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) *out = acc / static_cast<float>(n);
}

bool is_trainable(const std::string& name) {
  auto ends = [&](const char* s) {
    const size_t n = std::strlen(s);
    return name.size() >= n && name.compare(name.size() - n, n, s) == 0;
  };
  return !(ends(".moving_mean") || ends(".moving_variance"));
}

int wide_bn(int n) { return n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : 64); }

void free_program(TrainState* ts) {
  for (void* p : ts->allocs) cudaFree(p);
  ts->allocs.clear();
  ts->fwd.clear();
  ts->bwd.clear();
  ts->named.clear();
  ts->named_grad.clear();
  ts->mod_hi.clear();
  ts->bucket_lo.clear();
  ts->bucket_up.clear();
  ts->builder.reset();
  ts->batch = 0;
  ts->stats = nullptr;
  ts->stats_count = 0;
  if (ts->step_graph != nullptr) { cudaGraphExecDestroy(ts->step_graph); ts->step_graph = nullptr; }
  ts->steps_since_build = 0;
  ts->step_graph_broken = false;
}

struct TB {
  ishara_model* m = nullptr;
  TrainState* ts = nullptr;
  int B = 0, T = 0, D = 0, E = 0, M = 0;
  int rc = 0;
  int module_index = 0;
  bf16* gS[2] = {nullptr, nullptr};                 // residual-stream gradient ping-pong [M, D]
  bf16 *gW1 = nullptr, *gW2 = nullptr;              // wide gradient scratch [M, max(3D, 2D, E)]
  bf16 *gD1 = nullptr, *gD2 = nullptr, *tD = nullptr;  // [M, D] scratch
  float *lse = nullptr, *dsum = nullptr;            // attention backward scratch [B*H*T]
  size_t stats_used = 0;

  template <typename Tp>
  Tp* alloc(size_t count) {
    if (rc) return nullptr;
    void* d = nullptr;
    if (cudaMalloc(&d, count * sizeof(Tp) + 256) != cudaSuccess) {
      set_last_error("train: cudaMalloc of " + std::to_string(count * sizeof(Tp)) + " bytes failed");
      cudaGetLastError();
      rc = ISHARA_ERR_CUDA;
      return nullptr;
    }
    ts->allocs.push_back(d);
    return static_cast<Tp*>(d);
  }
  bf16* act(int cols, const std::string& name = "") {
    bf16* p = alloc<bf16>(static_cast<size_t>(M) * cols);
    if (!name.empty()) ts->named[name] = NamedBuf{p, M, cols, cols, 0};
    return p;
  }
  float* f32(size_t n) { return alloc<float>(n); }
  // fp64 statistics live in one arena so a single memset clears them per step; returned as an offset handle
  size_t stat(size_t n) {
    const size_t o = stats_used;
    stats_used += n;
    return o;
  }

  int pidx(const std::string& name) {
    auto it = m->index.find(name);
    if (it == m->index.end()) {
      set_last_error("train: unknown parameter " + name);
      rc = ISHARA_ERR_INVALID;
      return 0;
    }
    return it->second;
  }
  float* W(const std::string& name) { return ts->theta + ts->off[pidx(name)]; }
  int64_t touch_hi = 0;  // end offset of the highest gradient the current module's backward writes
  float* G(const std::string& name) {
    const int i = pidx(name);
    touch_hi = std::max(touch_hi, ts->off[i] + m->params[i].numel());
    return ts->grad + ts->off[i];
  }
  const bf16* WF(const std::string& base) {
    auto it = ts->wf.find(base);
    if (it == ts->wf.end()) { set_last_error("train: no forward pack for " + base); rc = ISHARA_ERR_INVALID; return nullptr; }
    return it->second;
  }
  const bf16* WB(const std::string& base) {
    auto it = ts->wb.find(base);
    if (it == ts->wb.end()) { set_last_error("train: no backward pack for " + base); rc = ISHARA_ERR_INVALID; return nullptr; }
    return it->second;
  }

  // ---- step factories ------------------------------------------------------------------------------
  // out[M, nout] = epi( A[M,K] @ Wt[N,K]^T )
  Step gemm(const bf16* A, int K, const bf16* Wt, int N, void* out, int nout, bool out_f32, bool row_mode, GemmEpi epi) {
    auto plan = std::make_shared<GemmPlan>();
    GemmPlan& p = *plan;
    p.M = M; p.N = N; p.K = K;
    p.out_f32 = out_f32;
    // ISHARA_TRAIN_BN: tile shape of the D-wide (N == 256) products. 0 = full-row kernel (192 tiles at B = 64: 1.3 waves
    // on 148 SMs), 128 = plain kernel with 128-column tiles (384 tiles: better tail), 256 = plain kernel, full width.
    static const int train_bn = getenv("ISHARA_TRAIN_BN") ? atoi(getenv("ISHARA_TRAIN_BN")) : 0;
    if (row_mode && N == 256 && (train_bn == 128 || train_bn == 256)) row_mode = false;
    p.row_mode = row_mode;
    p.block_n = out_f32 ? 64 : (row_mode ? N : (N == 256 && train_bn == 128 ? 128 : wide_bn(N)));
    epi.rows_per_seq = T;
    p.epi = epi;
    if (!rc) rc = gemm_plan_init(&p, A, K, Wt, out, nout, nout, nullptr, 0);
    const int sms = m->num_sms;
    return [plan, sms](cudaStream_t s) { return gemm_launch(*plan, sms, s); };
  }
  // full-row epilogue only for the D-wide stream tensors (bias / residual / row table; no activation there)
  bool rowable(int n) const { return n == D && (n == 256 || n == 128); }
  // y = A @ W^T(fwd pack) + bias [+ resid]   -> [M, N] bf16
  Step linear(const bf16* A, int K, const std::string& base, bool bias, int N, bf16* out, const bf16* resid = nullptr, int act = ACT_NONE,
              const float* rowtab = nullptr) {
    GemmEpi e;
    e.bias = bias ? W(base + ".bias") : nullptr;
    e.resid = resid;
    e.ld_resid = N;
    e.act = act;
    e.rowtab = rowtab;
    return gemm(A, K, WF(base), N, out, N, false, rowable(N), e);
  }
  // dx[M, I] = dy[M, O] @ W[I,O]^T(bwd pack) [+ resid]
  Step dgrad(const bf16* dy, int O, const std::string& base, int I, bf16* dx, const bf16* resid = nullptr) {
    GemmEpi e;
    e.resid = resid;
    e.ld_resid = I;
    return gemm(dy, O, WB(base), I, dx, I, false, rowable(I), e);
  }
  // dW += X^T Gy  (and, with_bias, dbias += column sums of Gy in the same launch)
  Step wgrad(const bf16* X, int I, int Ivalid, const bf16* Gy, int O, int Ovalid, const std::string& base, bool with_bias = false) {
    float* dW = G(base + ".kernel");
    float* db = with_bias ? G(base + ".bias") : nullptr;
    const int sms = m->num_sms;
    const int64_t MM = M;
    static const int use_tc = getenv("ISHARA_WGRAD_TC") ? atoi(getenv("ISHARA_WGRAD_TC")) : 1;
    if (use_tc && O % 64 == 0) {
      // tcgen05 path: MN-major operands via TMA; the bias gradient rides along (idle epilogue warps sum the G tiles)
      auto plan = std::make_shared<WgradTcPlan>();
      if (!rc) rc = wgrad_tc_plan_init(plan.get(), X, I, Gy, O, MM, I, O, sms);
      TrainState* t = ts;
      return [=](cudaStream_t s) {
        return side_run(t, s, [&](cudaStream_t q) { return wgrad_tc_launch(plan.get(), dW, Ovalid, db, Ivalid, Ovalid, q); });
      };
    }
    TrainState* t = ts;
    return [=](cudaStream_t s) {
      return side_run(t, s, [&](cudaStream_t q) { return wgrad_launch(X, I, Gy, O, dW, Ovalid, db, MM, I, O, Ivalid, Ovalid, sms, q); });
    };
  }
  Step ln_fwd(const bf16* x, bf16* out, const std::string& base, float eps) {
    const float *g = W(base + ".gamma"), *b = W(base + ".beta");
    const int64_t MM = M;
    const int DD = D;
    return [=](cudaStream_t s) { return layernorm_launch(x, out, g, b, eps, MM, DD, s); };
  }
  Step ln_bwd(const bf16* dy, const bf16* x, const std::string& base, float eps, const bf16* dresid, bf16* dx) {
    const float* g = W(base + ".gamma");
    float *dg = G(base + ".gamma"), *db = G(base + ".beta");
    const int64_t MM = M;
    const int DD = D;
    return [=](cudaStream_t s) { return ln_bwd_launch(dy, x, g, eps, dresid, dx, dg, db, MM, DD, s); };
  }
  // gradient snapshot for the debug interface (copies dS so later modules may overwrite the ping-pong buffer)
  void snap(std::vector<Step>& steps, const std::string& name, const bf16* src, int cols) {
    if (!ts->debug) return;
    bf16* dst = alloc<bf16>(static_cast<size_t>(M) * cols);
    ts->named_grad[name] = NamedBuf{dst, M, cols, cols, 0};
    const size_t bytes = static_cast<size_t>(M) * cols * sizeof(bf16);
    steps.push_back([=](cudaStream_t s) {
      return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s) == cudaSuccess ? 0 : 3;
    });
  }
  // forward: S_out = resid + drop(A @ W + bias). Returns the dropout site id (0 = none) for the backward side.
  uint32_t site_counter = 0;
  uint32_t branch_fwd(const bf16* A, int K, const std::string& base, bool bias, const bf16* resid, bf16* S_out, float p, bool per_sample) {
    if (p <= 0.f) {
      ts->fwd.push_back(linear(A, K, base, bias, D, S_out, resid));
      return 0;
    }
    const uint32_t site = ++site_counter;
    ts->fwd.push_back(linear(A, K, base, bias, D, tD));
    const bf16* src = tD;
    const int64_t MM = M;
    const int DD = D, TT = T;
    const uint64_t* keyp = ts->keys_dev + site;
    ts->fwd.push_back([=](cudaStream_t s) { return dropout_keyed_launch(src, resid, S_out, MM, DD, TT, p, keyp, per_sample ? 1 : 0, s); });
    return site;
  }
  // backward: returns the pointer holding dY = drop'(dOut) and appends the step producing it (if any)
  const bf16* branch_bwd(std::vector<Step>& steps, const bf16* dOut, float p, uint32_t site, bool per_sample) {
    if (site == 0) return dOut;
    bf16* dst = tD;
    const int64_t MM = M;
    const int DD = D, TT = T;
    const uint64_t* keyp = ts->keys_dev + site;
    steps.push_back([=](cudaStream_t s) { return dropout_keyed_launch(dOut, nullptr, dst, MM, DD, TT, p, keyp, per_sample ? 1 : 0, s); });
    return dst;
  }
  // in-place elementwise dropout of a wide tensor (same mask in forward and backward)
  uint32_t drop_inplace(std::vector<Step>& steps, bf16* x, int cols, float p, uint32_t site_in = 0) {
    if (p <= 0.f) return 0;
    const uint32_t site = site_in ? site_in : ++site_counter;
    const int64_t MM = M;
    const int TT = T;
    const uint64_t* keyp = ts->keys_dev + site;
    steps.push_back([=](cudaStream_t s) { return dropout_keyed_launch(x, nullptr, x, MM, cols, TT, p, keyp, 0, s); });
    return site;
  }

  bf16* din() const { return gS[module_index & 1]; }
  bf16* dout() const { return gS[(module_index + 1) & 1]; }
  void finish_module(std::vector<Step>& steps, const std::string& in_name) {
    // steps were appended in execution order for this module's backward; the global list runs modules in reverse
    if (ts->debug && !in_name.empty()) snap(steps, in_name, din(), D);
    auto packed = std::make_shared<std::vector<Step>>(std::move(steps));
    TrainState* t = ts;
    ts->bwd.push_back([packed, t](cudaStream_t s) {
      for (auto& st : *packed) {
        t->last_was_side = false;
        int r = st(s);
        // a weight gradient went to the side stream: the step after it (the data gradient of the same dY) runs beside
        // it, then the main stream waits - nothing later may overwrite the dY the side stream is still reading
        if (!r && !t->last_was_side) r = side_join(t, s);
        if (r) return r;
      }
      return side_join(t, s);  // module boundary: gradients final for the bucket exchange, scratch buffers reusable
    });
    ts->mod_hi.push_back(touch_hi);
    touch_hi = 0;
    ++module_index;
  }

  // ---- modules -------------------------------------------------------------------------------------
  struct BnSite {
    float *mean, *rstd, *scale, *shift, *coef;
    size_t sum, sumsq, sum1, sum2;
  };
  BnSite bn_site(int C) {
    BnSite b;
    b.mean = f32(C); b.rstd = f32(C); b.scale = f32(C); b.shift = f32(C); b.coef = f32(2 * static_cast<size_t>(C));
    b.sum = stat(C); b.sumsq = stat(C); b.sum1 = stat(C); b.sum2 = stat(C);
    return b;
  }
  Step bn_stats_step(const bf16* x, float* seqsum, const BnSite& bs, int C, const std::string& base, float momentum) {
    const float *g = W(base + ".gamma"), *bt = W(base + ".beta");
    float *mm = W(base + ".moving_mean"), *mv = W(base + ".moving_variance");
    TrainState* t = ts;
    const int BB = B, TT = T;
    const double count = static_cast<double>(M);
    return [=](cudaStream_t s) {
      int r = colstats_launch(x, seqsum, t->stats + bs.sum, t->stats + bs.sumsq, BB, TT, C, s);
      if (r) return r;
      return bn_finalize_launch(t->stats + bs.sum, t->stats + bs.sumsq, count, g, bt, 1e-3f, momentum, mm, mv, bs.mean, bs.rstd, bs.scale,
                                bs.shift, C, s);
    };
  }
  Step bn_bwd_step(const bf16* dG, const bf16* x, const float* sgate, const float* dm, const BnSite& bs, int C, const std::string& base,
                   bf16* dx) {
    const float* g = W(base + ".gamma");
    float *dg = G(base + ".gamma"), *db = G(base + ".beta");
    TrainState* t = ts;
    const int BB = B, TT = T;
    const double count = static_cast<double>(M);
    const float invT = 1.f / static_cast<float>(T);
    return [=](cudaStream_t s) {
      int r = bn_bwd_reduce_launch(dG, x, sgate, dm, invT, bs.mean, bs.rstd, t->stats + bs.sum1, t->stats + bs.sum2, BB, TT, C, s);
      if (r) return r;
      return bn_bwd_apply_launch(dG, x, sgate, dm, invT, bs.mean, bs.rstd, g, t->stats + bs.sum1, t->stats + bs.sum2, count, bs.coef, dx, dg, db,
                                 BB, TT, C, s);
    };
  }

  bf16* stem(const float* x_dev) {
    const ishara_config_t& c = m->cfg;
    const int fpad = m->fpad();
    bf16* XIN = act(fpad);
    bf16* Z = act(D, "stem.z");
    bf16* S0 = act(D, "stem");
    float* pe = f32(static_cast<size_t>(T) * D);
    {
      // positional_encoding c5:226-235: [sin | cos] halves, computed in fp32 like the reference
      std::vector<float> tab(static_cast<size_t>(T) * D);
      const int half = D / 2;
      const float depth = static_cast<float>(D) / 2.f;
      for (int t = 0; t < T; ++t)
        for (int i = 0; i < half; ++i) {
          const float ang = static_cast<float>(t) * (1.f / powf(10000.f, static_cast<float>(i) / depth));
          tab[static_cast<size_t>(t) * D + i] = sinf(ang);
          tab[static_cast<size_t>(t) * D + half + i] = cosf(ang);
        }
      if (!rc && cudaMemcpy(pe, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) rc = ISHARA_ERR_CUDA;
    }
    float* seqsum = f32(static_cast<size_t>(B) * D);
    BnSite bs = bn_site(D);
    const int64_t MM = M;
    const int F = c.features;
    ts->fwd.push_back([=](cudaStream_t s) { return cast_pad_launch(x_dev, XIN, MM, F, fpad, s); });
    ts->fwd.push_back(linear(XIN, fpad, "stem_conv", false, D, Z, nullptr, ACT_NONE, pe));
    ts->fwd.push_back(bn_stats_step(Z, seqsum, bs, D, "stem_bn", 0.95f));  // BatchNormalization(momentum=0.95, name='stem_bn') c7:17
    {
      const int DD = D, TT = T;
      ts->fwd.push_back([=](cudaStream_t s) { return affine_gate_add_launch(Z, bs.scale, bs.shift, nullptr, nullptr, S0, MM, DD, TT, s); });
    }
    std::vector<Step> bw;
    snap(bw, "stem", dout(), D);
    bw.push_back(bn_bwd_step(dout(), Z, nullptr, nullptr, bs, D, "stem_bn", gD1));
    snap(bw, "stem.z", gD1, D);
    bw.push_back(wgrad(XIN, fpad, c.features, gD1, D, D, "stem_conv"));
    finish_module(bw, "");
    return S0;
  }

  bf16* conv1d_block(const std::string& n, int k, const bf16* S_in) {
    const int C = 2 * D;
    bf16* Eb = act(C, n + ".e");
    bf16* Dc = act(C, n + ".d");
    bf16* Gb = act(C, n + ".g");
    bf16* S_out = act(D, n);
    float* seqsum = f32(static_cast<size_t>(B) * C);
    float* mmean = f32(static_cast<size_t>(B) * C);
    float* sgate = f32(static_cast<size_t>(B) * C);
    float* ds = f32(static_cast<size_t>(B) * C);
    float* dmm = f32(static_cast<size_t>(B) * C);
    BnSite bs = bn_site(C);
    const float* dww = W(n + "_dwconv.depthwise_kernel");
    const float* w5 = W(n + "_eca.kernel");
    const int BB = B, TT = T;
    const int64_t MM = M;
    const float invT = 1.f / static_cast<float>(T);
    ts->fwd.push_back(linear(S_in, D, n + "_expand_conv", true, C, Eb));
    {
      DwTrainArgs a;
      a.in = Eb; a.out = Dc; a.w = dww; a.B = B; a.T = T; a.C = C; a.k = k; a.pad_left = k - 1; a.pre_act = TACT_SWISH;
      ts->fwd.push_back([=](cudaStream_t s) { return dw_train_launch(a, s); });
    }
    ts->fwd.push_back(bn_stats_step(Dc, seqsum, bs, C, n + "_bn", 0.95f));
    ts->fwd.push_back([=](cudaStream_t s) { return eca_fwd_launch(seqsum, bs.scale, bs.shift, invT, w5, mmean, sgate, BB, C, s); });
    ts->fwd.push_back([=](cudaStream_t s) { return affine_gate_add_launch(Dc, bs.scale, bs.shift, sgate, nullptr, Gb, MM, C, TT, s); });
    const float p = ts->dropout;
    const uint32_t site = branch_fwd(Gb, C, n + "_project_conv", true, S_in, S_out, p, true);

    std::vector<Step> bw;
    const bf16* dOut = dout();
    snap(bw, n, dOut, D);
    const bf16* dY = branch_bwd(bw, dOut, p, site, true);
    bw.push_back(wgrad(Gb, C, C, dY, D, D, n + "_project_conv", true));
    bw.push_back(dgrad(dY, D, n + "_project_conv", C, gW1));  // dG
    snap(bw, n + ".g", gW1, C);
    {
      const bf16* g1 = gW1;
      bw.push_back([=](cudaStream_t s) { return seq_dot_launch(g1, Dc, bs.scale, bs.shift, ds, BB, TT, C, s); });
    }
    {
      float* dw5 = G(n + "_eca.kernel");
      bw.push_back([=](cudaStream_t s) { return eca_bwd_launch(ds, sgate, mmean, w5, dmm, dw5, BB, C, s); });
    }
    bw.push_back(bn_bwd_step(gW1, Dc, sgate, dmm, bs, C, n + "_bn", gW2));  // dDc
    snap(bw, n + ".d", gW2, C);
    {
      float* ddw = G(n + "_dwconv.depthwise_kernel");
      bf16* g2 = gW2;
      bw.push_back([=](cudaStream_t s) { return dw_wgrad_launch(g2, Eb, TACT_SWISH, ddw, nullptr, BB, TT, C, k, k - 1, s); });
      DwTrainArgs a;
      a.in = gW2; a.out = gW1; a.w = dww; a.B = B; a.T = T; a.C = C; a.k = k; a.pad_left = 0; a.flip = 1; a.mul_ref = Eb;
      bw.push_back([=](cudaStream_t s) { return dw_train_launch(a, s); });  // dE
    }
    snap(bw, n + ".e", gW1, C);
    bw.push_back(wgrad(S_in, D, D, gW1, C, C, n + "_expand_conv", true));
    bw.push_back(dgrad(gW1, C, n + "_expand_conv", D, din(), dOut));
    finish_module(bw, "");
    return S_out;
  }

  // x + drop(FFN(LN(x)))    (FeedForwardModule c5:237-247; Squeezeformer FFN c5:162-166)
  // branch_drop: SqueezeformerBlock wraps the branch in self.dropout (c5:190,205); ConformerBlock does not (c5:326,341)
  bf16* ffn(const std::string& base, const std::string& ln, const bf16* S_in, const std::string& out_name, bool branch_drop) {
    bf16* XN = act(D);
    bf16* U = act(E, base + ".u");
    bf16* Hh = act(E);
    bf16* S_out = act(D, out_name);
    const float p = ts->dropout, pb = branch_drop ? ts->dropout : 0.f;
    const int64_t nE = static_cast<int64_t>(M) * E;
    ts->fwd.push_back(ln_fwd(S_in, XN, ln, 1e-6f));
    ts->fwd.push_back(linear(XN, D, base + ".0", true, E, U));
    // swish and the inner dropout in one pass over the [M, E] tensor (forward and backward)
    const uint32_t site_h = p > 0.f ? ++site_counter : 0;
    const uint64_t* keyp_h = ts->keys_dev + site_h;
    if (site_h) ts->fwd.push_back([=](cudaStream_t s) { return act_drop_fwd_launch(U, Hh, nE, TACT_SWISH, p, keyp_h, s); });
    else ts->fwd.push_back([=](cudaStream_t s) { return act_fwd_launch(U, Hh, nE, TACT_SWISH, s); });
    const uint32_t site = branch_fwd(Hh, E, base + ".2", true, S_in, S_out, pb, false);

    std::vector<Step> bw;
    const bf16* dOut = dout();
    snap(bw, out_name, dOut, D);
    const bf16* dY = branch_bwd(bw, dOut, pb, site, false);
    bw.push_back(wgrad(Hh, E, E, dY, D, D, base + ".2", true));
    bw.push_back(dgrad(dY, D, base + ".2", E, gW1));  // dH
    {
      bf16* g1 = gW1;
      if (site_h) bw.push_back([=](cudaStream_t s) { return act_bwd_drop_launch(g1, U, g1, nE, TACT_SWISH, p, keyp_h, s); });  // dU
      else bw.push_back([=](cudaStream_t s) { return act_bwd_launch(g1, U, g1, nE, TACT_SWISH, s); });
    }
    snap(bw, base + ".u", gW1, E);
    bw.push_back(wgrad(XN, D, D, gW1, E, E, base + ".0", true));
    bw.push_back(dgrad(gW1, E, base + ".0", D, gD1));  // dXN
    bw.push_back(ln_bwd(gD1, S_in, ln, 1e-6f, dOut, din()));
    finish_module(bw, "");
    return S_out;
  }

  // x + drop(MHSA(LN(x)))   (MultiHeadSelfAttention c5:91-118)
  // p_attn: dropout on the attention probabilities (c5:113): SqueezeformerBlock passes its dropout, ConformerBlock its
  // own attn_dropout default 0.1 (c5:312,315)
  bf16* mhsa(const std::string& base, const std::string& ln, const bf16* S_in, const std::string& out_name, bool branch_drop, float p_attn) {
    const int H = m->cfg.num_heads, dh = D / H;
    bf16* XN = act(D);
    bf16* QKV = act(3 * D, base + ".qkv");
    bf16* O = act(D, base + ".o");
    bf16* S_out = act(D, out_name);
    float* lse_save = f32(static_cast<size_t>(B) * H * T);  // row log-sum-exp of this layer's scores, forward -> backward
    const float p = branch_drop ? ts->dropout : 0.f;
    const float scale = 1.f / std::sqrt(static_cast<float>(D));  // self.scale = dim ** -0.5 (c5:95)
    const uint32_t site_a = p_attn > 0.f ? ++site_counter : 0;
    const uint32_t thr_a = p_attn > 0.f ? dropout_thr16(p_attn) : 0;
    const float inv_a = p_attn > 0.f ? 1.f / (1.f - p_attn) : 1.f;
    const uint64_t* keyp_a = ts->keys_dev + site_a;
    ts->fwd.push_back(ln_fwd(S_in, XN, ln, 1e-6f));
    ts->fwd.push_back(linear(XN, D, base + ".qkv", false, 3 * D, QKV));
    {
      AttnArgs a;
      a.qkv = QKV; a.out = O; a.B = B; a.T = T; a.H = H; a.dh = dh; a.scale = scale;
      a.drop_thr16 = thr_a; a.drop_inv_keep = inv_a;
      a.lse_out = lse_save;
      ts->fwd.push_back([=](cudaStream_t s) {
        AttnArgs aa = a;
        if (site_a) aa.drop_key_ptr = keyp_a;
        return attention_launch(aa, s);
      });
    }
    const uint32_t site = branch_fwd(O, D, base + ".proj", false, S_in, S_out, p, false);

    std::vector<Step> bw;
    const bf16* dOut = dout();
    snap(bw, out_name, dOut, D);
    const bf16* dY = branch_bwd(bw, dOut, p, site, false);
    bw.push_back(wgrad(O, D, D, dY, D, D, base + ".proj"));
    bw.push_back(dgrad(dY, D, base + ".proj", D, gD1));  // dO
    snap(bw, base + ".o", gD1, D);
    {
      AttnBwdArgs a;
      a.qkv = QKV; a.o = O; a.dO = gD1; a.dqkv = gW1; a.lse2 = lse_save; a.have_lse = 1; a.dsum = dsum; a.B = B; a.T = T; a.H = H; a.dh = dh; a.scale = scale;
      a.drop_thr16 = thr_a; a.drop_inv_keep = inv_a;
      bw.push_back([=](cudaStream_t s) {
        AttnBwdArgs aa = a;
        if (site_a) aa.drop_key_ptr = keyp_a;
        return attention_bwd_launch(aa, s);
      });
    }
    snap(bw, base + ".qkv", gW1, 3 * D);
    bw.push_back(wgrad(XN, D, D, gW1, 3 * D, 3 * D, base + ".qkv"));
    bw.push_back(dgrad(gW1, 3 * D, base + ".qkv", D, gD2));  // dXN
    bw.push_back(ln_bwd(gD2, S_in, ln, 1e-6f, dOut, din()));
    finish_module(bw, "");
    return S_out;
  }

  // Squeezeformer ConvModule + SqueezeExcite (c5:120-153): x + SE(conv3(swish(dw(swish(conv1(LN x))))))
  bf16* sqz_conv(const std::string& n, const bf16* S_in) {
    const int tk = m->cfg.transformer_kernel_size, R = std::max(1, D / 8);
    bf16* XN = act(D);
    bf16* C1 = act(E, n + ".conv.c1");
    bf16* D2 = act(E, n + ".conv.d2");
    bf16* H2 = act(E);
    bf16* Z = act(D, n + ".conv.z");
    bf16* S_out = act(D, n + ".x3");
    float* zsum = f32(static_cast<size_t>(B) * D);
    float* dgate = f32(static_cast<size_t>(B) * D);
    SeTrainArgs se;
    se.zsum = zsum;
    se.fc1_w = W(n + ".conv.se.fc1.kernel"); se.fc1_b = W(n + ".conv.se.fc1.bias");
    se.fc2_w = W(n + ".conv.se.fc2.kernel"); se.fc2_b = W(n + ".conv.se.fc2.bias");
    se.g = f32(static_cast<size_t>(B) * D); se.a_pre = f32(static_cast<size_t>(B) * R); se.gate = f32(static_cast<size_t>(B) * D);
    se.dgate = dgate; se.dg = f32(static_cast<size_t>(B) * D);
    se.d_fc1_w = G(n + ".conv.se.fc1.kernel"); se.d_fc1_b = G(n + ".conv.se.fc1.bias");
    se.d_fc2_w = G(n + ".conv.se.fc2.kernel"); se.d_fc2_b = G(n + ".conv.se.fc2.bias");
    se.B = B; se.D = D; se.R = R; se.invT = 1.f / static_cast<float>(T);
    const float* dww = W(n + ".conv.conv2.depthwise_kernel");
    const int BB = B, TT = T, DD = D, EE = E;
    const int64_t MM = M, nE = static_cast<int64_t>(M) * E;
    ts->fwd.push_back(ln_fwd(S_in, XN, n + ".conv.norm", 1e-6f));
    ts->fwd.push_back(linear(XN, D, n + ".conv.conv1", true, E, C1));
    {
      DwTrainArgs a;
      a.in = C1; a.out = D2; a.w = dww; a.B = B; a.T = T; a.C = E; a.k = tk; a.pad_left = tk - 1; a.pre_act = TACT_SWISH;
      ts->fwd.push_back([=](cudaStream_t s) { return dw_train_launch(a, s); });
    }
    ts->fwd.push_back([=](cudaStream_t s) { return act_fwd_launch(D2, H2, nE, TACT_SWISH, s); });
    ts->fwd.push_back(linear(H2, E, n + ".conv.conv3", true, D, Z));
    ts->fwd.push_back([=](cudaStream_t s) { return colstats_launch(Z, zsum, nullptr, nullptr, BB, TT, DD, s); });
    ts->fwd.push_back([=](cudaStream_t s) { return se_fwd_launch(se, s); });
    ts->fwd.push_back([=](cudaStream_t s) { return affine_gate_add_launch(Z, nullptr, nullptr, se.gate, S_in, S_out, MM, DD, TT, s); });

    std::vector<Step> bw;
    const bf16* dOut = dout();
    snap(bw, n + ".x3", dOut, D);
    bw.push_back([=](cudaStream_t s) { return seq_dot_launch(dOut, Z, nullptr, nullptr, dgate, BB, TT, DD, s); });
    bw.push_back([=](cudaStream_t s) { return se_bwd_launch(se, s); });
    {
      bf16* g1 = gD1;
      bw.push_back([=](cudaStream_t s) { return gate_bias_launch(dOut, se.gate, se.dg, se.invT, g1, MM, DD, TT, s); });  // dZ
    }
    snap(bw, n + ".conv.z", gD1, D);
    bw.push_back(wgrad(H2, E, E, gD1, D, D, n + ".conv.conv3", true));
    bw.push_back(dgrad(gD1, D, n + ".conv.conv3", E, gW1));  // dH2
    {
      bf16* g1 = gW1;
      bw.push_back([=](cudaStream_t s) { return act_bwd_launch(g1, D2, g1, nE, TACT_SWISH, s); });  // dD2
      float* ddw = G(n + ".conv.conv2.depthwise_kernel");
      bw.push_back([=](cudaStream_t s) { return dw_wgrad_launch(g1, C1, TACT_SWISH, ddw, nullptr, BB, TT, EE, tk, tk - 1, s); });
      DwTrainArgs a;
      a.in = gW1; a.out = gW2; a.w = dww; a.B = B; a.T = T; a.C = E; a.k = tk; a.pad_left = 0; a.flip = 1; a.mul_ref = C1;
      bw.push_back([=](cudaStream_t s) { return dw_train_launch(a, s); });  // dC1
    }
    snap(bw, n + ".conv.c1", gW2, E);
    bw.push_back(wgrad(XN, D, D, gW2, E, E, n + ".conv.conv1", true));
    bw.push_back(dgrad(gW2, E, n + ".conv.conv1", D, gD1));  // dXN
    bw.push_back(ln_bwd(gD1, S_in, n + ".conv.norm", 1e-6f, dOut, din()));
    finish_module(bw, "");
    return S_out;
  }

  // Conformer ConvolutionModule (c5:249-309): LN_1e-3( pw2(BN(dw(GLU(pw1 x)))) + x )
  bf16* conf_conv(const std::string& n, const bf16* S_in) {
    const int tk = m->cfg.transformer_kernel_size, pl = (tk - 1) / 2;
    bf16* P1 = act(2 * D, n + ".conv.p1");
    bf16* GL = act(D, n + ".conv.gl");
    bf16* DWb = act(D, n + ".conv.dw");
    bf16* BNb = act(D, n + ".conv.bn");
    bf16* Rb = act(D, n + ".conv.r");
    bf16* S_out = act(D, n + ".x3");
    float* seqsum = f32(static_cast<size_t>(B) * D);
    BnSite bs = bn_site(D);
    const float* dww = W(n + ".conv.depthwise_conv.kernel");
    const float* dwb = W(n + ".conv.depthwise_conv.bias");
    const int BB = B, TT = T, DD = D;
    const int64_t MM = M;
    ts->fwd.push_back(linear(S_in, D, n + ".conv.pointwise_conv1", true, 2 * D, P1));
    ts->fwd.push_back([=](cudaStream_t s) { return glu_fwd_launch(P1, GL, MM, DD, s); });
    {
      DwTrainArgs a;
      a.in = GL; a.out = DWb; a.w = dww; a.bias = dwb; a.B = B; a.T = T; a.C = D; a.k = tk; a.pad_left = pl;
      ts->fwd.push_back([=](cudaStream_t s) { return dw_train_launch(a, s); });
    }
    ts->fwd.push_back(bn_stats_step(DWb, seqsum, bs, D, n + ".conv.batch_norm", 0.99f));
    ts->fwd.push_back([=](cudaStream_t s) { return affine_gate_add_launch(DWb, bs.scale, bs.shift, nullptr, nullptr, BNb, MM, DD, TT, s); });
    ts->fwd.push_back(linear(BNb, D, n + ".conv.pointwise_conv2", true, D, Rb, S_in));
    ts->fwd.push_back(ln_fwd(Rb, S_out, n + ".conv.layer_norm", 1e-3f));

    std::vector<Step> bw;
    const bf16* dOut = dout();
    snap(bw, n + ".x3", dOut, D);
    bw.push_back(ln_bwd(dOut, Rb, n + ".conv.layer_norm", 1e-3f, nullptr, gD1));  // dR
    snap(bw, n + ".conv.r", gD1, D);
    bw.push_back(wgrad(BNb, D, D, gD1, D, D, n + ".conv.pointwise_conv2", true));
    bw.push_back(dgrad(gD1, D, n + ".conv.pointwise_conv2", D, gD2));  // dBN
    bw.push_back(bn_bwd_step(gD2, DWb, nullptr, nullptr, bs, D, n + ".conv.batch_norm", gD2));  // dDW (in place)
    snap(bw, n + ".conv.dw", gD2, D);
    {
      float *ddw = G(n + ".conv.depthwise_conv.kernel"), *ddb = G(n + ".conv.depthwise_conv.bias");
      bf16* g2 = gD2;
      bw.push_back([=](cudaStream_t s) { return dw_wgrad_launch(g2, GL, TACT_NONE, ddw, ddb, BB, TT, DD, tk, pl, s); });
      DwTrainArgs a;
      a.in = gD2; a.out = gW2; a.w = dww; a.B = B; a.T = T; a.C = D; a.k = tk; a.pad_left = tk - 1 - pl; a.flip = 1;
      bw.push_back([=](cudaStream_t s) { return dw_train_launch(a, s); });  // dGL  (in gW2 viewed as [M, D])
    }
    {
      bf16 *g2 = gW2, *g1 = gW1;
      bw.push_back([=](cudaStream_t s) { return glu_bwd_launch(g2, P1, g1, MM, DD, s); });  // dP1 [M, 2D]
    }
    snap(bw, n + ".conv.p1", gW1, 2 * D);
    bw.push_back(wgrad(S_in, D, D, gW1, 2 * D, 2 * D, n + ".conv.pointwise_conv1", true));
    bw.push_back(dgrad(gW1, 2 * D, n + ".conv.pointwise_conv1", D, din(), gD1));  // + dR (residual path)
    finish_module(bw, "");
    return S_out;
  }

  void head(const bf16* S_in) {
    const ishara_config_t& c = m->cfg;
    const int V = c.num_classes, Vp = m->vpad(), C = 2 * D;
    bf16* HH = act(C, "head.h");
    bf16* dL = act(Vp);
    const float p = ts->dropout > 0.f ? 0.4f : 0.f;  // Dropout(0.4) c7:62
    const int64_t MM = M, nC = static_cast<int64_t>(M) * C;
    ts->fwd.push_back(linear(S_in, D, "top_conv", true, C, HH, nullptr, ACT_RELU));
    const uint32_t site = drop_inplace(ts->fwd, HH, C, p);
    {
      GemmEpi e;
      e.bias = W("classifier.bias");  // 60 floats; the 4 pad columns of the tile read past it into the next tensor and are clipped by the TMA store
      ts->fwd.push_back(gemm(HH, C, WF("classifier"), Vp, ts->logits, V, true, false, e));
    }
    ts->named["logits"] = NamedBuf{ts->logits, M, V, V, 1};

    std::vector<Step> bw;
    {
      TrainState* t = ts;
      const float alpha = 1.f / static_cast<float>(B);
      bw.push_back([=](cudaStream_t s) { return scale_cast_pad_launch(t->dlogits, dL, MM, V, Vp, alpha, s); });
    }
    bw.push_back(wgrad(HH, C, C, dL, Vp, V, "classifier", true));
    bw.push_back(dgrad(dL, Vp, "classifier", C, gW1));  // dHH
    {
      bf16* g1 = gW1;
      const uint64_t* keyp = ts->keys_dev + site;
      if (site) bw.push_back([=](cudaStream_t s) { return act_bwd_drop_launch(g1, HH, g1, nC, TACT_RELU, p, keyp, s); });
      else bw.push_back([=](cudaStream_t s) { return act_bwd_launch(g1, HH, g1, nC, TACT_RELU, s); });
    }
    snap(bw, "head.h", gW1, C);
    bw.push_back(wgrad(S_in, D, D, gW1, C, C, "top_conv", true));
    bw.push_back(dgrad(gW1, C, "top_conv", D, din()));
    finish_module(bw, "");
  }
};

int conv_k(const ishara_config_t& c, int j) { return c.kernel_sizes[j % c.num_kernel_sizes]; }

// ---- flat parameter storage + bf16 operand copies ------------------------------------------------------------
int upload_params(ishara_model* m, TrainState* ts) {
  std::vector<float> host(static_cast<size_t>(ts->n_total), 0.f);
  for (size_t i = 0; i < m->params.size(); ++i) {
    const Param& p = m->params[i];
    if (!p.set) { set_last_error("parameter not set: " + p.name); return ISHARA_ERR_STATE; }
    std::memcpy(host.data() + ts->off[i], p.data.data(), p.data.size() * sizeof(float));
  }
  ISHARA_CUDA_OK(cudaMemcpy(ts->theta, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

int init_storage(ishara_model* m, TrainState* ts) {
  const ishara_config_t& c = m->cfg;
  ts->off.assign(m->params.size(), 0);
  int64_t o = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (size_t i = 0; i < m->params.size(); ++i) {
      if (is_trainable(m->params[i].name) != (pass == 0)) continue;
      ts->off[i] = o;
      o += (m->params[i].numel() + 3) / 4 * 4;  // 16-byte aligned tensors (float4 loads, float2 atomics)
    }
    if (pass == 0) ts->n_train = o;
  }
  ts->n_total = o;
  const size_t bytes = static_cast<size_t>(o) * sizeof(float);
  ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ts->theta), bytes));
  ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ts->grad), bytes));
  ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ts->adam_m), bytes));
  ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ts->adam_v), bytes));
  ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ts->norm2), sizeof(double)));
  ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ts->skipped_dev), sizeof(int)));
  ISHARA_CUDA_OK(cudaMemset(ts->skipped_dev, 0, sizeof(int)));
  ISHARA_CUDA_OK(cudaMemset(ts->grad, 0, bytes));
  ISHARA_CUDA_OK(cudaMemset(ts->adam_m, 0, bytes));
  ISHARA_CUDA_OK(cudaMemset(ts->adam_v, 0, bytes));
  ISHARA_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&ts->loss_pinned), sizeof(float), cudaHostAllocDefault));
  int rc = upload_params(m, ts);
  if (rc) return rc;

  // bf16 operand copies of every dense kernel
  std::vector<RepackEntry> table;
  int max_tiles = 1;
  auto add = [&](const std::string& base, int I, int O, int Ipad, int Opad, bool need_bwd) -> int {
    const int idx = m->index.at(base + ".kernel");
    RepackEntry e;
    e.src = ts->theta + ts->off[idx];
    e.I = I; e.O = O; e.Ipad = Ipad; e.Opad = Opad;
    void* f = nullptr;
    ISHARA_CUDA_OK(cudaMalloc(&f, static_cast<size_t>(Opad) * Ipad * sizeof(bf16)));
    ts->wallocs.push_back(f);
    e.fwd = static_cast<bf16*>(f);
    ts->wf[base] = e.fwd;
    e.bwd = nullptr;
    if (need_bwd) {
      void* b = nullptr;
      ISHARA_CUDA_OK(cudaMalloc(&b, static_cast<size_t>(I) * Opad * sizeof(bf16)));
      ts->wallocs.push_back(b);
      e.bwd = static_cast<bf16*>(b);
      ts->wb[base] = e.bwd;
    }
    table.push_back(e);
    const int tiles = ((Ipad + 31) / 32) * ((Opad + 31) / 32);
    if (tiles > max_tiles) max_tiles = tiles;
    return 0;
  };
  const int D = c.dim, E = c.expansion_factor * c.dim;
  if ((rc = add("stem_conv", c.features, D, m->fpad(), D, false))) return rc;
  auto add_blocks = [&](const std::string& tag, int i) -> int {
    for (int j = 0; j < c.num_conv_per_block; ++j) {
      const std::string n = "conv" + tag + "_" + std::to_string(i) + "_" + std::to_string(j + 1);
      int r;
      if ((r = add(n + "_expand_conv", D, 2 * D, D, 2 * D, true))) return r;
      if ((r = add(n + "_project_conv", 2 * D, D, 2 * D, D, true))) return r;
    }
    return 0;
  };
  auto add_ffn = [&](const std::string& base) -> int {
    int r;
    if ((r = add(base + ".0", D, E, D, E, true))) return r;
    return add(base + ".2", E, D, E, D, true);
  };
  for (int i = 0; i < c.num_conv_squeeze_blocks; ++i) {
    if ((rc = add_blocks("squeeze", i))) return rc;
    const std::string n = "squeezeformer_" + std::to_string(i);
    if ((rc = add_ffn(n + ".ffn1"))) return rc;
    if ((rc = add_ffn(n + ".ffn2"))) return rc;
    if ((rc = add(n + ".mha.qkv", D, 3 * D, D, 3 * D, true))) return rc;
    if ((rc = add(n + ".mha.proj", D, D, D, D, true))) return rc;
    if ((rc = add(n + ".conv.conv1", D, E, D, E, true))) return rc;
    if ((rc = add(n + ".conv.conv3", E, D, E, D, true))) return rc;
  }
  for (int i = 0; i < c.num_conv_conform_blocks; ++i) {
    if ((rc = add_blocks("conform", i))) return rc;
    const std::string n = "conformer_" + std::to_string(i);
    if ((rc = add_ffn(n + ".ffn1"))) return rc;
    if ((rc = add_ffn(n + ".ffn2"))) return rc;
    if ((rc = add(n + ".mha.qkv", D, 3 * D, D, 3 * D, true))) return rc;
    if ((rc = add(n + ".mha.proj", D, D, D, D, true))) return rc;
    if ((rc = add(n + ".conv.pointwise_conv1", D, 2 * D, D, 2 * D, true))) return rc;
    if ((rc = add(n + ".conv.pointwise_conv2", D, D, D, D, true))) return rc;
  }
  if ((rc = add("top_conv", D, 2 * D, D, 2 * D, true))) return rc;
  if ((rc = add("classifier", 2 * D, c.num_classes, 2 * D, m->vpad(), true))) return rc;
  ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ts->repack_dev), table.size() * sizeof(RepackEntry)));
  ISHARA_CUDA_OK(cudaMemcpy(ts->repack_dev, table.data(), table.size() * sizeof(RepackEntry), cudaMemcpyHostToDevice));
  ts->repack_n = static_cast<int>(table.size());
  ts->repack_max_tiles = max_tiles;
  if ((rc = repack_launch(ts->repack_dev, ts->repack_n, ts->repack_max_tiles, m->stream))) return rc;
  ISHARA_CUDA_OK(cudaStreamSynchronize(m->stream));
  return 0;
}

int build_train_program(ishara_model* m, TrainState* ts, int batch, int labels_len) {
  free_program(ts);
  if (ts->keys_dev == nullptr) ISHARA_CUDA_OK(cudaMalloc(&ts->keys_dev, kMaxDropSites * sizeof(uint64_t)));
  const ishara_config_t& c = m->cfg;
  auto bp = std::make_shared<TB>();
  ts->builder = bp;
  TB& b = *bp;
  b.m = m; b.ts = ts; b.B = batch; b.T = c.frames; b.D = c.dim; b.E = c.expansion_factor * c.dim; b.M = batch * c.frames;
  const int D = c.dim, M = b.M, H = c.num_heads;
  const int wide = std::max(std::max(3 * D, 2 * D), b.E);
  b.gS[0] = b.act(D);
  b.gS[1] = b.act(D);
  b.gW1 = b.act(wide);
  b.gW2 = b.act(wide);
  b.gD1 = b.act(D);
  b.gD2 = b.act(D);
  b.tD = b.act(D);
  b.lse = b.f32(static_cast<size_t>(batch) * H * c.frames);
  b.dsum = b.f32(static_cast<size_t>(batch) * H * c.frames);
  ts->x_dev = b.f32(static_cast<size_t>(M) * c.features);
  ts->logits = b.f32(static_cast<size_t>(M) * c.num_classes);
  ts->dlogits = b.f32(static_cast<size_t>(M) * c.num_classes);
  ts->nll = b.f32(batch);
  ts->loss_dev = b.f32(4);
  ts->ctc_ws_bytes = ctc_workspace_bytes(batch, c.frames, labels_len);
  ts->ctc_ws = b.f32(ts->ctc_ws_bytes / sizeof(float));
  ts->labels_dev = b.alloc<int32_t>(static_cast<size_t>(batch) * labels_len);
  if (b.rc) return b.rc;

  const bf16* S = b.stem(ts->x_dev);
  auto conv_blocks = [&](const std::string& tag, int i) {
    for (int j = 0; j < c.num_conv_per_block; ++j)
      S = b.conv1d_block("conv" + tag + "_" + std::to_string(i) + "_" + std::to_string(j + 1), conv_k(c, j), S);
  };
  for (int i = 0; i < c.num_conv_squeeze_blocks && !b.rc; ++i) {
    conv_blocks("squeeze", i);
    const std::string n = "squeezeformer_" + std::to_string(i);
    S = b.ffn(n + ".ffn1", n + ".norm1", S, n + ".x1", true);
    S = b.mhsa(n + ".mha", n + ".norm2", S, n + ".x2", true, ts->dropout);
    S = b.sqz_conv(n, S);
    S = b.ffn(n + ".ffn2", n + ".norm3", S, n, true);
  }
  for (int i = 0; i < c.num_conv_conform_blocks && !b.rc; ++i) {
    conv_blocks("conform", i);
    const std::string n = "conformer_" + std::to_string(i);
    S = b.ffn(n + ".ffn1", n + ".layer_norm1", S, n + ".x1", false);
    S = b.mhsa(n + ".mha", n + ".layer_norm1", S, n + ".x2", false, ts->dropout > 0.f ? 0.1f : 0.f);  // layer_norm1 reused (c5:330)
    S = b.conf_conv(n, S);
    S = b.ffn(n + ".ffn2", n + ".layer_norm2", S, n, false);
  }
  if (!b.rc) b.head(S);
  if (b.rc) {
    free_program(ts);
    return b.rc;
  }
  ts->stats_count = b.stats_used;
  void* st = nullptr;
  ISHARA_CUDA_OK(cudaMalloc(&st, (ts->stats_count + 8) * sizeof(double)));
  ts->allocs.push_back(st);
  ts->stats = static_cast<double*>(st);
  ts->batch = batch;
  ts->labels_len = labels_len;
  {
    // exchange buckets (used only with a communicator): >= 1 M floats (4 MB) per all-reduce, so the ~30 MB of
    // gradients leave in a handful of NCCL calls spread over the backward pass
    const int n = static_cast<int>(ts->mod_hi.size());
    ts->bucket_lo.assign(n, 0);
    ts->bucket_up.assign(n, 0);
    static const long long min_elems = getenv("ISHARA_COMM_BUCKET") ? atoll(getenv("ISHARA_COMM_BUCKET")) : (1ll << 20);
    comm_bucket_plan(ts->mod_hi.data(), n, ts->n_train, min_elems, ts->bucket_lo.data(), ts->bucket_up.data());
  }
  return 0;
}

int ensure_train(ishara_model* m, int batch, int labels_len) {
  if (!m->finalized && m->train == nullptr) {
    // training needs the stream / device checks of finalize; the inference packs come for free
    int rc = model_finalize(m);
    if (rc) return rc;
  }
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  if (m->cfg.dim % 128 != 0 || m->cfg.dim > 512) {
    set_last_error("train: dim must be a multiple of 128, at most 512 (LayerNorm backward lane layout); got " + std::to_string(m->cfg.dim));
    return ISHARA_ERR_SHAPE;
  }
  if (m->train == nullptr) {
    auto ts = std::make_unique<TrainState>();
    int rc = init_storage(m, ts.get());
    if (rc) return rc;
    m->train = ts.release();
  }
  TrainState* ts = m->train;
  if (ts->batch != batch || ts->labels_len != labels_len || ts->fwd.empty()) {
    int rc = build_train_program(m, ts, batch, labels_len);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// entry points (wrapped by capi.cu)
// ------------------------------------------------------------------------------------------------
void train_destroy(ishara_model* m) {
  TrainState* ts = m->train;
  if (ts == nullptr) return;
  free_program(ts);
  for (void* p : ts->wallocs) cudaFree(p);
  cudaFree(ts->theta); cudaFree(ts->grad); cudaFree(ts->adam_m); cudaFree(ts->adam_v); cudaFree(ts->norm2);
  if (ts->skipped_dev) cudaFree(ts->skipped_dev);
  if (ts->slow) cudaFree(ts->slow);
  if (ts->repack_dev) cudaFree(ts->repack_dev);
  if (ts->loss_pinned) cudaFreeHost(ts->loss_pinned);
  if (ts->keys_dev) cudaFree(ts->keys_dev);
  if (ts->side) { cudaStreamSynchronize(ts->side); cudaStreamDestroy(ts->side); cudaEventDestroy(ts->ev_fork); cudaEventDestroy(ts->ev_side); }
  delete ts;
  m->train = nullptr;
}

// the captured step refers to the communicator's kernels (or to their absence): drop it whenever the exchange is set up or
// torn down - NCCL requires graphs that captured its collectives to be destroyed before the communicator
void train_drop_graph(ishara_model* m) {
  TrainState* ts = m->train;
  if (ts == nullptr) return;
  if (ts->step_graph != nullptr) {
    cudaDeviceSynchronize();
    cudaGraphExecDestroy(ts->step_graph);
    ts->step_graph = nullptr;
  }
  ts->steps_since_build = 0;
  ts->step_graph_broken = false;
}

// set_param after training started: the next train call re-uploads everything
void train_invalidate(ishara_model* m) {
  if (m->train != nullptr) train_destroy(m);
}

int train_configure(ishara_model* m, float dropout, uint64_t seed, int debug) {
  if (dropout < 0.f || dropout >= 1.f) { set_last_error("train_configure: dropout must be in [0, 1)"); return ISHARA_ERR_INVALID; }
  if (m->train == nullptr) {
    if (!m->finalized) { int rc = model_finalize(m); if (rc) return rc; }
    ISHARA_CUDA_OK(cudaSetDevice(m->device));
    if (m->cfg.dim % 128 != 0 || m->cfg.dim > 512) {
      set_last_error("train: dim must be a multiple of 128, at most 512 (LayerNorm backward lane layout); got " + std::to_string(m->cfg.dim));
      return ISHARA_ERR_SHAPE;
    }
    auto ts = std::make_unique<TrainState>();
    int rc = init_storage(m, ts.get());
    if (rc) return rc;
    m->train = ts.release();
  }
  TrainState* ts = m->train;
  if (ts->dropout != dropout || ts->debug != (debug != 0)) free_program(ts);  // dropout sites / taps are baked into the program
  ts->dropout = dropout;
  ts->seed_base = seed;
  ts->seed = seed;
  ts->fb_count = 0;
  ts->debug = debug != 0;
  return 0;
}

// Dropout seed of the n-th forward/backward after train_configure(seed): step 0 uses the configured seed itself (so a
// host can reproduce the masks of a freshly configured step from the seed alone), later steps mix the counter in.
// IsharaModel.dropout_masks(step=n) evaluates the same function.
uint64_t train_step_seed(uint64_t seed_base, int64_t n) {
  return n == 0 ? seed_base : seed_base ^ mix64(0x5eedull + static_cast<uint64_t>(n));
}

int train_counters(ishara_model* m, int64_t* fb_steps, int64_t* opt_steps, int64_t* skipped_steps) {
  if (m->train == nullptr) { set_last_error("train_counters: no training state"); return ISHARA_ERR_STATE; }
  if (fb_steps) *fb_steps = m->train->fb_count;
  if (opt_steps) *opt_steps = m->train->step;
  if (skipped_steps) {
    ISHARA_CUDA_OK(cudaSetDevice(m->device));
    ISHARA_CUDA_OK(cudaDeviceSynchronize());
    int sk = 0;
    ISHARA_CUDA_OK(cudaMemcpy(&sk, m->train->skipped_dev, sizeof(int), cudaMemcpyDeviceToHost));
    *skipped_steps = sk;
  }
  return 0;
}

int train_forward_backward_loss(ishara_model* m, float* loss_host, cudaStream_t stream);

// forward (training mode) + CTC + backward. x_dev fp32 [B,T,F]; labels_dev int32 [B,L] padded with blank.
int train_forward_backward(ishara_model* m, const float* x_dev, const int32_t* labels_dev, int batch, int labels_len, float* loss_host,
                           cudaStream_t stream) {
  if (x_dev == nullptr || labels_dev == nullptr || batch <= 0 || labels_len <= 0) { set_last_error("train: bad arguments"); return ISHARA_ERR_INVALID; }
  int rc = ensure_train(m, batch, labels_len);
  if (rc) return rc;
  TrainState* ts = m->train;
  const ishara_config_t& c = m->cfg;
  const size_t M = static_cast<size_t>(batch) * c.frames;
  ts->seed = train_step_seed(ts->seed_base, ts->fb_count);  // the step closures read ts->seed at launch time
  ++ts->fb_count;
  ISHARA_CUDA_OK(cudaMemcpyAsync(ts->x_dev, x_dev, M * c.features * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  ISHARA_CUDA_OK(cudaMemcpyAsync(ts->labels_dev, labels_dev, static_cast<size_t>(batch) * labels_len * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  // this step's dropout keys (one per site) go to the device table the kernels read: the launches below carry no
  // per-step host values and can be replayed from a graph
  if ((rc = dropout_keys_launch(ts->seed, ts->keys_dev, kMaxDropSites, stream))) return rc;
  const bool dp = m->comm != nullptr && m->comm_world > 1;
  // zero the accumulators, forward, CTC, mean loss, backward (with the gradient exchange when data-parallel)
  auto body = [&](cudaStream_t st) -> int {
    int r;
    ISHARA_CUDA_OK(cudaMemsetAsync(ts->grad, 0, static_cast<size_t>(ts->n_total) * sizeof(float), st));
    ISHARA_CUDA_OK(cudaMemsetAsync(ts->stats, 0, ts->stats_count * sizeof(double), st));
    for (auto& step : ts->fwd)
      if ((r = step(st))) { set_last_error(std::string("train forward: ") + get_last_error()); return r; }
    if ((r = ctc_loss_launch(ts->logits, ts->labels_dev, batch, c.frames, c.num_classes, labels_len, c.num_classes - 1, ts->nll, ts->dlogits, ts->ctc_ws, ts->ctc_ws_bytes, st)))
      return r;
    mean_kernel<<<1, 32, 0, st>>>(ts->nll, batch, ts->loss_dev);
    ISHARA_CUDA_OK(cudaGetLastError());
    note_launch();
    // the loss joins the exchange first (one float): nothing is read back to the host before the gradients are on their way
    if (dp && (r = comm_allreduce_after(m, ts->loss_dev, 1, st))) return r;
    for (int k = static_cast<int>(ts->bwd.size()) - 1; k >= 0; --k) {
      if ((r = ts->bwd[k](st))) { set_last_error(std::string("train backward: ") + get_last_error()); return r; }
      // everything in [bucket_lo[k], bucket_up[k]) is final now: sum it over the ranks on the communication stream while
      // the backward of the earlier modules keeps the SMs busy
      if (dp && ts->bucket_up[k] > ts->bucket_lo[k] &&
          (r = comm_allreduce_after(m, ts->grad + ts->bucket_lo[k], ts->bucket_up[k] - ts->bucket_lo[k], st)))
        return r;
    }
    if (dp && (r = comm_join(m, st))) return r;  // `st` continues only after every bucket has been reduced
    return 0;
  };
  // CUDA-graph replay: the ~490 launches of a step are captured on the second step of a program (every kernel has run
  // once: attributes set, lazy state built) and replayed afterwards - including, when data-parallel, the bucketed NCCL
  // all-reduces on the communication stream (forked / joined through events, so they become branches of the same graph).
  // Not in debug mode (snapshots).
  static const int use_graph = getenv("ISHARA_TRAIN_GRAPH") ? atoi(getenv("ISHARA_TRAIN_GRAPH")) : 1;
  static const int use_graph_dp = getenv("ISHARA_TRAIN_GRAPH_DP") ? atoi(getenv("ISHARA_TRAIN_GRAPH_DP")) : 1;  // NCCL all-reduces are captured with the step
  const bool graphable = use_graph && (!dp || use_graph_dp) && !ts->debug && !ts->step_graph_broken && m->stream != nullptr;
  bool done = false;
  if (graphable && ts->step_graph != nullptr) {
    ISHARA_CUDA_OK(cudaGraphLaunch(ts->step_graph, stream));
    note_launches(ts->step_graph_launches);
    done = true;
  } else if (graphable && ts->steps_since_build >= 1) {
    cudaGraph_t graph = nullptr;
    const uint64_t before = launch_count();
    if (cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      const int rcap = body(m->stream);
      const cudaError_t ce = cudaStreamEndCapture(m->stream, &graph);
      cudaGraphExec_t exec = nullptr;
      if (rcap == 0 && ce == cudaSuccess && graph != nullptr && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        ts->step_graph = exec;
        ts->step_graph_launches = static_cast<int>(launch_count() - before);
        cudaGraphDestroy(graph);
        ISHARA_CUDA_OK(cudaGraphLaunch(exec, stream));
        done = true;
      } else {
        if (graph != nullptr) cudaGraphDestroy(graph);
        cudaGetLastError();
        ts->step_graph_broken = true;  // capture is an optimisation: direct launches from now on
        ts->side_pending = false;
      }
    } else {
      cudaGetLastError();
      ts->step_graph_broken = true;
    }
  }
  if (!done && (rc = body(stream))) return rc;
  ++ts->steps_since_build;
  m->host_params_stale = true;  // BatchNorm moving statistics moved
  if (loss_host != nullptr) return train_forward_backward_loss(m, loss_host, stream);
  return 0;
}

// mean CTC loss of the last forward (stream sync)
int train_forward_backward_loss(ishara_model* m, float* loss_host, cudaStream_t stream) {
  if (m->train == nullptr || loss_host == nullptr) { set_last_error("train loss: no training state"); return ISHARA_ERR_STATE; }
  TrainState* ts = m->train;
  ISHARA_CUDA_OK(cudaMemcpyAsync(ts->loss_pinned, ts->loss_dev, sizeof(float), cudaMemcpyDeviceToHost, stream));
  ISHARA_CUDA_OK(cudaStreamSynchronize(stream));
  *loss_host = *ts->loss_pinned;
  if (m->comm != nullptr && m->comm_world > 1) *loss_host /= static_cast<float>(m->comm_world);  // summed over the ranks on the device
  return 0;
}

int train_grad_buffer(ishara_model* m, float** grad_dev, int64_t* numel) {
  if (m->train == nullptr) { set_last_error("train_grad_buffer: no training state (call train_forward_backward first)"); return ISHARA_ERR_STATE; }
  if (grad_dev) *grad_dev = m->train->grad;
  if (numel) *numel = m->train->n_train;
  return 0;
}

int train_apply(ishara_model* m, const ishara_adamw_t* opt, float grad_scale, cudaStream_t stream) {
  if (m->train == nullptr) { set_last_error("train_apply: no training state"); return ISHARA_ERR_STATE; }
  TrainState* ts = m->train;
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  AdamWArgs a;
  if (opt != nullptr) {
    a.lr = opt->lr; a.weight_decay = opt->weight_decay; a.beta1 = opt->beta1; a.beta2 = opt->beta2; a.eps = opt->eps; a.max_norm = opt->max_norm;
  }
  a.grad_scale = grad_scale;
  a.step = ++ts->step;
  a.skipped = ts->skipped_dev;
  int rc;
  ISHARA_CUDA_OK(cudaMemsetAsync(ts->norm2, 0, sizeof(double), stream));
  if ((rc = sqnorm_launch(ts->grad, ts->n_train, ts->norm2, stream))) return rc;
  if ((rc = adamw_launch(ts->theta, ts->grad, ts->adam_m, ts->adam_v, ts->n_train, ts->norm2, a, stream))) return rc;
  if ((rc = repack_launch(ts->repack_dev, ts->repack_n, ts->repack_max_tiles, stream))) return rc;
  m->host_params_stale = true;
  return 0;
}

// the reference's own optimiser (c7:68-69): RectifiedAdam(sma_threshold=4) wrapped in Lookahead(sync_period=5)
int train_apply_radam(ishara_model* m, const ishara_radam_lookahead_t* opt, float grad_scale, cudaStream_t stream) {
  if (m->train == nullptr) { set_last_error("train_apply_radam: no training state"); return ISHARA_ERR_STATE; }
  TrainState* ts = m->train;
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  RAdamArgs a;
  if (opt != nullptr) {
    a.lr = opt->lr; a.weight_decay = opt->weight_decay; a.beta1 = opt->beta1; a.beta2 = opt->beta2; a.eps = opt->eps;
    a.max_norm = opt->max_norm; a.sma_threshold = opt->sma_threshold; a.sync_period = opt->sync_period; a.slow_step = opt->slow_step_size;
  }
  if (ts->slow == nullptr) {
    // Lookahead's slow slot starts as a copy of the variables at the first step
    ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ts->slow), static_cast<size_t>(ts->n_total) * sizeof(float)));
    ISHARA_CUDA_OK(cudaMemcpyAsync(ts->slow, ts->theta, static_cast<size_t>(ts->n_total) * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  }
  a.grad_scale = grad_scale;
  a.step = ++ts->step;
  a.skipped = ts->skipped_dev;
  int rc;
  ISHARA_CUDA_OK(cudaMemsetAsync(ts->norm2, 0, sizeof(double), stream));
  if ((rc = sqnorm_launch(ts->grad, ts->n_train, ts->norm2, stream))) return rc;
  if ((rc = radam_lookahead_launch(ts->theta, ts->grad, ts->adam_m, ts->adam_v, ts->slow, ts->n_train, ts->norm2, a, stream))) return rc;
  if ((rc = repack_launch(ts->repack_dev, ts->repack_n, ts->repack_max_tiles, stream))) return rc;
  m->host_params_stale = true;
  return 0;
}

// ---- optimiser-state checkpoint (the reference saves weights every epoch, c9:10, and
// integration.py:912-958 saves optimiser state with them): flat fp32 slots in the order of the parameter table ----
int train_state_info(ishara_model* m, int64_t* numel, int64_t* opt_steps, int64_t* fb_steps, int32_t* has_slow) {
  if (m->train == nullptr) { set_last_error("train_state: no training state (call train_configure first)"); return ISHARA_ERR_STATE; }
  if (numel) *numel = m->train->n_total;
  if (opt_steps) *opt_steps = m->train->step;
  if (fb_steps) *fb_steps = m->train->fb_count;
  if (has_slow) *has_slow = m->train->slow != nullptr ? 1 : 0;
  return 0;
}
static float* state_slot(TrainState* ts, int which) {
  switch (which) {
    case 0: return ts->adam_m;
    case 1: return ts->adam_v;
    case 2: return ts->slow;
    case 3: return ts->theta;
  }
  return nullptr;
}
int train_state_get(ishara_model* m, int which, float* host_out, int64_t numel) {
  if (m->train == nullptr) { set_last_error("train_state_get: no training state"); return ISHARA_ERR_STATE; }
  TrainState* ts = m->train;
  float* src = state_slot(ts, which);
  if (src == nullptr) { set_last_error("train_state_get: slot not present (0 adam_m, 1 adam_v, 2 lookahead slow, 3 weights)"); return ISHARA_ERR_INVALID; }
  if (host_out == nullptr || numel != ts->n_total) { set_last_error("train_state_get: expected " + std::to_string(ts->n_total) + " elements"); return ISHARA_ERR_SHAPE; }
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  ISHARA_CUDA_OK(cudaDeviceSynchronize());
  ISHARA_CUDA_OK(cudaMemcpy(host_out, src, static_cast<size_t>(numel) * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}
int train_state_set(ishara_model* m, int which, const float* host_in, int64_t numel) {
  if (m->train == nullptr) { set_last_error("train_state_set: no training state"); return ISHARA_ERR_STATE; }
  TrainState* ts = m->train;
  if (which < 0 || which > 2) { set_last_error("train_state_set: slot 0 adam_m, 1 adam_v, 2 lookahead slow (weights go through set_param)"); return ISHARA_ERR_INVALID; }
  if (host_in == nullptr || numel != ts->n_total) { set_last_error("train_state_set: expected " + std::to_string(ts->n_total) + " elements"); return ISHARA_ERR_SHAPE; }
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  if (which == 2 && ts->slow == nullptr) ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ts->slow), static_cast<size_t>(ts->n_total) * sizeof(float)));
  ISHARA_CUDA_OK(cudaDeviceSynchronize());
  ISHARA_CUDA_OK(cudaMemcpy(state_slot(ts, which), host_in, static_cast<size_t>(numel) * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}
int train_state_set_counters(ishara_model* m, int64_t opt_steps, int64_t fb_steps) {
  if (m->train == nullptr) { set_last_error("train_state_set_counters: no training state"); return ISHARA_ERR_STATE; }
  if (opt_steps < 0 || fb_steps < 0) { set_last_error("train_state_set_counters: negative counter"); return ISHARA_ERR_INVALID; }
  m->train->step = static_cast<int>(opt_steps);
  m->train->fb_count = fb_steps;
  return 0;
}

// device masters -> host parameter table -> inference packs (so model(x) after training uses the new weights)
int train_sync(ishara_model* m) {
  if (m->train == nullptr || !m->host_params_stale) return 0;
  TrainState* ts = m->train;
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  ISHARA_CUDA_OK(cudaDeviceSynchronize());
  std::vector<float> host(static_cast<size_t>(ts->n_total));
  ISHARA_CUDA_OK(cudaMemcpy(host.data(), ts->theta, host.size() * sizeof(float), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < m->params.size(); ++i) {
    Param& p = m->params[i];
    std::memcpy(p.data.data(), host.data() + ts->off[i], p.data.size() * sizeof(float));
  }
  m->host_params_stale = false;
  return model_finalize(m);
}

int train_param_grad(ishara_model* m, const char* name, float* host_out, int64_t numel) {
  if (m->train == nullptr) { set_last_error("train_param_grad: no training state"); return ISHARA_ERR_STATE; }
  auto it = m->index.find(name ? name : "");
  if (it == m->index.end()) { set_last_error(std::string("unknown parameter: ") + (name ? name : "(null)")); return ISHARA_ERR_INVALID; }
  const Param& p = m->params[it->second];
  if (p.numel() != numel) { set_last_error("train_param_grad: size mismatch"); return ISHARA_ERR_SHAPE; }
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  ISHARA_CUDA_OK(cudaDeviceSynchronize());
  ISHARA_CUDA_OK(cudaMemcpy(host_out, m->train->grad + m->train->off[it->second], numel * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

// debug: fetch a named activation (want_grad = 0) or its gradient (want_grad = 1; needs train_configure(debug=1))
int train_fetch(ishara_model* m, const char* name, int want_grad, float* host_out, int64_t numel) {
  if (m->train == nullptr) { set_last_error("train_fetch: no training state"); return ISHARA_ERR_STATE; }
  auto& table = want_grad ? m->train->named_grad : m->train->named;
  auto it = table.find(name ? name : "");
  if (it == table.end()) { set_last_error(std::string("train_fetch: no tensor named ") + (name ? name : "(null)")); return ISHARA_ERR_INVALID; }
  const NamedBuf& nb = it->second;
  if (nb.rows * nb.cols != numel) { set_last_error("train_fetch: expected " + std::to_string(nb.rows * nb.cols) + " elements"); return ISHARA_ERR_SHAPE; }
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  ISHARA_CUDA_OK(cudaDeviceSynchronize());
  if (nb.dtype == 1) {
    ISHARA_CUDA_OK(cudaMemcpy(host_out, nb.ptr, numel * sizeof(float), cudaMemcpyDeviceToHost));
  } else {
    std::vector<uint16_t> tmp(static_cast<size_t>(numel));
    ISHARA_CUDA_OK(cudaMemcpy(tmp.data(), nb.ptr, numel * 2, cudaMemcpyDeviceToHost));
    for (int64_t j = 0; j < numel; ++j) {
      const uint32_t u = static_cast<uint32_t>(tmp[j]) << 16;
      std::memcpy(&host_out[j], &u, 4);
    }
  }
  return 0;
}

}  // namespace ishara
