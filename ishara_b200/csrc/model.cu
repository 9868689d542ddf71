// ishara_b200 — model runtime: parameter registry (Keras names/layouts), weight packing, workspace,
// and the launch program of get_model's forward (nb:conv-hybrid-model c7:12-65).
//
// HBM layout: activations are [B*T, C] bf16 row-major (channels-last, exactly the TF layout), the
// residual stream S lives in ONE buffer that the stream-producing GEMMs update in place (they read
// it only as the per-row residual), XN holds LayerNorm(S) for the next GEMM, H1/H2 hold the wide
// (2D / ef*D / 3D) intermediates. Weights are packed once: transposed to [N, K] (K contiguous, the
// UMMA "K-major" B operand), bf16, inference BatchNorm folded into the neighbouring linear op.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <unordered_map>
#include <vector>


#include "model_internal.h"

using namespace ishara;

namespace ishara {
// train.cu
void train_destroy(ishara_model* m);
int model_comm_destroy(ishara_model* m);
void train_invalidate(ishara_model* m);
int train_sync(ishara_model* m);

namespace {

// ------------------------------------------------------------------------------------------------
// parameter table
// ------------------------------------------------------------------------------------------------
void add_param(ishara_model* m, const std::string& name, std::vector<int64_t> shape) {
  Param p;
  p.name = name;
  p.shape = std::move(shape);
  m->index[name] = static_cast<int>(m->params.size());
  m->params.push_back(std::move(p));
}
void add_norm(ishara_model* m, const std::string& base, int64_t d, bool bn) {
  add_param(m, base + ".gamma", {d});
  add_param(m, base + ".beta", {d});
  if (bn) {
    add_param(m, base + ".moving_mean", {d});
    add_param(m, base + ".moving_variance", {d});
  }
}
void add_dense(ishara_model* m, const std::string& base, int64_t in, int64_t out, bool bias, bool conv1x1 = false) {
  if (conv1x1) add_param(m, base + ".kernel", {1, in, out});
  else add_param(m, base + ".kernel", {in, out});
  if (bias) add_param(m, base + ".bias", {out});
}

int conv_kernel_size(const ishara_config_t& c, int j) { return c.kernel_sizes[j % c.num_kernel_sizes]; }

void add_conv_blocks(ishara_model* m, const std::string& tag, int i) {
  const ishara_config_t& c = m->cfg;
  const int64_t D = c.dim;
  for (int j = 0; j < c.num_conv_per_block; ++j) {
    const std::string n = "conv" + tag + "_" + std::to_string(i) + "_" + std::to_string(j + 1);
    add_dense(m, n + "_expand_conv", D, 2 * D, true);
    add_param(m, n + "_dwconv.depthwise_kernel", {conv_kernel_size(c, j), 2 * D, 1});
    add_norm(m, n + "_bn", 2 * D, true);
    add_param(m, n + "_eca.kernel", {5, 1, 1});
    add_dense(m, n + "_project_conv", 2 * D, D, true);
  }
}
void add_ffn(ishara_model* m, const std::string& base, int64_t D, int64_t E) {
  add_dense(m, base + ".0", D, E, true);
  add_dense(m, base + ".2", E, D, true);
}

void build_param_table(ishara_model* m) {
  const ishara_config_t& c = m->cfg;
  const int64_t D = c.dim, E = static_cast<int64_t>(c.expansion_factor) * c.dim, tk = c.transformer_kernel_size;
  add_dense(m, "stem_conv", c.features, D, false);
  add_norm(m, "stem_bn", D, true);
  for (int i = 0; i < c.num_conv_squeeze_blocks; ++i) {
    add_conv_blocks(m, "squeeze", i);
    const std::string n = "squeezeformer_" + std::to_string(i);
    add_norm(m, n + ".norm1", D, false);
    add_ffn(m, n + ".ffn1", D, E);
    add_norm(m, n + ".norm2", D, false);
    add_dense(m, n + ".mha.qkv", D, 3 * D, false);
    add_dense(m, n + ".mha.proj", D, D, false);
    add_norm(m, n + ".conv.norm", D, false);
    add_dense(m, n + ".conv.conv1", D, E, true, true);
    add_param(m, n + ".conv.conv2.depthwise_kernel", {tk, E, 1});
    add_dense(m, n + ".conv.conv3", E, D, true, true);
    const int64_t R = std::max<int64_t>(1, D / 8);
    add_dense(m, n + ".conv.se.fc1", D, R, true);
    add_dense(m, n + ".conv.se.fc2", R, D, true);
    add_norm(m, n + ".norm3", D, false);
    add_ffn(m, n + ".ffn2", D, E);
  }
  for (int i = 0; i < c.num_conv_conform_blocks; ++i) {
    add_conv_blocks(m, "conform", i);
    const std::string n = "conformer_" + std::to_string(i);
    add_norm(m, n + ".layer_norm1", D, false);
    add_norm(m, n + ".layer_norm2", D, false);
    add_ffn(m, n + ".ffn1", D, E);
    add_dense(m, n + ".mha.qkv", D, 3 * D, false);
    add_dense(m, n + ".mha.proj", D, D, false);
    add_dense(m, n + ".conv.pointwise_conv1", D, 2 * D, true, true);
    add_param(m, n + ".conv.depthwise_conv.kernel", {tk, 1, D});
    add_param(m, n + ".conv.depthwise_conv.bias", {D});
    add_norm(m, n + ".conv.batch_norm", D, true);
    add_dense(m, n + ".conv.pointwise_conv2", D, D, true, true);
    add_norm(m, n + ".conv.layer_norm", D, false);
    add_ffn(m, n + ".ffn2", D, E);
  }
  add_dense(m, "top_conv", D, 2 * D, true);
  add_dense(m, "classifier", 2 * D, c.num_classes, true);
}

// ------------------------------------------------------------------------------------------------
// packing helpers
// ------------------------------------------------------------------------------------------------
const std::vector<float>& P(const ishara_model* m, const std::string& name) {
  return m->params[m->index.at(name)].data;
}

int upload(ishara_model* m, const std::string& key, const void* host, size_t bytes, void** out) {
  void* d = nullptr;
  ISHARA_CUDA_OK(cudaMalloc(&d, bytes < 256 ? 256 : bytes));
  ISHARA_CUDA_OK(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  m->wallocs.push_back(d);
  m->packed[key] = d;
  *out = d;
  return 0;
}

// kernel [K, N] fp32 (Keras Dense) -> bf16 [Npad, Kpad], row n = column perm[n] of the kernel scaled
// by col_scale[perm[n]]; zero padding.
int pack_linear(ishara_model* m, const std::string& key, const std::vector<float>& kernel, int K, int N, int Kpad,
                int Npad, const std::vector<float>* col_scale, const std::vector<int>* perm, bf16** out) {
  std::vector<uint16_t> h(static_cast<size_t>(Npad) * Kpad, 0);
  for (int n = 0; n < Npad; ++n) {
    const int src = perm ? (*perm)[n] : n;
    if (src < 0 || src >= N) continue;
    const float s = col_scale ? (*col_scale)[src] : 1.f;
    for (int k = 0; k < K; ++k) h[static_cast<size_t>(n) * Kpad + k] = f2bf(kernel[static_cast<size_t>(k) * N + src] * s);
  }
  void* d;
  int rc = upload(m, key, h.data(), h.size() * 2, &d);
  *out = static_cast<bf16*>(d);
  return rc;
}
int pack_f32(ishara_model* m, const std::string& key, const std::vector<float>& v, float** out) {
  void* d;
  int rc = upload(m, key, v.data(), v.size() * 4, &d);
  *out = static_cast<float*>(d);
  return rc;
}
// inference BatchNorm as y = x*s + o  (Keras default epsilon 1e-3; c5:73, c7:17, c5:281)
void bn_fold(const ishara_model* m, const std::string& base, std::vector<float>* s, std::vector<float>* o) {
  const auto &g = P(m, base + ".gamma"), &b = P(m, base + ".beta"), &mu = P(m, base + ".moving_mean"),
             &var = P(m, base + ".moving_variance");
  s->resize(g.size());
  o->resize(g.size());
  for (size_t i = 0; i < g.size(); ++i) {
    const double sc = static_cast<double>(g[i]) / std::sqrt(static_cast<double>(var[i]) + 1e-3);
    (*s)[i] = static_cast<float>(sc);
    (*o)[i] = static_cast<float>(static_cast<double>(b[i]) - static_cast<double>(mu[i]) * sc);
  }
}

int wide_block_n(int n) { return n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : 64); }

struct Packed {
  // lazily looked up packed device pointers by key
  ishara_model* m;
  template <typename T>
  T* get(const std::string& key) const {
    auto it = m->packed.find(key);
    return it == m->packed.end() ? nullptr : static_cast<T*>(it->second);
  }
};

int pack_ln(ishara_model* m, const std::string& base) {
  float* d;
  int rc;
  if ((rc = pack_f32(m, base + ".gamma", P(m, base + ".gamma"), &d))) return rc;
  return pack_f32(m, base + ".beta", P(m, base + ".beta"), &d);
}
int pack_dense(ishara_model* m, const std::string& base, int K, int N, bool bias) {
  bf16* w;
  int rc;
  if ((rc = pack_linear(m, base + ".w", P(m, base + ".kernel"), K, N, K, N, nullptr, nullptr, &w))) return rc;
  if (bias) {
    float* b;
    if ((rc = pack_f32(m, base + ".b", P(m, base + ".bias"), &b))) return rc;
  }
  return 0;
}

int pack_conv_blocks(ishara_model* m, const std::string& tag, int i) {
  const ishara_config_t& c = m->cfg;
  const int D = c.dim;
  int rc;
  for (int j = 0; j < c.num_conv_per_block; ++j) {
    const std::string n = "conv" + tag + "_" + std::to_string(i) + "_" + std::to_string(j + 1);
    if ((rc = pack_dense(m, n + "_expand_conv", D, 2 * D, true))) return rc;
    std::vector<float> s, o;
    bn_fold(m, n + "_bn", &s, &o);
    const int k = conv_kernel_size(c, j);
    std::vector<float> w = P(m, n + "_dwconv.depthwise_kernel");  // [k, 2D, 1]
    for (int t = 0; t < k; ++t)
      for (int ch = 0; ch < 2 * D; ++ch) w[static_cast<size_t>(t) * 2 * D + ch] *= s[ch];
    float* d;
    if ((rc = pack_f32(m, n + "_dw.w", w, &d))) return rc;
    if ((rc = pack_f32(m, n + "_dw.b", o, &d))) return rc;
    {
      std::vector<float> ws(2 * D, 0.f);
      for (int ch = 0; ch < 2 * D; ++ch) {
        double acc = 0.0;
        for (int t = 0; t < k; ++t) acc += w[static_cast<size_t>(t) * 2 * D + ch];
        ws[ch] = static_cast<float>(acc);
      }
      if ((rc = pack_f32(m, n + "_dw.wsum", ws, &d))) return rc;
    }
    if ((rc = pack_f32(m, n + "_eca.w", P(m, n + "_eca.kernel"), &d))) return rc;
    if ((rc = pack_dense(m, n + "_project_conv", 2 * D, D, true))) return rc;
  }
  return 0;
}

int pack_all(ishara_model* m) {
  for (void* p : m->wallocs) cudaFree(p);
  m->wallocs.clear();
  m->packed.clear();
  const ishara_config_t& c = m->cfg;
  const int D = c.dim, E = c.expansion_factor * c.dim, tk = c.transformer_kernel_size, T = c.frames;
  int rc;
  {
    // stem: BN(x@W + PE) = x@(W*s) + (PE*s + o)      (c7:14-17, positional_encoding c5:226-235)
    std::vector<float> s, o;
    bn_fold(m, "stem_bn", &s, &o);
    bf16* w;
    if ((rc = pack_linear(m, "stem.w", P(m, "stem_conv.kernel"), c.features, D, m->fpad(), D, &s, nullptr, &w))) return rc;
    std::vector<float> tab(static_cast<size_t>(T) * D);
    const int half = D / 2;
    const float depth = static_cast<float>(D) / 2.f;
    for (int t = 0; t < T; ++t) {
      for (int i = 0; i < half; ++i) {
        const float rate = 1.f / powf(10000.f, static_cast<float>(i) / depth);
        const float ang = static_cast<float>(t) * rate;
        tab[static_cast<size_t>(t) * D + i] = sinf(ang) * s[i] + o[i];
        tab[static_cast<size_t>(t) * D + half + i] = cosf(ang) * s[half + i] + o[half + i];
      }
    }
    float* d;
    if ((rc = pack_f32(m, "stem.tab", tab, &d))) return rc;
  }
  auto pack_ffn = [&](const std::string& base) -> int {
    int r;
    if ((r = pack_dense(m, base + ".0", D, E, true))) return r;
    return pack_dense(m, base + ".2", E, D, true);
  };
  for (int i = 0; i < c.num_conv_squeeze_blocks; ++i) {
    if ((rc = pack_conv_blocks(m, "squeeze", i))) return rc;
    const std::string n = "squeezeformer_" + std::to_string(i);
    for (const char* ln : {".norm1", ".norm2", ".norm3", ".conv.norm"})
      if ((rc = pack_ln(m, n + ln))) return rc;
    if ((rc = pack_ffn(n + ".ffn1"))) return rc;
    if ((rc = pack_ffn(n + ".ffn2"))) return rc;
    if ((rc = pack_dense(m, n + ".mha.qkv", D, 3 * D, false))) return rc;
    if ((rc = pack_dense(m, n + ".mha.proj", D, D, false))) return rc;
    if ((rc = pack_dense(m, n + ".conv.conv1", D, E, true))) return rc;
    float* d;
    if ((rc = pack_f32(m, n + ".conv.dw.w", P(m, n + ".conv.conv2.depthwise_kernel"), &d))) return rc;
    if ((rc = pack_dense(m, n + ".conv.conv3", E, D, true))) return rc;
    {  // the same kernel in its native [E, D] orientation for the SqueezeExcite gate (coalesced over output channels)
      const auto& k3 = P(m, n + ".conv.conv3.kernel");
      std::vector<uint16_t> h(k3.size());
      for (size_t i = 0; i < k3.size(); ++i) h[i] = f2bf(k3[i]);
      void* dkn;
      if ((rc = upload(m, n + ".conv.conv3.wkn", h.data(), h.size() * 2, &dkn))) return rc;
    }
    if ((rc = pack_f32(m, n + ".se.fc1.w", P(m, n + ".conv.se.fc1.kernel"), &d))) return rc;
    if ((rc = pack_f32(m, n + ".se.fc1.b", P(m, n + ".conv.se.fc1.bias"), &d))) return rc;
    if ((rc = pack_f32(m, n + ".se.fc2.w", P(m, n + ".conv.se.fc2.kernel"), &d))) return rc;
    if ((rc = pack_f32(m, n + ".se.fc2.b", P(m, n + ".conv.se.fc2.bias"), &d))) return rc;
  }
  for (int i = 0; i < c.num_conv_conform_blocks; ++i) {
    if ((rc = pack_conv_blocks(m, "conform", i))) return rc;
    const std::string n = "conformer_" + std::to_string(i);
    for (const char* ln : {".layer_norm1", ".layer_norm2", ".conv.layer_norm"})
      if ((rc = pack_ln(m, n + ln))) return rc;
    if ((rc = pack_ffn(n + ".ffn1"))) return rc;
    if ((rc = pack_ffn(n + ".ffn2"))) return rc;
    if ((rc = pack_dense(m, n + ".mha.qkv", D, 3 * D, false))) return rc;
    if ((rc = pack_dense(m, n + ".mha.proj", D, D, false))) return rc;
    {
      // pointwise_conv1 + GLU (c5:293-295): out = a * sigmoid(b), a = first D channels, b = last D.
      // Packed so that every BN-wide N tile holds [a-slice | matching b-slice].
      const int bn = wide_block_n(2 * D);
      std::vector<int> perm(2 * D);
      const int halfw = bn / 2;
      for (int tile = 0; tile < 2 * D / bn; ++tile)
        for (int j = 0; j < halfw; ++j) {
          perm[tile * bn + j] = tile * halfw + j;
          perm[tile * bn + halfw + j] = D + tile * halfw + j;
        }
      bf16* w;
      if ((rc = pack_linear(m, n + ".conv.pw1.w", P(m, n + ".conv.pointwise_conv1.kernel"), D, 2 * D, D, 2 * D, nullptr,
                            &perm, &w)))
        return rc;
      const auto& b = P(m, n + ".conv.pointwise_conv1.bias");
      std::vector<float> bp(2 * D);
      for (int j = 0; j < 2 * D; ++j) bp[j] = b[perm[j]];
      float* d;
      if ((rc = pack_f32(m, n + ".conv.pw1.b", bp, &d))) return rc;
    }
    {
      // grouped Conv1D 'same' + bias, then BN (momentum .99, eps 1e-3): y = (conv + b)*s + o
      std::vector<float> s, o;
      bn_fold(m, n + ".conv.batch_norm", &s, &o);
      std::vector<float> w = P(m, n + ".conv.depthwise_conv.kernel");  // [tk, 1, D]
      const auto& b = P(m, n + ".conv.depthwise_conv.bias");
      std::vector<float> bb(D);
      for (int t = 0; t < tk; ++t)
        for (int ch = 0; ch < D; ++ch) w[static_cast<size_t>(t) * D + ch] *= s[ch];
      for (int ch = 0; ch < D; ++ch) bb[ch] = b[ch] * s[ch] + o[ch];
      float* d;
      if ((rc = pack_f32(m, n + ".conv.dw.w", w, &d))) return rc;
      if ((rc = pack_f32(m, n + ".conv.dw.b", bb, &d))) return rc;
    }
    if ((rc = pack_dense(m, n + ".conv.pointwise_conv2", D, D, true))) return rc;
  }
  if ((rc = pack_dense(m, "top_conv", D, 2 * D, true))) return rc;
  {
    bf16* w;
    if ((rc = pack_linear(m, "classifier.w", P(m, "classifier.kernel"), 2 * D, c.num_classes, 2 * D, m->vpad(), nullptr,
                          nullptr, &w)))
      return rc;
    std::vector<float> b(m->vpad(), 0.f);
    const auto& src = P(m, "classifier.bias");
    for (int i = 0; i < c.num_classes; ++i) b[i] = src[i];
    float* d;
    if ((rc = pack_f32(m, "classifier.b", b, &d))) return rc;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// workspace + program
// ------------------------------------------------------------------------------------------------
template <typename T>
int ws_alloc(ishara_model* m, T** p, size_t count) {
  void* d = nullptr;
  ISHARA_CUDA_OK(cudaMalloc(&d, count * sizeof(T) + 256));
  m->wsallocs.push_back(d);
  *p = static_cast<T*>(d);
  return 0;
}

int ensure_workspace(ishara_model* m, int batch) {
  if (batch <= m->cap_batch) return 0;
  for (void* p : m->wsallocs) cudaFree(p);
  m->wsallocs.clear();
  m->taps.clear();
  m->program.clear();
  m->program_cache.clear();
  for (auto& kv : m->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  m->graphs.clear();
  m->program_batch = 0;
  const ishara_config_t& c = m->cfg;
  const size_t M = static_cast<size_t>(batch) * c.frames;
  const size_t D = c.dim, E = static_cast<size_t>(c.expansion_factor) * c.dim;
  const size_t W1 = std::max<size_t>(std::max<size_t>(2 * D, E), 3 * D), W2 = std::max<size_t>(2 * D, E);
  int rc;
  if ((rc = ws_alloc(m, &m->x_dev, M * c.features))) return rc;
  if ((rc = ws_alloc(m, &m->XIN, M * m->fpad()))) return rc;
  if ((rc = ws_alloc(m, &m->S, M * D))) return rc;
  if ((rc = ws_alloc(m, &m->XN, M * D))) return rc;
  if ((rc = ws_alloc(m, &m->H1, M * W1))) return rc;
  if ((rc = ws_alloc(m, &m->H2, M * W2))) return rc;
  if ((rc = ws_alloc(m, &m->O, M * D))) return rc;
  if ((rc = ws_alloc(m, &m->HEAD, M * 2 * D))) return rc;
  if ((rc = ws_alloc(m, &m->colsum, static_cast<size_t>(batch) * E))) return rc;
  if ((rc = ws_alloc(m, &m->gate, static_cast<size_t>(batch) * D))) return rc;
  if ((rc = ws_alloc(m, &m->logits_own, M * c.num_classes))) return rc;
  if ((rc = ws_alloc(m, &m->ids_dev, M))) return rc;
  if ((rc = ws_alloc(m, &m->lens_dev, static_cast<size_t>(batch)))) return rc;
  if ((rc = ws_alloc(m, &m->nll_dev, static_cast<size_t>(batch)))) return rc;
  if ((rc = ws_alloc(m, &m->mask_dev, M))) return rc;
  if ((rc = ws_alloc(m, &m->wbits_dev, M))) return rc;
  if ((rc = ws_alloc(m, &m->valid_dev, static_cast<size_t>(batch)))) return rc;
  if ((rc = ws_alloc(m, &m->user_mask_dev, M))) return rc;
  if ((rc = ws_alloc(m, &m->use_user_mask_dev, 4))) return rc;
  ISHARA_CUDA_OK(cudaMemset(m->use_user_mask_dev, 0, 4 * sizeof(int)));
  m->labels_dev = nullptr;
  m->labels_cap = 0;
  m->cap_batch = batch;
  return 0;
}

struct Builder {
  ishara_model* m;
  int B, M, D, E, T;
  std::vector<Op>& ops;
  Packed pk;
  int rc = 0;
  // mask_mode="propagated": true while the Keras mask is alive, i.e. up to and including the Conv1DBlocks in front of
  // the first ConformerBlock (ConformerBlock neither consumes nor forwards a mask: no supports_masking, c5:311-343)
  bool masked = false;

  void wide_gemm(const char* label, const bf16* A, int K, const std::string& wkey, const std::string& bkey, int N,
                 int act, bf16* out) {
    if (rc) return;
    Op op;
    op.kind = OP_GEMM;
    op.label = label;
    GemmPlan& p = op.gemm;
    p.M = M; p.N = N; p.K = K;
    p.block_n = wide_block_n(N);
    p.row_mode = false;
    p.out_f32 = false;
    p.epi.bias = bkey.empty() ? nullptr : pk.get<float>(bkey);
    p.epi.act = act;
    p.epi.rows_per_seq = T;
    const int nout = act == ACT_GLU ? N / 2 : N;
    rc = gemm_plan_init(&p, A, K, pk.get<bf16>(wkey), out, nout, nout, nullptr, 0);
    op.flops = 2.0 * M * N * K;
    op.bytes = 2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K + static_cast<double>(M) * nout);
    ops.push_back(op);
  }

  // S = [LN0]( A@W + bias [*gate] [+rowtab] [+S] ) ; XN = LN1(S)
  void stream_gemm(const char* label, const bf16* A, int K, const std::string& wkey, const std::string& bkey,
                   const float* gate, const float* rowtab, bool resid, LnRef ln0, LnRef ln1, int klogical = 0) {
    if (rc) return;
    if (klogical == 0) klogical = K;
    Op op;
    op.kind = OP_GEMM;
    op.label = label;
    GemmPlan& p = op.gemm;
    p.M = M; p.N = D; p.K = K;
    p.out_f32 = false;
    p.epi.bias = bkey.empty() ? nullptr : pk.get<float>(bkey);
    p.epi.gate = gate;
    p.epi.rowtab = rowtab;
    p.epi.resid = resid ? m->S : nullptr;
    p.epi.ld_resid = D;
    p.epi.rows_per_seq = T;
    const bool fused = (D == 256 || D == 128);
    static const int plain16 = getenv("ISHARA_STREAM_PLAIN16") ? atoi(getenv("ISHARA_STREAM_PLAIN16")) : 0;  // measured: no gain (60 vs 59 us), opt-in
    if (fused && plain16 && D == 256 && !ln0.g && !ln1.g) {
      // no LayerNorm to fuse: the plain (non-row) kernel with 16 epilogue warps handles bias / gate / rowtab / residual
      p.block_n = 256;
      p.row_mode = false;
      rc = gemm_plan_init(&p, A, K, pk.get<bf16>(wkey), m->S, D, D, nullptr, 0);
      op.flops = 2.0 * M * D * klogical;
      op.bytes = 2.0 * (static_cast<double>(M) * K + static_cast<double>(D) * K + static_cast<double>(M) * D * (1 + (resid ? 1 : 0))) +
                 (rowtab ? 4.0 * T * D : 0.0) + (gate ? 4.0 * B * D : 0.0);
      ops.push_back(op);
      return;
    }
    if (fused) {
      p.block_n = D;
      p.row_mode = true;
      p.epi.ln0_g = ln0.g; p.epi.ln0_b = ln0.b; p.epi.ln0_eps = ln0.eps;
      p.epi.ln1_g = ln1.g; p.epi.ln1_b = ln1.b; p.epi.ln1_eps = ln1.eps;
      rc = gemm_plan_init(&p, A, K, pk.get<bf16>(wkey), m->S, D, D, ln1.g ? m->XN : nullptr, D);
      op.flops = 2.0 * M * D * klogical;
      op.bytes = 2.0 * (static_cast<double>(M) * K + static_cast<double>(D) * K +
                        static_cast<double>(M) * D * (1 + (resid ? 1 : 0) + (ln1.g ? 1 : 0))) +
                 (rowtab ? 4.0 * T * D : 0.0) + (gate ? 4.0 * B * D : 0.0);
      ops.push_back(op);
    } else {
      p.block_n = wide_block_n(D);
      p.row_mode = false;
      rc = gemm_plan_init(&p, A, K, pk.get<bf16>(wkey), m->S, D, D, nullptr, 0);
      op.flops = 2.0 * M * D * klogical;
      op.bytes = 2.0 * (static_cast<double>(M) * K + static_cast<double>(D) * K + static_cast<double>(M) * D * (1 + (resid ? 1 : 0))) +
                 (rowtab ? 4.0 * T * D : 0.0) + (gate ? 4.0 * B * D : 0.0);
      ops.push_back(op);
      if (ln0.g) layernorm(m->S, m->S, ln0);
      if (ln1.g) layernorm(m->S, m->XN, ln1);
    }
  }
  void layernorm(const bf16* in, bf16* out, LnRef ln) {
    Op op;
    op.kind = OP_LN;
    op.label = "layernorm";
    op.ln_in = in; op.ln_out = out; op.ln = ln;
    op.ln_rows = M;
    op.flops = 8.0 * M * D;
    op.bytes = 4.0 * M * D;
    ops.push_back(op);
  }
  void dwconv(const char* label, const bf16* in, bf16* out, int C, int k, int pad_left, const std::string& wkey,
              const std::string& bkey, const std::string& ecakey, int post, float* colsum) {
    if (rc) return;
    Op op;
    op.kind = OP_DW;
    op.label = label;
    op.dw.in = in; op.dw.out = out;
    op.dw.w = pk.get<float>(wkey);
    op.dw.bias = bkey.empty() ? nullptr : pk.get<float>(bkey);
    op.dw.eca_w = ecakey.empty() ? nullptr : pk.get<float>(ecakey);
    op.dw.colsum = colsum;
    if (masked && (post == 2 || colsum != nullptr)) {  // ECA mean / SqueezeExcite pooling over the valid frames only
      op.dw.key_mask = m->mask_dev;
      op.dw.valid_cnt = m->valid_dev;
    }
    op.dw.B = B; op.dw.T = T; op.dw.C = C; op.dw.k = k; op.dw.pad_left = pad_left; op.dw.post = post;
    op.flops = 2.0 * M * C * k;
    op.bytes = 4.0 * M * C;  // bf16 in + bf16 out
    ops.push_back(op);
  }
  void attention(const bf16* qkv, bf16* out) {
    if (rc) return;
    Op op;
    op.kind = OP_ATTN;
    op.label = "attention";
    op.at.qkv = qkv; op.at.out = out;
    op.at.key_mask = masked ? m->mask_dev : nullptr;  // Softmax(mask): scores of padded keys += -1e9 (c5:109-112)
    op.at.B = B; op.at.T = T; op.at.H = m->cfg.num_heads; op.at.dh = D / m->cfg.num_heads;
    op.at.scale = 1.f / std::sqrt(static_cast<float>(D));  // self.scale = dim ** -0.5 (c5:95), NOT dh ** -0.5
    op.flops = 4.0 * M * T * D;             // QK^T and PV: 2 * (2 * T * T * D) per sequence
    op.bytes = 2.0 * M * (3.0 * D + D);     // qkv read + out write
    ops.push_back(op);
  }
  void tap(const std::string& name) {
    if (!m->debug_taps) return;
    Op op;
    op.kind = OP_TAP;
    op.label = "tap";
    op.tap = static_cast<int>(m->taps.size());
    bf16* buf = nullptr;
    if (ws_alloc(m, &buf, static_cast<size_t>(M) * D)) { rc = 3; return; }
    m->taps.push_back(buf);
    m->tap_names.push_back(name);
    ops.push_back(op);
  }
  LnRef ln(const std::string& base, float eps) {
    LnRef r;
    r.g = pk.get<float>(base + ".gamma");
    r.b = pk.get<float>(base + ".beta");
    r.eps = eps;
    return r;
  }

  void conv_blocks(const std::string& tag, int i, LnRef next_ln) {
    const ishara_config_t& c = m->cfg;
    for (int j = 0; j < c.num_conv_per_block; ++j) {
      const std::string n = "conv" + tag + "_" + std::to_string(i) + "_" + std::to_string(j + 1);
      const int k = conv_kernel_size(c, j);
      static const int fused_front = getenv("ISHARA_CONV1D_FUSED") ? atoi(getenv("ISHARA_CONV1D_FUSED")) : 0;  // measured: 118 us vs 55 + 66 us unfused => no gain yet, opt-in
      static const int fused_block = getenv("ISHARA_CONV1D_BLOCK") ? atoi(getenv("ISHARA_CONV1D_BLOCK")) : 1;
      const bool last_blk = j == c.num_conv_per_block - 1;
      if (fused_block && !fused_front && !rc && conv1d_block_applicable(D, T, k)) {
        // the whole block in one launch: the 2D-wide intermediate never leaves the SM (conv1d_block.cu)
        Op op;
        op.kind = OP_C1B;
        op.label = "conv1d.block_fused";
        Conv1dBlockPlan& p = op.c1b;
        p.B = B; p.T = T; p.k = k;
        p.bias_e = pk.get<float>(n + "_expand_conv.b");
        p.dw_w = pk.get<float>(n + "_dw.w");
        p.dw_b = pk.get<float>(n + "_dw.b");
        p.dw_wsum = pk.get<float>(n + "_dw.wsum");
        p.eca_w = pk.get<float>(n + "_eca.w");
        p.bias_p = pk.get<float>(n + "_project_conv.b");
        const LnRef nl = last_blk ? next_ln : LnRef();
        p.ln_g = nl.g; p.ln_b = nl.b; p.ln_eps = nl.eps;
        p.wbits = masked ? m->wbits_dev : nullptr;
        p.valid_cnt = masked ? m->valid_dev : nullptr;
        rc = conv1d_block_plan_init(&p, m->S, pk.get<bf16>(n + "_expand_conv.w"), pk.get<bf16>(n + "_project_conv.w"),
                                    nl.g ? m->XN : nullptr);
        op.flops = 2.0 * M * D * 2 * D * 2 + 2.0 * M * 2 * D * k;
        op.bytes = 2.0 * (static_cast<double>(M) * D * (2 + (nl.g ? 1 : 0)) + 4.0 * D * D);
        ops.push_back(op);
        tap(n);
        continue;
      }
      if (fused_front && !rc && conv1d_front_applicable(D, T, k)) {
        // expand GEMM + swish + causal depthwise + BatchNorm + ECA in one launch (conv1d_front.cu)
        Op op;
        op.kind = OP_C1F;
        op.label = "conv1d.front_fused";
        Conv1dFrontPlan& p = op.c1f;
        p.B = B; p.T = T; p.k = k;
        p.bias_e = pk.get<float>(n + "_expand_conv.b");
        p.dw_w = pk.get<float>(n + "_dw.w");
        p.dw_b = pk.get<float>(n + "_dw.b");
        p.eca_w = pk.get<float>(n + "_eca.w");
        p.out = m->H2;
        rc = conv1d_front_plan_init(&p, m->S, pk.get<bf16>(n + "_expand_conv.w"));
        op.flops = 2.0 * M * D * 2 * D + 2.0 * M * 2 * D * k;
        op.bytes = 2.0 * (static_cast<double>(M) * D + static_cast<double>(M) * 2 * D + 2.0 * D * D);
        ops.push_back(op);
      } else {
        wide_gemm("conv1d.expand", m->S, D, n + "_expand_conv.w", n + "_expand_conv.b", 2 * D, ACT_SWISH, m->H1);
        dwconv("conv1d.dw_bn_eca", m->H1, m->H2, 2 * D, k, k - 1, n + "_dw.w", n + "_dw.b", n + "_eca.w", 2, nullptr);
      }
      const bool last = j == c.num_conv_per_block - 1;
      stream_gemm("conv1d.project", m->H2, 2 * D, n + "_project_conv.w", n + "_project_conv.b", nullptr, nullptr, true,
                  LnRef(), last ? next_ln : LnRef());
      tap(n);
    }
  }
  void ffn(const std::string& base, LnRef next_ln) {
    static const int fused = getenv("ISHARA_FFN_FUSED") ? atoi(getenv("ISHARA_FFN_FUSED")) : 1;
    if (fused && !rc && ffn_applicable(D, E, M, m->num_sms)) {
      // one launch: the E-wide intermediate stays in TMEM / shared memory (ffn_tc.cu)
      Op op;
      op.kind = OP_FFN;
      op.label = "ffn.fused";
      FfnPlan& p = op.ffn;
      p.M = M; p.E = E;
      p.bias1 = pk.get<float>(base + ".0.b");
      p.epi.bias = pk.get<float>(base + ".2.b");
      p.epi.resid = m->S;
      p.epi.ld_resid = D;
      p.epi.rows_per_seq = T;
      p.epi.ln1_g = next_ln.g; p.epi.ln1_b = next_ln.b; p.epi.ln1_eps = next_ln.eps;
      rc = ffn_plan_init(&p, m->XN, pk.get<bf16>(base + ".0.w"), pk.get<bf16>(base + ".2.w"), m->S, next_ln.g ? m->XN : nullptr);
      op.flops = 4.0 * M * D * E;
      op.bytes = 2.0 * (static_cast<double>(M) * D * (3 + (next_ln.g ? 1 : 0)) + 2.0 * D * E);
      ops.push_back(op);
      return;
    }
    wide_gemm("ffn.up", m->XN, D, base + ".0.w", base + ".0.b", E, ACT_SWISH, m->H1);
    stream_gemm("ffn.down", m->H1, E, base + ".2.w", base + ".2.b", nullptr, nullptr, true, LnRef(), next_ln);
  }
  void mhsa(const std::string& base, LnRef next_ln) {
    wide_gemm("mhsa.qkv", m->XN, D, base + ".qkv.w", "", 3 * D, ACT_NONE, m->H1);
    attention(m->H1, m->O);
    stream_gemm("mhsa.proj", m->O, D, base + ".proj.w", "", nullptr, nullptr, true, LnRef(), next_ln);
  }
};

// Lanes: a large batch is cut into `lanes` independent sub-batches, each with its own slice of every workspace tensor and
// its own op list, launched on parallel streams (one forked CUDA graph). Every kernel of the path is per-sequence, so
// the results do not change; what changes is that the partial last wave of one lane's kernel (768 row tiles over 148 SMs
// = 5.2 waves, 256 clusters over ~47 cluster slots = 5.4 waves) and its launch ramp are filled by the other lane's work.
int lane_count(const ishara_model* m, int batch) {
  static const int want = getenv("ISHARA_LANES") ? std::max(1, std::min(kMaxLanes, atoi(getenv("ISHARA_LANES")))) : 2;
  static const int min_batch = getenv("ISHARA_LANES_MIN_BATCH") ? atoi(getenv("ISHARA_LANES_MIN_BATCH")) : 96;
  if (m->profile || m->debug_taps || batch < min_batch) return 1;  // per-op events and taps need one serial stream
  return std::min(want, batch);
}
void lane_range(int batch, int lanes, int l, int* b0, int* bl) {
  static const int frac0 = getenv("ISHARA_LANE0_PCT") ? atoi(getenv("ISHARA_LANE0_PCT")) : 50;  // two lanes: share of lane 0 (tuning)
  if (lanes == 2 && frac0 > 0 && frac0 < 100) {
    const int cut = std::max(1, std::min(batch - 1, static_cast<int>(static_cast<int64_t>(batch) * frac0 / 100)));
    *b0 = l == 0 ? 0 : cut;
    *bl = l == 0 ? cut : batch - cut;
    return;
  }
  const int lo = static_cast<int>(static_cast<int64_t>(batch) * l / lanes), hi = static_cast<int>(static_cast<int64_t>(batch) * (l + 1) / lanes);
  *b0 = lo;
  *bl = hi - lo;
}

int build_lane(ishara_model* m, int batch, float* logits, int lane);

int build_program(ishara_model* m, int batch, float* logits) {
  const ishara_config_t& c = m->cfg;
  m->program.clear();
  m->tap_names.clear();
  m->taps.clear();
  const int lanes = lane_count(m, batch);
  if (lanes == 1) return build_lane(m, batch, logits, 0);
  // point the handle's workspace pointers at this lane's slice while its ops are built, then restore them
  const size_t T = c.frames, D = c.dim, E = static_cast<size_t>(c.expansion_factor) * c.dim;
  const size_t W1 = std::max<size_t>(std::max<size_t>(2 * D, E), 3 * D), W2 = std::max<size_t>(2 * D, E);
  bf16 *const XIN = m->XIN, *const S = m->S, *const XN = m->XN, *const H1 = m->H1, *const H2 = m->H2, *const O = m->O, *const HEAD = m->HEAD;
  float *const colsum = m->colsum, *const gate = m->gate;
  uint8_t* const mask = m->mask_dev;
  uint16_t* const wbits = m->wbits_dev;
  int32_t* const valid = m->valid_dev;
  int rc = 0;
  for (int l = 0; l < lanes && rc == 0; ++l) {
    int b0, bl;
    lane_range(batch, lanes, l, &b0, &bl);
    const size_t r0 = static_cast<size_t>(b0) * T;
    m->XIN = XIN + r0 * m->fpad(); m->S = S + r0 * D; m->XN = XN + r0 * D; m->H1 = H1 + r0 * W1; m->H2 = H2 + r0 * W2;
    m->O = O + r0 * D; m->HEAD = HEAD + r0 * 2 * D; m->colsum = colsum + static_cast<size_t>(b0) * E; m->gate = gate + static_cast<size_t>(b0) * D;
    m->mask_dev = mask + r0; m->wbits_dev = wbits + r0; m->valid_dev = valid + b0;
    rc = build_lane(m, bl, logits + r0 * c.num_classes, l);
  }
  m->XIN = XIN; m->S = S; m->XN = XN; m->H1 = H1; m->H2 = H2; m->O = O; m->HEAD = HEAD; m->colsum = colsum; m->gate = gate;
  m->mask_dev = mask; m->wbits_dev = wbits; m->valid_dev = valid;
  if (rc) { m->program.clear(); return rc; }
  m->program_batch = batch;
  m->program_logits = logits;
  return 0;
}

int build_lane(ishara_model* m, int batch, float* logits, int lane) {
  const ishara_config_t& c = m->cfg;
  const size_t first_op = m->program.size();
  Builder b{m, batch, batch * c.frames, c.dim, c.expansion_factor * c.dim, c.frames, m->program, Packed{m}};
  const int D = c.dim, E = b.E, tk = c.transformer_kernel_size;

  // what the first module after the stem wants as its normalised input
  auto first_ln_of = [&](int sq_i, int cf_i) -> LnRef {
    if (sq_i < c.num_conv_squeeze_blocks) return b.ln("squeezeformer_" + std::to_string(sq_i) + ".norm1", 1e-6f);
    if (cf_i < c.num_conv_conform_blocks) return b.ln("conformer_" + std::to_string(cf_i) + ".layer_norm1", 1e-6f);
    return LnRef();
  };
  const bool has_conv = c.num_conv_per_block > 0;
  b.masked = m->mask_mode == 1;

  {
    // stem
    LnRef nl = has_conv ? LnRef() : first_ln_of(0, 0);
    b.stream_gemm("stem", m->XIN, m->fpad(), "stem.w", "", nullptr, b.pk.get<float>("stem.tab"), false, LnRef(), nl,
                  c.features);
    b.tap("stem");
  }
  for (int i = 0; i < c.num_conv_squeeze_blocks; ++i) {
    const std::string n = "squeezeformer_" + std::to_string(i);
    b.conv_blocks("squeeze", i, b.ln(n + ".norm1", 1e-6f));
    // next module's LN (only when there are no conv blocks in between)
    LnRef after = has_conv ? LnRef() : first_ln_of(i + 1, 0);
    b.ffn(n + ".ffn1", b.ln(n + ".norm2", 1e-6f));
    b.mhsa(n + ".mha", b.ln(n + ".conv.norm", 1e-6f));
    // ConvModule (c5:145-153): LN -> 1x1 -> swish -> causal DW -> swish -> 1x1 -> SE -> + x
    b.wide_gemm("sqz.conv1", m->XN, D, n + ".conv.conv1.w", n + ".conv.conv1.b", E, ACT_SWISH, m->H1);
    b.dwconv("sqz.dw_swish", m->H1, m->H2, E, tk, tk - 1, n + ".conv.dw.w", "", "", 1, m->colsum);
    {
      Op op;
      op.kind = OP_SEGATE;
      op.label = "sqz.se_gate";
      op.se.colsum = m->colsum;
      op.se.w3kn = b.pk.get<bf16>(n + ".conv.conv3.wkn");
      op.se.b3 = b.pk.get<float>(n + ".conv.conv3.b");
      op.se.fc1_w = b.pk.get<float>(n + ".se.fc1.w");
      op.se.fc1_b = b.pk.get<float>(n + ".se.fc1.b");
      op.se.fc2_w = b.pk.get<float>(n + ".se.fc2.w");
      op.se.fc2_b = b.pk.get<float>(n + ".se.fc2.b");
      op.se.gate = m->gate;
      op.se.B = batch; op.se.C = E; op.se.D = D; op.se.R = std::max(1, D / 8);
      op.se.inv_T = 1.f / static_cast<float>(c.frames);
      op.se.valid_cnt = b.masked ? m->valid_dev : nullptr;
      op.flops = 2.0 * batch * (static_cast<double>(E) * D + 2.0 * D * op.se.R);
      op.bytes = 4.0 * batch * (E + D) + 2.0 * E * D;
      m->program.push_back(op);
    }
    b.stream_gemm("sqz.conv3_se", m->H2, E, n + ".conv.conv3.w", n + ".conv.conv3.b", m->gate, nullptr, true, LnRef(),
                  b.ln(n + ".norm3", 1e-6f));
    b.ffn(n + ".ffn2", after);
    b.tap(n);
  }
  for (int i = 0; i < c.num_conv_conform_blocks; ++i) {
    const std::string n = "conformer_" + std::to_string(i);
    const LnRef ln1 = b.ln(n + ".layer_norm1", 1e-6f);
    b.conv_blocks("conform", i, ln1);
    b.masked = false;  // the mask dies at the first ConformerBlock
    LnRef after = has_conv ? LnRef() : first_ln_of(c.num_conv_squeeze_blocks, i + 1);
    b.ffn(n + ".ffn1", ln1);            // layer_norm1 is reused for the MHSA input (c5:324,330)
    b.mhsa(n + ".mha", LnRef());        // conv module consumes the raw stream
    // ConvolutionModule (c5:288-309)
    b.wide_gemm("cf.pw1_glu", m->S, D, n + ".conv.pw1.w", n + ".conv.pw1.b", 2 * D, ACT_GLU, m->H2);
    b.dwconv("cf.dw_bn", m->H2, m->O, D, tk, (tk - 1) / 2, n + ".conv.dw.w", n + ".conv.dw.b", "", 0, nullptr);
    b.stream_gemm("cf.pw2_ln", m->O, D, n + ".conv.pointwise_conv2.w", n + ".conv.pointwise_conv2.b", nullptr, nullptr,
                  true, b.ln(n + ".conv.layer_norm", 1e-3f), b.ln(n + ".layer_norm2", 1e-6f));
    b.ffn(n + ".ffn2", after);
    b.tap(n);
  }
  // head (c7:61-63)
  b.wide_gemm("head.top_conv", m->S, D, "top_conv.w", "top_conv.b", 2 * D, ACT_RELU, m->HEAD);
  if (!b.rc) {
    Op op;
    op.kind = OP_GEMM;
    op.label = "head.classifier";
    GemmPlan& p = op.gemm;
    p.M = b.M; p.N = m->vpad(); p.K = 2 * D;
    p.block_n = 64;
    p.out_f32 = true;
    p.epi.bias = b.pk.get<float>("classifier.b");
    p.epi.rows_per_seq = c.frames;
    b.rc = gemm_plan_init(&p, m->HEAD, 2 * D, b.pk.get<bf16>("classifier.w"), logits, c.num_classes, c.num_classes,
                          nullptr, 0);
    op.flops = 2.0 * b.M * c.num_classes * 2.0 * D;
    op.bytes = 2.0 * b.M * 2.0 * D + 2.0 * m->vpad() * 2.0 * D + 4.0 * b.M * c.num_classes;
    m->program.push_back(op);
  }
  if (b.rc) {
    m->program.clear();
    return b.rc;
  }
  for (size_t i = first_op; i < m->program.size(); ++i) m->program[i].lane = lane;
  m->program_batch = batch;
  m->program_logits = logits;
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// entry points used by capi.cu
// ------------------------------------------------------------------------------------------------
int model_create(const ishara_config_t* cfg, int device, ishara_model** out) {
  if (cfg == nullptr || out == nullptr) { set_last_error("null argument"); return ISHARA_ERR_INVALID; }
  const ishara_config_t& c = *cfg;
  if (c.dim <= 0 || c.dim % 64 != 0) { set_last_error("dim must be a positive multiple of 64"); return ISHARA_ERR_SHAPE; }
  if (c.num_heads <= 0 || c.dim % c.num_heads != 0) { set_last_error("dim must be divisible by num_heads"); return ISHARA_ERR_SHAPE; }
  const int dh = c.dim / c.num_heads;
  if (dh != 16 && dh != 32 && dh != 48 && dh != 64) { set_last_error("head dim must be 16/32/48/64"); return ISHARA_ERR_SHAPE; }
  if (c.num_kernel_sizes < 0 || c.num_kernel_sizes > 8 || (c.num_conv_per_block > 0 && c.num_kernel_sizes == 0)) {
    set_last_error("kernel_sizes: need 1..8 entries"); return ISHARA_ERR_SHAPE;
  }
  if (c.features % 4 != 0 || c.num_classes % 4 != 0 || c.frames <= 0 || c.expansion_factor <= 0) {
    set_last_error("features and num_classes must be multiples of 4; frames, expansion_factor > 0"); return ISHARA_ERR_SHAPE;
  }
  auto m = std::make_unique<ishara_model>();
  m->cfg = c;
  m->device = device;
  build_param_table(m.get());
  *out = m.release();
  return ISHARA_OK;
}

int model_destroy(ishara_model* m) {
  if (m == nullptr) return ISHARA_OK;
  if (m->finalized || !m->wsallocs.empty()) cudaSetDevice(m->device);
  train_destroy(m);
  model_comm_destroy(m);
  if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
  if (m->comm_ready) { cudaEventDestroy(m->comm_ready); cudaEventDestroy(m->comm_done); }
  for (void* p : m->wallocs) cudaFree(p);
  for (void* p : m->wsallocs) cudaFree(p);
  for (cudaEvent_t e : m->events) cudaEventDestroy(e);
  for (auto& kv : m->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (m->labels_dev) cudaFree(m->labels_dev);
  if (m->stream) cudaStreamDestroy(m->stream);
  if (m->lane_fork) {
    cudaEventDestroy(m->lane_fork);
    for (int l = 0; l < kMaxLanes - 1; ++l) { cudaStreamDestroy(m->lane_stream[l]); cudaEventDestroy(m->lane_join[l]); }
  }
  for (auto& sl : m->pipe) {
    if (sl.x) cudaFree(sl.x);
    if (sl.labels) cudaFree(sl.labels);
    if (sl.h2d_done) { cudaEventDestroy(sl.h2d_done); cudaEventDestroy(sl.x_free); cudaEventDestroy(sl.done); }
  }
  if (m->copy_stream) {
    cudaStreamDestroy(m->copy_stream);
    for (auto& e : m->copy_done) if (e) cudaEventDestroy(e);
  }
  delete m;
  return ISHARA_OK;
}

int model_finalize(ishara_model* m) {
  for (const Param& p : m->params)
    if (!p.set) { set_last_error("parameter not set: " + p.name); return ISHARA_ERR_STATE; }
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  int sms = 0;
  ISHARA_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device));
  int major = 0;
  ISHARA_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, m->device));
  if (major != 10) { set_last_error("ishara_b200 needs a Blackwell (sm_100a) device; no fallback exists"); return ISHARA_ERR_CUDA; }
  m->num_sms = sms;
  if (m->stream == nullptr) ISHARA_CUDA_OK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  if (m->lane_fork == nullptr) {
    ISHARA_CUDA_OK(cudaEventCreateWithFlags(&m->lane_fork, cudaEventDisableTiming));
    for (int l = 0; l < kMaxLanes - 1; ++l) {
      ISHARA_CUDA_OK(cudaStreamCreateWithFlags(&m->lane_stream[l], cudaStreamNonBlocking));
      ISHARA_CUDA_OK(cudaEventCreateWithFlags(&m->lane_join[l], cudaEventDisableTiming));
    }
  }
  if (m->copy_stream == nullptr) {
    ISHARA_CUDA_OK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    for (auto& e : m->copy_done) ISHARA_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  int rc = pack_all(m);
  if (rc) return rc;
  m->finalized = true;
  m->program.clear();
  m->program_cache.clear();
  for (auto& kv : m->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  m->graphs.clear();
  m->program_batch = 0;
  return ISHARA_OK;
}

int launch_program(ishara_model* m, const float* x_dev, int batch, cudaStream_t stream, bool prof);
static void drop_graphs(ishara_model* m) {
  for (auto& kv : m->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  m->graphs.clear();
}

int model_forward_impl(ishara_model* m, const float* x_dev, int batch, float* logits_dev, cudaStream_t stream);
int model_forward(ishara_model* m, const float* x_dev, int batch, float* logits_dev, cudaStream_t stream) {
  // propagated masks come from x itself (Masking(0.0)) unless forward_masked supplied one for this call
  if (m->mask_mode == 1 && m->use_user_mask_dev != nullptr) ISHARA_CUDA_OK(cudaMemsetAsync(m->use_user_mask_dev, 0, sizeof(int), stream));
  return model_forward_impl(m, x_dev, batch, logits_dev, stream);
}
int model_forward_impl(ishara_model* m, const float* x_dev, int batch, float* logits_dev, cudaStream_t stream) {
  if (!m->finalized) { set_last_error("forward before finalize"); return ISHARA_ERR_STATE; }
  if (m->host_params_stale) { int rcs = train_sync(m); if (rcs) return rcs; }  // weights moved by a training step
  if (batch <= 0 || x_dev == nullptr || logits_dev == nullptr) { set_last_error("forward: bad arguments"); return ISHARA_ERR_INVALID; }
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  int rc;
  if ((rc = ensure_workspace(m, batch))) return rc;
  if (m->program_batch != batch || m->program_logits != logits_dev || m->program.empty()) {
    if (m->debug_taps) {
      m->program_cache.clear();
      if ((rc = build_program(m, batch, logits_dev))) return rc;
    } else {
      if (!m->program.empty()) {
        if (m->program_cache.size() >= 16) { m->program_cache.clear(); drop_graphs(m); }
        m->program_cache[{m->program_batch, m->program_logits}] = std::move(m->program);
        m->program.clear();
      }
      auto it = m->program_cache.find({batch, logits_dev});
      if (it != m->program_cache.end()) {
        m->program = std::move(it->second);
        m->program_cache.erase(it);
        m->program_batch = batch;
        m->program_logits = logits_dev;
      } else if ((rc = build_program(m, batch, logits_dev))) {
        return rc;
      }
    }
  }
  const bool prof = m->profile;
  // CUDA graph path: the 70-odd launches of one forward are captured once per (program, input pointer) on the handle's
  // own stream and replayed on the caller's stream, which removes the CPU launch cost and most inter-kernel gaps.
  static const int use_graphs = getenv("ISHARA_GRAPH") ? atoi(getenv("ISHARA_GRAPH")) : 1;
  if (use_graphs && !prof && !m->debug_taps && !m->graphs_broken) {
    GraphKey key{batch, logits_dev, x_dev};
    auto it = m->graphs.find(key);
    if (it != m->graphs.end()) {
      if (it->second.exec != nullptr) {
        ISHARA_CUDA_OK(cudaGraphLaunch(it->second.exec, stream));
        note_launches(it->second.launches);
        return ISHARA_OK;
      }
      // second call with this key: every kernel has run once (attributes set, lazy state built) -> capture now
      cudaGraph_t graph = nullptr;
      const uint64_t before = launch_count();
      if (cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        rc = launch_program(m, x_dev, batch, m->stream, false);
        const cudaError_t ce = cudaStreamEndCapture(m->stream, &graph);
        cudaGraphExec_t exec = nullptr;
        if (rc == 0 && ce == cudaSuccess && graph != nullptr && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
          it->second.exec = exec;
          it->second.launches = static_cast<int>(launch_count() - before);
          cudaGraphDestroy(graph);
          if (m->graphs.size() > 32) {  // bound the cache (rotating caller buffers)
            for (auto& kv : m->graphs) if (kv.second.exec && kv.first != key) cudaGraphExecDestroy(kv.second.exec);
            GraphEntry keep = it->second;
            m->graphs.clear();
            m->graphs[key] = keep;
          }
          ISHARA_CUDA_OK(cudaGraphLaunch(exec, stream));
          return ISHARA_OK;
        }
        if (graph != nullptr) cudaGraphDestroy(graph);
      }
      cudaGetLastError();  // capture is an optimisation: fall back to direct launches for this handle
      m->graphs_broken = true;
    } else {
      m->graphs[key] = GraphEntry{};
    }
  }
  return launch_program(m, x_dev, batch, stream, prof);
}

int model_set_mask_mode(ishara_model* m, int mode) {
  if (mode != 0 && mode != 1) { set_last_error("mask_mode: 0 (dropped) or 1 (propagated)"); return ISHARA_ERR_INVALID; }
  if (m->mask_mode != mode) {
    m->mask_mode = mode;
    m->program.clear();
    m->program_cache.clear();
    drop_graphs(m);
    m->program_batch = 0;
  }
  return 0;
}

// forward with an explicit frame mask (uint8 [B, T], 1 = frame carries data) instead of Masking(0.0)'s any(x != 0); only
// meaningful with mask_mode = propagated. mask_dev == null: derive it from x.
int model_forward_masked(ishara_model* m, const float* x_dev, const uint8_t* mask_dev, int batch, float* logits_dev, cudaStream_t stream) {
  if (mask_dev != nullptr && m->mask_mode != 1) { set_last_error("forward_masked: a mask needs mask_mode = propagated (ishara_model_set_mask_mode)"); return ISHARA_ERR_STATE; }
  if (m->mask_mode == 1) {
    ISHARA_CUDA_OK(cudaSetDevice(m->device));
    int rc = ensure_workspace(m, batch);
    if (rc) return rc;
    const int flag = mask_dev != nullptr ? 1 : 0;
    // the flag and the mask copy are stream-ordered in front of the (possibly graph-replayed) program that reads them
    ISHARA_CUDA_OK(cudaMemsetAsync(m->use_user_mask_dev, 0, sizeof(int), stream));
    if (flag) {
      ISHARA_CUDA_OK(cudaMemcpyAsync(m->user_mask_dev, mask_dev, static_cast<size_t>(batch) * m->cfg.frames, cudaMemcpyDeviceToDevice, stream));
      ISHARA_CUDA_OK(cudaMemsetAsync(m->use_user_mask_dev, 1, 1, stream));  // little-endian int 1
    }
  }
  return model_forward_impl(m, x_dev, batch, logits_dev, stream);
}

int launch_program(ishara_model* m, const float* x_dev, int batch, cudaStream_t stream, bool prof) {
  const ishara_config_t& c = m->cfg;
  int rc;
  int lanes = 1;
  for (const Op& op : m->program) lanes = std::max(lanes, op.lane + 1);
  if (prof) {
    while (m->events.size() < m->program.size() + 2) {
      cudaEvent_t e;
      ISHARA_CUDA_OK(cudaEventCreate(&e));
      m->events.push_back(e);
    }
    ISHARA_CUDA_OK(cudaEventRecord(m->events[0], stream));
  }
  // lane 0 runs on the caller's stream; the others fork from it here and join it at the end (inside a capture this
  // becomes one graph with parallel branches)
  cudaStream_t ls[kMaxLanes] = {stream, nullptr, nullptr, nullptr};
  if (lanes > 1) {
    ISHARA_CUDA_OK(cudaEventRecord(m->lane_fork, stream));
    for (int l = 1; l < lanes; ++l) {
      ls[l] = m->lane_stream[l - 1];
      ISHARA_CUDA_OK(cudaStreamWaitEvent(ls[l], m->lane_fork, 0));
    }
  }
  for (int l = 0; l < lanes; ++l) {
    int b0, bl;
    lane_range(batch, lanes, l, &b0, &bl);
    const size_t r0 = static_cast<size_t>(b0) * c.frames;
    const float* xl = x_dev + r0 * c.features;
    if (m->mask_mode == 1 &&
        (rc = mask_prep_launch(xl, m->user_mask_dev + r0, m->use_user_mask_dev, bl, c.frames, c.features, m->mask_dev + r0, m->wbits_dev + r0,
                               m->valid_dev + b0, ls[l])))
      return rc;
    if ((rc = cast_pad_launch(xl, m->XIN + r0 * m->fpad(), static_cast<int64_t>(bl) * c.frames, c.features, m->fpad(), ls[l]))) return rc;
  }
  if (prof) ISHARA_CUDA_OK(cudaEventRecord(m->events[1], stream));
  const int64_t M = static_cast<int64_t>(batch) * c.frames;  // OP_TAP only exists in one-lane (debug) programs
  size_t op_index = 0;
  // interleave the lanes' launches so that neither stream runs far ahead of the other on the host side
  std::vector<size_t> next(lanes, 0);
  std::vector<std::vector<size_t>> by_lane(lanes);
  for (size_t i = 0; i < m->program.size(); ++i) by_lane[m->program[i].lane].push_back(i);
  for (size_t round = 0, left = m->program.size(); left > 0; ++round) {
    for (int l = 0; l < lanes; ++l) {
      if (round >= by_lane[l].size()) continue;
      const Op& op = m->program[by_lane[l][round]];
      cudaStream_t st = ls[l];
      --left;
      switch (op.kind) {
        case OP_GEMM: rc = gemm_launch(op.gemm, m->num_sms, st); break;
        case OP_DW: rc = dwconv_launch(op.dw, st); break;
        case OP_ATTN: rc = attention_launch(op.at, st); break;
        case OP_SEGATE: rc = se_gate_launch(op.se, st); break;
        case OP_FFN: rc = ffn_launch(op.ffn, m->num_sms, st); break;
        case OP_C1F: rc = conv1d_front_launch(op.c1f, st); break;
        case OP_C1B: rc = conv1d_block_launch(op.c1b, st); break;
        case OP_LN: rc = layernorm_launch(op.ln_in, op.ln_out, op.ln.g, op.ln.b, op.ln.eps, op.ln_rows, c.dim, st); break;
        case OP_TAP:
          rc = cudaMemcpyAsync(m->taps[op.tap], m->S, M * c.dim * sizeof(bf16), cudaMemcpyDeviceToDevice, st) == cudaSuccess ? 0 : 3;
          break;
      }
      if (rc) {
        set_last_error(std::string("forward: op '") + op.label + "' failed: " + get_last_error());
        return rc;
      }
      if (prof) ISHARA_CUDA_OK(cudaEventRecord(m->events[2 + op_index], st));
      ++op_index;
    }
  }
  for (int l = 1; l < lanes; ++l) {
    ISHARA_CUDA_OK(cudaEventRecord(m->lane_join[l - 1], ls[l]));
    ISHARA_CUDA_OK(cudaStreamWaitEvent(stream, m->lane_join[l - 1], 0));
  }
  return ISHARA_OK;
}

int model_set_profile(ishara_model* m, int on) {
  if (m->profile != (on != 0)) {  // profiled programs are built with one lane (per-op events need a serial stream)
    m->profile = on != 0;
    m->program.clear();
    m->program_cache.clear();
    drop_graphs(m);
    m->program_batch = 0;
  }
  return 0;
}
int model_profile_count(const ishara_model* m) { return m->program.empty() ? 0 : static_cast<int>(m->program.size()) + 1; }
// entry 0 = the fp32->bf16 input cast; entry 1+i = program[i]. ms = device time of that launch in the LAST
// profiled forward (event to event on the launch stream, so it includes the launch gap before the kernel).
int model_profile_entry(ishara_model* m, int i, const char** label, const char** kind, float* ms, double* flops,
                        double* bytes) {
  const int n = model_profile_count(m);
  if (i < 0 || i >= n) { set_last_error("profile_entry: index out of range"); return ISHARA_ERR_INVALID; }
  if (m->events.size() < static_cast<size_t>(n) + 1) { set_last_error("profile_entry: no profiled forward yet"); return ISHARA_ERR_STATE; }
  ISHARA_CUDA_OK(cudaEventSynchronize(m->events[n]));
  float t = 0.f;
  ISHARA_CUDA_OK(cudaEventElapsedTime(&t, m->events[i], m->events[i + 1]));
  const ishara_config_t& c = m->cfg;
  const double M = static_cast<double>(m->program_batch) * c.frames;
  static const char* kinds[] = {"gemm", "dwconv", "attention", "se_gate", "layernorm", "tap", "gemm", "gemm", "conv1d_block"};
  if (i == 0) {
    if (label) *label = "input.cast_pad";
    if (kind) *kind = "cast";
    if (flops) *flops = 0.0;
    if (bytes) *bytes = M * (4.0 * c.features + 2.0 * m->fpad());
  } else {
    const Op& op = m->program[i - 1];
    if (label) *label = op.label;
    if (kind) *kind = kinds[op.kind];
    if (flops) *flops = op.flops;
    if (bytes) *bytes = op.bytes;
  }
  if (ms) *ms = t;
  return 0;
}


// ------------------------------------------------------------------------------------------------
// accessors used by capi.cu (the struct stays private to this file)
// ------------------------------------------------------------------------------------------------
struct ModelView {
  const ishara_config_t* cfg;
  int device;
  cudaStream_t stream;
  cudaStream_t copy_stream;
  cudaEvent_t* copy_done;  // [8]
  float* x_dev;
  float* logits_own;
  int32_t* ids_dev;
  int32_t* lens_dev;
  float* nll_dev;
};

int model_view(ishara_model* m, int batch, ModelView* v) {
  if (!m->finalized) { set_last_error("model not finalized"); return ISHARA_ERR_STATE; }
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  int rc = ensure_workspace(m, batch);
  if (rc) return rc;
  v->cfg = &m->cfg;
  v->device = m->device;
  v->stream = m->stream;
  v->copy_stream = m->copy_stream;
  v->copy_done = m->copy_done;
  v->x_dev = m->x_dev;
  v->logits_own = m->logits_own;
  v->ids_dev = m->ids_dev;
  v->lens_dev = m->lens_dev;
  v->nll_dev = m->nll_dev;
  return 0;
}

// ---- pipelined host inference --------------------------------------------------------------------------------------
// submit: enqueue H2D (copy stream) + forward + decode [+ CTC] + D2H (main stream) for one batch and return at once;
// collect: wait for the oldest submitted batch. Up to two batches in flight, so batch i+1 uploads while batch i computes.
int model_infer_submit(ishara_model* m, const float* x_host, int batch, const int32_t* labels_host, int max_label_len, float* logits_host,
                       int32_t* ids_host, int32_t* lens_host, float* nll_host) {
  if (!m->finalized) { set_last_error("model not finalized"); return ISHARA_ERR_STATE; }
  if (m->pipe_submitted - m->pipe_collected >= 2) { set_last_error("infer_submit: two batches already in flight, collect one first"); return ISHARA_ERR_STATE; }
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  if (m->host_params_stale) { int rcs = train_sync(m); if (rcs) return rcs; }
  if (batch > m->cap_batch && m->pipe_submitted != m->pipe_collected) {
    set_last_error("infer_submit: a larger batch needs a bigger workspace; collect the batch in flight first");
    return ISHARA_ERR_STATE;
  }
  int rc;
  if ((rc = ensure_workspace(m, batch))) return rc;
  const ishara_config_t& c = m->cfg;
  const size_t M = static_cast<size_t>(batch) * c.frames;
  const int slot_i = static_cast<int>(m->pipe_submitted & 1);
  ishara_model::PipeSlot& sl = m->pipe[slot_i];
  if (sl.h2d_done == nullptr) {
    ISHARA_CUDA_OK(cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
    ISHARA_CUDA_OK(cudaEventCreateWithFlags(&sl.x_free, cudaEventDisableTiming));
    ISHARA_CUDA_OK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  }
  if (sl.x_cap < M * c.features) {
    if (m->pipe_submitted != m->pipe_collected) ISHARA_CUDA_OK(cudaDeviceSynchronize());
    if (sl.x) ISHARA_CUDA_OK(cudaFree(sl.x));
    sl.x = nullptr; sl.x_cap = 0;
    for (auto& kv : m->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);  // graphs hold the old input pointer
    m->graphs.clear();
    ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&sl.x), M * c.features * sizeof(float)));
    sl.x_cap = M * c.features;
  }
  const size_t nlab = labels_host != nullptr ? static_cast<size_t>(batch) * max_label_len : 0;
  if (sl.labels_cap < nlab) {
    if (sl.labels) ISHARA_CUDA_OK(cudaFree(sl.labels));
    sl.labels = nullptr; sl.labels_cap = 0;
    ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&sl.labels), nlab * sizeof(int32_t)));
    sl.labels_cap = nlab;
  }
  // upload on the copy stream, once the forward that last read this slot is done
  ISHARA_CUDA_OK(cudaStreamWaitEvent(m->copy_stream, sl.x_free, 0));
  ISHARA_CUDA_OK(cudaMemcpyAsync(sl.x, x_host, M * c.features * sizeof(float), cudaMemcpyHostToDevice, m->copy_stream));
  if (nlab) ISHARA_CUDA_OK(cudaMemcpyAsync(sl.labels, labels_host, nlab * sizeof(int32_t), cudaMemcpyHostToDevice, m->copy_stream));
  ISHARA_CUDA_OK(cudaEventRecord(sl.h2d_done, m->copy_stream));
  // compute + read-back on the main stream
  cudaStream_t s = m->stream;
  ISHARA_CUDA_OK(cudaStreamWaitEvent(s, sl.h2d_done, 0));
  if ((rc = model_forward(m, sl.x, batch, m->logits_own, s))) return rc;
  const int blank = c.num_classes - 1;
  if ((rc = greedy_decode_launch(m->logits_own, batch, c.frames, c.num_classes, blank, m->ids_dev, m->lens_dev, s))) return rc;
  if (nlab) {
    if ((rc = ctc_loss_launch(m->logits_own, sl.labels, batch, c.frames, c.num_classes, max_label_len, blank, m->nll_dev, nullptr, nullptr, 0, s))) return rc;
    ISHARA_CUDA_OK(cudaMemcpyAsync(nll_host, m->nll_dev, batch * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  ISHARA_CUDA_OK(cudaEventRecord(sl.x_free, s));  // every reader of this slot's x AND labels has been enqueued before this point
  ISHARA_CUDA_OK(cudaMemcpyAsync(ids_host, m->ids_dev, M * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  ISHARA_CUDA_OK(cudaMemcpyAsync(lens_host, m->lens_dev, batch * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (logits_host != nullptr)
    ISHARA_CUDA_OK(cudaMemcpyAsync(logits_host, m->logits_own, M * c.num_classes * sizeof(float), cudaMemcpyDeviceToHost, s));
  ISHARA_CUDA_OK(cudaEventRecord(sl.done, s));
  ++m->pipe_submitted;
  return 0;
}

int model_infer_collect(ishara_model* m) {
  if (m->pipe_submitted == m->pipe_collected) { set_last_error("infer_collect: nothing in flight"); return ISHARA_ERR_STATE; }
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  ISHARA_CUDA_OK(cudaEventSynchronize(m->pipe[m->pipe_collected & 1].done));
  ++m->pipe_collected;
  return 0;
}

int model_labels_buffer(ishara_model* m, size_t count, int32_t** out) {
  if (static_cast<size_t>(m->labels_cap) < count) {
    if (m->labels_dev) ISHARA_CUDA_OK(cudaFree(m->labels_dev));
    m->labels_dev = nullptr;
    m->labels_cap = 0;
    ISHARA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&m->labels_dev), count * sizeof(int32_t)));
    m->labels_cap = static_cast<int>(count);
  }
  *out = m->labels_dev;
  return 0;
}

int model_num_params(const ishara_model* m) { return static_cast<int>(m->params.size()); }

int model_param_info(const ishara_model* m, int idx, const char** name, int64_t* numel, int32_t* ndim, int64_t shape[4]) {
  if (idx < 0 || idx >= static_cast<int>(m->params.size())) { set_last_error("param index out of range"); return ISHARA_ERR_INVALID; }
  const Param& p = m->params[idx];
  if (name) *name = p.name.c_str();
  if (numel) *numel = p.numel();
  if (ndim) *ndim = static_cast<int32_t>(p.shape.size());
  if (shape) for (size_t i = 0; i < 4; ++i) shape[i] = i < p.shape.size() ? p.shape[i] : 1;
  return 0;
}

int model_set_param(ishara_model* m, const char* name, const float* data, int64_t numel) {
  if (name == nullptr || data == nullptr) { set_last_error("set_param: null argument"); return ISHARA_ERR_INVALID; }
  auto it = m->index.find(name);
  if (it == m->index.end()) { set_last_error(std::string("unknown parameter: ") + name); return ISHARA_ERR_INVALID; }
  Param& p = m->params[it->second];
  if (p.numel() != numel) {
    set_last_error(std::string("set_param ") + name + ": expected " + std::to_string(p.numel()) + " elements, got " + std::to_string(numel));
    return ISHARA_ERR_SHAPE;
  }
  if (m->host_params_stale) { int rcs = train_sync(m); if (rcs) return rcs; }
  train_invalidate(m);  // the next training call re-uploads the table
  p.data.assign(data, data + numel);
  p.set = true;
  m->finalized = false;
  return 0;
}

int model_get_param(const ishara_model* mc, const char* name, float* out, int64_t numel) {
  if (name == nullptr || out == nullptr) { set_last_error("get_param: null argument"); return ISHARA_ERR_INVALID; }
  ishara_model* m = const_cast<ishara_model*>(mc);
  if (m->host_params_stale) { int rcs = train_sync(m); if (rcs) return rcs; }
  auto it = m->index.find(name);
  if (it == m->index.end()) { set_last_error(std::string("unknown parameter: ") + name); return ISHARA_ERR_INVALID; }
  const Param& p = m->params[it->second];
  if (!p.set) { set_last_error(std::string("parameter not set: ") + name); return ISHARA_ERR_STATE; }
  if (p.numel() != numel) { set_last_error("get_param: size mismatch"); return ISHARA_ERR_SHAPE; }
  std::memcpy(out, p.data.data(), numel * sizeof(float));
  return 0;
}

int model_set_debug(ishara_model* m, int on) {
  m->debug_taps = on != 0;
  m->program.clear();
  m->program_cache.clear();
  for (auto& kv : m->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  m->graphs.clear();
  m->program_batch = 0;
  return 0;
}

int model_debug_fetch(ishara_model* m, const char* name, float* host_out, int64_t numel) {
  if (name == nullptr || host_out == nullptr) { set_last_error("debug_fetch: null argument"); return ISHARA_ERR_INVALID; }
  for (size_t i = 0; i < m->tap_names.size(); ++i) {
    if (m->tap_names[i] != name) continue;
    const int64_t n = static_cast<int64_t>(m->program_batch) * m->cfg.frames * m->cfg.dim;
    if (numel != n) { set_last_error("debug_fetch: expected " + std::to_string(n) + " elements"); return ISHARA_ERR_SHAPE; }
    std::vector<uint16_t> tmp(n);
    ISHARA_CUDA_OK(cudaStreamSynchronize(m->stream));
    ISHARA_CUDA_OK(cudaDeviceSynchronize());
    ISHARA_CUDA_OK(cudaMemcpy(tmp.data(), m->taps[i], n * 2, cudaMemcpyDeviceToHost));
    for (int64_t j = 0; j < n; ++j) {
      const uint32_t u = static_cast<uint32_t>(tmp[j]) << 16;
      std::memcpy(&host_out[j], &u, 4);
    }
    return 0;
  }
  set_last_error(std::string("debug_fetch: no tap named ") + name + " (enable ishara_model_set_debug before forward)");
  return ISHARA_ERR_INVALID;
}

}  // namespace ishara
