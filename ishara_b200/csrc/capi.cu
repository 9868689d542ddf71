// ishara_b200 — extern "C" surface (include/ishara_b200.h). Thin: argument checks + calls into model.cu
// and the kernel launchers. No torch, no Python types.
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../../include/ishara_b200.h"
#include "kernels.h"

// model.cu
struct ishara_model;
namespace ishara {
int model_create(const ishara_config_t* cfg, int device, ishara_model** out);
int model_destroy(ishara_model* m);
int model_finalize(ishara_model* m);
int model_forward(ishara_model* m, const float* x_dev, int batch, float* logits_dev, cudaStream_t stream);
}  // namespace ishara

// model.cu keeps the struct definition private; the accessors below are implemented there too.
namespace ishara {
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
int model_view(ishara_model* m, int batch, ModelView* v);                       // ensures workspace
int model_labels_buffer(ishara_model* m, size_t count, int32_t** out);
int model_num_params(const ishara_model* m);
int model_param_info(const ishara_model* m, int idx, const char** name, int64_t* numel, int32_t* ndim, int64_t shape[4]);
int model_set_param(ishara_model* m, const char* name, const float* data, int64_t numel);
int model_get_param(const ishara_model* m, const char* name, float* out, int64_t numel);
int model_set_debug(ishara_model* m, int on);
int model_set_profile(ishara_model* m, int on);
int model_profile_count(const ishara_model* m);
int model_profile_entry(ishara_model* m, int i, const char** label, const char** kind, float* ms, double* flops, double* bytes);
int model_debug_fetch(ishara_model* m, const char* name, float* host_out, int64_t numel);
int model_infer_submit(ishara_model* m, const float* x_host, int batch, const int32_t* labels_host, int max_label_len, float* logits_host,
                       int32_t* ids_host, int32_t* lens_host, float* nll_host);
int model_infer_collect(ishara_model* m);
// train.cu
int train_configure(ishara_model* m, float dropout, uint64_t seed, int debug);
int train_forward_backward(ishara_model* m, const float* x_dev, const int32_t* labels_dev, int batch, int labels_len, float* loss_host,
                           cudaStream_t stream);
int train_grad_buffer(ishara_model* m, float** grad_dev, int64_t* numel);
int train_apply(ishara_model* m, const ishara_adamw_t* opt, float grad_scale, cudaStream_t stream);
int train_sync(ishara_model* m);
int train_param_grad(ishara_model* m, const char* name, float* host_out, int64_t numel);
int train_fetch(ishara_model* m, const char* name, int want_grad, float* host_out, int64_t numel);
int train_forward_backward_loss(ishara_model* m, float* loss_host, cudaStream_t stream);
int train_apply_radam(ishara_model* m, const ishara_radam_lookahead_t* opt, float grad_scale, cudaStream_t stream);
int train_state_info(ishara_model* m, int64_t* numel, int64_t* opt_steps, int64_t* fb_steps, int32_t* has_slow);
int train_state_get(ishara_model* m, int which, float* host_out, int64_t numel);
int train_state_set(ishara_model* m, int which, const float* host_in, int64_t numel);
int train_state_set_counters(ishara_model* m, int64_t opt_steps, int64_t fb_steps);
int model_set_mask_mode(ishara_model* m, int mode);
int model_forward_masked(ishara_model* m, const float* x_dev, const uint8_t* mask_dev, int batch, float* logits_dev, cudaStream_t stream);
int train_counters(ishara_model* m, int64_t* fb_steps, int64_t* opt_steps, int64_t* skipped_steps);
int comm_unique_id(void* out128);
int model_comm_init(ishara_model* m, const void* id128, int rank, int world);
int model_comm_destroy(ishara_model* m);
int nccl_version();
void comm_bucket_plan(const int64_t* hi, int n, int64_t n_train, int64_t min_elems, int64_t* lo_out, int64_t* up_out);
}  // namespace ishara

using namespace ishara;

#define CAPI_CUDA_OK(expr)                                                                          \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                           \
      return ISHARA_ERR_CUDA;                                                                       \
    }                                                                                               \
  } while (0)

#define CHECK_HANDLE(m)                      \
  if ((m) == nullptr) {                      \
    set_last_error("null model handle");     \
    return ISHARA_ERR_INVALID;               \
  }

// H2D of x in chunks on the copy stream, forward of chunk c as soon as its copy has landed: the PCIe transfer of the
// 1.1 KB/frame fp32 input (108 MB at batch 256) overlaps the compute of the previous chunk instead of preceding it.
static int upload_and_forward(ishara_model* m, const ModelView& v, const float* x_host, int batch) {
  const ishara_config_t& c = *v.cfg;
  const size_t per_seq = static_cast<size_t>(c.frames) * c.features;
  // Chunk boundaries sit on whole waves of 128-row tiles (one wave = num_sms tiles), growing 1 : 2 : rest, so the chunked
  // forward runs the same number of tile rounds as the unchunked one while the first chunk's copy is short.
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, v.device);
  const double wave = static_cast<double>(sms) * 128.0 / c.frames;  // sequences per tile wave
  int bounds[9];
  int nchunk = 0;
  bounds[0] = 0;
  if (batch >= static_cast<int>(3 * wave)) {
    bounds[1] = static_cast<int>(wave);
    bounds[2] = static_cast<int>(3 * wave);
    bounds[3] = batch;
    nchunk = 3;
    if (batch - bounds[2] > static_cast<int>(3 * wave)) {  // very large batches: keep the tail chunks at ~3 waves
      nchunk = 2;
      while (batch - bounds[nchunk] > static_cast<int>(3 * wave) && nchunk < 7) {
        bounds[nchunk + 1] = bounds[nchunk] + static_cast<int>(3 * wave);
        ++nchunk;
      }
      bounds[nchunk + 1] = batch;
      ++nchunk;
    }
  } else {
    bounds[1] = batch;
    nchunk = 1;
  }
  int rc;
  for (int i = 0; i < nchunk; ++i) {
    const int b0 = bounds[i], nb = bounds[i + 1] - bounds[i];
    if (cudaMemcpyAsync(v.x_dev + b0 * per_seq, x_host + b0 * per_seq, nb * per_seq * sizeof(float), cudaMemcpyHostToDevice,
                        v.copy_stream) != cudaSuccess ||
        cudaEventRecord(v.copy_done[i], v.copy_stream) != cudaSuccess) {
      set_last_error(std::string("upload: ") + cudaGetErrorString(cudaGetLastError()));
      return ISHARA_ERR_CUDA;
    }
  }
  for (int i = 0; i < nchunk; ++i) {
    const int b0 = bounds[i], nb = bounds[i + 1] - bounds[i];
    if (cudaStreamWaitEvent(v.stream, v.copy_done[i], 0) != cudaSuccess) {
      set_last_error(std::string("upload: ") + cudaGetErrorString(cudaGetLastError()));
      return ISHARA_ERR_CUDA;
    }
    if ((rc = model_forward(m, v.x_dev + b0 * per_seq, nb,
                            v.logits_own + static_cast<size_t>(b0) * c.frames * c.num_classes, v.stream)))
      return rc;
  }
  return 0;
}

extern "C" {

const char* ishara_version(void) { return "ishara_b200 0.1.0 (sm_100a)"; }
const char* ishara_last_error(void) { return get_last_error(); }
int32_t ishara_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

ishara_status_t ishara_model_create(const ishara_config_t* cfg, int32_t device, ishara_model_t** out) {
  return static_cast<ishara_status_t>(model_create(cfg, device, reinterpret_cast<ishara_model**>(out)));
}
ishara_status_t ishara_model_destroy(ishara_model_t* m) {
  return static_cast<ishara_status_t>(model_destroy(reinterpret_cast<ishara_model*>(m)));
}
int32_t ishara_model_num_params(const ishara_model_t* m) {
  if (m == nullptr) return 0;
  return model_num_params(reinterpret_cast<const ishara_model*>(m));
}
ishara_status_t ishara_model_param_info(const ishara_model_t* m, int32_t index, const char** name, int64_t* numel,
                                        int32_t* ndim, int64_t shape[4]) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(
      model_param_info(reinterpret_cast<const ishara_model*>(m), index, name, numel, ndim, shape));
}
ishara_status_t ishara_model_set_param(ishara_model_t* m, const char* name, const float* host_data, int64_t numel) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(model_set_param(reinterpret_cast<ishara_model*>(m), name, host_data, numel));
}
ishara_status_t ishara_model_get_param(const ishara_model_t* m, const char* name, float* host_out, int64_t numel) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(model_get_param(reinterpret_cast<const ishara_model*>(m), name, host_out, numel));
}
ishara_status_t ishara_model_finalize(ishara_model_t* m) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(model_finalize(reinterpret_cast<ishara_model*>(m)));
}
ishara_status_t ishara_model_set_debug(ishara_model_t* m, int32_t on) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(model_set_debug(reinterpret_cast<ishara_model*>(m), on));
}
ishara_status_t ishara_model_debug_fetch(ishara_model_t* m, const char* name, float* host_out, int64_t numel) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(model_debug_fetch(reinterpret_cast<ishara_model*>(m), name, host_out, numel));
}

ishara_status_t ishara_model_forward(ishara_model_t* m, const float* x_dev, int32_t batch, float* logits_dev,
                                     void* stream) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(
      model_forward(reinterpret_cast<ishara_model*>(m), x_dev, batch, logits_dev, static_cast<cudaStream_t>(stream)));
}

ishara_status_t ishara_model_forward_host(ishara_model_t* mh, const float* x_host, int32_t batch, float* logits_host) {
  CHECK_HANDLE(mh);
  if (x_host == nullptr || logits_host == nullptr || batch <= 0) {
    set_last_error("forward_host: bad arguments");
    return ISHARA_ERR_INVALID;
  }
  ishara_model* m = reinterpret_cast<ishara_model*>(mh);
  ModelView v;
  int rc = model_view(m, batch, &v);
  if (rc) return static_cast<ishara_status_t>(rc);
  const size_t M = static_cast<size_t>(batch) * v.cfg->frames;
  if ((rc = upload_and_forward(m, v, x_host, batch))) return static_cast<ishara_status_t>(rc);
  CAPI_CUDA_OK(cudaMemcpyAsync(logits_host, v.logits_own, M * v.cfg->num_classes * sizeof(float), cudaMemcpyDeviceToHost,
                                 v.stream));
  CAPI_CUDA_OK(cudaStreamSynchronize(v.stream));
  return ISHARA_OK;
}

ishara_status_t ishara_model_infer_host(ishara_model_t* mh, const float* x_host, int32_t batch,
                                        const int32_t* labels_host, int32_t max_label_len, float* logits_host,
                                        int32_t* ids_host, int32_t* lens_host, float* nll_host) {
  CHECK_HANDLE(mh);
  if (x_host == nullptr || batch <= 0 || ids_host == nullptr || lens_host == nullptr) {
    set_last_error("infer_host: bad arguments");
    return ISHARA_ERR_INVALID;
  }
  if (labels_host != nullptr && (nll_host == nullptr || max_label_len <= 0)) {
    set_last_error("infer_host: labels need nll_host and max_label_len > 0");
    return ISHARA_ERR_INVALID;
  }
  ishara_model* m = reinterpret_cast<ishara_model*>(mh);
  ModelView v;
  int rc = model_view(m, batch, &v);
  if (rc) return static_cast<ishara_status_t>(rc);
  const ishara_config_t& c = *v.cfg;
  const size_t M = static_cast<size_t>(batch) * c.frames;
  const int blank = c.num_classes - 1;  // pad_token_idx = 59 (c1:4-7)
  int32_t* labels_dev = nullptr;
  if (labels_host != nullptr) {
    if ((rc = model_labels_buffer(m, static_cast<size_t>(batch) * max_label_len, &labels_dev)))
      return static_cast<ishara_status_t>(rc);
    CAPI_CUDA_OK(cudaMemcpyAsync(labels_dev, labels_host, static_cast<size_t>(batch) * max_label_len * sizeof(int32_t),
                                   cudaMemcpyHostToDevice, v.stream));
  }
  if ((rc = upload_and_forward(m, v, x_host, batch))) return static_cast<ishara_status_t>(rc);
  if ((rc = greedy_decode_launch(v.logits_own, batch, c.frames, c.num_classes, blank, v.ids_dev, v.lens_dev, v.stream)))
    return static_cast<ishara_status_t>(rc);
  if (labels_host != nullptr) {
    if ((rc = ctc_loss_launch(v.logits_own, labels_dev, batch, c.frames, c.num_classes, max_label_len, blank, v.nll_dev,
                              nullptr, nullptr, 0, v.stream)))
      return static_cast<ishara_status_t>(rc);
    CAPI_CUDA_OK(cudaMemcpyAsync(nll_host, v.nll_dev, batch * sizeof(float), cudaMemcpyDeviceToHost, v.stream));
  }
  CAPI_CUDA_OK(cudaMemcpyAsync(ids_host, v.ids_dev, M * sizeof(int32_t), cudaMemcpyDeviceToHost, v.stream));
  CAPI_CUDA_OK(cudaMemcpyAsync(lens_host, v.lens_dev, batch * sizeof(int32_t), cudaMemcpyDeviceToHost, v.stream));
  if (logits_host != nullptr)
    CAPI_CUDA_OK(cudaMemcpyAsync(logits_host, v.logits_own, M * c.num_classes * sizeof(float), cudaMemcpyDeviceToHost,
                                   v.stream));
  CAPI_CUDA_OK(cudaStreamSynchronize(v.stream));
  return ISHARA_OK;
}

ishara_status_t ishara_model_infer_submit(ishara_model_t* mh, const float* x_host, int32_t batch, const int32_t* labels_host,
                                          int32_t max_label_len, float* logits_host, int32_t* ids_host, int32_t* lens_host,
                                          float* nll_host) {
  CHECK_HANDLE(mh);
  if (x_host == nullptr || batch <= 0 || ids_host == nullptr || lens_host == nullptr) {
    set_last_error("infer_submit: bad arguments");
    return ISHARA_ERR_INVALID;
  }
  if (labels_host != nullptr && (nll_host == nullptr || max_label_len <= 0)) {
    set_last_error("infer_submit: labels need nll_host and max_label_len > 0");
    return ISHARA_ERR_INVALID;
  }
  return static_cast<ishara_status_t>(model_infer_submit(reinterpret_cast<ishara_model*>(mh), x_host, batch, labels_host, max_label_len,
                                                         logits_host, ids_host, lens_host, nll_host));
}
ishara_status_t ishara_model_infer_collect(ishara_model_t* mh) {
  CHECK_HANDLE(mh);
  return static_cast<ishara_status_t>(model_infer_collect(reinterpret_cast<ishara_model*>(mh)));
}

ishara_status_t ishara_ctc_loss(const float* logits_dev, const int32_t* labels_dev, int32_t batch, int32_t frames,
                                int32_t num_classes, int32_t max_label_len, int32_t blank, float* nll_dev,
                                float* grad_dev, void* stream) {
  if (logits_dev == nullptr || labels_dev == nullptr || nll_dev == nullptr) {
    set_last_error("ctc_loss: null pointer");
    return ISHARA_ERR_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* ws = nullptr;
  size_t ws_bytes = 0;
  if (grad_dev != nullptr) {
    // this entry point has no handle to own the alpha/beta workspace: a stream-ordered allocation on the caller's stream
    // and device lives exactly as long as the two kernels that use it
    ws_bytes = ctc_workspace_bytes(batch, frames, max_label_len);
    CAPI_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&ws), ws_bytes, s));
  }
  const int rc = ctc_loss_launch(logits_dev, labels_dev, batch, frames, num_classes, max_label_len, blank, nll_dev, grad_dev, ws,
                                 ws_bytes, s);
  if (ws != nullptr) cudaFreeAsync(ws, s);
  return static_cast<ishara_status_t>(rc);
}

ishara_status_t ishara_greedy_decode(const float* logits_dev, int32_t batch, int32_t frames, int32_t num_classes,
                                     int32_t blank, int32_t* ids_dev, int32_t* lens_dev, void* stream) {
  if (logits_dev == nullptr || ids_dev == nullptr || lens_dev == nullptr) {
    set_last_error("greedy_decode: null pointer");
    return ISHARA_ERR_INVALID;
  }
  return static_cast<ishara_status_t>(greedy_decode_launch(logits_dev, batch, frames, num_classes, blank, ids_dev,
                                                           lens_dev, static_cast<cudaStream_t>(stream)));
}

ishara_status_t ishara_op_gemm(const ishara_gemm_args_t* a, void* stream) {
  if (a == nullptr || a->a == nullptr || a->wt == nullptr || a->out0 == nullptr) {
    set_last_error("op_gemm: null pointer");
    return ISHARA_ERR_INVALID;
  }
  GemmPlan p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.block_n = a->block_n;
  p.out_f32 = a->out_f32 != 0;
  p.row_mode = a->row_mode != 0;
  p.epi.bias = a->bias;
  p.epi.gate = a->gate;
  p.epi.rowtab = a->rowtab;
  p.epi.resid = static_cast<const bf16*>(a->resid);
  p.epi.ln0_g = a->ln0_g; p.epi.ln0_b = a->ln0_b; p.epi.ln0_eps = a->ln0_eps;
  p.epi.ln1_g = a->ln1_g; p.epi.ln1_b = a->ln1_b; p.epi.ln1_eps = a->ln1_eps;
  p.epi.rows_per_seq = a->rows_per_seq > 0 ? a->rows_per_seq : 1;
  p.epi.act = a->act;
  const int nfull = a->act == ACT_GLU ? a->N / 2 : a->N;
  const int nout = (a->nout > 0 && a->nout <= nfull) ? a->nout : nfull;
  p.epi.ld_resid = nout;
  if ((a->ln0_g != nullptr || a->ln1_g != nullptr) && !p.row_mode) {
    set_last_error("op_gemm: LayerNorm fusion needs row_mode");
    return ISHARA_ERR_SHAPE;
  }
  if (a->ln1_g != nullptr && a->out1 == nullptr) {
    set_last_error("op_gemm: ln1 needs out1");
    return ISHARA_ERR_INVALID;
  }
  int rc = gemm_plan_init(&p, static_cast<const bf16*>(a->a), a->lda, static_cast<const bf16*>(a->wt), a->out0, nout, nout,
                          static_cast<bf16*>(a->out1), nout);
  if (rc) return static_cast<ishara_status_t>(rc);
  int dev = 0, sms = 148;
  CAPI_CUDA_OK(cudaGetDevice(&dev));
  CAPI_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return static_cast<ishara_status_t>(gemm_launch(p, sms, static_cast<cudaStream_t>(stream)));
}

ishara_status_t ishara_op_dwconv(const void* in_bf16, void* out_bf16, const float* w, const float* bias,
                                 const float* eca_w, float* colsum, int32_t B, int32_t T, int32_t C, int32_t k,
                                 int32_t pad_left, int32_t post, void* stream) {
  if (in_bf16 == nullptr || out_bf16 == nullptr || w == nullptr) {
    set_last_error("op_dwconv: null pointer");
    return ISHARA_ERR_INVALID;
  }
  DwConvArgs a;
  a.in = static_cast<const bf16*>(in_bf16);
  a.out = static_cast<bf16*>(out_bf16);
  a.w = w; a.bias = bias; a.eca_w = eca_w; a.colsum = colsum;
  a.B = B; a.T = T; a.C = C; a.k = k; a.pad_left = pad_left; a.post = post;
  return static_cast<ishara_status_t>(dwconv_launch(a, static_cast<cudaStream_t>(stream)));
}

ishara_status_t ishara_op_attention(const void* qkv_bf16, void* out_bf16, const uint8_t* key_mask, int32_t B, int32_t T,
                                    int32_t H, int32_t dh, float scale, void* stream) {
  if (qkv_bf16 == nullptr || out_bf16 == nullptr) {
    set_last_error("op_attention: null pointer");
    return ISHARA_ERR_INVALID;
  }
  AttnArgs a;
  a.qkv = static_cast<const bf16*>(qkv_bf16);
  a.out = static_cast<bf16*>(out_bf16);
  a.key_mask = key_mask;
  a.B = B; a.T = T; a.H = H; a.dh = dh; a.scale = scale;
  return static_cast<ishara_status_t>(attention_launch(a, static_cast<cudaStream_t>(stream)));
}

ishara_status_t ishara_op_relpos_attention(const void* qkv_bf16, const void* pos_bf16, const float* u_bias,
                                           const float* v_bias, void* out_bf16, const uint8_t* key_mask, int32_t B,
                                           int32_t T, int32_t H, int32_t dh, float scale, void* stream) {
  if (qkv_bf16 == nullptr || pos_bf16 == nullptr || u_bias == nullptr || v_bias == nullptr || out_bf16 == nullptr) {
    set_last_error("op_relpos_attention: null pointer");
    return ISHARA_ERR_INVALID;
  }
  AttnArgs a;
  a.qkv = static_cast<const bf16*>(qkv_bf16);
  a.pos = static_cast<const bf16*>(pos_bf16);
  a.u_bias = u_bias; a.v_bias = v_bias;
  a.out = static_cast<bf16*>(out_bf16);
  a.key_mask = key_mask;
  a.B = B; a.T = T; a.H = H; a.dh = dh; a.scale = scale;
  return static_cast<ishara_status_t>(attention_launch(a, static_cast<cudaStream_t>(stream)));
}

ishara_status_t ishara_op_time_reduce(const void* x_bf16, void* out_bf16, const float* w9_host, float bias, int32_t B,
                                      int32_t T, int32_t D, int32_t ldo, void* stream) {
  if (x_bf16 == nullptr || out_bf16 == nullptr || w9_host == nullptr) {
    set_last_error("op_time_reduce: null pointer");
    return ISHARA_ERR_INVALID;
  }
  TimeReduceArgs a;
  a.x = static_cast<const bf16*>(x_bf16);
  a.out = static_cast<bf16*>(out_bf16);
  for (int i = 0; i < 9; ++i) a.w[i] = w9_host[i];
  a.bias = bias;
  a.B = B; a.T = T; a.D = D; a.T2 = (T - 3) / 2 + 1; a.D2 = (D - 3) / 2 + 1; a.ldo = ldo;
  return static_cast<ishara_status_t>(time_reduce_launch(a, static_cast<cudaStream_t>(stream)));
}

ishara_status_t ishara_op_upsample_add(const void* y_bf16, const void* rec_bf16, void* out_bf16, int32_t B, int32_t T2,
                                       int32_t T, int32_t D, void* stream) {
  if (y_bf16 == nullptr || rec_bf16 == nullptr || out_bf16 == nullptr) {
    set_last_error("op_upsample_add: null pointer");
    return ISHARA_ERR_INVALID;
  }
  return static_cast<ishara_status_t>(upsample_add_launch(static_cast<const bf16*>(y_bf16), static_cast<const bf16*>(rec_bf16),
                                                          static_cast<bf16*>(out_bf16), B, T2, T, D,
                                                          static_cast<cudaStream_t>(stream)));
}

ishara_status_t ishara_op_conv2d_subsample(const float* x, void* out_bf16, const float* w1, const float* b1,
                                           const float* w2, const float* b2, int32_t B, int32_t T, int32_t F, int32_t C,
                                           int32_t ldo, void* stream) {
  if (x == nullptr || out_bf16 == nullptr || w1 == nullptr || b1 == nullptr || w2 == nullptr || b2 == nullptr) {
    set_last_error("op_conv2d_subsample: null pointer");
    return ISHARA_ERR_INVALID;
  }
  Conv2dSubsampleArgs a;
  a.x = x; a.out = static_cast<bf16*>(out_bf16);
  a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2;
  a.B = B; a.T = T; a.F = F; a.C = C;
  a.T4 = (((T - 3) / 2 + 1) - 3) / 2 + 1;
  a.F4 = (((F - 3) / 2 + 1) - 3) / 2 + 1;
  a.ldo = ldo;
  return static_cast<ishara_status_t>(conv2d_subsample_launch(a, static_cast<cudaStream_t>(stream)));
}

ishara_status_t ishara_op_layernorm(const void* x_bf16, void* out_bf16, const float* gamma, const float* beta, float eps,
                                    int64_t M, int32_t D, void* stream) {
  if (x_bf16 == nullptr || out_bf16 == nullptr || gamma == nullptr || beta == nullptr) {
    set_last_error("op_layernorm: null pointer");
    return ISHARA_ERR_INVALID;
  }
  return static_cast<ishara_status_t>(layernorm_launch(static_cast<const bf16*>(x_bf16), static_cast<bf16*>(out_bf16),
                                                       gamma, beta, eps, M, D, static_cast<cudaStream_t>(stream)));
}

ishara_status_t ishara_op_cast_pad(const float* x, void* out_bf16, int64_t M, int32_t F, int32_t Fpad, void* stream) {
  if (x == nullptr || out_bf16 == nullptr) {
    set_last_error("op_cast_pad: null pointer");
    return ISHARA_ERR_INVALID;
  }
  return static_cast<ishara_status_t>(
      cast_pad_launch(x, static_cast<bf16*>(out_bf16), M, F, Fpad, static_cast<cudaStream_t>(stream)));
}

ishara_status_t ishara_ids_to_text(const int32_t* ids_host, const int32_t* lens_host, int32_t batch, int32_t frames,
                                   const char* table, int32_t table_len, char* out, int64_t* offsets) {
  if (ids_host == nullptr || lens_host == nullptr || table == nullptr || out == nullptr || offsets == nullptr || batch < 0) {
    set_last_error("ids_to_text: null pointer");
    return ISHARA_ERR_INVALID;
  }
  int64_t n = 0;
  for (int b = 0; b < batch; ++b) {
    offsets[b] = n;
    const int32_t* row = ids_host + static_cast<size_t>(b) * frames;
    const int len = lens_host[b] < frames ? lens_host[b] : frames;
    for (int i = 0; i < len; ++i) {
      const int32_t id = row[i];
      if (id >= 0 && id < table_len) out[n++] = table[id];
    }
  }
  offsets[batch] = n;
  return ISHARA_OK;
}

ishara_status_t ishara_device_malloc(int32_t device, int64_t bytes, void** out_dev) {
  if (out_dev == nullptr || bytes < 0) { set_last_error("device_malloc: bad arguments"); return ISHARA_ERR_INVALID; }
  CAPI_CUDA_OK(cudaSetDevice(device));
  CAPI_CUDA_OK(cudaMalloc(out_dev, static_cast<size_t>(bytes > 0 ? bytes : 1)));
  return ISHARA_OK;
}
ishara_status_t ishara_device_free(int32_t device, void* dev_ptr) {
  if (dev_ptr == nullptr) return ISHARA_OK;
  CAPI_CUDA_OK(cudaSetDevice(device));
  CAPI_CUDA_OK(cudaFree(dev_ptr));
  return ISHARA_OK;
}
ishara_status_t ishara_host_malloc_pinned(int64_t bytes, void** out_host) {
  if (out_host == nullptr || bytes < 0) { set_last_error("host_malloc_pinned: bad arguments"); return ISHARA_ERR_INVALID; }
  CAPI_CUDA_OK(cudaMallocHost(out_host, static_cast<size_t>(bytes > 0 ? bytes : 1)));
  return ISHARA_OK;
}
ishara_status_t ishara_host_free_pinned(void* host_ptr) {
  if (host_ptr == nullptr) return ISHARA_OK;
  CAPI_CUDA_OK(cudaFreeHost(host_ptr));
  return ISHARA_OK;
}
ishara_status_t ishara_memcpy_async(void* dst, const void* src, int64_t bytes, int32_t kind, void* stream) {
  if (dst == nullptr || src == nullptr || bytes < 0 || kind < 1 || kind > 3) {
    set_last_error("memcpy_async: bad arguments");
    return ISHARA_ERR_INVALID;
  }
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  CAPI_CUDA_OK(cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), k, static_cast<cudaStream_t>(stream)));
  return ISHARA_OK;
}
ishara_status_t ishara_stream_synchronize(int32_t device, void* stream) {
  CAPI_CUDA_OK(cudaSetDevice(device));
  CAPI_CUDA_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return ISHARA_OK;
}

ishara_status_t ishara_edit_distances(const char* const* preds, const int32_t* pred_lens, const char* const* targets,
                                      const int32_t* target_lens, int32_t n, int32_t* out) {
  if (n < 0 || (n > 0 && (preds == nullptr || pred_lens == nullptr || targets == nullptr || target_lens == nullptr || out == nullptr))) {
    set_last_error("edit_distances: null argument");
    return ISHARA_ERR_INVALID;
  }
  std::vector<int32_t> row;
  for (int32_t i = 0; i < n; ++i) {
    const char *a = preds[i], *b = targets[i];
    const int la = pred_lens[i], lb = target_lens[i];
    if (la < 0 || lb < 0 || (la > 0 && a == nullptr) || (lb > 0 && b == nullptr)) {
      set_last_error("edit_distances: bad string");
      return ISHARA_ERR_INVALID;
    }
    row.resize(static_cast<size_t>(lb) + 1);
    for (int j = 0; j <= lb; ++j) row[j] = j;
    for (int p = 1; p <= la; ++p) {
      int32_t diag = row[0];
      row[0] = p;
      for (int j = 1; j <= lb; ++j) {
        const int32_t sub = diag + (a[p - 1] == b[j - 1] ? 0 : 1);
        diag = row[j];
        const int32_t del = row[j] + 1, ins = row[j - 1] + 1;
        row[j] = sub < del ? (sub < ins ? sub : ins) : (del < ins ? del : ins);
      }
    }
    out[i] = row[lb];
  }
  return ISHARA_OK;
}

ishara_status_t ishara_preprocess(const float* frames_dev, const int32_t* offsets_dev, int32_t batch, int32_t max_frames,
                                  const float* mean_dev, const float* std_dev, int32_t frame_len, int32_t filter_frames, float* out_dev,
                                  void* stream) {
  if (offsets_dev == nullptr || mean_dev == nullptr || std_dev == nullptr || out_dev == nullptr || (frames_dev == nullptr && max_frames > 0)) {
    set_last_error("preprocess: null argument");
    return ISHARA_ERR_INVALID;
  }
  return static_cast<ishara_status_t>(preprocess_launch(frames_dev, offsets_dev, batch, max_frames, mean_dev, std_dev, frame_len,
                                                        filter_frames, out_dev, static_cast<cudaStream_t>(stream)));
}

// ---- training step (train.cu) ---------------------------------------------------------------------
ishara_status_t ishara_model_train_configure(ishara_model_t* m, float dropout_rate, uint64_t seed, int32_t debug) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_configure(reinterpret_cast<ishara_model*>(m), dropout_rate, seed, debug));
}
ishara_status_t ishara_model_train_forward_backward(ishara_model_t* m, const float* x_dev, const int32_t* labels_dev, int32_t batch,
                                                    int32_t labels_len, float* loss_host, void* stream) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_forward_backward(reinterpret_cast<ishara_model*>(m), x_dev, labels_dev, batch, labels_len,
                                                             loss_host, static_cast<cudaStream_t>(stream)));
}
ishara_status_t ishara_model_train_grad_buffer(ishara_model_t* m, float** grad_dev, int64_t* numel) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_grad_buffer(reinterpret_cast<ishara_model*>(m), grad_dev, numel));
}
ishara_status_t ishara_model_train_apply(ishara_model_t* m, const ishara_adamw_t* opt, float grad_scale, void* stream) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_apply(reinterpret_cast<ishara_model*>(m), opt, grad_scale, static_cast<cudaStream_t>(stream)));
}
ishara_status_t ishara_model_train_step(ishara_model_t* m, const float* x_dev, const int32_t* labels_dev, int32_t batch, int32_t labels_len,
                                        const ishara_adamw_t* opt, float* loss_host, void* stream) {
  CHECK_HANDLE(m);
  ishara_model* mm = reinterpret_cast<ishara_model*>(m);
  int rc = train_forward_backward(mm, x_dev, labels_dev, batch, labels_len, nullptr, static_cast<cudaStream_t>(stream));
  if (rc) return static_cast<ishara_status_t>(rc);
  rc = train_apply(mm, opt, 1.f, static_cast<cudaStream_t>(stream));
  if (rc) return static_cast<ishara_status_t>(rc);
  if (loss_host != nullptr) {
    // the loss of THIS step's forward; read back after the update was enqueued so the copy overlaps nothing critical
    rc = train_forward_backward_loss(mm, loss_host, static_cast<cudaStream_t>(stream));
  }
  return static_cast<ishara_status_t>(rc);
}
ishara_status_t ishara_model_train_step_host(ishara_model_t* mh, const float* x_host, const int32_t* labels_host, int32_t batch,
                                             int32_t labels_len, const ishara_adamw_t* opt, float* loss_host) {
  CHECK_HANDLE(mh);
  if (x_host == nullptr || labels_host == nullptr || batch <= 0 || labels_len <= 0) {
    set_last_error("train_step_host: bad arguments");
    return ISHARA_ERR_INVALID;
  }
  ishara_model* m = reinterpret_cast<ishara_model*>(mh);
  ModelView v;
  int rc = model_view(m, batch, &v);
  if (rc) return static_cast<ishara_status_t>(rc);
  int32_t* labels_dev = nullptr;
  if ((rc = model_labels_buffer(m, static_cast<size_t>(batch) * labels_len, &labels_dev))) return static_cast<ishara_status_t>(rc);
  const size_t xbytes = static_cast<size_t>(batch) * v.cfg->frames * v.cfg->features * sizeof(float);
  CAPI_CUDA_OK(cudaMemcpyAsync(v.x_dev, x_host, xbytes, cudaMemcpyHostToDevice, v.stream));
  CAPI_CUDA_OK(cudaMemcpyAsync(labels_dev, labels_host, static_cast<size_t>(batch) * labels_len * sizeof(int32_t), cudaMemcpyHostToDevice, v.stream));
  return ishara_model_train_step(mh, v.x_dev, labels_dev, batch, labels_len, opt, loss_host, v.stream);
}
ishara_status_t ishara_model_train_sync(ishara_model_t* m) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_sync(reinterpret_cast<ishara_model*>(m)));
}
ishara_status_t ishara_model_train_param_grad(ishara_model_t* m, const char* name, float* host_out, int64_t numel) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_param_grad(reinterpret_cast<ishara_model*>(m), name, host_out, numel));
}
ishara_status_t ishara_model_train_fetch(ishara_model_t* m, const char* name, int32_t want_grad, float* host_out, int64_t numel) {
  CHECK_HANDLE(m);
  if (name == nullptr || host_out == nullptr) { set_last_error("train_fetch: null argument"); return ISHARA_ERR_INVALID; }
  return static_cast<ishara_status_t>(train_fetch(reinterpret_cast<ishara_model*>(m), name, want_grad, host_out, numel));
}

ishara_status_t ishara_model_train_apply_radam(ishara_model_t* m, const ishara_radam_lookahead_t* opt, float grad_scale, void* stream) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_apply_radam(reinterpret_cast<ishara_model*>(m), opt, grad_scale, static_cast<cudaStream_t>(stream)));
}
ishara_status_t ishara_model_train_state_info(ishara_model_t* m, int64_t* numel, int64_t* opt_steps, int64_t* fb_steps, int32_t* has_slow) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_state_info(reinterpret_cast<ishara_model*>(m), numel, opt_steps, fb_steps, has_slow));
}
ishara_status_t ishara_model_train_state_get(ishara_model_t* m, int32_t which, float* host_out, int64_t numel) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_state_get(reinterpret_cast<ishara_model*>(m), which, host_out, numel));
}
ishara_status_t ishara_model_train_state_set(ishara_model_t* m, int32_t which, const float* host_in, int64_t numel) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_state_set(reinterpret_cast<ishara_model*>(m), which, host_in, numel));
}
ishara_status_t ishara_model_train_state_set_counters(ishara_model_t* m, int64_t opt_steps, int64_t fb_steps) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_state_set_counters(reinterpret_cast<ishara_model*>(m), opt_steps, fb_steps));
}
ishara_status_t ishara_model_set_mask_mode(ishara_model_t* m, int32_t mode) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(model_set_mask_mode(reinterpret_cast<ishara_model*>(m), mode));
}
ishara_status_t ishara_model_forward_masked(ishara_model_t* m, const float* x_dev, const uint8_t* mask_dev, int32_t batch, float* logits_dev,
                                            void* stream) {
  CHECK_HANDLE(m);
  if (x_dev == nullptr || logits_dev == nullptr || batch <= 0) { set_last_error("forward_masked: bad arguments"); return ISHARA_ERR_INVALID; }
  return static_cast<ishara_status_t>(
      model_forward_masked(reinterpret_cast<ishara_model*>(m), x_dev, mask_dev, batch, logits_dev, static_cast<cudaStream_t>(stream)));
}
ishara_status_t ishara_model_train_loss(ishara_model_t* m, float* loss_host, void* stream) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_forward_backward_loss(reinterpret_cast<ishara_model*>(m), loss_host, static_cast<cudaStream_t>(stream)));
}
ishara_status_t ishara_model_train_counters(ishara_model_t* m, int64_t* fb_steps, int64_t* opt_steps, int64_t* skipped_steps) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(train_counters(reinterpret_cast<ishara_model*>(m), fb_steps, opt_steps, skipped_steps));
}
ishara_status_t ishara_comm_unique_id(void* out_id128) {
  if (out_id128 == nullptr) { set_last_error("comm_unique_id: null pointer"); return ISHARA_ERR_INVALID; }
  return static_cast<ishara_status_t>(comm_unique_id(out_id128));
}
ishara_status_t ishara_model_comm_init(ishara_model_t* m, const void* id128, int32_t rank, int32_t world) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(model_comm_init(reinterpret_cast<ishara_model*>(m), id128, rank, world));
}
ishara_status_t ishara_model_comm_destroy(ishara_model_t* m) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(model_comm_destroy(reinterpret_cast<ishara_model*>(m)));
}
int32_t ishara_nccl_version(void) { return nccl_version(); }
ishara_status_t ishara_comm_bucket_plan(const int64_t* hi, int32_t n, int64_t n_train, int64_t min_elems, int64_t* lo_out,
                                        int64_t* up_out) {
  if (hi == nullptr || lo_out == nullptr || up_out == nullptr || n <= 0 || n_train < 0) {
    set_last_error("comm_bucket_plan: bad arguments");
    return ISHARA_ERR_INVALID;
  }
  comm_bucket_plan(hi, n, n_train, min_elems, lo_out, up_out);
  return ISHARA_OK;
}

ishara_status_t ishara_model_set_profile(ishara_model_t* m, int32_t on) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(model_set_profile(reinterpret_cast<ishara_model*>(m), on));
}
int32_t ishara_model_profile_count(const ishara_model_t* m) {
  if (m == nullptr) return 0;
  return model_profile_count(reinterpret_cast<const ishara_model*>(m));
}
ishara_status_t ishara_model_profile_entry(ishara_model_t* m, int32_t index, const char** label, const char** kind,
                                           float* ms, double* flops, double* bytes) {
  CHECK_HANDLE(m);
  return static_cast<ishara_status_t>(
      model_profile_entry(reinterpret_cast<ishara_model*>(m), index, label, kind, ms, flops, bytes));
}
uint64_t ishara_launch_count(void) { return launch_count(); }

void* ishara_model_stream(ishara_model_t* m) {
  if (m == nullptr) return nullptr;
  ModelView v;
  if (model_view(reinterpret_cast<ishara_model*>(m), 1, &v)) return nullptr;
  return v.stream;
}

}  // extern "C"
