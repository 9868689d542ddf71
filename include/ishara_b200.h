/* ishara_b200 — C ABI of the B200-native (sm_100a) Ishara encoder hot path.
 *
 * Drop-in boundary for the reference's model call. The reference has no FFI of its
 * own — its boundary is the Keras model call inside the notebooks — so every entry point below names
 * the reference call site it replaces:
 *
 *   ishara_model_create / set_param / finalize   <- get_model(...)            nb:conv-hybrid-model c7:1-72
 *                                                   model.load_weights / save_weights           c9:10
 *   ishara_model_forward[_host]                  <- model(x) / model(x, training=False)   c7:82, c9:15, c13:17
 *   ishara_ctc_loss                              <- CTCLoss(labels, logits)                     c6:1-13
 *   ishara_greedy_decode                         <- decode_phrase / decode_batch_predictions    c8:4-20
 *   ishara_model_train_*                         <- model.fit's inner step: model(x, training=True), CTCLoss,
 *                                                   gradients, clip-norm + AdamW        c12:1, c7:67-70,
 *                                                   integration.py:675-679,750
 *   ishara_preprocess                            <- pre_process00 + pre_process1 (landmark gather, hand-frame filter,
 *                                                   resize_pad, normalise)       c3:1-115, c13:9-15
 *   ishara_op_*                                  <- operator-level building blocks (tests, P-rows of §8a)
 *
 * Conventions
 *   - plain pointers and sizes only; all tensors are dense, row-major, C order.
 *   - every function returns an ishara_status_t (0 = ok); ishara_last_error() gives a thread-local
 *     message for the last failure on the calling thread. Nothing aborts.
 *   - "dev" pointers are device memory on the model's device; "host" pointers are host memory
 *     (pinned or pageable). The caller owns every buffer; the library only owns its handle
 *     (packed weights, workspace, TMA descriptors) and never retains caller pointers past a call.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it without a host sync unless
 *     the function name ends in _host.
 *   - a handle is bound to one device and is not re-entrant; distinct handles may be driven from
 *     distinct threads. There is no CPU fallback: without a Blackwell GPU every compute call fails
 *     with ISHARA_ERR_CUDA.
 */
#ifndef ISHARA_B200_H_
#define ISHARA_B200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define ISHARA_API __attribute__((visibility("default")))
#else
#define ISHARA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ISHARA_OK = 0,
  ISHARA_ERR_INVALID = 1,   /* null handle / bad enum / unknown parameter name */
  ISHARA_ERR_SHAPE = 2,     /* shape, alignment or size mismatch */
  ISHARA_ERR_CUDA = 3,      /* CUDA runtime / driver error (message has the details) */
  ISHARA_ERR_STATE = 4,     /* call order (e.g. forward before finalize, missing parameters) */
  ISHARA_ERR_COMM = 5       /* NCCL missing / NCCL error (data-parallel exchange) */
} ishara_status_t;

typedef struct ishara_model ishara_model_t;

/* get_model(...) keyword arguments (c7:1-11) plus the two module-level globals it closes over
 * (INPUT_SHAPE c1:27 and len(char_to_num) c1:4-7). */
typedef struct {
  int32_t dim;                      /* 256 */
  int32_t num_conv_squeeze_blocks;  /* 2 */
  int32_t num_conv_conform_blocks;  /* 2 */
  int32_t num_conv_per_block;       /* 3 */
  int32_t kernel_sizes[8];          /* {11,5,3} */
  int32_t num_kernel_sizes;         /* 3 */
  int32_t num_heads;                /* 8 */
  int32_t expansion_factor;         /* 2 */
  int32_t transformer_kernel_size;  /* 15 */
  int32_t frames;                   /* T = INPUT_SHAPE[0] (384) */
  int32_t features;                 /* INPUT_SHAPE[1] (276) */
  int32_t num_classes;              /* 60 */
} ishara_config_t;

ISHARA_API const char* ishara_version(void);
ISHARA_API const char* ishara_last_error(void);
/* number of CUDA devices visible, or 0 (never fails) */
ISHARA_API int32_t ishara_device_count(void);

/* ---- model lifetime ------------------------------------------------------------------------- */
ISHARA_API ishara_status_t ishara_model_create(const ishara_config_t* cfg, int32_t device, ishara_model_t** out);
ISHARA_API ishara_status_t ishara_model_destroy(ishara_model_t* m);

/* Parameter enumeration in canonical order (Keras layer names). */
ISHARA_API int32_t ishara_model_num_params(const ishara_model_t* m);
ISHARA_API ishara_status_t ishara_model_param_info(const ishara_model_t* m, int32_t index, const char** name,
                                        int64_t* numel, int32_t* ndim, int64_t shape[4]);
/* Upload one parameter in Keras layout (Dense [in,out]; Conv1D [k,in/groups,out]; depthwise [k,C,1]),
 * fp32, from host memory. */
ISHARA_API ishara_status_t ishara_model_set_param(ishara_model_t* m, const char* name, const float* host_data, int64_t numel);
/* Read back the fp32 master copy. */
ISHARA_API ishara_status_t ishara_model_get_param(const ishara_model_t* m, const char* name, float* host_out, int64_t numel);
/* Pack weights for the kernels (bf16, K-major, BatchNorm folded). Must follow set_param of every
 * parameter and precede forward. May be called again after parameters change. */
ISHARA_API ishara_status_t ishara_model_finalize(ishara_model_t* m);

/* ---- hot path -------------------------------------------------------------------------------- */
/* logits[B,T,num_classes] fp32 = model(x[B,T,features] fp32), inference mode (BatchNorm moving
 * statistics, dropout off). Device pointers. */
ISHARA_API ishara_status_t ishara_model_forward(ishara_model_t* m, const float* x_dev, int32_t batch, float* logits_dev,
                                     void* stream);
/* mask_mode: 0 = "dropped" (default; the reference AS EXECUTED: the Keras mask of Masking(0.0), c7:13, is lost at the
 * TFOpLambda `x + pe`, c7:16, so ECA / SqueezeExcite average over all T frames and the softmax is unmasked), 1 =
 * "propagated" (the authors' apparent intent: mask_t = any(x[t,:] != 0) reaches ECA's and SqueezeExcite's
 * GlobalAveragePooling1D(mask), c5:8-9,129-130, and Softmax(mask) of the SqueezeformerBlock MHSA, c5:109-112; the first
 * ConformerBlock ends it because ConformerBlock has no supports_masking, c5:311-343). Inference only. */
ISHARA_API ishara_status_t ishara_model_set_mask_mode(ishara_model_t* m, int32_t mode);
/* forward with an explicit frame mask: mask_dev uint8 [B,T], 1 = the frame carries data, or NULL = derive it from x
 * like Masking(0.0). A non-NULL mask needs mask_mode = 1. (SURVEY.md §8b `..._forward(model, x, mask_or_null, ...)`.) */
ISHARA_API ishara_status_t ishara_model_forward_masked(ishara_model_t* m, const float* x_dev, const uint8_t* mask_dev, int32_t batch,
                                                       float* logits_dev, void* stream);
/* Same through host buffers: H2D copy, forward, D2H copy, stream sync (the reference-facing call). */
ISHARA_API ishara_status_t ishara_model_forward_host(ishara_model_t* m, const float* x_host, int32_t batch, float* logits_host);
/* Whole inference step through host buffers: forward + greedy decode (+ CTC loss when labels != NULL).
 * ids_host int32 [B,T], lens_host int32 [B], nll_host fp32 [B] (or NULL), logits_host optional (or NULL). */
ISHARA_API ishara_status_t ishara_model_infer_host(ishara_model_t* m, const float* x_host, int32_t batch,
                                        const int32_t* labels_host, int32_t max_label_len, float* logits_host,
                                        int32_t* ids_host, int32_t* lens_host, float* nll_host);

/* Pipelined form of infer_host for back-to-back batches (serving / evaluation loops, c9:15-26 run over a dataset):
 * submit enqueues the upload (copy stream), forward, decode, optional CTC and the read-back for one batch and returns
 * immediately; collect blocks until the OLDEST submitted batch is complete. Up to two batches may be in flight, so the
 * H2D copy of batch i+1 runs under the kernels of batch i. All host buffers (pinned for overlap) must stay valid and
 * untouched until the matching collect returns. */
ISHARA_API ishara_status_t ishara_model_infer_submit(ishara_model_t* m, const float* x_host, int32_t batch, const int32_t* labels_host,
                                                     int32_t max_label_len, float* logits_host, int32_t* ids_host,
                                                     int32_t* lens_host, float* nll_host);
ISHARA_API ishara_status_t ishara_model_infer_collect(ishara_model_t* m);

/* CTCLoss (c6:1-13): per-sequence negative log-likelihood nll[B] (the reference returns their mean) and,
 * when grad_dev != NULL, d nll_b / d logits [B,T,V]. labels int32 [B,L] padded with `blank`. */
ISHARA_API ishara_status_t ishara_ctc_loss(const float* logits_dev, const int32_t* labels_dev, int32_t batch, int32_t frames,
                                int32_t num_classes, int32_t max_label_len, int32_t blank, float* nll_dev,
                                float* grad_dev, void* stream);
/* decode_phrase (c8:4-12) for a batch: ids_dev int32 [B,T] (first lens[b] entries valid, rest -1). */
ISHARA_API ishara_status_t ishara_greedy_decode(const float* logits_dev, int32_t batch, int32_t frames, int32_t num_classes,
                                     int32_t blank, int32_t* ids_dev, int32_t* lens_dev, void* stream);

/* ---- training step ----------------------------------------------------
 * Keras training-mode forward (BatchNormalization on biased batch statistics over (B,T) + moving-average update,
 * dropout at the reference's sites), CTCLoss (mean over the batch), gradients of every trainable tensor, then
 * global-norm clipping and AdamW. Master weights, gradients and Adam moments are fp32 on the device; activations
 * and activation gradients are bf16. dim must be a multiple of 128 (<= 512). The handle's inference path sees the trained weights
 * after ishara_model_train_sync (called implicitly by forward / get_param / infer when weights are stale). */
typedef struct {
  float lr;           /* 4.5e-3  (integration.py:675-679) */
  float weight_decay; /* 0.08, decoupled, applied to every trainable tensor (torch.optim.AdamW default grouping) */
  float beta1;        /* 0.9 */
  float beta2;        /* 0.999 */
  float eps;          /* 1e-8 */
  float max_norm;     /* global-norm clip, 1.0 (integration.py:750); <= 0 disables */
} ishara_adamw_t;

/* dropout_rate = get_model's dropout_rate (0 disables every dropout site; > 0 also enables the head's fixed 0.4,
 * c7:62 and ConformerBlock's default attention dropout 0.1, c5:312); masks are a counter-based hash of (seed, site,
 * element) so forward and backward agree and a host can reproduce them. debug != 0 keeps named activation
 * gradients for ishara_model_train_fetch. */
ISHARA_API ishara_status_t ishara_model_train_configure(ishara_model_t* m, float dropout_rate, uint64_t seed, int32_t debug);
/* x_dev fp32 [batch, T, F], labels_dev int32 [batch, labels_len] padded with the blank (num_classes-1). Leaves the
 * mean-reduced gradients in the flat buffer; loss_host (optional) receives mean CTC loss after a stream sync. */
ISHARA_API ishara_status_t ishara_model_train_forward_backward(ishara_model_t* m, const float* x_dev, const int32_t* labels_dev,
                                                                int32_t batch, int32_t labels_len, float* loss_host, void* stream);
/* mean CTC loss of the LAST forward_backward (mean over all ranks with a communicator): read back after a stream sync.
 * Lets a caller enqueue forward_backward (loss_host = NULL) and train_apply back to back and look at the loss afterwards. */
ISHARA_API ishara_status_t ishara_model_train_loss(ishara_model_t* m, float* loss_host, void* stream);
/* flat fp32 gradient buffer of all trainable tensors (device pointer, element count). With ishara_model_comm_init the
 * library reduces it over the ranks itself; without, a host-side exchange may all-reduce it before train_apply. */
ISHARA_API ishara_status_t ishara_model_train_grad_buffer(ishara_model_t* m, float** grad_dev, int64_t* numel);
ISHARA_API ishara_status_t ishara_model_train_apply(ishara_model_t* m, const ishara_adamw_t* opt, float grad_scale, void* stream);
/* forward_backward + apply in one call (single GPU) */
ISHARA_API ishara_status_t ishara_model_train_step(ishara_model_t* m, const float* x_dev, const int32_t* labels_dev, int32_t batch,
                                                    int32_t labels_len, const ishara_adamw_t* opt, float* loss_host, void* stream);
/* same with host buffers (H2D inside), on the handle's stream; returns after a sync */
ISHARA_API ishara_status_t ishara_model_train_step_host(ishara_model_t* m, const float* x_host, const int32_t* labels_host,
                                                         int32_t batch, int32_t labels_len, const ishara_adamw_t* opt, float* loss_host);
/* The optimiser the reference itself compiles the model with (nb:conv-hybrid-model c7:68-69):
 * tfa.optimizers.Lookahead(tfa.optimizers.RectifiedAdam(sma_threshold=4), sync_period=5). tensorflow_addons is a
 * third-party dependency that is neither vendored nor pinned (Dockerfile:20-21); the arithmetic restated here is its
 * published algorithm (rectified_adam.py with total_steps = 0, lookahead.py with slow_step_size = 0.5). Same role as
 * train_apply; the two must not be mixed on one handle (they share the moment buffers and the step counter).
 * Note: WeightDecayCallback (c11:58-65) assigns `weight_decay` on the Lookahead wrapper, which does not reach the
 * wrapped RectifiedAdam (built with weight_decay = 0): the reference as executed decays nothing, hence the default 0. */
typedef struct {
  float lr;             /* 1e-3 (tfa default; the per-epoch value comes from the LR schedule c11:51-55) */
  float weight_decay;   /* 0 */
  float beta1;          /* 0.9 */
  float beta2;          /* 0.999 */
  float eps;            /* 1e-7 */
  float max_norm;       /* global-norm clip; <= 0 disables (Keras default: none) */
  float sma_threshold;  /* 4 (c7:68) */
  int32_t sync_period;  /* 5 (c7:69) */
  float slow_step_size; /* 0.5 */
} ishara_radam_lookahead_t;
ISHARA_API ishara_status_t ishara_model_train_apply_radam(ishara_model_t* m, const ishara_radam_lookahead_t* opt, float grad_scale,
                                                          void* stream);
/* ---- optimiser-state checkpoint (model.save_weights c9:10, integration.py:912-958) --------
 * Flat fp32 slots in the order of the parameter table (ishara_model_param_info gives names and sizes; every slot has
 * `numel` = the sum of all parameter sizes): which = 0 Adam m, 1 Adam v, 2 Lookahead slow weights (only after a RAdam
 * step), 3 master weights (read only; weights are restored through set_param). Counters: optimiser steps (bias
 * correction, Lookahead phase) and forward/backward passes (dropout stream). Host buffers; synchronises the device. */
ISHARA_API ishara_status_t ishara_model_train_state_info(ishara_model_t* m, int64_t* numel, int64_t* opt_steps, int64_t* fb_steps,
                                                         int32_t* has_slow);
ISHARA_API ishara_status_t ishara_model_train_state_get(ishara_model_t* m, int32_t which, float* host_out, int64_t numel);
ISHARA_API ishara_status_t ishara_model_train_state_set(ishara_model_t* m, int32_t which, const float* host_in, int64_t numel);
ISHARA_API ishara_status_t ishara_model_train_state_set_counters(ishara_model_t* m, int64_t opt_steps, int64_t fb_steps);
/* device masters -> parameter table (get_param) -> inference packs */
ISHARA_API ishara_status_t ishara_model_train_sync(ishara_model_t* m);
/* gradient of one named parameter (Keras layout) after train_forward_backward */
ISHARA_API ishara_status_t ishara_model_train_param_grad(ishara_model_t* m, const char* name, float* host_out, int64_t numel);
/* named activation (want_grad = 0) or its gradient (want_grad = 1, debug mode) as fp32 */
ISHARA_API ishara_status_t ishara_model_train_fetch(ishara_model_t* m, const char* name, int32_t want_grad, float* host_out, int64_t numel);

/* forward/backward passes since train_configure (the dropout noise of pass n is a function of (seed, n)), optimiser
 * steps taken, and optimiser steps SKIPPED because the global gradient norm was not finite (an infeasible CTC
 * alignment gives +inf loss and NaN gradients: the update is dropped instead of poisoning weights and moments). Any
 * pointer may be NULL. Synchronises the device. */
ISHARA_API ishara_status_t ishara_model_train_counters(ishara_model_t* m, int64_t* fb_steps, int64_t* opt_steps, int64_t* skipped_steps);

/* ---- data-parallel exchange (SURVEY.md §8b `ishara_model_comm_init`, §8e) ---------------------------
 * One process per GPU. Rank 0 creates a 128-byte NCCL unique id (ishara_comm_unique_id), every rank receives it out of
 * band (ishara_b200/parallel.py: torch.distributed broadcast or a plain TCP rendezvous) and calls comm_init. From then on
 * ishara_model_train_forward_backward sums the gradients of all ranks INSIDE the library: bucketed per module and
 * issued on the handle's communication stream as soon as a module's backward has finished, so the NVLink transfer
 * overlaps the rest of the backward pass; train_apply(grad_scale = 1/world) waits for it on the device. The loss
 * returned by the training entry points is the mean over all ranks. BatchNorm statistics stay per rank (no SyncBN in
 * the reference). Reference counterpart: nn.DataParallel, integration.py:1058-1060. NCCL is bound at run time
 * (libnccl.so.2, or $ISHARA_NCCL_LIB); without it these calls return ISHARA_ERR_COMM. */
ISHARA_API ishara_status_t ishara_comm_unique_id(void* out_id128);
ISHARA_API ishara_status_t ishara_model_comm_init(ishara_model_t* m, const void* id128, int32_t rank, int32_t world);
ISHARA_API ishara_status_t ishara_model_comm_destroy(ishara_model_t* m);
/* NCCL_VERSION_CODE of the bound library, 0 if NCCL is not available */
ISHARA_API int32_t ishara_nccl_version(void);
/* The bucket plan as pure host logic (no GPU needed; the CPU tests pin it): hi[k] = end offset of the highest gradient
 * the backward of module k writes, modules run their backward in reverse order; lo_out/up_out[k] = the range of the
 * flat gradient buffer that becomes final - and is all-reduced - right after module k's backward (empty if lo == up).
 * The ranges tile [0, n_train) exactly once; ranges smaller than min_elems are merged into the next one. */
ISHARA_API ishara_status_t ishara_comm_bucket_plan(const int64_t* hi, int32_t n, int64_t n_train, int64_t min_elems,
                                                    int64_t* lo_out, int64_t* up_out);

/* ---- landmark preprocessing -----------------------------------------------
 * The step in front of the model call: TFLiteModel.__call__ c13:9-15 = pre_process00 (c3:57-101) + pre_process1
 * (c3:103-115). frames_dev fp32 [total_frames, 276] holds the batch's sequences back to back in the reference's SEL_COLS
 * order (c1:22-26); offsets_dev int32 [batch+1] are frame offsets (a sequence may be empty: it becomes one zero frame,
 * c13:11); mean_dev / std_dev fp32 [276] are the per-group statistics laid out in OUTPUT column order (lip, rhand,
 * lhand, rpose, lpose; landmark-major, xyz-minor); out_dev fp32 [batch, frame_len, 276] is the model input.
 * filter_frames != 0 applies the hand-frame filter (inference path); 0 = training path (pre_process1 only). */
ISHARA_API ishara_status_t ishara_preprocess(const float* frames_dev, const int32_t* offsets_dev, int32_t batch, int32_t max_frames,
                                             const float* mean_dev, const float* std_dev, int32_t frame_len, int32_t filter_frames,
                                             float* out_dev, void* stream);

/* ---- scorer -----------------------------------------------------------------
 * Levenshtein distances of n (prediction, target) byte-string pairs, the `distance` call of the evaluation loop
 * c18:1-15 (score = (len(target) - distance) / len(target)). Host function; strings need not be NUL-terminated. */
ISHARA_API ishara_status_t ishara_edit_distances(const char* const* preds, const int32_t* pred_lens, const char* const* targets,
                                                 const int32_t* target_lens, int32_t n, int32_t* out_distances);

/* ---- operator-level entry points (device pointers; bf16 tensors are raw uint16 bit patterns) --- */
typedef struct {
  const void* a;        /* bf16 [M,K], row pitch lda */
  const void* wt;       /* bf16 [N,K] (weight transposed, K contiguous) */
  void* out0;           /* bf16 or fp32 [M,Nout] */
  void* out1;           /* bf16 [M,Nout] or NULL (LayerNorm'd copy) */
  const float* bias;    /* [N] or NULL */
  const float* gate;    /* [M/rows_per_seq, N] or NULL */
  const float* rowtab;  /* [rows_per_seq, N] or NULL */
  const void* resid;    /* bf16 [M,Nout] or NULL */
  const float* ln0_g; const float* ln0_b; float ln0_eps;
  const float* ln1_g; const float* ln1_b; float ln1_eps;
  int32_t M, N, K, lda;
  int32_t nout;         /* logical output columns (<= N, or <= N/2 for GLU); 0 = all */
  int32_t rows_per_seq;
  int32_t act;          /* 0 none, 1 swish, 2 relu, 3 GLU (Nout = N/2) */
  int32_t block_n;      /* 64, 128, 256 */
  int32_t out_f32;      /* 0 | 1 */
  int32_t row_mode;     /* 1: full-row epilogue (residual + LayerNorm fusion), needs N == block_n */
} ishara_gemm_args_t;
ISHARA_API ishara_status_t ishara_op_gemm(const ishara_gemm_args_t* args, void* stream);

/* depthwise temporal conv; post: 0 none, 1 swish, 2 ECA (needs eca_w[5]); see csrc/kernels.h DwConvArgs */
ISHARA_API ishara_status_t ishara_op_dwconv(const void* in_bf16, void* out_bf16, const float* w, const float* bias,
                                 const float* eca_w, float* colsum, int32_t B, int32_t T, int32_t C, int32_t k,
                                 int32_t pad_left, int32_t post, void* stream);
/* attention core on per-head-interleaved qkv bf16 [B*T, 3*H*dh] -> out bf16 [B*T, H*dh] */
ISHARA_API ishara_status_t ishara_op_attention(const void* qkv_bf16, void* out_bf16, const uint8_t* key_mask, int32_t B, int32_t T,
                                    int32_t H, int32_t dh, float scale, void* stream);
/* Transformer-XL relative-position attention core (squeezeformer/attention.py:25-110): qkv as above (biases already
 * added), pos bf16 [2T-1, H*dh] = pos_proj(pos_emb), u_bias / v_bias fp32 [H*dh]; the reference's _relative_shift is
 * folded into index arithmetic. key_mask uint8 [B,T], 1 = keep (the reference's mask is True = masked). */
ISHARA_API ishara_status_t ishara_op_relpos_attention(const void* qkv_bf16, const void* pos_bf16, const float* u_bias,
                                           const float* v_bias, void* out_bf16, const uint8_t* key_mask, int32_t B,
                                           int32_t T, int32_t H, int32_t dh, float scale, void* stream);
/* TimeReductionLayer (squeezeformer/convolution.py:241-269): Conv2d(1->1,3,stride 2)+bias+Swish over the [T,D] plane;
 * x bf16 [B,T,D] -> out bf16 [B,(T-3)/2+1, ldo] (first (D-3)/2+1 columns valid, rest zero). w9 = 3x3 weight (host). */
ISHARA_API ishara_status_t ishara_op_time_reduce(const void* x_bf16, void* out_bf16, const float* w9_host, float bias, int32_t B,
                                      int32_t T, int32_t D, int32_t ldo, void* stream);
/* recover step (squeezeformer/modules.py:137-142 + encoder.py:157-162): out[b,t,:] = y[b,t/2,:] + rec[b,t,:], t < 2*T2 */
ISHARA_API ishara_status_t ishara_op_upsample_add(const void* y_bf16, const void* rec_bf16, void* out_bf16, int32_t B, int32_t T2,
                                       int32_t T, int32_t D, void* stream);
/* DepthwiseConv2dSubsampling (squeezeformer/convolution.py:39-73): x fp32 [B,T,F] -> bf16 [B,T4,ldo], column c*F4+f.
 * w1/b1: Conv2d(1->C,3,s2) [C,9]/[C]; w2/b2: depthwise Conv2d(C,3,s2) [C,9]/[C]; all device fp32. */
ISHARA_API ishara_status_t ishara_op_conv2d_subsample(const float* x, void* out_bf16, const float* w1, const float* b1,
                                           const float* w2, const float* b2, int32_t B, int32_t T, int32_t F, int32_t C,
                                           int32_t ldo, void* stream);
ISHARA_API ishara_status_t ishara_op_layernorm(const void* x_bf16, void* out_bf16, const float* gamma, const float* beta,
                                    float eps, int64_t M, int32_t D, void* stream);
ISHARA_API ishara_status_t ishara_op_cast_pad(const float* x, void* out_bf16, int64_t M, int32_t F, int32_t Fpad, void* stream);


/* num_to_char_fn + "".join (c8:1-2, c8:18) for a whole batch on the host: ids_host int32 [B, frames] / lens_host [B] as
 * returned by ishara_greedy_decode; table[id] is the character of id (ids outside [0, table_len) map to "" like
 * num_to_char.get(x, "")). Writes the concatenated text to out (capacity >= B*frames) and B+1 offsets. */
ISHARA_API ishara_status_t ishara_ids_to_text(const int32_t* ids_host, const int32_t* lens_host, int32_t batch, int32_t frames,
                                   const char* table, int32_t table_len, char* out, int64_t* offsets);

/* ---- raw buffers (so a host without PyTorch can own device / pinned memory; DLPack producers in
 * ishara_b200/_dlpack.py sit on top of these) ------------------------------------------------- */
ISHARA_API ishara_status_t ishara_device_malloc(int32_t device, int64_t bytes, void** out_dev);
ISHARA_API ishara_status_t ishara_device_free(int32_t device, void* dev_ptr);
ISHARA_API ishara_status_t ishara_host_malloc_pinned(int64_t bytes, void** out_host);
ISHARA_API ishara_status_t ishara_host_free_pinned(void* host_ptr);
/* kind: 1 host->device, 2 device->host, 3 device->device. Asynchronous on `stream`. */
ISHARA_API ishara_status_t ishara_memcpy_async(void* dst, const void* src, int64_t bytes, int32_t kind, void* stream);
ISHARA_API ishara_status_t ishara_stream_synchronize(int32_t device, void* stream);
/* the handle's own stream (used by the *_host entry points) */
ISHARA_API void* ishara_model_stream(ishara_model_t* m);

/* ---- measurement hooks (bench.py) ---------------------------------------------------------------
 * With profiling on, every forward records one CUDA event per launch on the launch stream; entries are the
 * launches of the LAST forward in order (entry 0 = the input cast). ms = device time event-to-event; flops /
 * bytes = the algorithmic cost of that launch. kind: "gemm" | "dwconv" | "attention" | ... */
ISHARA_API ishara_status_t ishara_model_set_profile(ishara_model_t* m, int32_t on);
ISHARA_API int32_t ishara_model_profile_count(const ishara_model_t* m);
ISHARA_API ishara_status_t ishara_model_profile_entry(ishara_model_t* m, int32_t index, const char** label,
                                           const char** kind, float* ms, double* flops, double* bytes);
/* kernels launched by this library in this process so far */
ISHARA_API uint64_t ishara_launch_count(void);

/* ---- debugging aid: copy an internal activation of the last forward (bf16 -> fp32) to the host ---
 * Enable with ishara_model_set_debug(m, 1) before the forward. Names: "stem", a Conv1DBlock name
 * ("convsqueeze_0_1", ...), "squeezeformer_<i>", "conformer_<i>": the residual stream after that module. */
ISHARA_API ishara_status_t ishara_model_set_debug(ishara_model_t* m, int32_t on);
ISHARA_API ishara_status_t ishara_model_debug_fetch(ishara_model_t* m, const char* name, float* host_out, int64_t numel);

#ifdef __cplusplus
}
#endif
#endif /* ISHARA_B200_H_ */
