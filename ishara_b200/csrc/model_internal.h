// ishara_b200 — private definition of the model handle, shared by model.cu (inference program), train.cu (training
// step) and nothing else. Not part of the C ABI.
#pragma once
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ishara_b200.h"
#include "kernels.h"

namespace ishara {

struct Param {
  std::string name;
  std::vector<int64_t> shape;
  std::vector<float> data;
  bool set = false;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

struct LnRef {
  const float* g = nullptr;
  const float* b = nullptr;
  float eps = 0.f;
};

enum OpKind { OP_GEMM, OP_DW, OP_ATTN, OP_SEGATE, OP_LN, OP_TAP, OP_FFN, OP_C1F, OP_C1B };
struct Op {
  OpKind kind;
  GemmPlan gemm;
  FfnPlan ffn;
  Conv1dFrontPlan c1f;
  Conv1dBlockPlan c1b;
  DwConvArgs dw;
  AttnArgs at;
  SeGateArgs se;
  // OP_LN
  const bf16* ln_in = nullptr;
  bf16* ln_out = nullptr;
  LnRef ln;
  int64_t ln_rows = 0;  // rows of this op's (lane's) stream
  int tap = -1;
  int lane = 0;  // sub-batch this op belongs to (model.cu: lanes)
  const char* label = "";
  // algorithmic cost of one launch: useful flops and compulsory HBM bytes (operands read once,
  // results written once; weights counted once per launch)
  double flops = 0.0, bytes = 0.0;
};

inline uint16_t f2bf(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);                      // round to nearest even
  return static_cast<uint16_t>(u >> 16);
}

}  // namespace ishara

namespace ishara { struct TrainState; }
using ishara::bf16;
constexpr int kMaxLanes = 4;

struct GraphKey {
  int batch;
  float* logits;
  const float* x;
  bool operator<(const GraphKey& o) const {
    if (batch != o.batch) return batch < o.batch;
    if (logits != o.logits) return logits < o.logits;
    return x < o.x;
  }
  bool operator!=(const GraphKey& o) const { return batch != o.batch || logits != o.logits || x != o.x; }
};
struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  int launches = 0;
};

struct ishara_model {
  ishara_config_t cfg;
  int device = 0;
  int num_sms = 148;
  std::vector<ishara::Param> params;
  std::unordered_map<std::string, int> index;
  bool finalized = false;
  bool debug_taps = false;

  std::vector<void*> wallocs;  // packed weights
  std::unordered_map<std::string, void*> packed;

  // workspace
  int cap_batch = 0;
  std::vector<void*> wsallocs;
  bf16 *XIN = nullptr, *S = nullptr, *XN = nullptr, *H1 = nullptr, *H2 = nullptr, *O = nullptr, *HEAD = nullptr;
  float *colsum = nullptr, *gate = nullptr;
  // mask_mode="propagated" (Keras Masking reaching ECA / SE / Softmax, c5:8-9,109-112,129-130): written by mask_prep_launch
  int mask_mode = 0;                   // 0 = dropped (the reference as executed), 1 = propagated
  uint8_t* mask_dev = nullptr;         // [cap_batch * frames] 1 = frame carries data
  uint16_t* wbits_dev = nullptr;       // [cap_batch * frames] validity of the 16 frames starting at each frame
  int32_t* valid_dev = nullptr;        // [cap_batch] valid frames per sequence
  uint8_t* user_mask_dev = nullptr;    // [cap_batch * frames] copy of the caller's mask (forward_masked)
  int* use_user_mask_dev = nullptr;    // device flag read by mask_prep: 1 = take user_mask_dev, 0 = derive the mask from x
  float* logits_own = nullptr;
  int32_t *ids_dev = nullptr, *lens_dev = nullptr, *labels_dev = nullptr;
  float* nll_dev = nullptr;
  int labels_cap = 0;
  float* x_dev = nullptr;
  std::vector<bf16*> taps;
  std::vector<std::string> tap_names;

  std::vector<ishara::Op> program;
  bool profile = false;               // record one CUDA event per op during forward
  std::vector<cudaEvent_t> events;    // [0] before cast_pad, [1] after it, [2+i] after program[i]
  int program_batch = 0;
  float* program_logits = nullptr;
  // programs built for other (batch, logits pointer) pairs: the chunked host path alternates between a few of them
  std::map<std::pair<int, float*>, std::vector<ishara::Op>> program_cache;
  std::map<GraphKey, GraphEntry> graphs;  // captured forwards; dropped whenever programs are rebuilt
  bool graphs_broken = false;
  cudaStream_t stream = nullptr;
  cudaStream_t lane_stream[kMaxLanes - 1] = {nullptr};  // lanes 1.. of a forward (lane 0 runs on the caller's stream)
  cudaEvent_t lane_fork = nullptr, lane_join[kMaxLanes - 1] = {nullptr};
  cudaStream_t copy_stream = nullptr;     // H2D of the chunked host path
  cudaEvent_t copy_done[8] = {nullptr};
  // pipelined host inference (model_infer_submit / _collect): two input slots so the H2D copy of batch i+1 runs under
  // the kernels of batch i
  struct PipeSlot {
    float* x = nullptr;
    int32_t* labels = nullptr;
    size_t x_cap = 0, labels_cap = 0;
    cudaEvent_t h2d_done = nullptr, x_free = nullptr, done = nullptr;
  } pipe[2];
  uint64_t pipe_submitted = 0, pipe_collected = 0;
  // data-parallel exchange (comm.cu): NCCL communicator bound at run time, its own stream, two events
  void* comm = nullptr;                   // ncclComm_t
  int comm_rank = 0, comm_world = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t comm_ready = nullptr, comm_done = nullptr;
  ishara::TrainState* train = nullptr;    // training step state (train.cu); null until the first train call
  bool host_params_stale = false;         // device master weights are newer than params[].data (after a train step)

  int fpad() const { return (cfg.features + 63) / 64 * 64; }
  int vpad() const { return (cfg.num_classes + 63) / 64 * 64; }
};

