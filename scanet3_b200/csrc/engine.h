// Internal engine structures (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/scanet_b200.h"
#include "kernels.cuh"

namespace sb {

constexpr int64_t kArenaTailFloats = 1024;  // flags + scalars of the device-synchronised averaging (group.cu)

struct Shape {
  int rank = 0;
  int64_t d[5] = {0, 0, 0, 0, 0};
  int64_t numel() const {
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) n *= d[i];
    return n;
  }
};

struct Param {
  std::string path;
  int rank = 0;
  int64_t dims[4] = {0, 0, 0, 0};
  int role = 0;  // sb_param_role
  int agg = 0;   // sb_aggregation
  int64_t numel = 0;
  int64_t padded = 0;  // numel rounded up to 32 floats (128 B: vector loads + TMA alignment)
  int64_t offset = 0;  // floats from the arena base
  sb_initializer init{};
  int layer = -1;
  int weight_param = -1;  // role OPT: index of the weight this state belongs to
  int slot = -1;          // role OPT: state slot 0..2
  int reg_kind = 0;       // sb_regularization of a trainable weight (models/Regularization.scala)
  float reg_lambda = 0.f;
};

struct LayerPlan {
  sb_layer_desc d{};
  Shape in, out;        // incl. batch
  int buf_in = 0;       // activation buffer ids
  int buf_out = 0;
  bool inplace = false;
  bool need_dx = false; // some trainable parameter lives in an earlier layer
  int p_w = -1;         // param indices (kind specific)
  int p_aux[16];        // BN: beta,gamma,mm,mv ; LSTM: per gate kernel,recurrent,bias
  ConvGeom cg{};
  PoolGeom pg{};
  // BatchNorm
  bool bn_fused = false;
  float* bn_save_mean = nullptr;
  float* bn_save_var = nullptr;
  // fusion / fast-path decisions
  bool fused_into_prev = false;   // this layer's forward+backward is executed by the previous op
  bool skip_fwd = false;          // forward only is executed by the previous op's epilogue (ReLU after a tensor-core conv)
  bool skip_bwd = false;          // backward is folded into a neighbour (ReLU into the max-pool backward, head layers)
  bool fuse_relu_fwd = false;     // small-K conv: apply the following in-place ReLU in the store
  bool pool_relu_bwd = false;     // max-pool backward also applies the mask of the in-place ReLU that produced its input
  float fused_thr = 0.f;
  int fuse_conv_wgrad = -1;       // Pool2D: its backward also computes the filter gradient of this (small-K) conv layer
  int fuse_pool_fwd = -1;         // Conv2D (small K): its forward kernel also produces this Pool2D layer's output
  unsigned char* pool_map = nullptr;  // Pool2D (2x2 stride 1) whose forward is fused into the producing conv: one winner byte per
                                      // output and channel (first-max index | ReLU bit); the backward routes dy with it and the
                                      // conv activations are neither written nor read
  int bn_pool = -1;               // BatchNorm: forward and backward also do this Pool2D(2x2, stride 2) layer (bn_pool.cu); its pool_map holds the winners
  bool in_bn = false;             // Pool2D: executed by the BatchNorm in front of it
  bool bn_relu_bwd = false;       // ... and the backward of the in-place ReLU in front of the BatchNorm (fused_thr)
  int fuse_head_dgrad = -1;       // Pool2D: its backward also forms the input gradient of this classifier-head Dense layer
  bool dgrad_in_pool = false;     // Dense (head): dX is produced by the pool's backward kernel; only dW is computed here
  bool wgrad_in_pool = false;     // Conv2D: the filter gradient is produced by the following pool's backward kernel
  bool fuse_act_bwd = false;      // Bias: also applies the backward of the in-place activation that follows (one pass)
  bool head_skip = false;         // Bias / Softmax of the fused classifier head: executed by the head kernel when training
  bool fast_smallk = false;       // Conv2D with KH*KW*Cin <= 32: register-tiled direct kernel instead of the implicit GEMM
  bool fast_pool = false;         // 128-bit max-pool kernels (C % 4 == 0)
  bool fast_skinny = false;       // Dense with <= 16 outputs and a long reduction: split-K skinny kernels
  int tc_kind = 0;                // 0 none, otherwise a tcgen05 kernel variant
  void* tc_plan = nullptr;        // opaque plan of the tensor-core path (TMA descriptors, buffers)
  void* head_tc = nullptr;        // opaque plan of the tensor-core classifier-head forward (tc.cu)
  void* head_bwd_tc = nullptr;    // ... and of its backward (dW and dX in one kernel)
  void* lstm = nullptr;           // opaque LSTM workspace
};

struct ProfEntry {
  std::string name;
  double total_ms = 0;
  long long launches = 0;
  double bytes = 0;
  double flops = 0;
};
struct ProfPending {
  int entry;
  cudaEvent_t a, b;
};

}  // namespace sb

struct sb_engine {
  int device = -1;
  bool plan_only = false;
  std::vector<sb_layer_desc> layers_desc;
  sb_model_desc model{};
  sb_train_desc train{};
  sb::Shape in_shape, label_shape, out_shape;  // incl. batch
  int64_t x_per_row = 0, y_per_row = 0;

  std::vector<sb::Param> params;
  std::unordered_map<std::string, int> by_path;
  std::vector<sb::LayerPlan> L;
  int64_t nW = 0;      // floats in the trainable-weight region (padded)
  int64_t nS = 0;      // floats in the state region
  int n_slots = 0;     // optimizer state slots
  int64_t arena_floats = 0;

  float* arena = nullptr;
  float* grads = nullptr;
  std::vector<float*> acts, dacts;
  std::vector<int64_t> act_numel;
  float* by = nullptr;  // label staging
  float* ds_x = nullptr;
  float* ds_y = nullptr;
  int64_t ds_rows = 0;
  sb::StepState* st = nullptr;
  sb::StepState* st_host = nullptr;  // pinned mirror
  float* ws = nullptr;
  float* ws_side = nullptr;       // workspace of kernels that run on the backward side branch next to users of `ws`
  size_t ws_side_floats = 0;
  double* est_acc = nullptr;      // 4 running sums of sb_estimate
  float* head_scratch = nullptr;  // per-CTA partials + completion ticket of the fused classifier-head kernel
  size_t ws_floats = 0;
  cudaStream_t stream = nullptr;
  // host-fed pipeline (sb_train_step_async): the H2D copy of batch t+1 runs on `cstream` into the other staging buffer while
  // batch t computes; the step's prologue gathers from its staging buffer
  // backward fork: a conv layer's filter gradient (tensor-core kernel + reduce) runs on `bstream` next to the input-gradient
  // chain (dgrad -> pool backward ...), joined before the penalties / optimizer; inside a captured graph this is a parallel branch
  cudaStream_t bstream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_mid = nullptr;
  // Early optimizer (SB_EARLY_OPT=0 turns it off): the side branch of layer `early_opt_layer` also updates every parameter from arena
  // offset `early_opt_off` on (all their gradients exist by then), once that layer's dgrad has read the weights; the tail of the step
  // then only updates the first layer's parameters [0, early_opt_off)
  int early_opt_layer = -1;
  int64_t early_opt_off = 0;
  bool early_opt_now = false, early_opt_done = false;
  bool bwd_forked = false;
  sb::ReduceBatch pending_reduce;  // filter-gradient partial sums of this backward pass, reduced in one launch before the optimizer
  bool pdl_early = false;   // the shared kernels (prologue, optimizer, head loss, partial reduce) trigger their dependents early
  cudaStream_t cstream = nullptr;
  float* stage_x[2] = {nullptr, nullptr};
  float* stage_y[2] = {nullptr, nullptr};
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  cudaGraphExec_t g_staged[2] = {nullptr, nullptr};
  long long nodes_staged[2] = {0, 0};
  // sb_train_batches_async: each staging buffer holds `stage_batches` consecutive host batches; one graph runs them all
  int stage_batches = 1;
  cudaGraphExec_t g_group[2] = {nullptr, nullptr};
  long long nodes_group[2] = {0, 0};
  float* loss_hist[2] = {nullptr, nullptr};  // per staging buffer: pre-update losses of the group's steps
  float* pending_loss_host = nullptr;        // read-back of the latest group, queued behind the NEXT group's H2D copy
  int pending_loss_par = 0;
  int pending_loss_n = 0;
  cudaGraphExec_t g_cursor[2] = {nullptr, nullptr};  // one host-fed step taking its batch through the device-side cursor (epoch tails)
  long long nodes_cursor[2] = {0, 0};
  int stage_parity = 0;
  cudaGraphExec_t g_resident = nullptr, g_hostfed = nullptr, g_loss = nullptr;
  cudaGraphExec_t g_resident_multi = nullptr;  // kStepsPerGraph resident steps in one graph (fewer graph boundaries per epoch)
  long long nodes_resident_multi = 0;
  long long nodes_resident = 0, nodes_hostfed = 0, nodes_loss = 0;
  long long launches = 0;
  long long host_iter_mirror = -1;  // device st->iter as far as the host knows (-1 unknown)
  bool params_ready = false, opt_ready = false;
  bool device_sync_used = false;  // took part in sb_group_average_device: host sync points also read StepState.err
  bool softmax_cce_fused = false;
  bool has_reg = false;      // some weight carries an L1/L2 penalty
  bool head_fused = false;   // last layers Dense >> Bias >> Softmax with CCE: one fused bias+softmax+loss+dz+dbias kernel
  int head_chunks = 0;       // split-K partials of the head's Dense waiting in `ws` for the head kernel (0: none)
  bool fast = true;          // vectorised / fused bandwidth kernels (SB_DISABLE_FAST=1 turns them off for debugging)

  bool profiling = false;
  std::vector<sb::ProfEntry> prof;
  std::vector<sb::ProfPending> prof_pending;
  std::unordered_map<std::string, int> prof_by_name;
};
