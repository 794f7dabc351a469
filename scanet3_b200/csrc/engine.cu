// scanet-b200 engine: model descriptor -> plan (flat arena, activation buffers, op schedule) ->
// eager or CUDA-graph execution of the reference's train step.  Reference lines cited are under
// /root/reference/src/main/scala/scanet/.
#include "engine.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>

#include "common.cuh"
#include "tc.h"

using namespace sb;

// -------------------------------------------------------------------------------------------------
// error plumbing
// -------------------------------------------------------------------------------------------------
static thread_local std::string tl_error;
namespace sb {
struct ApiError : std::runtime_error {
  int code;
  ApiError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_error(const std::string& m) { tl_error = m; }
bool fork_enabled() {  // SB_BWD_FORK=0: the conv filter gradient stays in the main stream
  static const bool on = [] {
    const char* v = getenv("SB_BWD_FORK");
    return !(v && v[0] == '0');
  }();
  return on;
}
static thread_local int tl_prewait_ok = 0;
int prewait_ok() { return tl_prewait_ok; }
void set_prewait_ok(int ok) { tl_prewait_ok = ok; }
static thread_local int tl_pdl_early = 0;
int pdl_early_hint() { return tl_pdl_early; }
void set_pdl_early_hint(int on) { tl_pdl_early = on; }
bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("SB_PDL");
    return !(v && v[0] == '0');
  }();
  return on;
}
}  // namespace sb
#define SB_FAIL(code, msg) throw ::sb::ApiError((code), (msg))
#define SB_API_BEGIN try {
#define SB_API_END                      \
  }                                     \
  catch (const ::sb::ApiError& e) {     \
    tl_error = e.what();                \
    return e.code;                      \
  }                                     \
  catch (const ::sb::CudaError& e) {    \
    tl_error = e.what();                \
    return SB_ERR_CUDA;                 \
  }                                     \
  catch (const std::exception& e) {     \
    tl_error = e.what();                \
    return SB_ERR_INVALID_ARG;          \
  }                                     \
  return SB_OK;

extern "C" const char* sb_last_error(void) { return tl_error.c_str(); }
// internal hooks of wire.cpp (host-only translation unit above the public ABI); not declared in the public header
extern "C" void sb_internal_set_last_error(const char* msg) { tl_error = msg ? msg : ""; }
extern "C" int sb_abi_version(void) { return SB_ABI_VERSION; }
extern "C" int sb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

namespace sb {
void trace_set_kernels_simt(unsigned long long*); void trace_set_kernels_fast(unsigned long long*); void trace_set_tc(unsigned long long*);
void trace_set_tc_gemm(unsigned long long*); void trace_set_lstm(unsigned long long*); void trace_set_rnn(unsigned long long*);
void trace_set_group(unsigned long long*); void trace_set_lstm_persist(unsigned long long*); void trace_set_bn_pool(unsigned long long*);
}  // namespace sb
// SB_TRACE timeline (common.cuh): sb_trace_begin points every translation unit's kernels at a fresh device buffer,
// sb_trace_read copies the stamps back (count first) and switches tracing off again.
static unsigned long long* g_trace_dev = nullptr;
static void trace_point_all(unsigned long long* p) {
  sb::trace_set_kernels_simt(p); sb::trace_set_kernels_fast(p); sb::trace_set_tc(p); sb::trace_set_tc_gemm(p);
  sb::trace_set_lstm(p); sb::trace_set_rnn(p); sb::trace_set_group(p); sb::trace_set_lstm_persist(p); sb::trace_set_bn_pool(p);
}
extern "C" int sb_internal_trace_begin(void) {
  SB_API_BEGIN
  if (!g_trace_dev) SB_CUDA(cudaMalloc((void**)&g_trace_dev, (size_t)sb::kTraceSlots * 8));
  SB_CUDA(cudaMemset(g_trace_dev, 0, (size_t)sb::kTraceSlots * 8));
  trace_point_all(g_trace_dev);
  SB_CUDA(cudaDeviceSynchronize());
  SB_API_END
}
extern "C" int sb_internal_trace_read(unsigned long long* out, int cap) {
  SB_API_BEGIN
  if (!g_trace_dev || !out || cap < 1) SB_FAIL(SB_ERR_INVALID_ARG, "no trace in progress");
  SB_CUDA(cudaDeviceSynchronize());
  SB_CUDA(cudaMemcpy(out, g_trace_dev, (size_t)std::min(cap, sb::kTraceSlots) * 8, cudaMemcpyDeviceToHost));
  trace_point_all(nullptr);
  SB_API_END
}

static void need_device(const sb_engine* e) {
  if (e == nullptr) SB_FAIL(SB_ERR_INVALID_ARG, "null engine");
  if (e->plan_only)
    SB_FAIL(SB_ERR_CUDA, "plan-only engine (device < 0): no CUDA device bound; this library has no CPU fallback");
  sb::set_pdl_early_hint(e->pdl_early ? 1 : 0);  // every compute entry point passes here: the launches that follow are this model's
  sb::set_prewait_ok(0);                          // what ran last on the stream is unknown at an API boundary
}
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    SB_CUDA(cudaGetDevice(&prev));
    if (prev != dev) SB_CUDA(cudaSetDevice(dev));
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// -------------------------------------------------------------------------------------------------
// plan
// -------------------------------------------------------------------------------------------------
static int64_t pad32(int64_t n) { return (n + 31) / 32 * 32; }

static void conv_pad(int padding, int in, int k, int s, int* out, int* before) {
  // TF padding arithmetic (math/nn/AllKernels.scala:39-65)
  if (padding == SB_PAD_VALID) {
    *out = (in - k + 1 + s - 1) / s;
    *before = 0;
  } else {
    *out = (in + s - 1) / s;
    int total = std::max((*out - 1) * s + k - in, 0);
    *before = total / 2;
  }
}

static int add_param(sb_engine* e, const std::string& path, std::initializer_list<int64_t> dims, int role, int agg,
                     const sb_initializer& init, int layer) {
  Param p;
  p.path = path;
  p.rank = (int)dims.size();
  int i = 0;
  p.numel = 1;
  for (int64_t d : dims) {
    p.dims[i++] = d;
    p.numel *= d;
  }
  p.padded = pad32(p.numel);
  p.role = role;
  p.agg = agg;
  p.init = init;
  p.layer = layer;
  if (e->by_path.count(path)) SB_FAIL(SB_ERR_INVALID_ARG, "duplicate param path " + path);
  e->by_path[path] = (int)e->params.size();
  e->params.push_back(p);
  return (int)e->params.size() - 1;
}

static void set_reg(sb_engine* e, int pi, const sb_layer_desc& d) {
  if (d.reg_kind == SB_REG_ZERO || d.reg_lambda == 0.f) return;
  if (d.reg_kind != SB_REG_L1 && d.reg_kind != SB_REG_L2) SB_FAIL(SB_ERR_INVALID_ARG, "unknown regularization kind");
  e->params[pi].reg_kind = d.reg_kind;
  e->params[pi].reg_lambda = d.reg_lambda;
  e->has_reg = true;
}

static const char* kGateNames[4] = {"forget", "input", "gate", "output"};

static int opt_slots(int kind) {
  switch (kind) {
    case SB_OPT_SGD: case SB_OPT_ADAGRAD: case SB_OPT_RMSPROP: return 1;
    case SB_OPT_ADAM: case SB_OPT_ADADELTA: case SB_OPT_ADAMAX: case SB_OPT_NADAM: return 2;
    case SB_OPT_AMSGRAD: return 3;
  }
  SB_FAIL(SB_ERR_INVALID_ARG, "unknown optimizer kind");
}
static const char* opt_slot_name(int kind, int slot) {
  // state paths `<weightPath>/<stateName>` (Optimizer.scala:118-121,187)
  static const char* names[8][3] = {{"velocity", "", ""},        {"momentum", "velocity", ""},
                                    {"grad_acc", "", ""},        {"avg_grad", "", ""},
                                    {"avg_grad", "avg_delta", ""}, {"momentum_1", "momentum_2", ""},
                                    {"momentum", "velocity", ""}, {"momentum", "velocity", "corrected_velocity"}};
  return names[kind][slot];
}

static void build_plan(sb_engine* e) {
  const sb_model_desc& m = e->model;
  const sb_train_desc& t = e->train;
  if (m.n_layers <= 0) SB_FAIL(SB_ERR_INVALID_ARG, "requirement failed: at least 1 layer is required");
  if (t.batch <= 0) SB_FAIL(SB_ERR_INVALID_ARG, "batch must be positive");
  if (m.input_rank < 0 || m.input_rank > 3 || m.label_rank < 0 || m.label_rank > 3)
    SB_FAIL(SB_ERR_INVALID_ARG, "input/label rank must be in [0,3]");
  if (t.loss < SB_LOSS_MSE || t.loss > SB_LOSS_CCE) SB_FAIL(SB_ERR_INVALID_ARG, "unknown loss");
  e->n_slots = opt_slots(t.optimizer);

  Shape cur;
  cur.rank = m.input_rank + 1;
  cur.d[0] = t.batch;
  for (int i = 0; i < m.input_rank; ++i) cur.d[i + 1] = m.input_dims[i];
  e->in_shape = cur;
  e->label_shape.rank = m.label_rank + 1;
  e->label_shape.d[0] = t.batch;
  for (int i = 0; i < m.label_rank; ++i) e->label_shape.d[i + 1] = m.label_dims[i];
  e->x_per_row = cur.numel() / t.batch;
  e->y_per_row = e->label_shape.numel() / t.batch;

  int cur_buf = 0;
  std::vector<int64_t> buf_numel{cur.numel()};
  std::vector<char> output_needed{0};
  bool any_trainable_before = false;
  e->L.resize(m.n_layers);
  for (int li = 0; li < m.n_layers; ++li) {
    LayerPlan& L = e->L[li];
    L.d = m.layers[li];
    for (int k = 0; k < 16; ++k) L.p_aux[k] = -1;
    L.in = cur;
    L.buf_in = cur_buf;
    L.need_dx = any_trainable_before;
    const sb_layer_desc& d = L.d;
    if (d.reg_kind != SB_REG_ZERO && d.reg_lambda != 0.f && d.kind != SB_LAYER_DENSE && d.kind != SB_LAYER_BIAS)
      SB_FAIL(SB_ERR_UNSUPPORTED, "regularisation is accelerated for Dense kernels and Bias layers only");
    const std::string pre = std::to_string(li) + "/";
    Shape out = cur;
    bool new_buf = true;
    switch (d.kind) {
      case SB_LAYER_DENSE: {
        if (cur.rank != 2) SB_FAIL(SB_ERR_INVALID_ARG, "Dense requires input Shape(batch, features)");
        if (d.units <= 0) SB_FAIL(SB_ERR_INVALID_ARG, "Dense outputs must be positive");
        if (!d.trainable) SB_FAIL(SB_ERR_UNSUPPORTED, "frozen layers are outside the accelerated path");
        L.p_w = add_param(e, pre + "weights", {cur.d[1], (int64_t)d.units}, SB_ROLE_WEIGHT, SB_AGG_AVG, d.init[0], li);
        set_reg(e, L.p_w, d);
        out.d[1] = d.units;
        any_trainable_before = true;
        break;
      }
      case SB_LAYER_BIAS: {
        if (cur.d[cur.rank - 1] != d.units) SB_FAIL(SB_ERR_INVALID_ARG, "Bias features must equal the last input dimension");
        if (!d.trainable) SB_FAIL(SB_ERR_UNSUPPORTED, "frozen layers are outside the accelerated path");
        L.p_w = add_param(e, pre + "weights", {(int64_t)d.units}, SB_ROLE_WEIGHT, SB_AGG_AVG, d.init[2], li);
        set_reg(e, L.p_w, d);
        new_buf = (cur_buf == 0) || output_needed[cur_buf];
        any_trainable_before = true;
        break;
      }
      case SB_LAYER_ACTIVATE: {
        if (d.activation < SB_ACT_IDENTITY || d.activation > SB_ACT_RELU) SB_FAIL(SB_ERR_INVALID_ARG, "unknown activation");
        if (d.activation == SB_ACT_SOFTMAX && cur.rank != 2) SB_FAIL(SB_ERR_INVALID_ARG, "Softmax requires a rank-2 input");
        new_buf = (cur_buf == 0) || output_needed[cur_buf];
        break;
      }
      case SB_LAYER_FLATTEN: {
        if (cur.rank < 2) SB_FAIL(SB_ERR_INVALID_ARG, "requirement failed: rank should be >= 2");
        out.rank = 2;
        out.d[1] = cur.numel() / cur.d[0];
        new_buf = false;  // a view
        break;
      }
      case SB_LAYER_CONV2D: {
        if (cur.rank != 4) SB_FAIL(SB_ERR_INVALID_ARG, "Conv2D input should have a shape (NHWC)");
        if (!d.trainable) SB_FAIL(SB_ERR_UNSUPPORTED, "frozen layers are outside the accelerated path");
        if (d.units <= 0 || d.window_h <= 0 || d.window_w <= 0 || d.stride_h <= 0 || d.stride_w <= 0)
          SB_FAIL(SB_ERR_INVALID_ARG, "Conv2D filters/kernel/strides must be positive");
        if (d.padding != SB_PAD_VALID && d.padding != SB_PAD_SAME && d.padding != SB_PAD_EXPLICIT)
          SB_FAIL(SB_ERR_INVALID_ARG, "unknown padding kind");
        ConvGeom& g = L.cg;
        g.N = (int)cur.d[0]; g.H = (int)cur.d[1]; g.W = (int)cur.d[2]; g.Cin = (int)cur.d[3];
        g.Cout = d.units; g.KH = d.window_h; g.KW = d.window_w; g.SH = d.stride_h; g.SW = d.stride_w;
        if (d.padding == SB_PAD_EXPLICIT) {
          // Explicit(height = (top, bottom), width = (left, right)) (math/nn/AllKernels.scala:67-102).  The kernels need the
          // leading pads only: rows / columns past the input are zero whichever pad they belong to; bottom / right enter
          // through the output size.  TF floors where the reference's `size` ceils: equal exactly when the padding "fits".
          if (d.pad_top < 0 || d.pad_bottom < 0 || d.pad_left < 0 || d.pad_right < 0) SB_FAIL(SB_ERR_INVALID_ARG, "negative explicit padding");
          const int th = g.H + d.pad_top + d.pad_bottom - g.KH, tw = g.W + d.pad_left + d.pad_right - g.KW;
          if (th < 0 || tw < 0) SB_FAIL(SB_ERR_INVALID_ARG, "Conv2D output would be empty");
          if (th % g.SH || tw % g.SW)
            SB_FAIL(SB_ERR_INVALID_ARG, "explicit padding does not fit the strides (the reference's shape and TensorFlow's would differ)");
          g.OH = th / g.SH + 1; g.PT = d.pad_top;
          g.OW = tw / g.SW + 1; g.PL = d.pad_left;
        } else {
          conv_pad(d.padding, g.H, g.KH, g.SH, &g.OH, &g.PT);
          conv_pad(d.padding, g.W, g.KW, g.SW, &g.OW, &g.PL);
        }
        if (g.OH <= 0 || g.OW <= 0) SB_FAIL(SB_ERR_INVALID_ARG, "Conv2D output would be empty");
        L.p_w = add_param(e, pre + "weights", {(int64_t)g.KH, (int64_t)g.KW, (int64_t)g.Cin, (int64_t)g.Cout}, SB_ROLE_WEIGHT,
                          SB_AGG_AVG, d.init[0], li);
        out.d[1] = g.OH; out.d[2] = g.OW; out.d[3] = g.Cout;
        any_trainable_before = true;
        break;
      }
      case SB_LAYER_POOL2D: {
        if (cur.rank != 4) SB_FAIL(SB_ERR_INVALID_ARG, "Pool2D input should have a shape (NHWC)");
        if (d.padding != SB_PAD_VALID && d.padding != SB_PAD_SAME)
          SB_FAIL(SB_ERR_INVALID_ARG, "requirement failed: only [Same, Valid] paddings are supported");
        if (d.window_h <= 0 || d.window_w <= 0 || d.stride_h <= 0 || d.stride_w <= 0)
          SB_FAIL(SB_ERR_INVALID_ARG, "Pool2D window/strides must be positive");
        PoolGeom& g = L.pg;
        g.N = (int)cur.d[0]; g.H = (int)cur.d[1]; g.W = (int)cur.d[2]; g.C = (int)cur.d[3];
        g.KH = d.window_h; g.KW = d.window_w; g.SH = d.stride_h; g.SW = d.stride_w;
        conv_pad(d.padding, g.H, g.KH, g.SH, &g.OH, &g.PT);
        conv_pad(d.padding, g.W, g.KW, g.SW, &g.OW, &g.PL);
        if (g.OH <= 0 || g.OW <= 0) SB_FAIL(SB_ERR_INVALID_ARG, "Pool2D output would be empty");
        out.d[1] = g.OH; out.d[2] = g.OW;
        break;
      }
      case SB_LAYER_BATCHNORM: {
        if (cur.rank != 2 && cur.rank != 4) SB_FAIL(SB_ERR_UNSUPPORTED, "BatchNorm supports rank-2 and rank-4 (NHWC) inputs, axes=[-1]");
        if (!d.trainable) SB_FAIL(SB_ERR_UNSUPPORTED, "frozen layers are outside the accelerated path");
        int64_t C = cur.d[cur.rank - 1];
        L.bn_fused = (cur.rank == 4) && d.bn_fused != 0;
        if (d.bn_fused == 1 && cur.rank != 4) SB_FAIL(SB_ERR_INVALID_ARG, "Cannot use fused implementation with this input shape");
        // param shape = input shape with every non-axes dim set to 1 (layer/BatchNorm.scala:79-87)
        auto bn_param = [&](const char* name, int role, const sb_initializer& in) {
          if (cur.rank == 2) return add_param(e, pre + name, {1, C}, role, SB_AGG_AVG, in, li);
          return add_param(e, pre + name, {1, 1, 1, C}, role, SB_AGG_AVG, in, li);
        };
        L.p_aux[0] = bn_param("beta", SB_ROLE_WEIGHT, d.init[1]);
        L.p_aux[1] = bn_param("gamma", SB_ROLE_WEIGHT, d.init[0]);
        L.p_aux[2] = bn_param("moving_mean", SB_ROLE_STATE, d.init[2]);
        L.p_aux[3] = bn_param("moving_variance", SB_ROLE_STATE, d.init[3]);
        any_trainable_before = true;
        break;
      }
      case SB_LAYER_LSTM: {
        if (cur.rank != 3) SB_FAIL(SB_ERR_INVALID_ARG, "requirement failed: RNN requires input to have Shape(batch, time, features)");
        if (!d.trainable) SB_FAIL(SB_ERR_UNSUPPORTED, "frozen layers are outside the accelerated path");
        if (d.units <= 0) SB_FAIL(SB_ERR_INVALID_ARG, "requirement failed: RNN Cell should contain at least one unit");
        int64_t F = cur.d[2], U = d.units;
        for (int gi = 0; gi < 4; ++gi) {
          std::string gp = pre + kGateNames[gi] + "/";
          L.p_aux[gi * 3 + 0] = add_param(e, gp + "0/kernel_weights", {F, U}, SB_ROLE_WEIGHT, SB_AGG_AVG, d.init[0], li);
          L.p_aux[gi * 3 + 1] = add_param(e, gp + "0/recurrent_weights", {U, U}, SB_ROLE_WEIGHT, SB_AGG_AVG, d.init[1], li);
          if (d.use_bias)
            L.p_aux[gi * 3 + 2] = add_param(e, gp + "1/weights", {U}, SB_ROLE_WEIGHT, SB_AGG_AVG, gi == 0 ? d.init[3] : d.init[2], li);
        }
        if (d.stateful) {  // `if (stateful) state ++ weights` (layer/RNN.scala:32-35; LSTMCell states :253-254, 302-304)
          sb_initializer zeros{};
          zeros.kind = SB_INIT_ZEROS;
          L.p_aux[12] = add_param(e, pre + "cell_state", {cur.d[0], U}, SB_ROLE_STATE, SB_AGG_NONE, zeros, li);
          L.p_aux[13] = add_param(e, pre + "hidden_state", {cur.d[0], U}, SB_ROLE_STATE, SB_AGG_NONE, zeros, li);
        }
        if (d.return_sequence) {  // (batch, time, units): `joinAlong(outputs, 1)` (layer/RNN.scala:69-71)
          out.rank = 3;
          out.d[1] = cur.d[1];
          out.d[2] = U;
        } else {
          out.rank = 2;
          out.d[1] = U;
        }
        any_trainable_before = true;
        break;
      }
      case SB_LAYER_SIMPLE_RNN: {
        if (cur.rank != 3) SB_FAIL(SB_ERR_INVALID_ARG, "requirement failed: RNN requires input to have Shape(batch, time, features)");
        if (!d.trainable) SB_FAIL(SB_ERR_UNSUPPORTED, "frozen layers are outside the accelerated path");
        if (d.units <= 0) SB_FAIL(SB_ERR_INVALID_ARG, "requirement failed: RNN Cell should contain at least one unit");
        if (d.activation == SB_ACT_SOFTMAX) SB_FAIL(SB_ERR_UNSUPPORTED, "Softmax inside a SimpleRNN cell is outside the accelerated path");
        if (!d.use_bias && d.activation == SB_ACT_IDENTITY)
          SB_FAIL(SB_ERR_UNSUPPORTED, "a bare SimpleRNNCell (no bias, identity activation) names its parameters without the cell index");
        const int64_t F = cur.d[2], U = d.units;
        // the composed cell `SimpleRNNCell >> Bias >> act` prefixes its leaves' parameters with their index (layer/Composed.scala:18-27)
        L.p_aux[0] = add_param(e, pre + "0/kernel_weights", {F, U}, SB_ROLE_WEIGHT, SB_AGG_AVG, d.init[0], li);
        L.p_aux[1] = add_param(e, pre + "0/recurrent_weights", {U, U}, SB_ROLE_WEIGHT, SB_AGG_AVG, d.init[1], li);
        if (d.use_bias) L.p_aux[2] = add_param(e, pre + "1/weights", {U}, SB_ROLE_WEIGHT, SB_AGG_AVG, d.init[2], li);
        if (d.stateful) {  // SimpleRNNCell.State (layer/RNN.scala:170-172)
          sb_initializer zeros{};
          zeros.kind = SB_INIT_ZEROS;
          L.p_aux[3] = add_param(e, pre + "0/state", {cur.d[0], U}, SB_ROLE_STATE, SB_AGG_NONE, zeros, li);
        }
        if (d.return_sequence) {
          out.rank = 3;
          out.d[1] = cur.d[1];
          out.d[2] = U;
        } else {
          out.rank = 2;
          out.d[1] = U;
        }
        any_trainable_before = true;
        break;
      }
      default:
        SB_FAIL(SB_ERR_INVALID_ARG, "unknown layer kind " + std::to_string(d.kind));
    }
    L.out = out;
    if (new_buf) {
      buf_numel.push_back(out.numel());
      output_needed.push_back(0);
      cur_buf = (int)buf_numel.size() - 1;
      L.inplace = false;
    } else {
      L.inplace = true;
    }
    if (d.kind == SB_LAYER_ACTIVATE) output_needed[cur_buf] = 1;
    L.buf_out = cur_buf;
    cur = out;
  }
  e->out_shape = cur;
  // early programmatic-launch trigger in the kernels shared by every model: measured per BASELINE config (common.cuh) -- the CNN
  // step gains 6 %, the latency-bound fp32-GEMM MLP step loses 14 %
  e->pdl_early = false;
  for (const LayerPlan& Lp : e->L)
    if (Lp.d.kind == SB_LAYER_CONV2D || Lp.d.kind == SB_LAYER_POOL2D) e->pdl_early = true;
  if (cur.numel() != e->label_shape.numel())
    SB_FAIL(SB_ERR_INVALID_ARG, "model output " + std::to_string(cur.numel()) + " elements per batch does not match the labels " +
                                    std::to_string(e->label_shape.numel()));
  if (cur_buf == 0) SB_FAIL(SB_ERR_INVALID_ARG, "the model has no computing layer");
  e->act_numel = buf_numel;

  // arena layout: [trainable weights | state | opt slot 0 | opt slot 1 | opt slot 2]; weights first so that the
  // fused optimizer is one element-wise pass over nW floats and the gradient buffer mirrors the weight region.
  int64_t off = 0;
  for (auto& p : e->params)
    if (p.role == SB_ROLE_WEIGHT) { p.offset = off; off += p.padded; }
  e->nW = off;
  for (auto& p : e->params)
    if (p.role == SB_ROLE_STATE) { p.offset = off; off += p.padded; }
  e->nS = off - e->nW;
  const int n_model = (int)e->params.size();
  for (int slot = 0; slot < e->n_slots; ++slot)
    for (int wi = 0; wi < n_model; ++wi) {
      if (e->params[wi].role != SB_ROLE_WEIGHT) continue;
      const Param w = e->params[wi];
      sb_initializer in{};
      in.kind = SB_INIT_ZEROS;
      if (t.optimizer == SB_OPT_ADAMAX || t.optimizer == SB_OPT_NADAM) {  // Const(initAcc) (Adamax.scala:32-40)
        in.kind = SB_INIT_CONST;
        in.a = t.init_acc;
      }
      int pi;
      switch (w.rank) {
        case 1: pi = add_param(e, w.path + "/" + opt_slot_name(t.optimizer, slot), {w.dims[0]}, SB_ROLE_OPT, SB_AGG_AVG, in, w.layer); break;
        case 2: pi = add_param(e, w.path + "/" + opt_slot_name(t.optimizer, slot), {w.dims[0], w.dims[1]}, SB_ROLE_OPT, SB_AGG_AVG, in, w.layer); break;
        default: pi = add_param(e, w.path + "/" + opt_slot_name(t.optimizer, slot), {w.dims[0], w.dims[1], w.dims[2], w.dims[3]}, SB_ROLE_OPT, SB_AGG_AVG, in, w.layer); break;
      }
      Param& op = e->params[pi];
      op.weight_param = wi;
      op.slot = slot;
      op.offset = e->nW + e->nS + (int64_t)slot * e->nW + w.offset;
    }
  e->arena_floats = e->nW + e->nS + (int64_t)e->n_slots * e->nW;
}

// -------------------------------------------------------------------------------------------------
// allocation
// -------------------------------------------------------------------------------------------------
static float* dmalloc_floats(int64_t n) {
  void* p = nullptr;
  SB_CUDA(cudaMalloc(&p, (size_t)std::max<int64_t>(n, 32) * sizeof(float)));
  return (float*)p;
}

// Decide which layers run on the vectorised / fused kernels of kernels_fast.cu.  Runs after tc_plan_layers, which owns
// the contractions that fit tcgen05 (and already folded the ReLU that follows them into its epilogue).
// Early optimizer (cfg2: 89.3 -> 87.8 us per step, bit-identical; profiles/r2_notes.md 17): eligible when the tensor-core conv layer closest to the input that forks a side branch
// has only parameter-free layers between itself and the FIRST trainable layer -- then every gradient but the first layer's is complete
// when that side branch has reduced its filter gradient, and the arena from that layer's first parameter on can be updated there.
static void plan_early_optimizer(sb_engine* e) {
  e->early_opt_layer = -1;
  const char* off_ = getenv("SB_EARLY_OPT");  // =0: one optimizer launch at the end of the step
  if ((off_ && off_[0] == '0') || e->has_reg || !fork_enabled()) return;
  const int nl = (int)e->L.size();
  int first_tr = -1;
  for (const Param& p : e->params)
    if (p.role == SB_ROLE_WEIGHT && (first_tr < 0 || p.layer < first_tr)) first_tr = p.layer;
  if (first_tr < 0) return;
  int cand = -1;
  for (int li = first_tr + 1; li < nl; ++li) {
    const LayerPlan& L = e->L[li];
    bool has_params = false;
    for (const Param& p : e->params)
      if (p.role == SB_ROLE_WEIGHT && p.layer == li) has_params = true;
    if (!has_params) continue;
    if (L.d.kind == SB_LAYER_CONV2D && L.tc_kind == 1 && L.need_dx && !L.wgrad_in_pool) cand = li;
    break;  // the first parameterised layer behind the first trainable one decides
  }
  if (cand < 0) return;
  int64_t off = -1;
  for (const Param& p : e->params)
    if (p.role == SB_ROLE_WEIGHT && p.layer >= cand && (off < 0 || p.offset < off)) off = p.offset;
  for (const Param& p : e->params)  // arena order must be model order: nothing of the first layer behind `off`
    if (p.role == SB_ROLE_WEIGHT && p.layer < cand && p.offset >= off) return;
  if (off <= 0 || off % 4) return;
  e->early_opt_layer = cand;
  e->early_opt_off = off;
}

static void plan_fusions(sb_engine* e) {
  const char* off = getenv("SB_DISABLE_FAST");
  e->fast = !(off && off[0] == '1');
  const int nl = (int)e->L.size();
  auto inplace_relu = [&](int li) {
    return li >= 0 && li < nl && e->L[li].d.kind == SB_LAYER_ACTIVATE && e->L[li].d.activation == SB_ACT_RELU && e->L[li].inplace;
  };
  // the classifier head  Dense >> Bias >> Softmax  under CategoricalCrossentropy (every BASELINE classifier ends so)
  const LayerPlan& last = e->L[nl - 1];
  if (e->train.loss == SB_LOSS_CCE && last.d.kind == SB_LAYER_ACTIVATE && last.d.activation == SB_ACT_SOFTMAX && last.inplace &&
      last.out.rank == 2) {
    e->softmax_cce_fused = true;
    if (e->fast && nl >= 2 && e->L[nl - 2].d.kind == SB_LAYER_BIAS && e->L[nl - 2].inplace && last.out.d[1] <= 32 &&
        std::max<int64_t>(std::max<int64_t>(1, 32 / last.out.d[1]), (last.out.d[0] + 1023) / 1024) * last.out.d[1] <= 128) {
      e->head_fused = true;
      e->L[nl - 2].head_skip = true;
    }
    e->L[nl - 1].head_skip = true;
  }
  if (!e->fast) return;
  for (int li = 0; li < nl; ++li) {
    LayerPlan& L = e->L[li];
    switch (L.d.kind) {
      case SB_LAYER_CONV2D: {
        const ConvGeom& g = L.cg;
        if (!L.tc_kind && g.KH * g.KW * g.Cin <= 32 && g.Cout % 4 == 0 && g.Cout <= 256) {
          L.fast_smallk = true;
          if (inplace_relu(li + 1) && !e->L[li + 1].skip_fwd) {
            L.fuse_relu_fwd = true;
            L.fused_thr = e->L[li + 1].d.relu_threshold;
            e->L[li + 1].skip_fwd = true;
          }
        }
        break;
      }
      case SB_LAYER_BATCHNORM: {
        // BatchNorm >> Pool2D(2x2, stride 2, VALID, max): one forward (normalise + pool + winner bytes) and one backward (route by
        // the bytes, column sums, dx, and the ReLU in front of the BatchNorm) -- the normalised activations and the pool's / ReLU's
        // gradients never go to memory (SB_NO_BN_POOL=1: the separate kernels)
        const char* off = getenv("SB_NO_BN_POOL");
        if ((off && off[0] == '1') || li + 1 >= nl || L.in.rank != 4 || !L.need_dx) break;
        LayerPlan& PL = e->L[li + 1];
        if (PL.d.kind != SB_LAYER_POOL2D || (PL.d.pool_honor_reduce && PL.d.pool_reduce == SB_POOL_AVG) || PL.buf_in != L.buf_out ||
            L.buf_in == L.buf_out || !PL.need_dx || !bn_pool_ok(PL.pg))
          break;
        void* mp = nullptr;
        SB_CUDA(cudaMalloc(&mp, (size_t)PL.pg.N * PL.pg.OH * PL.pg.OW * PL.pg.C));
        PL.pool_map = (unsigned char*)mp;
        PL.skip_fwd = PL.skip_bwd = true;
        PL.in_bn = true;
        L.bn_pool = li + 1;
        if (inplace_relu(li - 1) && e->L[li - 1].buf_out == L.buf_in && e->L[li - 1].need_dx && !e->L[li - 1].skip_bwd) {
          L.bn_relu_bwd = true;
          L.fused_thr = e->L[li - 1].d.relu_threshold;
          e->L[li - 1].skip_bwd = true;
        }
        break;
      }
      case SB_LAYER_POOL2D: {
        if (L.in_bn) break;  // executed by the BatchNorm in front of it
        if (L.d.pool_honor_reduce && L.d.pool_reduce == SB_POOL_AVG) break;  // average pooling: the generic kernels, no fusions
        L.fast_pool = L.pg.C % 4 == 0;
        // dx = [x > thr] * maxpool_grad: the mask of the in-place ReLU that produced this layer's input
        if (L.need_dx && inplace_relu(li - 1) && e->L[li - 1].buf_out == L.buf_in) {
          L.pool_relu_bwd = true;
          L.fused_thr = e->L[li - 1].d.relu_threshold;
          e->L[li - 1].skip_bwd = true;
        }
        // conv(small K, first trainable layer) >> [in-place ReLU] >> pool(2x2, stride 1): the pool backward holds dA for
        // every pixel, so it also accumulates the conv's filter gradient and dA never goes to memory
        {
          const int ci = L.pool_relu_bwd ? li - 2 : li - 1;
          const PoolGeom& g = L.pg;
          const char* off = getenv("SB_NO_POOL_WGRAD_FUSION");
          if (!(off && off[0] == '1') && L.need_dx && L.fast_pool && ci >= 0 && e->L[ci].d.kind == SB_LAYER_CONV2D && e->L[ci].fast_smallk &&
              !e->L[ci].need_dx && e->L[ci].buf_out == L.buf_in && g.KH == 2 && g.KW == 2 && g.SH == 1 && g.SW == 1 && g.PT == 0 &&
              g.PL == 0 && e->L[ci].cg.SH == 1 && e->L[ci].cg.SW == 1) {
            const ConvGeom& cg = e->L[ci].cg;
            const int shape = cg.KH * 100 + cg.KW * 10 + cg.Cin, C4 = g.C / 4;
            if ((shape == 331 || shape == 333 || shape == 551) && C4 <= 32 && (C4 & (C4 - 1)) == 0) {
              L.fuse_conv_wgrad = ci;
              e->L[ci].wgrad_in_pool = true;
            }
          }
        }
        // forward twin: conv(small K) >> [fused ReLU] >> pool(2x2, stride 1) in one pass (the conv output is written, never re-read)
        {
          const int ci = (li >= 1 && e->L[li - 1].d.kind == SB_LAYER_ACTIVATE && e->L[li - 1].skip_fwd) ? li - 2 : li - 1;
          const PoolGeom& g = L.pg;
          const char* off = getenv("SB_NO_CONV_POOL_FUSION");
          if (!(off && off[0] == '1') && L.fast_pool && ci >= 0 && e->L[ci].d.kind == SB_LAYER_CONV2D && e->L[ci].fast_smallk &&
              (ci == li - 1 || e->L[ci].fuse_relu_fwd) && e->L[ci].buf_out == L.buf_in && g.KH == 2 && g.KW == 2 && g.SH == 1 &&
              g.SW == 1 && g.PT == 0 && g.PL == 0 && g.OH == g.H - 1 && g.OW == g.W - 1 && e->L[ci].cg.SH == 1 && e->L[ci].cg.SW == 1) {
            const ConvGeom& cg = e->L[ci].cg;
            const int shape = cg.KH * 100 + cg.KW * 10 + cg.Cin;
            if (shape == 331 || shape == 333) {
              e->L[ci].fuse_pool_fwd = li;
              L.skip_fwd = true;
              // Winner map: when nothing but this pool (and the ReLU folded into its backward) ever reads the conv activations,
              // the fused forward records one byte per pool output instead of storing them (SB_NO_POOL_MAP=1: the activation path)
              const char* nm = getenv("SB_NO_POOL_MAP");
              const bool relu_between = ci == li - 2;
              if (!(nm && nm[0] == '1') && L.need_dx && pool22_map_ok(g) && (!relu_between || L.pool_relu_bwd)) {
                void* mp = nullptr;
                SB_CUDA(cudaMalloc(&mp, (size_t)g.N * g.OH * g.OW * g.C));
                L.pool_map = (unsigned char*)mp;
              }
            }
          }
        }
        // tensor-core twin: conv (tcgen05 halo fprop, ReLU in its epilogue) >> pool(2x2, stride 1, VALID): the conv's eight epilogue
        // warps pool the staged tile and record the winner bytes; the conv activations are never written
        if (!L.skip_fwd && L.fast_pool) {
          const int ci = (li >= 1 && e->L[li - 1].d.kind == SB_LAYER_ACTIVATE && e->L[li - 1].skip_fwd) ? li - 2 : li - 1;
          const PoolGeom& g = L.pg;
          const bool relu_between = ci == li - 2;
          const char* nm = getenv("SB_NO_POOL_MAP");  // =1: every pool runs on the activations (A/B and the bit-identity test)
          if (!(nm && nm[0] == '1') && ci >= 0 && e->L[ci].d.kind == SB_LAYER_CONV2D && e->L[ci].tc_kind == 1 && e->L[ci].buf_out == L.buf_in &&
              pool22_map_ok(g) && (!relu_between || (L.pool_relu_bwd || !L.need_dx)) && (ci == li - 1 || relu_between)) {
            unsigned char* mp = nullptr;
            if (L.need_dx) SB_CUDA(cudaMalloc((void**)&mp, (size_t)g.N * g.OH * g.OW * g.C));
            if (tc_conv_enable_pool_fusion(e, e->L[ci], e->acts[L.buf_out], mp)) {
              e->L[ci].fuse_pool_fwd = li;
              L.skip_fwd = true;
              L.pool_map = mp;
            } else {
              cudaFree(mp);
            }
          }
        }
        break;
      }
      case SB_LAYER_BIAS: {
        // backward of the in-place activation that follows folds into this layer's column sum (one pass over g and y)
        const int nx = li + 1;
        if (L.inplace && !L.head_skip && nx < nl && e->L[nx].d.kind == SB_LAYER_ACTIVATE && e->L[nx].inplace && !e->L[nx].head_skip &&
            !e->L[nx].skip_bwd && e->L[nx].d.activation != SB_ACT_SOFTMAX && e->L[nx].d.activation != SB_ACT_IDENTITY &&
            e->L[nx].need_dx && L.d.units % 4 == 0) {
          L.fuse_act_bwd = true;
          e->L[nx].skip_bwd = true;
        }
        break;
      }
      case SB_LAYER_DENSE:
        L.fast_skinny = L.out.d[1] <= 16 && L.in.d[1] >= 512 && L.in.d[1] % 4 == 0;
        break;
      default:
        break;
    }
  }
  // Second pass (needs fast_skinny of later layers), OFF by default -- SB_POOL_MAP_ALL=1 turns it on: winner maps for 2x2 stride-1
  // pools whose forward is the standalone kernel, and the classifier head's input gradient folded into the backward of the pool
  // that feeds it (Pool2D >> Flatten >> Dense(N <= 16)).  Measured on cfg2 (profiles/r2_notes.md): recording the map costs the
  // standalone pool forward more than the map-driven backward saves (975 k vs 1019 k samples/s), and the fused pool-backward /
  // head-dgrad kernel is issue-bound at 23 us (985 k): both stay parity-tested (bit-identical to the default path) for the day the
  // map comes for free out of the tensor-core conv's epilogue.
  const char* nm = getenv("SB_NO_POOL_MAP");
  const char* all = getenv("SB_POOL_MAP_ALL");
  const char* nh = getenv("SB_NO_HEAD_DGRAD_FUSION");
  for (int li = 0; li < nl; ++li) {
    LayerPlan& L = e->L[li];
    if (L.d.kind != SB_LAYER_POOL2D || !L.fast_pool || !L.need_dx || !pool22_map_ok(L.pg) || (nm && nm[0] == '1')) continue;
    if (!(all && all[0] == '1')) continue;
    if (!L.pool_map && !L.skip_fwd) {  // the standalone forward kernel records the map
      void* mp = nullptr;
      SB_CUDA(cudaMalloc(&mp, (size_t)L.pg.N * L.pg.OH * L.pg.OW * L.pg.C));
      L.pool_map = (unsigned char*)mp;
    }
    if (!L.pool_map || L.fuse_conv_wgrad >= 0 || (nh && nh[0] == '1')) continue;
    if (li + 2 < nl && e->L[li + 1].d.kind == SB_LAYER_FLATTEN && e->L[li + 2].d.kind == SB_LAYER_DENSE && e->L[li + 2].fast_skinny &&
        !e->L[li + 2].tc_kind && e->L[li + 2].buf_in == L.buf_out && e->L[li + 2].out.d[1] <= 16 &&
        e->L[li + 2].in.d[1] == (int64_t)L.pg.OH * L.pg.OW * L.pg.C && 2 * L.pg.OW * (L.pg.C / 4) <= 704 && L.pg.W * (L.pg.C / 4) <= 704) {
      L.fuse_head_dgrad = li + 2;
      e->L[li + 2].dgrad_in_pool = true;
      if (!e->ws_side) {
        e->ws_side_floats = (size_t)e->L[li + 2].in.d[1] * e->L[li + 2].out.d[1] * 8;
        e->ws_side = dmalloc_floats((int64_t)e->ws_side_floats);
      }
    }
  }
}

static void allocate(sb_engine* e) {
  SB_CUDA(cudaSetDevice(e->device));
  // the main stream outranks the backward side branch: filter gradients fill SMs the input-gradient chain leaves idle, they must
  // not delay it (stream priorities are captured into the graph's kernel nodes)
  int prio_lo = 0, prio_hi = 0;
  SB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  SB_CUDA(cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prio_hi));
  SB_CUDA(cudaStreamCreateWithPriority(&e->bstream, cudaStreamNonBlocking, prio_lo));  // created outside any capture
  SB_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  SB_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  SB_CUDA(cudaEventCreateWithFlags(&e->ev_mid, cudaEventDisableTiming));
  // the arena allocation carries a tail of cross-GPU flags / scalars behind the tensors (one IPC handle maps both)
  e->arena = dmalloc_floats(e->arena_floats + sb::kArenaTailFloats);
  SB_CUDA(cudaMemsetAsync(e->arena, 0, (size_t)(e->arena_floats + sb::kArenaTailFloats) * 4, e->stream));
  e->grads = dmalloc_floats(e->nW);
  SB_CUDA(cudaMemsetAsync(e->grads, 0, (size_t)std::max<int64_t>(e->nW, 32) * 4, e->stream));
  const int nb = (int)e->act_numel.size();
  e->acts.assign(nb, nullptr);
  e->dacts.assign(nb, nullptr);
  for (int i = 0; i < nb; ++i) {
    e->acts[i] = dmalloc_floats(pad32(e->act_numel[i]));
    if (i > 0) e->dacts[i] = dmalloc_floats(pad32(e->act_numel[i]));
  }
  e->by = dmalloc_floats(pad32(e->label_shape.numel()));
  void* p = nullptr;
  SB_CUDA(cudaMalloc(&p, sizeof(StepState)));
  e->st = (StepState*)p;
  SB_CUDA(cudaMemsetAsync(e->st, 0, sizeof(StepState), e->stream));
  SB_CUDA(cudaMallocHost(&p, sizeof(StepState) * 4));
  e->st_host = (StepState*)p;
  memset(e->st_host, 0, sizeof(StepState) * 4);
  e->ws_floats = (size_t)16 << 20;  // 64 MB: split-K partials, column-reduction partials
  e->ws = dmalloc_floats((int64_t)e->ws_floats);
  e->head_scratch = dmalloc_floats((int64_t)kHeadScratchFloats);
  SB_CUDA(cudaMemsetAsync(e->head_scratch, 0, kHeadScratchFloats * 4, e->stream));
  for (auto& L : e->L) {
    if (L.d.kind == SB_LAYER_BATCHNORM) {
      int64_t C = L.in.d[L.in.rank - 1];
      L.bn_save_mean = dmalloc_floats(C);
      L.bn_save_var = dmalloc_floats(C);
    }
  }
  tc_plan_layers(e);  // tensor-core fast paths (TMA descriptors, padded-grid buffers)
  plan_fusions(e);    // vectorised / fused bandwidth kernels
  plan_early_optimizer(e);
  lstm_plan(e);       // LSTM workspaces (saved gate activations for BPTT)
  rnn_plan(e);        // SimpleRNN workspaces
  SB_CUDA(cudaStreamSynchronize(e->stream));
}

// -------------------------------------------------------------------------------------------------
// profiling scopes (eager mode only)
// -------------------------------------------------------------------------------------------------
namespace {
struct Prof {
  sb_engine* e;
  int idx = -1;
  cudaEvent_t a = nullptr, b = nullptr;
  long long launches0 = 0;
  Prof(sb_engine* e_, const char* name, double bytes, double flops) : e(e_) {
    if (!e->profiling) return;
    auto it = e->prof_by_name.find(name);
    if (it == e->prof_by_name.end()) {
      e->prof_by_name[name] = idx = (int)e->prof.size();
      ProfEntry pe;
      pe.name = name;
      e->prof.push_back(pe);
    } else {
      idx = it->second;
    }
    e->prof[idx].bytes += bytes;
    e->prof[idx].flops += flops;
    launches0 = tl_launch_count;
    SB_CUDA(cudaEventCreate(&a));
    SB_CUDA(cudaEventCreate(&b));
    SB_CUDA(cudaEventRecord(a, e->stream));
  }
  ~Prof() {
    if (idx < 0) return;
    cudaEventRecord(b, e->stream);
    e->prof[idx].launches += tl_launch_count - launches0;
    e->prof_pending.push_back(ProfPending{idx, a, b});
  }
};
}  // namespace

static void prof_resolve(sb_engine* e) {
  if (e->prof_pending.empty()) return;
  SB_CUDA(cudaStreamSynchronize(e->stream));
  for (auto& pp : e->prof_pending) {
    float ms = 0;
    SB_CUDA(cudaEventElapsedTime(&ms, pp.a, pp.b));
    e->prof[pp.entry].total_ms += ms;
    cudaEventDestroy(pp.a);
    cudaEventDestroy(pp.b);
  }
  e->prof_pending.clear();
}

// -------------------------------------------------------------------------------------------------
// the train step
// -------------------------------------------------------------------------------------------------
static float* P(sb_engine* e, int pi) { return e->arena + e->params[pi].offset; }
static float* G(sb_engine* e, int pi) { return e->grads + e->params[pi].offset; }

enum FwdMode { FWD_TRAIN = 0, FWD_LOSS_ONLY = 1, FWD_INFER = 2 };

// per-kernel timing collects a layer's deferred reduction so that it is its own class; it must then run what the product path runs
static void reduce_each(ReduceBatch& rb, cudaStream_t s) {
  for (int i = 0; i < rb.count; ++i) partial_reduce(rb.partial[i], rb.out[i], rb.n[i], rb.parts[i], s);
  rb.count = 0;
}

static void run_forward(sb_engine* e, int mode) {
  cudaStream_t s = e->stream;
  const int nl = (int)e->L.size();
  for (int li = 0; li < nl; ++li) {
    LayerPlan& L = e->L[li];
    if (L.fused_into_prev || L.skip_fwd) continue;
    if (L.head_skip && mode != FWD_INFER) continue;  // executed by the loss kernel
    float* in = e->acts[L.buf_in];
    float* out = e->acts[L.buf_out];
    const int64_t n_out = L.out.numel();
    switch (L.d.kind) {
      case SB_LAYER_DENSE: {
        const int M = (int)L.in.d[0], K = (int)L.in.d[1], N = (int)L.out.d[1];
        const double db = 4.0 * ((double)M * K + (double)K * N + (double)M * N);
        if (L.tc_kind) {
          Prof pr(e, "dense_fwd_tf32_tcgen05", db, 2.0 * M * N * K);
          if (tc_dense_fwd(e, L, li)) break;
        }
        if (L.fast_skinny) {
          Prof pr(e, "dense_head_fwd", db, 2.0 * M * N * K);
          e->head_chunks = 0;
          if (e->head_fused && li == nl - 3 && mode != FWD_INFER) {
            // split-K partials stay in the workspace: the head kernel (next launch) sums them in chunk order
            if (tc_head_fwd(e, L, in, P(e, L.p_w), e->ws, e->ws_floats, &e->head_chunks)) break;
            if (dense_head_fwd(in, P(e, L.p_w), e->ws, &e->head_chunks, M, N, K, e->ws_floats, s)) break;
          } else if (dense_skinny_fwd(in, P(e, L.p_w), out, M, N, K, e->ws, e->ws_floats, s)) {
            break;
          }
        }
        Prof pr(e, "dense_fwd_f32", db, 2.0 * M * N * K);
        gemm_nn(in, P(e, L.p_w), out, M, N, K, e->ws, e->ws_floats, s);
        break;
      }
      case SB_LAYER_BIAS: {
        Prof pr(e, "bias_add", 8.0 * n_out, 0);
        if (!L.inplace) SB_CUDA(cudaMemcpyAsync(out, in, (size_t)n_out * 4, cudaMemcpyDeviceToDevice, s));
        bias_add(out, P(e, L.p_w), n_out / L.d.units, L.d.units, s);
        break;
      }
      case SB_LAYER_ACTIVATE: {
        Prof pr(e, "act_fwd", 8.0 * n_out, 0);
        if (!L.inplace) SB_CUDA(cudaMemcpyAsync(out, in, (size_t)n_out * 4, cudaMemcpyDeviceToDevice, s));
        if (L.d.activation == SB_ACT_SOFTMAX) softmax_fwd(out, (int)L.out.d[0], (int)L.out.d[1], s);
        else act_fwd(out, L.d.activation, L.d.relu_threshold, n_out, s);
        break;
      }
      case SB_LAYER_FLATTEN:
        break;
      case SB_LAYER_CONV2D: {
        const ConvGeom& g = L.cg;
        const double cbytes = 4.0 * ((double)L.in.numel() + n_out + e->params[L.p_w].numel);
        const double cflops = 2.0 * g.N * g.OH * g.OW * (double)g.Cout * g.KH * g.KW * g.Cin;
        if (L.tc_kind) {
          Prof pr(e, "conv2d_fwd_tf32_tcgen05", cbytes, cflops);
          if (tc_conv_fwd(e, L, li)) break;
        }
        if (L.fast_smallk && L.fuse_pool_fwd >= 0) {
          LayerPlan& PLy = e->L[L.fuse_pool_fwd];
          Prof pr(e, "conv2d_smallk_pool_fwd", cbytes + 4.0 * (double)PLy.out.numel(), cflops);
          if (conv_smallk_pool_fwd(in, P(e, L.p_w), out, e->acts[PLy.buf_out], g, L.fuse_relu_fwd, L.fused_thr, PLy.pool_map, s)) break;
          SB_FAIL(SB_ERR_CUDA, "conv_smallk_pool_fwd rejected a planned shape");
        }
        if (L.fast_smallk) {
          Prof pr(e, "conv2d_smallk_fwd", cbytes, cflops);
          conv_smallk_fwd(in, P(e, L.p_w), out, g, L.fuse_relu_fwd, L.fused_thr, s);
          break;
        }
        Prof pr(e, "conv2d_fwd_f32", cbytes, cflops);
        conv2d_fwd(in, P(e, L.p_w), out, g, e->ws, e->ws_floats, s);
        break;
      }
      case SB_LAYER_POOL2D: {
        Prof pr(e, "pool_fwd", 4.0 * ((double)L.in.numel() + n_out), 0);
        // the reference layer ignores `reduce` (layer/Pool2D.scala:36-42): always max, unless the descriptor asks to honour it
        if (L.fast_pool) pool_max_fwd_v4(in, out, L.pg, s, L.pool_map, L.pool_relu_bwd, L.fused_thr);
        else pool_fwd(in, out, L.pg, (L.d.pool_honor_reduce && L.d.pool_reduce == SB_POOL_AVG) ? SB_POOL_AVG : SB_POOL_MAX, s);
        break;
      }
      case SB_LAYER_BATCHNORM: {
        const int C = (int)L.in.d[L.in.rank - 1];
        const long long rows = L.in.numel() / C;
        if (L.bn_pool >= 0) {
          LayerPlan& PL = e->L[L.bn_pool];
          Prof pr(e, "batchnorm_pool_fwd", 4.0 * (2.0 * L.in.numel() + 1.25 * PL.out.numel()), 0);
          const float *mean = P(e, L.p_aux[2]), *var = P(e, L.p_aux[3]);  // `freeze` / predict: the moving statistics
          if (mode != FWD_INFER) {
            if (!bn_stats(in, rows, C, L.bn_save_mean, L.bn_save_var, P(e, L.p_aux[2]), P(e, L.p_aux[3]), L.d.bn_momentum, L.bn_fused, L.d.bn_mode,
                          mode == FWD_TRAIN, e->ws, e->ws_floats, s))
              SB_FAIL(SB_ERR_CUDA, "bn_stats rejected a planned shape");
            mean = L.bn_save_mean;
            var = L.bn_save_var;
          }
          if (!bn_apply_pool_fwd(in, e->acts[PL.buf_out], mode != FWD_INFER ? PL.pool_map : nullptr, P(e, L.p_aux[1]), P(e, L.p_aux[0]), mean, var,
                                 L.d.bn_epsilon, L.bn_fused, PL.pg, s))
            SB_FAIL(SB_ERR_CUDA, "bn_apply_pool_fwd rejected a planned shape");
          break;
        }
        Prof pr(e, "batchnorm_fwd", 4.0 * (3.0 * L.in.numel()), 0);
        if (mode == FWD_INFER) {
          bn_fwd_infer(in, out, P(e, L.p_aux[1]), P(e, L.p_aux[0]), P(e, L.p_aux[2]), P(e, L.p_aux[3]), rows, C, L.d.bn_epsilon,
                       L.bn_fused, s);
        } else {
          bn_fwd(in, out, P(e, L.p_aux[1]), P(e, L.p_aux[0]), P(e, L.p_aux[2]), P(e, L.p_aux[3]), L.bn_save_mean, L.bn_save_var,
                 rows, C, L.d.bn_epsilon, L.d.bn_momentum, L.bn_fused, L.d.bn_mode, mode == FWD_TRAIN, e->ws, e->ws_floats, s);
        }
        break;
      }
      case SB_LAYER_LSTM: {
        Prof pr(e, "lstm_fwd", 0, 0);
        lstm_forward(e, L, li, mode == FWD_TRAIN);
        break;
      }
      case SB_LAYER_SIMPLE_RNN: {
        Prof pr(e, "rnn_fwd", 0, 0);
        rnn_forward(e, L, li, mode == FWD_TRAIN);
        break;
      }
    }
  }
}

// Side branch of the backward pass: work that only the optimizer waits for (filter gradients) is queued on `bstream` behind the
// point the main stream has reached; successive branches simply queue up there (in order), and the main stream joins the last
// one at the end of run_backward.  Inside a captured graph these become parallel branches.
static void side_branch_begin(sb_engine* e) {
  SB_CUDA(cudaEventRecord(e->ev_fork, e->stream));
  SB_CUDA(cudaStreamWaitEvent(e->bstream, e->ev_fork, 0));
}
static void side_branch_end(sb_engine* e) {
  SB_CUDA(cudaEventRecord(e->ev_join, e->bstream));
  e->bwd_forked = true;
}

static void run_loss(sb_engine* e, bool with_grad) {
  const LayerPlan& last = e->L.back();
  float* pred = e->acts[last.buf_out];
  float* dpred = with_grad ? e->dacts[last.buf_out] : nullptr;
  const long long rows = e->label_shape.d[0];
  const long long cols = e->label_shape.numel() / rows;
  if (e->head_fused) {
    const LayerPlan& bias = e->L[e->L.size() - 2];
    Prof pr(e, "head_bias_softmax_cce", 4.0 * 4.0 * (double)rows * cols, 0);
    // With a backward pass behind it, the final sum of the per-CTA loss / bias-gradient partials moves to the side branch: only dz
    // is on the critical path (SB_HEAD_DEFER=0 keeps the in-kernel last-CTA combine)
    static const bool defer_on = [] { const char* v = getenv("SB_HEAD_DEFER"); return !(v && v[0] == '0'); }();
    const bool defer = defer_on && with_grad && fork_enabled() && !e->profiling;
    int n_ctas = 0;
    if (!head_bias_softmax_cce(pred, e->head_chunks > 0 ? e->ws : nullptr, e->head_chunks, P(e, bias.p_w), e->by, dpred,
                               with_grad ? G(e, bias.p_w) : nullptr, e->st, (int)rows, (int)cols, e->head_scratch, e->stream,
                               defer ? &n_ctas : nullptr))
      SB_FAIL(SB_ERR_CUDA, "fused classifier head rejected a planned shape");
    if (defer) {
      side_branch_begin(e);
      head_loss_combine(e->head_scratch, n_ctas, (int)cols, (int)rows, e->st, G(e, bias.p_w), e->bstream);
      side_branch_end(e);
    }
  } else if (e->softmax_cce_fused) {
    Prof pr(e, "softmax_cce_fused", 4.0 * 4.0 * (double)rows * cols, 0);
    softmax_cce_fwd_bwd(pred, e->by, dpred, e->st, (int)rows, (int)cols, e->stream);
  } else {
    Prof pr(e, "loss_fwd_bwd", 4.0 * 3.0 * (double)rows * cols, 0);
    loss_fwd_bwd(e->train.loss, pred, e->by, dpred, e->st, rows, cols, e->stream);
  }
}

// SB_MERGE_REDUCE=1 (experiment, off): the filter-gradient partial sums of the step in ONE launch before the optimizer instead of
// one reduction right behind each kernel.  Measured on cfg2: 1097 k -> 975 k samples/s -- conv2's reduction (590 k threads, ~8 us
// in the step) leaves the side branch, where it overlapped pool1's backward, for the serial tail.
static ReduceBatch* reduce_batch(sb_engine* e) {
  static const bool on = [] { const char* v = getenv("SB_MERGE_REDUCE"); return v && v[0] == '1'; }();
  return (on && !e->profiling) ? &e->pending_reduce : nullptr;
}

static void optimizer_range(sb_engine* e, int64_t off, int64_t n, int advance_batch, cudaStream_t s);

static void run_backward(sb_engine* e) {
  cudaStream_t s = e->stream;
  e->pending_reduce.count = 0;
  const int nl = (int)e->L.size();
  for (int li = nl - 1; li >= 0; --li) {
    LayerPlan& L = e->L[li];
    if (L.fused_into_prev || L.head_skip || L.skip_bwd) continue;
    float* in = e->acts[L.buf_in];
    float* out = e->acts[L.buf_out];
    float* dout = e->dacts[L.buf_out];
    float* din = e->dacts[L.buf_in];  // null for the input buffer
    const int64_t n_out = L.out.numel();
    switch (L.d.kind) {
      case SB_LAYER_DENSE: {
        const int M = (int)L.in.d[0], K = (int)L.in.d[1], N = (int)L.out.d[1];
        const double db = 4.0 * ((double)M * K + (double)K * N + (double)M * N);
        bool wdone = false, ddone = false;
        if (L.tc_kind) {
          {
            Prof pr(e, "dense_wgrad_tf32_tcgen05", db, 2.0 * M * N * K);
            wdone = tc_dense_wgrad(e, L, li);
          }
          if (L.need_dx && din) {
            Prof pr(e, "dense_dgrad_tf32_tcgen05", db, 2.0 * M * N * K);
            ddone = tc_dense_dgrad(e, L, li);
          }
        }
        if (L.fast_skinny && !wdone && L.dgrad_in_pool) {
          // dX is formed inside the backward kernel of the pool that feeds this head (pool_bwd_head_dgrad): only the weight gradient
          // is left, and nothing but the optimizer waits for it -- side branch (its own workspace: `ws` is in use on the main stream)
          const bool fork = fork_enabled() && !e->profiling;
          Prof pr(e, "dense_head_wgrad", db, 2.0 * M * N * K);
          if (fork) side_branch_begin(e);
          wdone = ddone = tc_head_bwd(e, L, in, dout, P(e, L.p_w), G(e, L.p_w), nullptr, fork ? e->bstream : s) ||
                          dense_head_bwd(in, dout, P(e, L.p_w), G(e, L.p_w), nullptr, M, N, K, e->ws_side, e->ws_side_floats,
                                         fork ? e->bstream : s);
          if (fork) side_branch_end(e);
          if (!wdone) SB_FAIL(SB_ERR_CUDA, "dense_head_bwd rejected a planned shape");
          break;
        }
        if (L.fast_skinny && !wdone) {
          const bool dg = L.need_dx && din;
          Prof pr(e, "dense_head_bwd", db + (dg ? 4.0 * M * K : 0.0), (dg ? 4.0 : 2.0) * M * N * K);
          wdone = ddone = tc_head_bwd(e, L, in, dout, P(e, L.p_w), G(e, L.p_w), dg ? din : nullptr, s) ||
                          dense_head_bwd(in, dout, P(e, L.p_w), G(e, L.p_w), dg ? din : nullptr, M, N, K, e->ws, e->ws_floats, s);
        }
        if (!wdone) {
          Prof pr(e, "dense_wgrad_f32", db, 2.0 * M * N * K);
          gemm_tn(in, dout, G(e, L.p_w), K, N, M, e->ws, e->ws_floats, s);  // dW = X^T . G
        }
        if (!ddone && L.need_dx && din) {
          Prof pr(e, "dense_dgrad_f32", db, 2.0 * M * N * K);
          gemm_nt(dout, P(e, L.p_w), din, M, K, N, e->ws, e->ws_floats, s);  // dX = G . W^T
        }
        break;
      }
      case SB_LAYER_BIAS: {
        if (L.fuse_act_bwd) {
          const LayerPlan& A = e->L[li + 1];
          Prof pr(e, "act_bwd_bias_grad", 12.0 * n_out, 0);
          if (!act_bwd_col_sum(dout, e->acts[A.buf_out], A.d.activation, A.d.relu_threshold, G(e, L.p_w), n_out / L.d.units,
                               L.d.units, e->ws, e->ws_floats, s))
            SB_FAIL(SB_ERR_CUDA, "act_bwd_col_sum rejected a planned shape");
          break;
        }
        Prof pr(e, "bias_grad", 4.0 * n_out, 0);
        if (e->fast && act_bwd_col_sum(dout, nullptr, 0, 0.f, G(e, L.p_w), n_out / L.d.units, L.d.units, e->ws, e->ws_floats, s)) {
          if (!L.inplace && L.need_dx && din) SB_CUDA(cudaMemcpyAsync(din, dout, (size_t)n_out * 4, cudaMemcpyDeviceToDevice, s));
          break;
        }
        col_sum(dout, G(e, L.p_w), n_out / L.d.units, L.d.units, e->ws, e->ws_floats, s);
        if (!L.inplace && L.need_dx && din) SB_CUDA(cudaMemcpyAsync(din, dout, (size_t)n_out * 4, cudaMemcpyDeviceToDevice, s));
        break;
      }
      case SB_LAYER_ACTIVATE: {
        if (!L.need_dx) break;
        Prof pr(e, "act_bwd", 12.0 * n_out, 0);
        float* g = dout;
        if (!L.inplace) {
          if (!din) break;
          SB_CUDA(cudaMemcpyAsync(din, dout, (size_t)n_out * 4, cudaMemcpyDeviceToDevice, s));
          g = din;
        }
        if (L.d.activation == SB_ACT_SOFTMAX) softmax_bwd(g, out, (int)L.out.d[0], (int)L.out.d[1], s);
        else act_bwd(g, out, L.d.activation, L.d.relu_threshold, n_out, s);
        break;
      }
      case SB_LAYER_FLATTEN:
        break;
      case SB_LAYER_CONV2D: {
        const ConvGeom& g = L.cg;
        const double fl = 2.0 * g.N * g.OH * g.OW * (double)g.Cout * g.KH * g.KW * g.Cin;
        const double cb = 4.0 * ((double)L.in.numel() + n_out + e->params[L.p_w].numel);
        bool wdone = L.wgrad_in_pool, ddone = false;  // wgrad_in_pool: done by the following pool's backward kernel
        if (L.tc_kind) {
          // With an input-gradient chain behind it, the filter gradient (tensor-core kernel + reduce) moves to a side branch
          // that starts AFTER the dgrad kernel: both want every SM's shared memory, whereas the pool / activation backward
          // kernels that follow dgrad leave room for it (not while per-kernel timing is on).
          const bool fork = fork_enabled() && !e->profiling && L.need_dx && din && !L.wgrad_in_pool;
          // SB_FORK_MODE: 1 (default) = the side branch starts BEFORE dgrad, as soon as dY exists; 0 = after dgrad; 2 = no side
          // branch.  Round 1 measured "after" ahead (both kernels want every SM's shared memory); with the dgrad kernel folded and
          // the head's combine gone from the main stream the filter gradient + its reduce had become the tail of the step (in-situ
          // timeline: they finished 9 us after the main stream's last kernel): 1050 k -> 1090 k samples/s, no branch 1022 k.
          static const int fork_mode = [] { const char* v = getenv("SB_FORK_MODE"); return v ? atoi(v) : 1; }();
          const bool early_opt = fork && fork_mode == 1 && e->early_opt_now && li == e->early_opt_layer;
          if (early_opt) {
            // filter gradient + reduction on the side branch, dgrad on the main stream, and -- once dgrad has read this layer's
            // weights -- the optimizer update of every parameter from this layer on, still on the side branch
            side_branch_begin(e);
            wdone = tc_conv_wgrad(e, L, li, e->bstream, nullptr);
            {
              Prof pr(e, "conv2d_dgrad_tf32_tcgen05", cb, fl);
              ddone = tc_conv_dgrad(e, L, li);
            }
            SB_CUDA(cudaEventRecord(e->ev_mid, e->stream));
            SB_CUDA(cudaStreamWaitEvent(e->bstream, e->ev_mid, 0));
            optimizer_range(e, e->early_opt_off, e->nW - e->early_opt_off, 0, e->bstream);
            e->early_opt_done = true;
            side_branch_end(e);
          } else if (fork && fork_mode == 1) {
            side_branch_begin(e);
            wdone = tc_conv_wgrad(e, L, li, e->bstream, reduce_batch(e));
            side_branch_end(e);
          }
          if (!fork || fork_mode == 2) {
            // per-kernel timing: the deterministic reduction of the per-CTA partials is its own class, not half of this one
            ReduceBatch rb;
            {
              Prof pr(e, "conv2d_wgrad_tf32_tcgen05", cb, fl);
              wdone = tc_conv_wgrad(e, L, li, nullptr, e->profiling ? &rb : nullptr);
            }
            if (rb.count) {
              Prof pr(e, "partial_reduce", 4.0 * ((double)rb.n[0] * rb.parts[0] + rb.n[0]), 0);
              reduce_each(rb, s);  // the launches the product path makes, not the merged variant
            }
          }
          if (L.need_dx && din && !ddone) {
            Prof pr(e, "conv2d_dgrad_tf32_tcgen05", cb, fl);
            ddone = tc_conv_dgrad(e, L, li);
          }
          if (fork && fork_mode == 0) {
            side_branch_begin(e);
            wdone = tc_conv_wgrad(e, L, li, e->bstream);
            side_branch_end(e);
            if (!wdone) SB_FAIL(SB_ERR_CUDA, "tensor-core filter gradient rejected a planned shape");
          }
        }
        if (!wdone && L.fast_smallk) {
          Prof pr(e, "conv2d_smallk_wgrad", cb, fl);
          wdone = conv_smallk_wgrad(in, dout, G(e, L.p_w), g, e->ws, e->ws_floats, s);
        }
        if (!wdone) {
          Prof pr(e, "conv2d_wgrad_f32", cb, fl);
          conv2d_wgrad(in, dout, G(e, L.p_w), g, e->ws, e->ws_floats, s);
        }
        if (!ddone && L.need_dx && din) {
          Prof pr(e, "conv2d_dgrad_f32", cb, fl);
          conv2d_dgrad(dout, P(e, L.p_w), din, g, e->ws, e->ws_floats, s);
        }
        break;
      }
      case SB_LAYER_POOL2D: {
        if (!L.need_dx || !din) break;
        if (L.fuse_conv_wgrad >= 0) {
          LayerPlan& CL = e->L[L.fuse_conv_wgrad];
          ReduceBatch rb;
          bool ok;
          {
            Prof pr(e, "pool_relu_bwd_conv_wgrad", 4.0 * ((double)L.in.numel() + n_out + CL.in.numel()),
                    2.0 * CL.cg.N * CL.cg.OH * CL.cg.OW * (double)CL.cg.Cout * CL.cg.KH * CL.cg.KW * CL.cg.Cin);
            ok = pool_bwd_conv_wgrad(in, dout, din, L.pg, L.pool_relu_bwd, L.fused_thr, e->acts[CL.buf_in], CL.cg, G(e, CL.p_w), 0, e->ws,
                                     e->ws_floats, s, L.pool_map, e->profiling ? &rb : reduce_batch(e));
          }
          if (rb.count) {  // per-kernel timing: the reduction of the per-CTA partials is its own class
            Prof pr(e, "partial_reduce", 4.0 * ((double)rb.n[0] * rb.parts[0] + rb.n[0]), 0);
            reduce_each(rb, s);  // the launches the product path makes, not the merged variant
          }
          if (ok) break;
          SB_FAIL(SB_ERR_CUDA, "pool_bwd_conv_wgrad rejected a planned shape");
        }
        if (L.fuse_head_dgrad >= 0) {
          const LayerPlan& D = e->L[L.fuse_head_dgrad];
          const int Nc = (int)D.out.d[1];
          Prof pr(e, "pool_bwd_head_dgrad", 4.0 * ((double)L.in.numel() + (double)D.in.d[1] * Nc) + (double)n_out, 2.0 * (double)n_out * Nc);
          if (pool_bwd_head_dgrad(L.pool_map, e->dacts[D.buf_out], P(e, D.p_w), din, L.pg, Nc, L.pool_relu_bwd, s)) break;
          SB_FAIL(SB_ERR_CUDA, "pool_bwd_head_dgrad rejected a planned shape");
        }
        Prof pr(e, L.pool_relu_bwd ? "pool_relu_bwd" : "pool_bwd", 4.0 * (2.0 * L.in.numel() + 2.0 * n_out), 0);
        if (L.fast_pool) pool_max_bwd_v4(in, dout, din, L.pg, L.pool_relu_bwd, L.fused_thr, s, L.pool_map);
        else if (L.pool_relu_bwd) pool_bwd_relu(in, dout, din, L.pg, L.fused_thr, s);
        else pool_bwd(in, dout, din, L.pg, (L.d.pool_honor_reduce && L.d.pool_reduce == SB_POOL_AVG) ? SB_POOL_AVG : SB_POOL_MAX, s);
        break;
      }
      case SB_LAYER_BATCHNORM: {
        const int C = (int)L.in.d[L.in.rank - 1];
        const long long rows = L.in.numel() / C;
        if (L.bn_pool >= 0) {
          LayerPlan& PL = e->L[L.bn_pool];
          Prof pr(e, "batchnorm_pool_bwd", 4.0 * (3.0 * L.in.numel() + 2.5 * PL.out.numel()), 0);
          if (!bn_pool_bwd(in, e->dacts[PL.buf_out], PL.pool_map, din, P(e, L.p_aux[1]), L.bn_save_mean, L.bn_save_var, G(e, L.p_aux[1]),
                           G(e, L.p_aux[0]), PL.pg, L.d.bn_epsilon, L.bn_fused, L.d.bn_mode, L.bn_relu_bwd, L.fused_thr, e->ws, e->ws_floats, s))
            SB_FAIL(SB_ERR_CUDA, "bn_pool_bwd rejected a planned shape");
          break;
        }
        Prof pr(e, "batchnorm_bwd", 4.0 * (5.0 * L.in.numel()), 0);
        bn_bwd(in, dout, (L.need_dx && din) ? din : nullptr, P(e, L.p_aux[1]), L.bn_save_mean, L.bn_save_var, G(e, L.p_aux[1]),
               G(e, L.p_aux[0]), rows, C, L.d.bn_epsilon, L.bn_fused, L.d.bn_mode, e->ws, e->ws_floats, s);
        break;
      }
      case SB_LAYER_LSTM: {
        Prof pr(e, "lstm_bwd", 0, 0);
        lstm_backward(e, L, li);
        break;
      }
      case SB_LAYER_SIMPLE_RNN: {
        Prof pr(e, "rnn_bwd", 0, 0);
        rnn_backward(e, L, li);
        break;
      }
    }
  }
  if (e->bwd_forked) {  // every gradient is in place before the penalties / the optimizer / a read-back
    SB_CUDA(cudaStreamWaitEvent(s, e->ev_join, 0));
    e->bwd_forked = false;
  }
  // the filter-gradient partials collected above (both streams have joined): one launch for all of them
  const long long before = tl_launch_count;
  partial_reduce_multi(e->pending_reduce, s);
  (void)before;
}

// loss += penalty(theta) and (with_grad) grad += dPenalty/dw, tensors in model order (models/Model.scala:92-119)
static void run_reg(sb_engine* e, bool with_grad) {
  if (!e->has_reg) return;
  int n_reg = 0, seen = 0;
  for (const Param& p : e->params)
    if (p.role == SB_ROLE_WEIGHT && p.reg_kind) ++n_reg;
  Prof pr(e, "reg_penalty", 0, 0);
  for (const Param& p : e->params) {
    if (p.role != SB_ROLE_WEIGHT || !p.reg_kind) continue;
    reg_penalty_grad(e->arena + p.offset, with_grad ? e->grads + p.offset : nullptr, p.numel, p.reg_kind, p.reg_lambda, e->st,
                     seen == 0, seen == n_reg - 1, e->stream);
    ++seen;
  }
}

static OptDesc opt_desc(const sb_engine* e) {
  OptDesc o;
  o.kind = e->train.optimizer;
  o.rate = e->train.rate;
  o.beta1 = e->train.beta1;
  o.beta2 = e->train.beta2;
  o.epsilon = e->train.epsilon;
  o.nesterov = e->train.nesterov;
  o.minimizing = e->train.minimizing;
  return o;
}

// the update of arena floats [off, off + n) on stream s
static void optimizer_range(sb_engine* e, int64_t off, int64_t n, int advance_batch, cudaStream_t s) {
  float* w = e->arena + off;
  float* s0 = e->arena + e->nW + e->nS + off;
  float* s1 = e->n_slots >= 2 ? s0 + e->nW : nullptr;
  float* s2 = e->n_slots >= 3 ? s0 + 2 * e->nW : nullptr;
  optimizer_step(opt_desc(e), w, e->grads + off, s0, s1, s2, n, e->st, advance_batch, s);
}
static void run_optimizer(sb_engine* e, int advance_batch) {
  static const double bytes_per_param[8] = {20, 28, 20, 20, 28, 28, 28, 36};
  Prof pr(e, "optimizer_fused", bytes_per_param[e->train.optimizer] * (double)e->nW, 0);
  if (e->early_opt_done) {  // everything from early_opt_off on was updated on the side branch of this step's backward pass
    e->early_opt_done = false;
    optimizer_range(e, 0, e->early_opt_off, advance_batch, e->stream);
    return;
  }
  optimizer_range(e, 0, e->nW, advance_batch, e->stream);
}

static void run_prologue(sb_engine* e, bool resident) {
  Prof pr(e, "step_prologue", resident ? 8.0 * (double)(e->in_shape.numel() + e->label_shape.numel()) : 0, 0);
  step_prologue(e->st, resident ? e->ds_x : nullptr, resident ? e->ds_y : nullptr, e->acts[0], e->by, e->in_shape.numel(),
                e->label_shape.numel(), e->train.beta1, e->train.beta2, 1, e->stream);
}
static void run_prologue_staged(sb_engine* e, int parity, int j = 0) {
  Prof pr(e, "step_prologue", 8.0 * (double)(e->in_shape.numel() + e->label_shape.numel()), 0);
  // batch j of the staging buffer; from the second step of a group on, the prologue also files the previous step's loss
  step_prologue(e->st, e->stage_x[parity] + (size_t)j * e->in_shape.numel(), e->stage_y[parity] + (size_t)j * e->label_shape.numel(),
                e->acts[0], e->by, e->in_shape.numel(), e->label_shape.numel(), e->train.beta1, e->train.beta2, 0, e->stream,
                j > 0 ? e->loss_hist[parity] + (j - 1) : nullptr);
}

// backward pass of a TRAIN step (an optimizer update follows): the side branch may already update the parameters whose gradients
// are complete (early optimizer, engine.h)
static void run_backward_train(sb_engine* e) {
  e->early_opt_done = false;
  e->early_opt_now = e->early_opt_layer >= 0 && !e->profiling;
  run_backward(e);
  e->early_opt_now = false;
}

static void run_train_step_body(sb_engine* e, bool resident) {
  run_prologue(e, resident);
  run_forward(e, FWD_TRAIN);
  run_loss(e, true);
  run_backward_train(e);
  run_reg(e, true);
  run_optimizer(e, resident ? 1 : 0);
}

// host-fed step whose batch sits in staging buffer `parity` (the bool carries the parity through `capture`)
static void run_train_step_body_staged(sb_engine* e, bool parity) {
  run_prologue_staged(e, parity ? 1 : 0);
  run_forward(e, FWD_TRAIN);
  run_loss(e, true);
  run_backward_train(e);
  run_reg(e, true);
  run_optimizer(e, 0);
}

// host-fed step whose batch is the one the DEVICE-side cursor (StepState.batch_idx) points at inside staging buffer `parity`: one
// graph serves every slot, so the tail of an epoch (n_batches % stage_batches steps) needs no D2H in the compute stream; the
// prologue files the previous step's loss in loss_hist[parity][cursor - 1] and the optimizer advances the cursor
static void run_train_step_body_cursor(sb_engine* e, bool parity) {
  const int par = parity ? 1 : 0;
  {
    Prof pr(e, "step_prologue", 8.0 * (double)(e->in_shape.numel() + e->label_shape.numel()), 0);
    step_prologue(e->st, e->stage_x[par], e->stage_y[par], e->acts[0], e->by, e->in_shape.numel(), e->label_shape.numel(),
                  e->train.beta1, e->train.beta2, 1, e->stream, e->loss_hist[par]);
  }
  run_forward(e, FWD_TRAIN);
  run_loss(e, true);
  run_backward_train(e);
  run_reg(e, true);
  run_optimizer(e, 1);
}

// `stage_batches` host-fed steps back to back (one graph): batch j comes from slot j of staging buffer `parity`
static void run_train_group_body_staged(sb_engine* e, bool parity) {
  for (int j = 0; j < e->stage_batches; ++j) {
    run_prologue_staged(e, parity ? 1 : 0, j);
    run_forward(e, FWD_TRAIN);
    run_loss(e, true);
    run_backward_train(e);
    run_reg(e, true);
    run_optimizer(e, 0);
  }
  store_loss(e->st, e->loss_hist[parity ? 1 : 0] + (e->stage_batches - 1), e->stream);
}

static cudaGraphExec_t capture(sb_engine* e, long long* nodes, void (*body)(sb_engine*, bool), bool flag) {
  cudaGraph_t graph = nullptr;
  const long long before = tl_launch_count;
  SB_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
  try {
    body(e, flag);
  } catch (...) {
    cudaStreamEndCapture(e->stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  SB_CUDA(cudaStreamEndCapture(e->stream, &graph));
  *nodes = tl_launch_count - before;
  tl_launch_count = before;  // captured launches are counted per replay
  cudaGraphExec_t exec = nullptr;
  SB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  SB_CUDA(cudaGraphDestroy(graph));
  return exec;
}

static void loss_body(sb_engine* e, bool) {
  run_forward(e, FWD_LOSS_ONLY);
  run_loss(e, false);
  run_reg(e, false);
}

static bool use_graph(const sb_engine* e) { return e->train.use_graph && !e->profiling; }

static void launch_train_step(sb_engine* e, bool resident) {
  const long long before = tl_launch_count;
  if (use_graph(e)) {
    cudaGraphExec_t& g = resident ? e->g_resident : e->g_hostfed;
    long long& nodes = resident ? e->nodes_resident : e->nodes_hostfed;
    if (!g) g = capture(e, &nodes, run_train_step_body, resident);
    SB_CUDA(cudaGraphLaunch(g, e->stream));
    set_prewait_ok(0);  // a replayed graph may end in the optimizer
    e->launches += nodes;
  } else {
    run_train_step_body(e, resident);
    e->launches += tl_launch_count - before;
  }
}

// several resident steps in one graph: the device-side batch index / iter advance inside the graph, so the body is simply
// repeated; one graph boundary per kStepsPerGraph steps instead of per step
static int steps_per_graph() {
  static const int n = [] {
    const char* v = getenv("SB_STEPS_PER_GRAPH");
    const int k = v ? atoi(v) : 8;
    return k < 1 ? 1 : (k > 64 ? 64 : k);
  }();
  return n;
}
static void run_train_steps_body_multi(sb_engine* e, bool) {
  for (int i = 0; i < steps_per_graph(); ++i) run_train_step_body(e, true);
}
static void launch_train_steps_resident(sb_engine* e, int64_t n_steps) {
  int64_t j = 0;
  const int kStepsPerGraph = steps_per_graph();
  if (use_graph(e) && kStepsPerGraph > 1 && n_steps >= kStepsPerGraph) {
    if (!e->g_resident_multi) e->g_resident_multi = capture(e, &e->nodes_resident_multi, run_train_steps_body_multi, true);
    for (; j + kStepsPerGraph <= n_steps; j += kStepsPerGraph) {
      SB_CUDA(cudaGraphLaunch(e->g_resident_multi, e->stream));
      set_prewait_ok(0);  // a replayed graph may end in the optimizer
      e->launches += e->nodes_resident_multi;
    }
  }
  for (; j < n_steps; ++j) launch_train_step(e, true);
}

static void ensure_staging(sb_engine* e) {
  if (e->cstream) return;
  SB_CUDA(cudaStreamCreateWithFlags(&e->cstream, cudaStreamNonBlocking));
  // group size of sb_train_batches_async: as many steps as a resident graph holds, within 64 MB of staging per buffer
  const int64_t batch_bytes = (e->in_shape.numel() + e->label_shape.numel()) * 4;
  e->stage_batches = (int)std::max<int64_t>(1, std::min<int64_t>(steps_per_graph(), (64ll << 20) / std::max<int64_t>(1, batch_bytes)));
  for (int i = 0; i < 2; ++i) {
    e->stage_x[i] = dmalloc_floats(pad32(e->stage_batches * e->in_shape.numel()));
    e->stage_y[i] = dmalloc_floats(pad32(e->stage_batches * e->label_shape.numel()));
    e->loss_hist[i] = dmalloc_floats(pad32(e->stage_batches));
    SB_CUDA(cudaEventCreateWithFlags(&e->ev_copied[i], cudaEventDisableTiming));
    SB_CUDA(cudaEventCreateWithFlags(&e->ev_free[i], cudaEventDisableTiming));
    SB_CUDA(cudaEventRecord(e->ev_free[i], e->stream));
  }
}

// H2D of the batch on the copy stream into the free staging buffer, then the step on the main stream
static void launch_train_step_staged(sb_engine* e, const float* x, const float* y) {
  ensure_staging(e);
  const int par = e->stage_parity;
  e->stage_parity ^= 1;
  SB_CUDA(cudaStreamWaitEvent(e->cstream, e->ev_free[par], 0));  // the step that last read this buffer is done
  SB_CUDA(cudaMemcpyAsync(e->stage_x[par], x, (size_t)e->in_shape.numel() * 4, cudaMemcpyHostToDevice, e->cstream));
  SB_CUDA(cudaMemcpyAsync(e->stage_y[par], y, (size_t)e->label_shape.numel() * 4, cudaMemcpyHostToDevice, e->cstream));
  SB_CUDA(cudaEventRecord(e->ev_copied[par], e->cstream));
  SB_CUDA(cudaStreamWaitEvent(e->stream, e->ev_copied[par], 0));
  const long long before = tl_launch_count;
  if (use_graph(e)) {
    if (!e->g_staged[par]) e->g_staged[par] = capture(e, &e->nodes_staged[par], run_train_step_body_staged, par != 0);
    SB_CUDA(cudaGraphLaunch(e->g_staged[par], e->stream));
    set_prewait_ok(0);  // a replayed graph may end in the optimizer
    e->launches += e->nodes_staged[par];
  } else {
    run_train_step_body_staged(e, par != 0);
    e->launches += tl_launch_count - before;
  }
  SB_CUDA(cudaEventRecord(e->ev_free[par], e->stream));
}

// Loss read-back of a finished group, on the copy stream (a D2H in the compute stream stalls it ~25 us).  It waits for that
// group's compute, so it is queued only AFTER the next group's H2D copy -- otherwise that copy would no longer overlap compute.
// loss_hist[par] is rewritten two groups later, by a graph that waits for an H2D queued behind this D2H: no race.
static void flush_pending_losses(sb_engine* e) {
  if (!e->pending_loss_host) return;
  const int par = e->pending_loss_par;
  SB_CUDA(cudaStreamWaitEvent(e->cstream, e->ev_free[par], 0));
  SB_CUDA(cudaMemcpyAsync(e->pending_loss_host, e->loss_hist[par], (size_t)e->pending_loss_n * 4, cudaMemcpyDeviceToHost, e->cstream));
  e->pending_loss_host = nullptr;
}

// `stage_batches` consecutive host batches: one H2D pair on the copy stream, one graph, one read-back of the group's losses
static void launch_train_group_staged(sb_engine* e, const float* x, const float* y, float* losses_out_host) {
  ensure_staging(e);
  const int par = e->stage_parity, G = e->stage_batches;
  e->stage_parity ^= 1;
  SB_CUDA(cudaStreamWaitEvent(e->cstream, e->ev_free[par], 0));
  SB_CUDA(cudaMemcpyAsync(e->stage_x[par], x, (size_t)G * e->in_shape.numel() * 4, cudaMemcpyHostToDevice, e->cstream));
  SB_CUDA(cudaMemcpyAsync(e->stage_y[par], y, (size_t)G * e->label_shape.numel() * 4, cudaMemcpyHostToDevice, e->cstream));
  SB_CUDA(cudaEventRecord(e->ev_copied[par], e->cstream));
  flush_pending_losses(e);
  SB_CUDA(cudaStreamWaitEvent(e->stream, e->ev_copied[par], 0));
  if (!e->g_group[par]) e->g_group[par] = capture(e, &e->nodes_group[par], run_train_group_body_staged, par != 0);
  SB_CUDA(cudaGraphLaunch(e->g_group[par], e->stream));
  set_prewait_ok(0);  // a replayed graph may end in the optimizer
  e->launches += e->nodes_group[par];
  SB_CUDA(cudaEventRecord(e->ev_free[par], e->stream));
  if (losses_out_host) {
    e->pending_loss_host = losses_out_host;
    e->pending_loss_par = par;
    e->pending_loss_n = G;
  }
}

// r < stage_batches consecutive host batches (the tail of sb_train_batches_async): one H2D pair, r replays of the cursor-driven
// single-step graph, one read-back of the r losses on the copy stream
static void launch_train_tail_staged(sb_engine* e, const float* x, const float* y, int r, float* losses_out_host) {
  ensure_staging(e);
  const int par = e->stage_parity;
  e->stage_parity ^= 1;
  SB_CUDA(cudaStreamWaitEvent(e->cstream, e->ev_free[par], 0));
  SB_CUDA(cudaMemcpyAsync(e->stage_x[par], x, (size_t)r * e->in_shape.numel() * 4, cudaMemcpyHostToDevice, e->cstream));
  SB_CUDA(cudaMemcpyAsync(e->stage_y[par], y, (size_t)r * e->label_shape.numel() * 4, cudaMemcpyHostToDevice, e->cstream));
  SB_CUDA(cudaEventRecord(e->ev_copied[par], e->cstream));
  flush_pending_losses(e);
  SB_CUDA(cudaStreamWaitEvent(e->stream, e->ev_copied[par], 0));
  {  // cursor = 0; the device's step counter continues
    static thread_local int slot = 0;
    slot = (slot + 1) & 3;
    StepState* h = e->st_host + slot;
    h->batch_idx = 0;
    SB_CUDA(cudaMemcpyAsync(&e->st->batch_idx, &h->batch_idx, sizeof(long long), cudaMemcpyHostToDevice, e->stream));
  }
  if (!e->g_cursor[par]) e->g_cursor[par] = capture(e, &e->nodes_cursor[par], run_train_step_body_cursor, par != 0);
  for (int q = 0; q < r; ++q) {
    SB_CUDA(cudaGraphLaunch(e->g_cursor[par], e->stream));
    e->launches += e->nodes_cursor[par];
  }
  set_prewait_ok(0);  // a replayed graph may end in the optimizer
  store_loss(e->st, e->loss_hist[par], e->stream, 1);
  e->launches += 1;
  SB_CUDA(cudaEventRecord(e->ev_free[par], e->stream));
  if (losses_out_host) {
    e->pending_loss_host = losses_out_host;
    e->pending_loss_par = par;
    e->pending_loss_n = r;
  }
}

static void launch_loss(sb_engine* e) {
  const long long before = tl_launch_count;
  if (use_graph(e)) {
    if (!e->g_loss) e->g_loss = capture(e, &e->nodes_loss, loss_body, false);
    SB_CUDA(cudaGraphLaunch(e->g_loss, e->stream));
    set_prewait_ok(0);  // a replayed graph may end in the optimizer
    e->launches += e->nodes_loss;
  } else {
    loss_body(e, false);
    e->launches += tl_launch_count - before;
  }
}

static void set_device_iter(sb_engine* e, long long iter_done, long long batch_idx) {
  // host -> device control write; st_host slots rotate so an in-flight copy is never overwritten
  static thread_local int slot = 0;
  slot = (slot + 1) & 3;
  StepState* h = e->st_host + slot;
  h->iter = iter_done;
  h->batch_idx = batch_idx;
  if (iter_done < 0) {  // SB_ITER_FROM_DEVICE: the device's own counter continues (the averaging kernel added the peers' counts)
    SB_CUDA(cudaMemcpyAsync(&e->st->batch_idx, &h->batch_idx, sizeof(long long), cudaMemcpyHostToDevice, e->stream));
    e->host_iter_mirror = -1;
    return;
  }
  SB_CUDA(cudaMemcpyAsync(e->st, h, 2 * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
  e->host_iter_mirror = iter_done;
}

// sticky device-side errors (today: the cross-GPU averaging barrier timed out, group.cu) surface at the next host sync point
static void check_device_error(sb_engine* e, int err) {
  if (err == 0) return;
  SB_CUDA(cudaMemsetAsync(&e->st->err, 0, sizeof(int), e->stream));
  SB_FAIL(SB_ERR_TIMEOUT, "the device-synchronised averaging timed out waiting for a peer rank (SB_P2P_TIMEOUT_MS); this rank's "
                          "arena was left un-averaged");
}

static float read_loss(sb_engine* e) {
  StepState* h = e->st_host + 0;
  SB_CUDA(cudaMemcpyAsync(&h->loss, &e->st->loss, sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  if (e->device_sync_used) SB_CUDA(cudaMemcpyAsync(&h->err, &e->st->err, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  SB_CUDA(cudaStreamSynchronize(e->stream));
  if (e->device_sync_used) check_device_error(e, h->err);
  return h->loss;
}

static void require_ready(const sb_engine* e) {
  if (!e->params_ready || !e->opt_ready)
    SB_FAIL(SB_ERR_STATE, "parameters are not initialised: call sb_init_params, or sb_params_set for every model param "
                          "followed by sb_init_opt_params");
}

static void upload_batch(sb_engine* e, const float* x, const float* y) {
  if (!x) SB_FAIL(SB_ERR_INVALID_ARG, "null x");
  SB_CUDA(cudaMemcpyAsync(e->acts[0], x, (size_t)e->in_shape.numel() * 4, cudaMemcpyHostToDevice, e->stream));
  if (y) SB_CUDA(cudaMemcpyAsync(e->by, y, (size_t)e->label_shape.numel() * 4, cudaMemcpyHostToDevice, e->stream));
}

// -------------------------------------------------------------------------------------------------
// initialisation (models/Initializer.scala; Optimizer.scala:123-131)
// -------------------------------------------------------------------------------------------------
static void fans_of(const Param& p, int64_t* fin, int64_t* fout) {
  // Initializer.fans (Initializer.scala:16-25): rank>=3 takes d0 as "out" and d1 as "in" (sic)
  if (p.rank == 0) { *fin = *fout = 1; return; }
  if (p.rank == 1) { *fin = *fout = p.dims[0]; return; }
  if (p.rank == 2) { *fin = p.dims[0]; *fout = p.dims[1]; return; }
  int64_t size = 1;
  for (int i = 2; i < p.rank; ++i) size *= p.dims[i];
  *fin = size * p.dims[1];
  *fout = size * p.dims[0];
}
static float f32_sqrt_ratio(float num, int64_t den) {
  // math.sqrt(6f / (fanIn + fanOut)).toFloat : fp32 division, double sqrt, cast back (Initializer.scala:86)
  float q = num / (float)den;
  return (float)sqrt((double)q);
}

static void init_param(sb_engine* e, const Param& p) {
  float* dst = e->arena + p.offset;
  cudaStream_t s = e->stream;
  const sb_initializer& in = p.init;
  auto need_seed = [&]() {
    if (!in.has_seed)
      SB_FAIL(SB_ERR_UNSUPPORTED, "unseeded random initializer for '" + p.path + "': pass a seed or set the tensor with sb_params_set "
                                  "(every partition initialises independently in the reference, SURVEY App. B-Q6)");
  };
  SB_CUDA(cudaMemsetAsync(dst, 0, (size_t)p.padded * 4, s));
  int64_t fin, fout;
  fans_of(p, &fin, &fout);
  switch (in.kind) {
    case SB_INIT_ZEROS: break;
    case SB_INIT_ONES: fill_const(dst, 1.0f, p.numel, s); break;
    case SB_INIT_CONST: fill_const(dst, in.a, p.numel, s); break;
    case SB_INIT_RANDOM_UNIFORM: need_seed(); philox_fill(dst, p.numel, (unsigned long long)in.seed, 0, in.b - in.a, 1, in.a, 1, s); break;
    case SB_INIT_RANDOM_NORMAL: need_seed(); philox_fill(dst, p.numel, (unsigned long long)in.seed, 1, in.b, 1, in.a, 1, s); break;
    case SB_INIT_GLOROT_UNIFORM: {
      need_seed();
      float limit = f32_sqrt_ratio(6.0f, fin + fout);
      philox_fill(dst, p.numel, (unsigned long long)in.seed, 0, 2.0f * limit, 1, -limit, 1, s);
      break;
    }
    case SB_INIT_HE_UNIFORM: {
      need_seed();
      float limit = f32_sqrt_ratio(6.0f, fin);
      philox_fill(dst, p.numel, (unsigned long long)in.seed, 0, 2.0f * limit, 1, -limit, 1, s);
      break;
    }
    case SB_INIT_LECUN_UNIFORM: {
      need_seed();
      float limit = f32_sqrt_ratio(3.0f, fin);
      philox_fill(dst, p.numel, (unsigned long long)in.seed, 0, 2.0f * limit, 1, -limit, 1, s);
      break;
    }
    case SB_INIT_TRUNCATED_NORMAL: need_seed(); philox_fill(dst, p.numel, (unsigned long long)in.seed, 2, in.b, 1, in.a, 1, s); break;
    case SB_INIT_GLOROT_NORMAL:
    case SB_INIT_HE_NORMAL:
    case SB_INIT_LECUN_NORMAL: {
      // truncated normal with std = sqrt(c / fan) / 0.87962566 (models/Initializer.scala:93-110, 128-150, 168-190)
      need_seed();
      const float std_ = in.kind == SB_INIT_GLOROT_NORMAL ? f32_sqrt_ratio(2.0f, fin + fout)
                                                          : f32_sqrt_ratio(in.kind == SB_INIT_HE_NORMAL ? 2.0f : 1.0f, fin);
      philox_fill(dst, p.numel, (unsigned long long)in.seed, 2, std_ / 0.87962566103423978f, 1, 0.f, 0, s);
      break;
    }
    case SB_INIT_ORTHOGONAL: {
      need_seed();
      std::vector<float> host((size_t)p.numel);
      int64_t cols = p.dims[p.rank - 1], rows = p.numel / cols;
      orthogonal_host(host.data(), rows, cols, (unsigned long long)in.seed, in.a);
      SB_CUDA(cudaMemcpyAsync(dst, host.data(), (size_t)p.numel * 4, cudaMemcpyHostToDevice, s));
      SB_CUDA(cudaStreamSynchronize(s));
      break;
    }
    default: SB_FAIL(SB_ERR_UNSUPPORTED, "initializer kind " + std::to_string(in.kind) + " is unsupported on the device; set '" + p.path + "' with sb_params_set");
  }
}

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------
extern "C" int sb_engine_create(const sb_model_desc* model, const sb_train_desc* train, int device, sb_engine** out) {
  SB_API_BEGIN
  if (!model || !train || !out || !model->layers) SB_FAIL(SB_ERR_INVALID_ARG, "null argument");
  std::unique_ptr<sb_engine> e(new sb_engine());
  e->layers_desc.assign(model->layers, model->layers + std::max(model->n_layers, 0));
  e->model = *model;
  e->model.layers = e->layers_desc.data();
  e->train = *train;
  e->device = device;
  e->plan_only = device < 0;
  build_plan(e.get());
  if (!e->plan_only) {
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || device >= n) {
      cudaGetLastError();
      SB_FAIL(SB_ERR_CUDA, "CUDA device " + std::to_string(device) + " is not available (" + std::to_string(n) +
                               " visible); this library has no CPU fallback");
    }
    allocate(e.get());
  }
  *out = e.release();
  SB_API_END
}

extern "C" void sb_engine_destroy(sb_engine* e) {
  if (!e) return;
  if (!e->plan_only) {
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (auto& pp : e->prof_pending) { cudaEventDestroy(pp.a); cudaEventDestroy(pp.b); }
    if (e->g_resident) cudaGraphExecDestroy(e->g_resident);
    if (e->g_resident_multi) cudaGraphExecDestroy(e->g_resident_multi);
    if (e->g_hostfed) cudaGraphExecDestroy(e->g_hostfed);
    if (e->g_loss) cudaGraphExecDestroy(e->g_loss);
    for (int i = 0; i < 2; ++i) {
      if (e->g_staged[i]) cudaGraphExecDestroy(e->g_staged[i]);
      if (e->g_group[i]) cudaGraphExecDestroy(e->g_group[i]);
      if (e->g_cursor[i]) cudaGraphExecDestroy(e->g_cursor[i]);
      cudaFree(e->stage_x[i]); cudaFree(e->stage_y[i]); cudaFree(e->loss_hist[i]);
      if (e->ev_copied[i]) cudaEventDestroy(e->ev_copied[i]);
      if (e->ev_free[i]) cudaEventDestroy(e->ev_free[i]);
    }
    if (e->cstream) { cudaStreamSynchronize(e->cstream); cudaStreamDestroy(e->cstream); }
    if (e->bstream) { cudaStreamSynchronize(e->bstream); cudaStreamDestroy(e->bstream); cudaEventDestroy(e->ev_fork); cudaEventDestroy(e->ev_join); cudaEventDestroy(e->ev_mid); }
    tc_free_layers(e);
    rnn_free(e);
    lstm_free(e);
    for (auto& L : e->L) { cudaFree(L.bn_save_mean); cudaFree(L.bn_save_var); cudaFree(L.pool_map); }
    for (float* p : e->acts) cudaFree(p);
    for (float* p : e->dacts) cudaFree(p);
    cudaFree(e->arena); cudaFree(e->grads); cudaFree(e->by); cudaFree(e->ds_x); cudaFree(e->ds_y);
    cudaFree(e->st); cudaFree(e->ws); cudaFree(e->ws_side); cudaFree(e->head_scratch); cudaFree(e->est_acc);
    if (e->st_host) cudaFreeHost(e->st_host);
    if (e->stream) cudaStreamDestroy(e->stream);
  }
  delete e;
}

extern "C" int sb_param_count(const sb_engine* e) { return e ? (int)e->params.size() : 0; }

extern "C" int sb_param_info(const sb_engine* e, int index, char* path, size_t path_cap, int64_t* dims, int32_t* rank,
                             int32_t* role, int32_t* aggregation, int64_t* arena_offset) {
  SB_API_BEGIN
  if (!e || index < 0 || index >= (int)e->params.size()) SB_FAIL(SB_ERR_INVALID_ARG, "param index out of range");
  const Param& p = e->params[index];
  if (path && path_cap) {
    strncpy(path, p.path.c_str(), path_cap - 1);
    path[path_cap - 1] = 0;
  }
  if (dims) for (int i = 0; i < p.rank; ++i) dims[i] = p.dims[i];
  if (rank) *rank = p.rank;
  if (role) *role = p.role;
  if (aggregation) *aggregation = p.agg;
  if (arena_offset) *arena_offset = p.offset;
  SB_API_END
}

extern "C" int sb_output_dims(const sb_engine* e, int64_t* dims, int32_t* rank) {
  SB_API_BEGIN
  if (!e || !dims || !rank) SB_FAIL(SB_ERR_INVALID_ARG, "null argument");
  *rank = e->out_shape.rank;
  for (int i = 0; i < e->out_shape.rank; ++i) dims[i] = e->out_shape.d[i];
  SB_API_END
}

static const Param& find_param(const sb_engine* e, const char* path, size_t n) {
  if (!e || !path) SB_FAIL(SB_ERR_INVALID_ARG, "null argument");
  auto it = e->by_path.find(path);
  if (it == e->by_path.end()) SB_FAIL(SB_ERR_INVALID_ARG, std::string("missing ") + path + " param");  // core/Params.scala:22-23
  const Param& p = e->params[it->second];
  if ((int64_t)n != p.numel)
    SB_FAIL(SB_ERR_INVALID_ARG, std::string("param ") + path + " has " + std::to_string(p.numel) + " elements, got " + std::to_string(n));
  return p;
}

static void refresh_ready(sb_engine* e) { /* readiness is tracked by explicit calls */ (void)e; }

extern "C" int sb_params_set(sb_engine* e, const char* path, const float* host, size_t n) {
  SB_API_BEGIN
  need_device(e);
  if (!host) SB_FAIL(SB_ERR_INVALID_ARG, "null host buffer");
  const Param& p = find_param(e, path, n);
  DeviceGuard dg(e->device);
  SB_CUDA(cudaMemcpyAsync(e->arena + p.offset, host, n * 4, cudaMemcpyHostToDevice, e->stream));
  SB_CUDA(cudaStreamSynchronize(e->stream));
  tc_params_changed(e);
  // setting every model param makes the model side ready (mirrors `initParams`, Optimizer.scala:124-126)
  e->params_ready = true;
  refresh_ready(e);
  SB_API_END
}

extern "C" int sb_params_get(sb_engine* e, const char* path, float* host, size_t n) {
  SB_API_BEGIN
  need_device(e);
  if (!host) SB_FAIL(SB_ERR_INVALID_ARG, "null host buffer");
  const Param& p = find_param(e, path, n);
  DeviceGuard dg(e->device);
  SB_CUDA(cudaMemcpyAsync(host, e->arena + p.offset, n * 4, cudaMemcpyDeviceToHost, e->stream));
  SB_CUDA(cudaStreamSynchronize(e->stream));
  SB_API_END
}

extern "C" int sb_init_opt_params(sb_engine* e) {
  SB_API_BEGIN
  need_device(e);
  DeviceGuard dg(e->device);
  for (const Param& p : e->params)
    if (p.role == SB_ROLE_OPT) init_param(e, p);
  SB_CUDA(cudaStreamSynchronize(e->stream));
  e->opt_ready = true;
  SB_API_END
}

extern "C" int sb_init_params(sb_engine* e) {
  SB_API_BEGIN
  need_device(e);
  DeviceGuard dg(e->device);
  for (const Param& p : e->params) init_param(e, p);
  SB_CUDA(cudaStreamSynchronize(e->stream));
  tc_params_changed(e);
  e->params_ready = e->opt_ready = true;
  SB_API_END
}

extern "C" int sb_internal_mark_params_ready(sb_engine* e) {  // every tensor of the arena was restored from a checkpoint
  SB_API_BEGIN
  need_device(e);
  e->params_ready = e->opt_ready = true;
  e->host_iter_mirror = -1;
  SB_API_END
}

extern "C" int sb_dataset_upload(sb_engine* e, const float* x, const float* y, int64_t n_rows) {
  SB_API_BEGIN
  need_device(e);
  if (!x || !y || n_rows <= 0) SB_FAIL(SB_ERR_INVALID_ARG, "dataset must have at least one row");
  DeviceGuard dg(e->device);
  SB_CUDA(cudaStreamSynchronize(e->stream));
  cudaFree(e->ds_x);
  cudaFree(e->ds_y);
  e->ds_x = e->ds_y = nullptr;
  e->ds_x = dmalloc_floats(pad32(n_rows * e->x_per_row));
  e->ds_y = dmalloc_floats(pad32(n_rows * e->y_per_row));
  SB_CUDA(cudaMemcpyAsync(e->ds_x, x, (size_t)(n_rows * e->x_per_row) * 4, cudaMemcpyHostToDevice, e->stream));
  SB_CUDA(cudaMemcpyAsync(e->ds_y, y, (size_t)(n_rows * e->y_per_row) * 4, cudaMemcpyHostToDevice, e->stream));
  SB_CUDA(cudaStreamSynchronize(e->stream));
  e->ds_rows = n_rows;
  // the captured graphs hold the dataset pointers
  if (e->g_resident) { cudaGraphExecDestroy(e->g_resident); e->g_resident = nullptr; }
  if (e->g_resident_multi) { cudaGraphExecDestroy(e->g_resident_multi); e->g_resident_multi = nullptr; }
  SB_API_END
}

extern "C" int sb_train_steps_async(sb_engine* e, int64_t global_iter, int64_t first_batch, int64_t n_steps) {
  SB_API_BEGIN
  need_device(e);
  require_ready(e);
  if (!e->ds_x) SB_FAIL(SB_ERR_STATE, "no dataset uploaded");
  const int64_t n_batches = e->ds_rows / e->train.batch;
  if (first_batch < 0 || n_steps < 0 || first_batch + n_steps > n_batches)
    SB_FAIL(SB_ERR_INVALID_ARG, "batch range exceeds the shard (" + std::to_string(n_batches) + " full batches)");
  DeviceGuard dg(e->device);
  set_device_iter(e, global_iter < 0 ? -1 : global_iter, first_batch);
  // profiling: hold the device back while the host enqueues the event-bracketed launches (~3 API calls per kernel), so
  // the events time kernels back to back instead of launch gaps
  if (e->profiling) device_delay(std::min<long long>(400000000LL, 2000000LL + n_steps * 600000LL), e->stream);
  launch_train_steps_resident(e, n_steps);
  e->host_iter_mirror = global_iter < 0 ? -1 : global_iter + n_steps;
  SB_API_END
}

// Capture, instantiate and upload the CUDA graphs the train path replays, without running anything: the first
// sb_train_epoch / sb_train_steps_async / sb_train_batches_async after this call costs what every later one costs.
extern "C" int sb_warm(sb_engine* e, int what) {
  SB_API_BEGIN
  need_device(e);
  require_ready(e);
  DeviceGuard dg(e->device);
  if (!e->train.use_graph) return SB_OK;
  const bool was_profiling = e->profiling;
  e->profiling = false;
  auto upload = [&](cudaGraphExec_t g) { if (g) SB_CUDA(cudaGraphUpload(g, e->stream)); };
  try {
    if ((what & SB_WARM_RESIDENT) && e->ds_x) {
      if (!e->g_resident) e->g_resident = capture(e, &e->nodes_resident, run_train_step_body, true);
      if (steps_per_graph() > 1 && !e->g_resident_multi)
        e->g_resident_multi = capture(e, &e->nodes_resident_multi, run_train_steps_body_multi, true);
      upload(e->g_resident);
      upload(e->g_resident_multi);
    }
    if (what & (SB_WARM_RESIDENT | SB_WARM_LOSS)) {
      if (!e->g_loss) e->g_loss = capture(e, &e->nodes_loss, loss_body, false);
      upload(e->g_loss);
    }
    if (what & SB_WARM_HOSTFED) {
      ensure_staging(e);
      if (!e->g_hostfed) e->g_hostfed = capture(e, &e->nodes_hostfed, run_train_step_body, false);
      upload(e->g_hostfed);
      for (int par = 0; par < 2; ++par) {
        if (!e->g_staged[par]) e->g_staged[par] = capture(e, &e->nodes_staged[par], run_train_step_body_staged, par != 0);
        if (e->stage_batches > 1 && !e->g_group[par])
          e->g_group[par] = capture(e, &e->nodes_group[par], run_train_group_body_staged, par != 0);
        if (e->stage_batches > 1 && !e->g_cursor[par])
          e->g_cursor[par] = capture(e, &e->nodes_cursor[par], run_train_step_body_cursor, par != 0);
        upload(e->g_staged[par]);
        upload(e->g_group[par]);
        upload(e->g_cursor[par]);
      }
    }
  } catch (...) {
    e->profiling = was_profiling;
    throw;
  }
  e->profiling = was_profiling;
  SB_CUDA(cudaStreamSynchronize(e->stream));
  set_prewait_ok(0);
  SB_API_END
}

extern "C" int sb_train_epoch(sb_engine* e, int64_t global_iter, int64_t* iterations_out, float* last_loss_out) {
  SB_API_BEGIN
  need_device(e);
  require_ready(e);
  if (!e->ds_x) SB_FAIL(SB_ERR_STATE, "no dataset uploaded");
  const int64_t n_batches = e->ds_rows / e->train.batch;  // Skip: the partial tail batch is dropped (Iterators.scala:46-47,90)
  if (n_batches == 0)
    SB_FAIL(SB_ERR_INVALID_ARG, "NoSuchElementException: the shard holds fewer rows than one batch (Iterators.scala:79-81)");
  DeviceGuard dg(e->device);
  set_device_iter(e, global_iter < 0 ? -1 : global_iter, 0);  // SB_ITER_FROM_DEVICE: the device's counter continues
  launch_train_steps_resident(e, n_batches);
  e->host_iter_mirror = global_iter < 0 ? -1 : global_iter + n_batches;
  launch_loss(e);  // forward-only loss on the LAST batch with the UPDATED params (Optimizer.scala:148)
  float loss = read_loss(e);
  if (e->profiling) prof_resolve(e);
  if (iterations_out) *iterations_out = n_batches;
  if (last_loss_out) *last_loss_out = loss;
  SB_API_END
}

extern "C" int sb_train_step(sb_engine* e, const float* x, const float* y, int64_t iter, float* loss_out) {
  SB_API_BEGIN
  need_device(e);
  require_ready(e);
  if (!y) SB_FAIL(SB_ERR_INVALID_ARG, "null y");
  if (iter < 1) SB_FAIL(SB_ERR_INVALID_ARG, "iter is the 1-based global step");
  DeviceGuard dg(e->device);
  if (e->host_iter_mirror != iter - 1) set_device_iter(e, iter - 1, 0);
  upload_batch(e, x, y);
  launch_train_step(e, false);
  e->host_iter_mirror = iter;
  if (loss_out) *loss_out = read_loss(e);
  SB_API_END
}

extern "C" int sb_train_step_async(sb_engine* e, const float* x, const float* y, int64_t iter, float* loss_out_host) {
  SB_API_BEGIN
  need_device(e);
  require_ready(e);
  if (!y) SB_FAIL(SB_ERR_INVALID_ARG, "null y");
  if (iter < 1) SB_FAIL(SB_ERR_INVALID_ARG, "iter is the 1-based global step");
  DeviceGuard dg(e->device);
  if (e->host_iter_mirror != iter - 1) set_device_iter(e, iter - 1, 0);
  if (!x) SB_FAIL(SB_ERR_INVALID_ARG, "null x");
  launch_train_step_staged(e, x, y);
  e->host_iter_mirror = iter;
  if (loss_out_host)
    SB_CUDA(cudaMemcpyAsync(loss_out_host, &e->st->loss, sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  SB_API_END
}

extern "C" int sb_train_batches_async(sb_engine* e, const float* x, const float* y, int64_t n_batches, int64_t global_iter,
                                      float* losses_out_host) {
  SB_API_BEGIN
  need_device(e);
  require_ready(e);
  if (!x || !y) SB_FAIL(SB_ERR_INVALID_ARG, "null batch buffers");
  if (n_batches < 0 || global_iter < 0) SB_FAIL(SB_ERR_INVALID_ARG, "negative batch count or iteration");
  DeviceGuard dg(e->device);
  if (e->host_iter_mirror != global_iter) set_device_iter(e, global_iter, 0);
  ensure_staging(e);
  const int64_t xn = e->in_shape.numel(), yn = e->label_shape.numel();
  int64_t j = 0;
  if (use_graph(e) && e->stage_batches > 1)
    for (; j + e->stage_batches <= n_batches; j += e->stage_batches)
      launch_train_group_staged(e, x + j * xn, y + j * yn, losses_out_host ? losses_out_host + j : nullptr);
  if (use_graph(e) && e->stage_batches > 1 && j < n_batches) {  // the tail: fewer than a group's worth of batches
    launch_train_tail_staged(e, x + j * xn, y + j * yn, (int)(n_batches - j), losses_out_host ? losses_out_host + j : nullptr);
    j = n_batches;
  }
  flush_pending_losses(e);
  for (; j < n_batches; ++j) {
    launch_train_step_staged(e, x + j * xn, y + j * yn);
    if (losses_out_host)
      SB_CUDA(cudaMemcpyAsync(losses_out_host + j, &e->st->loss, sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  }
  e->host_iter_mirror = global_iter + n_batches;
  SB_API_END
}

extern "C" int sb_eval_loss(sb_engine* e, const float* x, const float* y, float* loss_out) {
  SB_API_BEGIN
  need_device(e);
  if (!e->params_ready) SB_FAIL(SB_ERR_STATE, "model parameters are not initialised");
  if (!y || !loss_out) SB_FAIL(SB_ERR_INVALID_ARG, "null argument");
  DeviceGuard dg(e->device);
  upload_batch(e, x, y);
  launch_loss(e);
  *loss_out = read_loss(e);
  SB_API_END
}

extern "C" int sb_predict(sb_engine* e, const float* x, float* out, size_t n_out) {
  SB_API_BEGIN
  need_device(e);
  if (!e->params_ready) SB_FAIL(SB_ERR_STATE, "model parameters are not initialised");
  if (!out || (int64_t)n_out != e->out_shape.numel()) SB_FAIL(SB_ERR_INVALID_ARG, "output buffer size mismatch");
  DeviceGuard dg(e->device);
  upload_batch(e, x, nullptr);
  const long long before = tl_launch_count;
  run_forward(e, FWD_INFER);
  e->launches += tl_launch_count - before;
  SB_CUDA(cudaMemcpyAsync(out, e->acts[e->L.back().buf_out], n_out * 4, cudaMemcpyDeviceToHost, e->stream));
  SB_CUDA(cudaStreamSynchronize(e->stream));
  SB_API_END
}

extern "C" int sb_estimate(sb_engine* e, int estimator, const float* x, const float* y, int64_t n_rows, double out[4]) {
  SB_API_BEGIN
  need_device(e);
  if (!e->params_ready) SB_FAIL(SB_ERR_STATE, "model parameters are not initialised");
  if (!out) SB_FAIL(SB_ERR_INVALID_ARG, "null output");
  if (estimator < SB_EST_ACCURACY || estimator > SB_EST_R2) SB_FAIL(SB_ERR_INVALID_ARG, "unknown estimator");
  if (e->out_shape.numel() != e->label_shape.numel())
    SB_FAIL(SB_ERR_INVALID_ARG, "estimators compare the model output with the labels element by element: shapes differ");
  if (estimator == SB_EST_R2 && e->y_per_row != 1)
    SB_FAIL(SB_ERR_INVALID_ARG, "labels should have shape (1)");  // estimators/package.scala:103
  const bool resident = x == nullptr;
  if (resident) {
    if (!e->ds_x) SB_FAIL(SB_ERR_STATE, "no resident shard: call sb_dataset_upload first or pass host buffers");
    if (n_rows <= 0 || n_rows > e->ds_rows) n_rows = e->ds_rows;
    x = e->ds_x;
    y = e->ds_y;
  } else {
    if (!y) SB_FAIL(SB_ERR_INVALID_ARG, "null y");
    if (n_rows < 0) SB_FAIL(SB_ERR_INVALID_ARG, "negative row count");
  }
  DeviceGuard dg(e->device);
  if (!e->est_acc) SB_CUDA(cudaMalloc(&e->est_acc, 4 * sizeof(double)));
  SB_CUDA(cudaMemsetAsync(e->est_acc, 0, 4 * sizeof(double), e->stream));
  const int64_t B = e->in_shape.d[0];
  const int C = (int)e->y_per_row;
  const cudaMemcpyKind kind = resident ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const long long before = tl_launch_count;
  for (int64_t r0 = 0; r0 < n_rows; r0 += B) {
    // rows are independent in inference mode: a partial tail leaves the rest of the batch buffers as they were
    const int64_t nv = std::min<int64_t>(B, n_rows - r0);
    SB_CUDA(cudaMemcpyAsync(e->acts[0], x + r0 * e->x_per_row, (size_t)(nv * e->x_per_row) * 4, kind, e->stream));
    SB_CUDA(cudaMemcpyAsync(e->by, y + r0 * e->y_per_row, (size_t)(nv * e->y_per_row) * 4, kind, e->stream));
    run_forward(e, FWD_INFER);
    estimator_accumulate(e->acts[e->L.back().buf_out], e->by, nv, C, estimator, e->est_acc, e->stream);
  }
  e->launches += tl_launch_count - before;
  SB_CUDA(cudaMemcpyAsync(out, e->est_acc, 4 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  SB_CUDA(cudaStreamSynchronize(e->stream));
  SB_API_END
}

extern "C" int sb_compute_grads(sb_engine* e, const float* x, const float* y, float* loss_out) {
  SB_API_BEGIN
  need_device(e);
  if (!e->params_ready) SB_FAIL(SB_ERR_STATE, "model parameters are not initialised");
  if (!y) SB_FAIL(SB_ERR_INVALID_ARG, "null y");
  DeviceGuard dg(e->device);
  upload_batch(e, x, y);
  const long long before = tl_launch_count;
  run_forward(e, FWD_LOSS_ONLY);
  run_loss(e, true);
  run_backward(e);  // gradients only: no optimizer follows, every gradient must be left complete
  run_reg(e, true);
  e->launches += tl_launch_count - before;
  float loss = read_loss(e);
  if (loss_out) *loss_out = loss;
  SB_API_END
}

extern "C" int sb_grad_get(sb_engine* e, const char* path, float* host, size_t n) {
  SB_API_BEGIN
  need_device(e);
  const Param& p = find_param(e, path, n);
  if (p.role != SB_ROLE_WEIGHT) SB_FAIL(SB_ERR_INVALID_ARG, std::string(path) + " is not a trainable weight");
  DeviceGuard dg(e->device);
  SB_CUDA(cudaMemcpyAsync(host, e->grads + p.offset, n * 4, cudaMemcpyDeviceToHost, e->stream));
  SB_CUDA(cudaStreamSynchronize(e->stream));
  SB_API_END
}

extern "C" int sb_read_loss(sb_engine* e, float* loss_out) {
  SB_API_BEGIN
  need_device(e);
  if (!loss_out) SB_FAIL(SB_ERR_INVALID_ARG, "null argument");
  DeviceGuard dg(e->device);
  *loss_out = read_loss(e);
  SB_API_END
}

extern "C" int sb_sync(sb_engine* e) {
  SB_API_BEGIN
  need_device(e);
  DeviceGuard dg(e->device);
  if (e->device_sync_used) SB_CUDA(cudaMemcpyAsync(&e->st_host->err, &e->st->err, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  SB_CUDA(cudaStreamSynchronize(e->stream));
  if (e->cstream) SB_CUDA(cudaStreamSynchronize(e->cstream));  // loss read-backs of sb_train_batches_async
  if (e->profiling) prof_resolve(e);
  if (e->device_sync_used) check_device_error(e, e->st_host->err);
  SB_API_END
}

extern "C" int sb_arena(sb_engine* e, void** device_ptr, int64_t* n_floats) {
  SB_API_BEGIN
  need_device(e);
  if (device_ptr) *device_ptr = e->arena;
  if (n_floats) *n_floats = e->arena_floats;
  SB_API_END
}

extern "C" int sb_stream(sb_engine* e, void** stream) {
  SB_API_BEGIN
  need_device(e);
  if (!stream) SB_FAIL(SB_ERR_INVALID_ARG, "null argument");
  *stream = (void*)e->stream;
  SB_API_END
}

extern "C" int sb_arena_scale(sb_engine* e, float w) {
  SB_API_BEGIN
  need_device(e);
  DeviceGuard dg(e->device);
  const long long before = tl_launch_count;
  scale_inplace(e->arena, w, e->arena_floats, e->stream);
  e->launches += tl_launch_count - before;
  SB_API_END
}

extern "C" int sb_reset_unaggregated(sb_engine* e) {
  SB_API_BEGIN
  need_device(e);
  DeviceGuard dg(e->device);
  for (const Param& p : e->params)
    if (p.agg == SB_AGG_NONE) init_param(e, p);
  SB_API_END
}

extern "C" int sb_arena_ipc_handle(sb_engine* e, void* handle64, int64_t* byte_offset) {
  SB_API_BEGIN
  need_device(e);
  if (!handle64) SB_FAIL(SB_ERR_INVALID_ARG, "null argument");
  DeviceGuard dg(e->device);
  cudaIpcMemHandle_t h;
  SB_CUDA(cudaIpcGetMemHandle(&h, e->arena));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  if (byte_offset) *byte_offset = 0;
  SB_API_END
}

extern "C" int sb_profile_enable(sb_engine* e, int on) {
  SB_API_BEGIN
  need_device(e);
  DeviceGuard dg(e->device);
  if (!on && e->profiling) prof_resolve(e);
  e->profiling = on != 0;
  SB_API_END
}
extern "C" int sb_profile_count(const sb_engine* e) { return e ? (int)e->prof.size() : 0; }
extern "C" int sb_profile_get(const sb_engine* ce, int index, char* name, size_t name_cap, double* total_ms, int64_t* launches,
                              double* bytes, double* flops) {
  SB_API_BEGIN
  sb_engine* e = const_cast<sb_engine*>(ce);
  need_device(e);
  if (index < 0 || index >= (int)e->prof.size()) SB_FAIL(SB_ERR_INVALID_ARG, "profile index out of range");
  DeviceGuard dg(e->device);
  prof_resolve(e);
  const ProfEntry& p = e->prof[index];
  if (name && name_cap) {
    strncpy(name, p.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (total_ms) *total_ms = p.total_ms;
  if (launches) *launches = p.launches;
  if (bytes) *bytes = p.bytes;
  if (flops) *flops = p.flops;
  SB_API_END
}
extern "C" int sb_profile_reset(sb_engine* e) {
  SB_API_BEGIN
  need_device(e);
  DeviceGuard dg(e->device);
  prof_resolve(e);
  for (auto& p : e->prof) { p.total_ms = 0; p.launches = 0; p.bytes = 0; p.flops = 0; }
  SB_API_END
}
extern "C" int64_t sb_launch_count(const sb_engine* e) { return e ? e->launches : 0; }
