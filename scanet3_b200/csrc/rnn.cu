// RNN(SimpleRNNCell >> Bias >> Activate) (models/layer/RNN.scala:27-206).
//
//   z_t   = x_t . Wk + z_{t-1} . Wr          (z_{-1} = 0)
//   out_t = act(z_t + b)
//
// NOTE the recurrent state is z_t, the PRE-bias, PRE-activation sum: `SimpleRNNCell.build` returns `Params(State -> result)`
// before Bias / Activate are composed onto the cell (RNN.scala:190-196) -- a reference quirk, reproduced (oracle/layers.py
// SimpleRNNCell, pinned by RNNLayerSpec.scala:17-34 and by finite differences).  Parameter paths as the composed cell names them:
// "0/kernel_weights" [F,U], "0/recurrent_weights" [U,U], "1/weights" [U] (bias).
//
// Forward: one GEMM X_all . Wk for every step, then per step one [B,U]x[U,U] product and a pointwise kernel.  Backward (BPTT):
//   dpre_t = act'(out_t) * dOut_t ;  dz_t = dpre_t + dz_{t+1} . Wr^T ;  db = sum dpre ;  dWk = X_all^T . dZ_all ;
//   dWr = Z_prev_all^T . dZ_all ;  dX_all = dZ_all . Wk^T.
// The recurrence is a chain of small products by nature; contractions run on the fp32 FFMA kernels in every precision mode.
#include "common.cuh"
#include "engine.h"
#include "tc.h"

namespace sb {
namespace {

struct RnnWs {
  int B = 0, T = 0, F = 0, U = 0;
  float* x_tm = nullptr;     // [T, B, F]
  float* z = nullptr;        // [T+1, B, U]: slot 0 is z_{-1} = 0, slot t+1 is z_t (so Z_prev_all is one [T*B, U] matrix)
  float* out = nullptr;      // [T, B, U] activated outputs
  float* rec = nullptr;      // [B, U] z_{t-1} . Wr, later dz_{t+1} . Wr^T
  float* dz = nullptr;       // [T, B, U]
  float* dpre = nullptr;     // [T, B, U] (bias gradient source)
  float* dout_tm = nullptr;  // [T, B, U] (returnSequence)
  float* dx_tm = nullptr;    // [T, B, F]
};

__device__ __forceinline__ float rnn_actf(float x, int act, float thr) {
  switch (act) {
    case SB_ACT_SIGMOID: return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
    case SB_ACT_TANH: return tanhf(x);
    case SB_ACT_RELU: return fmaxf(thr, x);
    default: return x;
  }
}
__device__ __forceinline__ float rnn_actg(float g, float y, int act, float thr) {  // derivative through the output y
  switch (act) {
    case SB_ACT_SIGMOID: return __fmul_rn(__fmul_rn(y, __fsub_rn(1.f, y)), g);
    case SB_ACT_TANH: return __fmul_rn(__fsub_rn(1.f, __fmul_rn(y, y)), g);
    case SB_ACT_RELU: return y > thr ? g : 0.f;
    default: return g;
  }
}

__global__ void rnn_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int D0, int D1, int F) {
  pdl_prologue();
  const long long n = (long long)D0 * D1 * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long r = i / F;
    const int d1 = (int)(r % D1), d0 = (int)(r / D1);
    out[((long long)d1 * D0 + d0) * F + f] = in[i];  // in[d0][d1][f] -> out[d1][d0][f]
  }
}

// z_t = xw_t + rec (rec absent at t = 0) ; out_t = act(z_t + b).  z_t aliases xw_t (in place).
__global__ void rnn_step_fwd_kernel(float* __restrict__ z_t, const float* __restrict__ rec, const float* __restrict__ bias,
                                    float* __restrict__ out_t, long long n, int U, int act, float thr) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float z = rec ? __fadd_rn(z_t[i], rec[i]) : z_t[i];
    z_t[i] = z;
    out_t[i] = rnn_actf(bias ? __fadd_rn(z, bias[i % U]) : z, act, thr);
  }
}

// dpre_t = act'(out_t) * dOut_t (zero when this step has no output gradient) ; dz_t = dpre_t + rec (absent at t = T-1)
__global__ void rnn_step_bwd_kernel(const float* __restrict__ out_t, const float* __restrict__ dout_t, const float* __restrict__ rec,
                                    float* __restrict__ dpre_t, float* __restrict__ dz_t, long long n, int act, float thr) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float dp = dout_t ? rnn_actg(dout_t[i], out_t[i], act, thr) : 0.f;
    dpre_t[i] = dp;
    dz_t[i] = rec ? __fadd_rn(dp, rec[i]) : dp;
  }
}

int grid_of(long long n) {
  long long b = (n + 255) / 256;
  if (b < 1) b = 1;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b > cap ? cap : b);
}
float* dalloc(long long floats) {
  void* p = nullptr;
  SB_CUDA(cudaMalloc(&p, (size_t)(floats < 32 ? 32 : floats) * sizeof(float)));
  return (float*)p;
}

RnnWs* ws_of(sb_engine* e, LayerPlan& L) {
  if (L.lstm) return (RnnWs*)L.lstm;  // the RNN layers share the opaque workspace slot
  RnnWs* w = new RnnWs();
  w->B = (int)L.in.d[0]; w->T = (int)L.in.d[1]; w->F = (int)L.in.d[2]; w->U = L.d.units;
  const long long B = w->B, T = w->T, F = w->F, U = w->U;
  w->x_tm = dalloc(T * B * F);
  w->z = dalloc((T + 1) * B * U);
  SB_CUDA(cudaMemsetAsync(w->z, 0, (size_t)B * U * 4, e->stream));  // z_{-1} = 0 (Zeros state, stateless RNN: RNN.scala:66-67)
  w->out = dalloc(T * B * U);
  w->rec = dalloc(B * U);
  w->dz = dalloc(T * B * U);
  w->dpre = dalloc(T * B * U);
  if (L.d.return_sequence) w->dout_tm = dalloc(T * B * U);
  if (L.need_dx) w->dx_tm = dalloc(T * B * F);
  L.lstm = w;
  return w;
}

}  // namespace

void rnn_plan(sb_engine* e) {
  for (auto& L : e->L)
    if (L.d.kind == SB_LAYER_SIMPLE_RNN) ws_of(e, L);  // allocate outside any stream capture
}

void rnn_forward(sb_engine* e, LayerPlan& L, int, bool store_state) {
  RnnWs* w = ws_of(e, L);
  cudaStream_t s = e->stream;
  const int B = w->B, T = w->T, F = w->F, U = w->U;
  const long long BU = (long long)B * U;
  const float* Wk = e->arena + e->params[L.p_aux[0]].offset;
  const float* Wr = e->arena + e->params[L.p_aux[1]].offset;
  const float* b = L.d.use_bias ? e->arena + e->params[L.p_aux[2]].offset : nullptr;
  launch_k(rnn_transpose_kernel, grid_of((long long)B * T * F), 256, 0, s, e->acts[L.buf_in], w->x_tm, B, T, F);
  SB_LAUNCHED();
  gemm_nn(w->x_tm, Wk, w->z + BU, T * B, U, F, e->ws, e->ws_floats, s);  // X_all . Wk straight into the z slots 1..T
  const bool stateful = L.d.stateful != 0;
  float* state = stateful ? e->arena + e->params[L.p_aux[3]].offset : nullptr;
  // stateful: z_{-1} is the state parameter (slot 0 keeps it for the backward pass: dWr needs z_{-1} too); zeros otherwise
  if (stateful) SB_CUDA(cudaMemcpyAsync(w->z, state, (size_t)BU * 4, cudaMemcpyDeviceToDevice, s));
  for (int t = 0; t < T; ++t) {
    float* z_t = w->z + (long long)(t + 1) * BU;
    const bool rec = t > 0 || stateful;
    if (rec) gemm_nn(w->z + (long long)t * BU, Wr, w->rec, B, U, U, e->ws, e->ws_floats, s);
    launch_k(rnn_step_fwd_kernel, grid_of(BU), 256, 0, s, z_t, rec ? w->rec : nullptr, b, w->out + (long long)t * BU, BU, U,
             L.d.activation, L.d.relu_threshold);
    SB_LAUNCHED();
  }
  // a train step hands the last state on with the parameters (`nextParams = nextWeights ++ nextState`, Optimizer.scala:141-144)
  if (stateful && store_state)
    SB_CUDA(cudaMemcpyAsync(state, w->z + (long long)T * BU, (size_t)BU * 4, cudaMemcpyDeviceToDevice, s));
  if (L.d.return_sequence) {
    launch_k(rnn_transpose_kernel, grid_of((long long)T * BU), 256, 0, s, w->out, e->acts[L.buf_out], T, B, U);
    SB_LAUNCHED();
  } else {
    SB_CUDA(cudaMemcpyAsync(e->acts[L.buf_out], w->out + (long long)(T - 1) * BU, (size_t)BU * 4, cudaMemcpyDeviceToDevice, s));
  }
}

void rnn_backward(sb_engine* e, LayerPlan& L, int) {
  RnnWs* w = ws_of(e, L);
  cudaStream_t s = e->stream;
  const int B = w->B, T = w->T, F = w->F, U = w->U;
  const long long BU = (long long)B * U;
  const float* Wk = e->arena + e->params[L.p_aux[0]].offset;
  const float* Wr = e->arena + e->params[L.p_aux[1]].offset;
  const bool seq = L.d.return_sequence != 0;
  if (seq) {
    launch_k(rnn_transpose_kernel, grid_of((long long)T * BU), 256, 0, s, e->dacts[L.buf_out], w->dout_tm, B, T, U);
    SB_LAUNCHED();
  }
  for (int t = T - 1; t >= 0; --t) {
    const float* dout_t = seq ? w->dout_tm + (long long)t * BU : (t == T - 1 ? e->dacts[L.buf_out] : nullptr);
    float* dz_t = w->dz + (long long)t * BU;
    launch_k(rnn_step_bwd_kernel, grid_of(BU), 256, 0, s, w->out + (long long)t * BU, dout_t, t < T - 1 ? w->rec : nullptr,
             w->dpre + (long long)t * BU, dz_t, BU, L.d.activation, L.d.relu_threshold);
    SB_LAUNCHED();
    if (t > 0) gemm_nt(dz_t, Wr, w->rec, B, U, U, e->ws, e->ws_floats, s);  // dz_t . Wr^T -> the state gradient of step t-1
  }
  gemm_tn(w->x_tm, w->dz, e->grads + e->params[L.p_aux[0]].offset, F, U, T * B, e->ws, e->ws_floats, s);  // dWk = X_all^T . dZ_all
  gemm_tn(w->z, w->dz, e->grads + e->params[L.p_aux[1]].offset, U, U, T * B, e->ws, e->ws_floats, s);     // dWr = Z_prev_all^T . dZ_all
  if (L.d.use_bias) col_sum(w->dpre, e->grads + e->params[L.p_aux[2]].offset, (long long)T * B, U, e->ws, e->ws_floats, s);
  float* din = e->dacts[L.buf_in];
  if (L.need_dx && din && w->dx_tm) {
    gemm_nt(w->dz, Wk, w->dx_tm, T * B, F, U, e->ws, e->ws_floats, s);  // dX_t = dz_t . Wk^T
    launch_k(rnn_transpose_kernel, grid_of((long long)B * T * F), 256, 0, s, w->dx_tm, din, T, B, F);
    SB_LAUNCHED();
  }
}

void rnn_free(sb_engine* e) {
  for (auto& L : e->L) {
    if (!L.lstm || L.d.kind != SB_LAYER_SIMPLE_RNN) continue;
    RnnWs* w = (RnnWs*)L.lstm;
    for (float* p : {w->x_tm, w->z, w->out, w->rec, w->dz, w->dpre, w->dout_tm, w->dx_tm}) cudaFree(p);
    delete w;
    L.lstm = nullptr;
  }
}

}  // namespace sb

namespace sb {
void trace_set_rnn(unsigned long long* p) { sb_trace_set_local(p); }  // SB_TRACE timeline buffer of this translation unit
}  // namespace sb
