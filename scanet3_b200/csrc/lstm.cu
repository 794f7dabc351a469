// RNN(LSTMCell) macro-op of the scanet-b200 engine (layer/RNN.scala:41-73 wrapper, :257-339 cell).
//
// Reference semantics (stateless, returnSequence = false): x[B,T,F] is walked in time order from zero (c, h);
// per gate G in (forget, input, gate, output):  z_G = (x_t . Wk_G + h_{t-1} . Wr_G) + b_G ; f,i,o = recurrent activation,
// g = activation ; c_t = c_{t-1} * f + i * g ; h_t = o * act(c_t) ; the layer output is h_{T-1}.
//
// Device layout.  The per-gate tensors keep the reference's paths in the arena; every step starts by packing them into
//   Wr_cat[U, 4U], Wk_cat[F, 4U], b_cat[4U]   (column block G*U.. of gate G; 1 MB at U = 256)
// so that one GEMM per time step produces all four gates:  Z_t[B,4U] = H_{t-1}[B,U] . Wr_cat.  The gate kernel adds the input
// projection (inline for F <= 8, from a pre-computed X . Wk_cat otherwise) and the bias, applies the activations IN PLACE
// (Z_t becomes the saved gate activations) and writes c_t, h_t.  Saved for BPTT: acts[T,B,4U], c[T,B,U], h[T,B,U].
// Backward walks t = T-1..0: the gate kernel turns acts_t into dZ_t IN PLACE and produces dc_{t-1};
// dH_{t-1} = dZ_t . Wr_cat^T is one GEMM (K = 4U).  The weight gradients are three contractions over the WHOLE sequence:
//   dWr_cat = H_prev_all^T . dZ_all (K = T*B), dWk_cat = X_tm^T . dZ_all, db_cat = column sums of dZ_all,
// then unpacked into the per-gate gradient tensors.  All reductions run in fixed orders.
#include <stdexcept>
#include <string>

#include "common.cuh"
#include "engine.h"
#include "tc.h"

namespace sb {

namespace {

struct LstmWs {
  int B = 0, T = 0, F = 0, U = 0;
  float* wr_cat = nullptr;   // [U, 4U]
  float* wk_cat = nullptr;   // [F, 4U]
  float* b_cat = nullptr;    // [4U]
  float* x_tm = nullptr;     // [T, B, F] time-major copy of the input
  float* acts = nullptr;     // [T, B, 4U] gate activations, overwritten by dZ in the backward pass
  float* c = nullptr;        // [T, B, U]
  float* h = nullptr;        // [T+1, B, U]: slot 0 is h_{-1} = 0, slot t+1 is h_t  (so H_prev_all is one [T*B, U] matrix)
  float* xw = nullptr;       // [T, B, 4U] X . Wk_cat (only when F > 8)
  float* dh = nullptr;       // [B, U]
  float* dc = nullptr;       // [B, U]
  float* dwr_cat = nullptr;  // [U, 4U]
  float* dwk_cat = nullptr;  // [F, 4U]
  float* db_cat = nullptr;   // [4U]
  float* dx_tm = nullptr;    // [T, B, F] (only when the input gradient is needed)
  // tcgen05 TF32 plans (SB_PREC_TF32): one plan per contraction serves every time step through the tensor maps' slice coordinate
  TcGemm* g_fwd = nullptr;   // Z_t = H_{t-1} . Wr_cat           A slice t of h, C slice t of acts
  TcGemm* g_bwd = nullptr;   // dH_{t-1} = dZ_t . Wr_cat^T       A slice t of acts
  TcGemm* g_dwr = nullptr;   // dWr_cat = H_prev_all^T . dZ_all  (split-K over T*B)
};

__device__ __forceinline__ float actf(float x, int act) {
  switch (act) {
    case 1: return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
    case 2: return tanhf(x);
    case 4: return fmaxf(0.f, x);
    default: return x;
  }
}
// derivative through the OUTPUT y (models/Activation.scala + math/alg/AllKernels.scala:391-422): sigmoid (s(1-s))g, tanh (1-y^2)g
__device__ __forceinline__ float actg(float g, float y, int act) {
  switch (act) {
    case 1: return __fmul_rn(__fmul_rn(y, __fsub_rn(1.f, y)), g);
    case 2: return __fmul_rn(__fsub_rn(1.f, __fmul_rn(y, y)), g);
    case 4: return y > 0.f ? g : 0.f;
    default: return g;
  }
}

// pack / unpack: per-gate [rows, U] blocks <-> cat[rows, 4U].  ptrs[g] = gate g's tensor (null => zeros / skipped)
struct Gate4 {
  float* p[4];
};
__global__ void lstm_pack_kernel(Gate4 src, float* __restrict__ cat, int rows, int U) {
  pdl_prologue();
  const long long n = (long long)rows * 4 * U;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % (4 * U)), r = (int)(i / (4 * U));
    const int g = col / U, j = col - g * U;
    cat[i] = src.p[g] ? src.p[g][(long long)r * U + j] : 0.f;
  }
}
__global__ void lstm_unpack_kernel(const float* __restrict__ cat, Gate4 dst, int rows, int U) {
  pdl_prologue();
  const long long n = (long long)rows * 4 * U;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % (4 * U)), r = (int)(i / (4 * U));
    const int g = col / U, j = col - g * U;
    if (dst.p[g]) dst.p[g][(long long)r * U + j] = cat[i];
  }
}
// [B,T,F] <-> [T,B,F]
__global__ void lstm_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int D0, int D1, int F) {
  pdl_prologue();
  const long long n = (long long)D0 * D1 * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long r = i / F;
    const int d1 = (int)(r % D1), d0 = (int)(r / D1);
    out[((long long)d1 * D0 + d0) * F + f] = in[i];  // in[d0][d1][f] -> out[d1][d0][f]
  }
}

// Forward gates of one time step.  z (in/out): recurrent product (ignored when first), becomes the activations.
template <int FI>  // FI > 0: inline input projection with F = FI features ; FI == 0: xw holds X_t . Wk_cat
__global__ void __launch_bounds__(256) lstm_gates_fwd_kernel(float* __restrict__ z, const float* __restrict__ x_t,
                                                             const float* __restrict__ wk_cat, const float* __restrict__ b_cat,
                                                             const float* __restrict__ xw_t, const float* __restrict__ c_prev,
                                                             float* __restrict__ c_out, float* __restrict__ h_out, int B, int U,
                                                             int first, int act, int ract, int has_bias) {
  pdl_prologue();
  const long long n = (long long)B * U;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % U);
    const long long b = i / U;
    float a[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = g * U + j;
      float xp;
      if (FI > 0) {
        xp = 0.f;
#pragma unroll
        for (int f = 0; f < FI; ++f) xp = fmaf(__ldg(x_t + b * FI + f), __ldg(wk_cat + (long long)f * 4 * U + col), xp);
      } else {
        xp = xw_t[b * 4 * U + col];
      }
      float v = first ? xp : __fadd_rn(xp, z[b * 4 * U + col]);  // (x.Wk + h.Wr)
      if (has_bias) v = __fadd_rn(v, __ldg(b_cat + col));          // + b
      a[g] = actf(v, g == 2 ? act : ract);
      z[b * 4 * U + col] = a[g];
    }
    const float cp = first ? 0.f : c_prev[i];
    const float c = __fadd_rn(__fmul_rn(cp, a[0]), __fmul_rn(a[1], a[2]));  // c_prev*f + i*g
    c_out[i] = c;
    h_out[i] = __fmul_rn(a[3], actf(c, act));                                // o * act(c)
  }
}

// Backward gates of one time step (oracle/layers.py LSTMCell.step_bwd):
//   do = dh*ac ; dc = act'(dh*o ; ac) + dc_next ; df = dc*c_prev ; di = dc*g ; dg = dc*i ; dz_G = act_G'(d_G ; a_G) ; dc_prev = dc*f
__global__ void __launch_bounds__(256) lstm_gates_bwd_kernel(float* __restrict__ acts_dz, const float* __restrict__ c_t,
                                                             const float* __restrict__ c_prev, const float* __restrict__ dh,
                                                             float* __restrict__ dc_io, int B, int U, int first, int last, int act,
                                                             int ract) {
  pdl_prologue();
  const long long n = (long long)B * U;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % U);
    const long long b = i / U;
    float* row = acts_dz + b * 4 * U;
    const float f = row[j], ig = row[U + j], g = row[2 * U + j], o = row[3 * U + j];
    const float ac = actf(c_t[i], act);
    const float dhv = dh[i];
    const float d_o = __fmul_rn(dhv, ac);
    float dc = actg(__fmul_rn(dhv, o), ac, act);
    if (!last) dc = __fadd_rn(dc, dc_io[i]);
    const float cp = first ? 0.f : c_prev[i];
    row[j] = actg(__fmul_rn(dc, cp), f, ract);
    row[U + j] = actg(__fmul_rn(dc, g), ig, ract);
    row[2 * U + j] = actg(__fmul_rn(dc, ig), g, act);
    row[3 * U + j] = actg(d_o, o, ract);
    dc_io[i] = __fmul_rn(dc, f);
  }
}

int ew_grid(long long n) {
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

LstmWs* ws_of(sb_engine* e, LayerPlan& L) {
  if (L.lstm) return (LstmWs*)L.lstm;
  LstmWs* w = new LstmWs();
  w->B = (int)L.in.d[0]; w->T = (int)L.in.d[1]; w->F = (int)L.in.d[2]; w->U = L.d.units;
  const long long B = w->B, T = w->T, F = w->F, U = w->U;
  w->wr_cat = dalloc(U * 4 * U);
  w->wk_cat = dalloc(F * 4 * U);
  w->b_cat = dalloc(4 * U);
  w->x_tm = dalloc(T * B * F);
  w->acts = dalloc(T * B * 4 * U);
  w->c = dalloc(T * B * U);
  w->h = dalloc((T + 1) * B * U);
  SB_CUDA(cudaMemsetAsync(w->h, 0, (size_t)B * U * 4, e->stream));  // h_{-1} = 0 (models/layer/RNN.scala:32-35: Zeros state)
  if (F > 8) w->xw = dalloc(T * B * 4 * U);
  w->dh = dalloc(B * U);
  w->dc = dalloc(B * U);
  w->dwr_cat = dalloc(U * 4 * U);
  w->dwk_cat = dalloc(F * 4 * U);
  w->db_cat = dalloc(4 * U);
  if (L.need_dx) w->dx_tm = dalloc(T * B * F);
  if (e->train.precision != SB_PREC_FP32 && B * U * 4 * U >= (1 << 22)) {
    const int x3 = e->train.precision == SB_PREC_3XTF32 ? 1 : 0;
    w->g_fwd = tc_gemm_create(0, 1, w->h, (int)T + 1, w->wr_cat, 1, w->acts, (int)T, (int)B, (int)(4 * U), (int)U, 0, nullptr, 0, e->device, x3);
    w->g_bwd = tc_gemm_create(0, 0, w->acts, (int)T, w->wr_cat, 1, w->dh, 1, (int)B, (int)U, (int)(4 * U), 0, nullptr, 0, e->device, x3);
    w->g_dwr = tc_gemm_create(1, 1, w->h, 1, w->acts, 1, w->dwr_cat, 1, (int)U, (int)(4 * U), (int)(T * B), 1, e->ws, e->ws_floats,
                              e->device, x3);
  }
  L.lstm = w;
  return w;
}

Gate4 gate_ptrs(sb_engine* e, const LayerPlan& L, int which, bool grads) {
  Gate4 g;
  for (int gi = 0; gi < 4; ++gi) {
    const int pi = L.p_aux[gi * 3 + which];
    g.p[gi] = pi < 0 ? nullptr : (grads ? e->grads : e->arena) + e->params[pi].offset;
  }
  return g;
}

template <int FI>
void launch_gates_fwd(LstmWs* w, float* z, const float* x_t, const float* xw_t, const float* c_prev, float* c_out, float* h_out,
                      int first, int act, int ract, int has_bias, cudaStream_t s) {
  const long long n = (long long)w->B * w->U;
  launch_k(lstm_gates_fwd_kernel<FI>, ew_grid(n), 256, 0, s, z, x_t, w->wk_cat, w->b_cat, xw_t, c_prev, c_out, h_out, w->B, w->U, first,
                                                       act, ract, has_bias);
  SB_LAUNCHED();
}

}  // namespace

void lstm_plan(sb_engine* e) {
  for (auto& L : e->L)
    if (L.d.kind == SB_LAYER_LSTM) ws_of(e, L);  // allocate outside any stream capture
}

void lstm_forward(sb_engine* e, LayerPlan& L, int) {
  LstmWs* w = ws_of(e, L);
  cudaStream_t s = e->stream;
  const int B = w->B, T = w->T, F = w->F, U = w->U;
  const long long BU = (long long)B * U;
  const int act = L.d.activation, ract = L.d.recurrent_activation;
  // pack the per-gate tensors (the optimizer updated them since the last step)
  launch_k(lstm_pack_kernel, ew_grid((long long)U * 4 * U), 256, 0, s, gate_ptrs(e, L, 1, false), w->wr_cat, U, U);
  SB_LAUNCHED();
  launch_k(lstm_pack_kernel, ew_grid((long long)F * 4 * U), 256, 0, s, gate_ptrs(e, L, 0, false), w->wk_cat, F, U);
  SB_LAUNCHED();
  if (L.d.use_bias) {
    launch_k(lstm_pack_kernel, ew_grid(4LL * U), 256, 0, s, gate_ptrs(e, L, 2, false), w->b_cat, 1, U);
    SB_LAUNCHED();
  }
  launch_k(lstm_transpose_kernel, ew_grid((long long)B * T * F), 256, 0, s, e->acts[L.buf_in], w->x_tm, B, T, F);
  SB_LAUNCHED();
  if (w->xw) gemm_nn(w->x_tm, w->wk_cat, w->xw, T * B, 4 * U, F, e->ws, e->ws_floats, s);
  for (int t = 0; t < T; ++t) {
    float* z = w->acts + (long long)t * B * 4 * U;
    const float* h_prev = w->h + (long long)t * BU;
    float* h_t = w->h + (long long)(t + 1) * BU;
    float* c_t = w->c + (long long)t * BU;
    const float* c_prev = t ? w->c + (long long)(t - 1) * BU : nullptr;
    if (t > 0) {  // h_{-1} = 0: no product at t = 0
      if (w->g_fwd) tc_gemm_run(w->g_fwd, t, 0, t, s);
      else gemm_nn(h_prev, w->wr_cat, z, B, 4 * U, U, e->ws, e->ws_floats, s);
    }
    const float* x_t = w->x_tm + (long long)t * B * F;
    const float* xw_t = w->xw ? w->xw + (long long)t * B * 4 * U : nullptr;
    const int first = t == 0;
    switch (w->xw ? 0 : F) {
      case 0: launch_gates_fwd<0>(w, z, x_t, xw_t, c_prev, c_t, h_t, first, act, ract, L.d.use_bias, s); break;
      case 1: launch_gates_fwd<1>(w, z, x_t, xw_t, c_prev, c_t, h_t, first, act, ract, L.d.use_bias, s); break;
      case 2: launch_gates_fwd<2>(w, z, x_t, xw_t, c_prev, c_t, h_t, first, act, ract, L.d.use_bias, s); break;
      case 3: launch_gates_fwd<3>(w, z, x_t, xw_t, c_prev, c_t, h_t, first, act, ract, L.d.use_bias, s); break;
      case 4: launch_gates_fwd<4>(w, z, x_t, xw_t, c_prev, c_t, h_t, first, act, ract, L.d.use_bias, s); break;
      case 5: launch_gates_fwd<5>(w, z, x_t, xw_t, c_prev, c_t, h_t, first, act, ract, L.d.use_bias, s); break;
      case 6: launch_gates_fwd<6>(w, z, x_t, xw_t, c_prev, c_t, h_t, first, act, ract, L.d.use_bias, s); break;
      case 7: launch_gates_fwd<7>(w, z, x_t, xw_t, c_prev, c_t, h_t, first, act, ract, L.d.use_bias, s); break;
      default: launch_gates_fwd<8>(w, z, x_t, xw_t, c_prev, c_t, h_t, first, act, ract, L.d.use_bias, s); break;
    }
  }
  // the layer output is the last hidden state (returnSequence = false)
  SB_CUDA(cudaMemcpyAsync(e->acts[L.buf_out], w->h + (long long)T * BU, (size_t)BU * 4, cudaMemcpyDeviceToDevice, s));
}

void lstm_backward(sb_engine* e, LayerPlan& L, int) {
  LstmWs* w = ws_of(e, L);
  cudaStream_t s = e->stream;
  const int B = w->B, T = w->T, F = w->F, U = w->U;
  const long long BU = (long long)B * U;
  const int act = L.d.activation, ract = L.d.recurrent_activation;
  const float* dh = e->dacts[L.buf_out];  // dLoss/dh_{T-1}
  for (int t = T - 1; t >= 0; --t) {
    float* a = w->acts + (long long)t * B * 4 * U;
    const float* c_t = w->c + (long long)t * BU;
    const float* c_prev = t ? w->c + (long long)(t - 1) * BU : nullptr;
    launch_k(lstm_gates_bwd_kernel, ew_grid(BU), 256, 0, s, a, c_t, c_prev, dh, w->dc, B, U, t == 0, t == T - 1, act, ract);
    SB_LAUNCHED();
    if (t > 0) {
      if (w->g_bwd) tc_gemm_run(w->g_bwd, t, 0, 0, s);  // dH_{t-1} = dZ_t . Wr_cat^T
      else gemm_nt(a, w->wr_cat, w->dh, B, U, 4 * U, e->ws, e->ws_floats, s);
      dh = w->dh;
    }
  }
  // weight gradients over the whole sequence (acts now holds dZ_all[T*B, 4U]; h slots 0..T-1 are H_prev_all[T*B, U])
  if (w->g_dwr) tc_gemm_run(w->g_dwr, 0, 0, 0, s);
  else gemm_tn(w->h, w->acts, w->dwr_cat, U, 4 * U, T * B, e->ws, e->ws_floats, s);
  gemm_tn(w->x_tm, w->acts, w->dwk_cat, F, 4 * U, T * B, e->ws, e->ws_floats, s);
  launch_k(lstm_unpack_kernel, ew_grid((long long)U * 4 * U), 256, 0, s, w->dwr_cat, gate_ptrs(e, L, 1, true), U, U);
  SB_LAUNCHED();
  launch_k(lstm_unpack_kernel, ew_grid((long long)F * 4 * U), 256, 0, s, w->dwk_cat, gate_ptrs(e, L, 0, true), F, U);
  SB_LAUNCHED();
  if (L.d.use_bias) {
    col_sum(w->acts, w->db_cat, (long long)T * B, 4 * U, e->ws, e->ws_floats, s);
    launch_k(lstm_unpack_kernel, ew_grid(4LL * U), 256, 0, s, w->db_cat, gate_ptrs(e, L, 2, true), 1, U);
    SB_LAUNCHED();
  }
  float* din = e->dacts[L.buf_in];
  if (L.need_dx && din && w->dx_tm) {
    gemm_nt(w->acts, w->wk_cat, w->dx_tm, T * B, F, 4 * U, e->ws, e->ws_floats, s);  // dX_t = dZ_t . Wk_cat^T
    launch_k(lstm_transpose_kernel, ew_grid((long long)B * T * F), 256, 0, s, w->dx_tm, din, T, B, F);
    SB_LAUNCHED();
  }
}

void lstm_free(sb_engine* e) {
  for (auto& L : e->L) {
    if (!L.lstm) continue;
    LstmWs* w = (LstmWs*)L.lstm;
    for (float* p : {w->wr_cat, w->wk_cat, w->b_cat, w->x_tm, w->acts, w->c, w->h, w->xw, w->dh, w->dc, w->dwr_cat, w->dwk_cat,
                     w->db_cat, w->dx_tm})
      cudaFree(p);
    tc_gemm_destroy(w->g_fwd); tc_gemm_destroy(w->g_bwd); tc_gemm_destroy(w->g_dwr);
    delete w;
    L.lstm = nullptr;
  }
}

}  // namespace sb
