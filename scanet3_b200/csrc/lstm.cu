// RNN(LSTMCell) macro-op of the scanet-b200 engine (layer/RNN.scala:41-73 wrapper, :257-339 cell).
//
// Reference semantics (stateless, returnSequence = false): x[B,T,F] is walked in time order from zero (c, h);
// per gate G in (forget, input, gate, output):  z_G = (x_t . Wk_G + h_{t-1} . Wr_G) + b_G ; f,i,o = recurrent activation,
// g = activation ; c_t = c_{t-1} * f + i * g ; h_t = o * act(c_t) ; the layer output is h_{T-1}.
//
// Device layout.  The per-gate tensors keep the reference's paths in the arena; every step starts by packing them into
//   Wr_cat[U, 4U], Wk_cat[F, 4U], b_cat[4U]   (column 4*j + G = gate G of unit j: the four gates of a unit are adjacent, so a
//   thread that owns 4k consecutive columns of a GEMM row owns k whole units; 1 MB at U = 256)
// so that one GEMM per time step produces all four gates:  Z_t[B,4U] = H_{t-1}[B,U] . Wr_cat.  The gate kernel adds the input
// projection (inline for F <= 8, from a pre-computed X . Wk_cat otherwise) and the bias, applies the activations IN PLACE
// (Z_t becomes the saved gate activations) and writes c_t, h_t.  Saved for BPTT: acts[T,B,4U], c[T,B,U], h[T,B,U].
// Backward walks t = T-1..0: the gate kernel turns acts_t into dZ_t IN PLACE and produces dc_{t-1};
// dH_{t-1} = dZ_t . Wr_cat^T is one GEMM (K = 4U).  The weight gradients are three contractions over the WHOLE sequence:
//   dWr_cat = H_prev_all^T . dZ_all (K = T*B), dWk_cat = X_tm^T . dZ_all, db_cat = column sums of dZ_all,
// then unpacked into the per-gate gradient tensors.  All reductions run in fixed orders.
#include <stdlib.h>

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
  float* dwkb_cat = nullptr; // [1 + F][4U]: db_cat row followed by dWk_cat rows (F <= 8: one fused pass)
  float* dx_tm = nullptr;    // [T, B, F] (only when the input gradient is needed)
  float* dout_tm = nullptr;  // [T, B, U] time-major copy of dLoss/d(output sequence) (returnSequence only)
  float* c_init = nullptr;   // [B, U] stateful: c_{-1} as it was when the forward pass started (the backward pass needs it after
                             // the train step has stored the new state)
  // tcgen05 TF32 plans (SB_PREC_TF32): one plan per contraction serves every time step through the tensor maps' slice coordinate
  TcGemm* g_fwd = nullptr;   // Z_t = H_{t-1} . Wr_cat           A slice t of h, C slice t of acts
  TcGemm* g_bwd = nullptr;   // dH_{t-1} = dZ_t . Wr_cat^T       A slice t of acts
  TcGemm* g_dwr = nullptr;   // dWr_cat = H_prev_all^T . dZ_all  (split-K over T*B)
  // Gate math inside the GEMM epilogues (tc.h LstmEpi).  Measured on cfg3 (B200): 5.68 ms/step vs 1.80 ms with separate gate
  // kernels -- the epilogue's row-per-thread TMEM layout makes every c / dc / acts access uncoalesced and concentrates the
  // pointwise work on 128 threads of <= 128 CTAs.  Kept selectable (SB_LSTM_FUSED_EPILOGUE=1) and parity-tested; off by default.
  bool fused_epilogue = false;
  bool fused_fwd = false;
  float* bwd_ws = nullptr;   // split-K partials of the dH product
  size_t bwd_ws_floats = 0;
  bool bwd_fold = true;      // the gate backward kernel sums the split-K partials itself (SB_LSTM_BWD_FOLD=0: a separate partial_reduce launch)
  bool fast_math = false;    // TF32 mode: approximate exp / reciprocal in the gate activations (actf_t); SB_LSTM_EXACT_GATES=1 turns it off
  void* persist = nullptr;   // lstm_persist.cu: one launch for all T forward steps (TF32 mode, F <= 8, B % 128 == 0, <= 148 CTAs)
};

__device__ __forceinline__ float actf(float x, int act) {
  switch (act) {
    case 1: return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
    case 2: return tanhf(x);
    case 4: return fmaxf(0.f, x);
    default: return x;
  }
}
// TF32 mode only (FAST): sigmoid and tanh through ex2.approx / rcp.approx -- relative error ~1e-6 (absolute ~1e-7 for tanh near
// zero), three orders below the tf32 products they are applied to, and 4-5x fewer instructions than expf + IEEE division / tanhf:
// the gate kernels of a time step are instruction-bound, not memory-bound (profiles/r2_notes.md 13).  FP32 and 3xTF32 keep the exact forms.
template <bool FAST>
__device__ __forceinline__ float actf_t(float x, int act) {
  if (!FAST) return actf(x, act);
  switch (act) {
    case 1: return __fdividef(1.f, 1.f + __expf(-x));
    case 2: return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x));  // tanh x = 1 - 2 / (1 + e^{2x}); saturates cleanly (e^{2x} -> inf: 1, -> 0: -1)
    case 4: return fmaxf(0.f, x);
    default: return x;
  }
}
// TF32 mode: values that are ONLY ever read as kind::tf32 MMA operands of the recurrence (h_t, dZ_t, the packed Wr_cat) are stored
// rounded to the NEAREST tf32 value.  kind::tf32 truncates its operands, i.e. it shrinks every product by ~2^-11 on average, and
// through 64 steps of BPTT that bias compounds (measured on config 3: gradient norm 0.86 of the float64 oracle's, cosine 0.991);
// with operands that are already tf32 values the truncation is exact and the rounding error is unbiased.
template <bool FAST>
__device__ __forceinline__ float tf32_rn(float x) {
  if (!FAST) return x;
  unsigned u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
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
__global__ void lstm_pack_kernel(Gate4 src, float* __restrict__ cat, int rows, int U, int round_tf32) {
  pdl_prologue_trigger();
  const long long n = (long long)rows * 4 * U;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % (4 * U)), r = (int)(i / (4 * U));
    const int j = col >> 2, g = col & 3;
    const float v = src.p[g] ? src.p[g][(long long)r * U + j] : 0.f;
    cat[i] = round_tf32 ? tf32_rn<true>(v) : v;  // Wr_cat in TF32 mode: an MMA operand only (see tf32_rn)
  }
}
__global__ void lstm_unpack_kernel(const float* __restrict__ cat, Gate4 dst, int rows, int U) {
  pdl_prologue_trigger();
  const long long n = (long long)rows * 4 * U;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % (4 * U)), r = (int)(i / (4 * U));
    const int j = col >> 2, g = col & 3;
    if (dst.p[g]) dst.p[g][(long long)r * U + j] = cat[i];
  }
}
// [B,T,F] <-> [T,B,F]
__global__ void lstm_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int D0, int D1, int F) {
  pdl_prologue_trigger();
  const long long n = (long long)D0 * D1 * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long r = i / F;
    const int d1 = (int)(r % D1), d0 = (int)(r / D1);
    out[((long long)d1 * D0 + d0) * F + f] = in[i];  // in[d0][d1][f] -> out[d1][d0][f]
  }
}

// Forward gates of one time step.  z (in/out): recurrent product (ignored when first), becomes the activations.
template <int FI, bool FAST>  // FI > 0: inline input projection with F = FI features ; FI == 0: xw holds X_t . Wk_cat
__global__ void __launch_bounds__(256) lstm_gates_fwd_kernel(float* __restrict__ z, const float* __restrict__ x_t,
                                                             const float* __restrict__ wk_cat, const float* __restrict__ b_cat,
                                                             const float* __restrict__ xw_t, const float* __restrict__ c_prev,
                                                             float* __restrict__ c_out, float* __restrict__ h_out, int B, int U,
                                                             int first, int act, int ract, int has_bias) {
  pdl_prologue_trigger();
  const unsigned n = (unsigned)B * (unsigned)U;  // the launcher keeps B * 4U below 2^31: 32-bit index math
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned b = i / (unsigned)U;
    const int j = (int)(i - b * (unsigned)U);
    float a[4];
    float4* zq = reinterpret_cast<float4*>(z + (size_t)b * 4 * U + 4 * j);  // the unit's four gates are adjacent
    const float4 zr = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *zq;
    const float zin[4] = {zr.x, zr.y, zr.z, zr.w};
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = 4 * j + g;
      float xp;
      if (FI > 0) {
        xp = 0.f;
#pragma unroll
        for (int f = 0; f < FI; ++f) xp = fmaf(__ldg(x_t + (size_t)b * FI + f), __ldg(wk_cat + (size_t)f * 4 * U + col), xp);
      } else {
        xp = xw_t[(size_t)b * 4 * U + col];
      }
      float v = first ? xp : __fadd_rn(xp, zin[g]);           // (x.Wk + h.Wr)
      if (has_bias) v = __fadd_rn(v, __ldg(b_cat + col));       // + b
      a[g] = actf_t<FAST>(v, g == 2 ? act : ract);
    }
    *zq = make_float4(a[0], a[1], a[2], a[3]);
    const float cp = first ? 0.f : c_prev[i];
    const float c = __fadd_rn(__fmul_rn(cp, a[0]), __fmul_rn(a[1], a[2]));  // c_prev*f + i*g
    c_out[i] = c;
    h_out[i] = tf32_rn<FAST>(__fmul_rn(a[3], actf_t<FAST>(c, act)));         // o * act(c)
  }
}

// Backward gates of one time step (oracle/layers.py LSTMCell.step_bwd):
//   do = dh*ac ; dc = act'(dh*o ; ac) + dc_next ; df = dc*c_prev ; di = dc*g ; dg = dc*i ; dz_G = act_G'(d_G ; a_G) ; dc_prev = dc*f
template <bool FAST>
__global__ void __launch_bounds__(256) lstm_gates_bwd_kernel(float* __restrict__ acts_dz, const float* __restrict__ c_t,
                                                             const float* __restrict__ c_prev, const float* __restrict__ dh,
                                                             const float* __restrict__ dh_seq, float* __restrict__ dc_io, int B, int U,
                                                             int first, int last, int act, int ract, int dh_parts) {
  pdl_prologue_trigger();
  const unsigned n = (unsigned)B * (unsigned)U;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned b = i / (unsigned)U;
    const int j = (int)(i - b * (unsigned)U);
    float4* q = reinterpret_cast<float4*>(acts_dz + (size_t)b * 4 * U + 4 * j);
    const float4 a4 = *q;
    const float f = a4.x, ig = a4.y, g = a4.z, o = a4.w;
    const float ac = actf_t<FAST>(c_t[i], act);
    // returnSequence: this step's own output gradient joins the recurrent one (`dout + dstate`, oracle RNN.backward)
    // dh_parts > 1: dh is the split-K partials [dh_parts][B*U] of the dH product, summed here in split order (the order partial_reduce uses)
    float dhp[8];  // all loads in flight first (the plan keeps dh_parts <= 8), then the ordered sum
#pragma unroll
    for (int sp = 0; sp < 8; ++sp) dhp[sp] = sp < dh_parts ? __ldg(dh + (size_t)sp * n + i) : 0.f;
    float dhr = dhp[0];
#pragma unroll
    for (int sp = 1; sp < 8; ++sp)
      if (sp < dh_parts) dhr = __fadd_rn(dhr, dhp[sp]);
    const float dhv = dh_seq ? __fadd_rn(dh_seq[i], dhr) : dhr;
    const float d_o = __fmul_rn(dhv, ac);
    float dc = actg(__fmul_rn(dhv, o), ac, act);
    if (!last) dc = __fadd_rn(dc, dc_io[i]);
    const float cp = first ? 0.f : c_prev[i];
    *q = make_float4(tf32_rn<FAST>(actg(__fmul_rn(dc, cp), f, ract)), tf32_rn<FAST>(actg(__fmul_rn(dc, g), ig, ract)),
                     tf32_rn<FAST>(actg(__fmul_rn(dc, ig), g, act)), tf32_rn<FAST>(actg(d_o, o, ract)));
    dc_io[i] = __fmul_rn(dc, f);
  }
}


// Bias and input-kernel gradients in ONE pass over dZ_all[R, 4U] (R = T*B):  db[c] = sum_r dZ[r][c] ;  dWk[f][c] = sum_r X[r][f] dZ[r][c]
// (F <= 8 input features: the "GEMM" has M = F rows, i.e. it is a weighted column sum).  Block = 32 column quads x 8 row lanes
// over a chunk of rows; partial[chunk][1 + F][4U], summed in chunk order by partial_reduce.
template <int FI>
__global__ void __launch_bounds__(256) lstm_dwk_db_kernel(const float* __restrict__ dz, const float* __restrict__ x, float* __restrict__ partial,
                                                          long long rows, int C, long long rows_per_chunk) {
  pdl_prologue_trigger();
  __shared__ float4 red[8][33];
  const int c4 = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c4 * 4 < C;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  float4 acc[FI + 1];
#pragma unroll
  for (int f = 0; f <= FI; ++f) acc[f] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
#pragma unroll 2
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(dz + r * C + c4 * 4));
      acc[0].x = __fadd_rn(acc[0].x, g.x); acc[0].y = __fadd_rn(acc[0].y, g.y); acc[0].z = __fadd_rn(acc[0].z, g.z); acc[0].w = __fadd_rn(acc[0].w, g.w);
#pragma unroll
      for (int f = 0; f < FI; ++f) {
        const float xv = __ldg(x + r * FI + f);
        acc[f + 1].x = fmaf(xv, g.x, acc[f + 1].x); acc[f + 1].y = fmaf(xv, g.y, acc[f + 1].y);
        acc[f + 1].z = fmaf(xv, g.z, acc[f + 1].z); acc[f + 1].w = fmaf(xv, g.w, acc[f + 1].w);
      }
    }
  }
#pragma unroll
  for (int f = 0; f <= FI; ++f) {
    red[threadIdx.y][threadIdx.x] = acc[f];
    __syncthreads();
    if (threadIdx.y == 0 && ok) {
      float4 t = red[0][threadIdx.x];
      for (int j = 1; j < 8; ++j) {
        const float4 v = red[j][threadIdx.x];
        t.x = __fadd_rn(t.x, v.x); t.y = __fadd_rn(t.y, v.y); t.z = __fadd_rn(t.z, v.z); t.w = __fadd_rn(t.w, v.w);
      }
      *reinterpret_cast<float4*>(partial + ((long long)blockIdx.y * (FI + 1) + f) * C + c4 * 4) = t;
    }
    __syncthreads();
  }
}
template <int FI>
bool launch_dwk_db(const float* dz, const float* x, float* out, long long rows, int C, float* ws, size_t ws_floats, cudaStream_t s) {
  const int col_blocks = (C / 4 + 31) / 32;
  long long chunks = (4LL * kNumSMs + col_blocks - 1) / col_blocks;
  if (chunks > (rows + 31) / 32) chunks = (rows + 31) / 32;
  if (chunks > (long long)(ws_floats / ((size_t)(FI + 1) * C))) chunks = (long long)(ws_floats / ((size_t)(FI + 1) * C));
  if (chunks < 1) return false;
  const long long rpc = (rows + chunks - 1) / chunks;
  chunks = (rows + rpc - 1) / rpc;
  launch_k(lstm_dwk_db_kernel<FI>, dim3(col_blocks, (unsigned)chunks), dim3(32, 8), 0, s, dz, x, ws, rows, C, rpc);
  SB_LAUNCHED();
  partial_reduce(ws, out, (FI + 1) * C, (int)chunks, s);
  return true;
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
  if (B * 4 * U >= (1ll << 31)) { delete w; throw CudaError("LSTM: batch x 4 x units must stay below 2^31 (32-bit element indices in the gate kernels)"); }
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
  if (F <= 8) w->dwkb_cat = dalloc((F + 1) * 4 * U);
  if (L.need_dx) w->dx_tm = dalloc(T * B * F);
  if (L.d.return_sequence) w->dout_tm = dalloc(T * B * U);
  if (L.d.stateful) w->c_init = dalloc(B * U);
  if (e->train.precision != SB_PREC_FP32 && B * U * 4 * U >= (1 << 22) && U % 32 == 0) {
    const int x3 = e->train.precision == SB_PREC_3XTF32 ? 1 : 0;
    w->g_fwd = tc_gemm_create(0, 1, w->h, (int)T + 1, w->wr_cat, 1, w->acts, (int)T, (int)B, (int)(4 * U), (int)U, 0, nullptr, 0, e->device, x3);
    // dH_{t-1} = dZ_t . Wr_cat^T sits on the serial path of BPTT and has few output tiles (cfg3: 16) over a long K (4U): deterministic
    // split-K over a private workspace (SB_LSTM_BWD_SPLITK=0: one CTA per tile, 32-column tiles -- 13 us per step instead of ~5)
    {
      const char* sk = getenv("SB_LSTM_BWD_SPLITK");
      const char* fe0 = getenv("SB_LSTM_FUSED_EPILOGUE");  // the fused-epilogue variant cannot split
      if (!(sk && sk[0] == '0') && !(fe0 && fe0[0] == '1') && !x3) {
        { const char* bf = getenv("SB_LSTM_BWD_FOLD"); w->bwd_fold = !(bf && bf[0] == '0'); }
        w->bwd_ws_floats = (size_t)8 * B * U;
        w->bwd_ws = dalloc((long long)w->bwd_ws_floats);
      }
    }
    w->g_bwd = tc_gemm_create(0, 0, w->acts, (int)T, w->wr_cat, 1, w->dh, 1, (int)B, (int)U, (int)(4 * U), w->bwd_ws ? 2 : 0, w->bwd_ws,
                              w->bwd_ws_floats, e->device, x3);
    w->g_dwr = tc_gemm_create(1, 1, w->h, 1, w->acts, 1, w->dwr_cat, 1, (int)U, (int)(4 * U), (int)(T * B), 1, e->ws, e->ws_floats,
                              e->device, x3);
    { const char* ex = getenv("SB_LSTM_EXACT_GATES"); w->fast_math = !x3 && w->g_fwd && !(ex && ex[0] == '1'); }
    const char* fe = getenv("SB_LSTM_FUSED_EPILOGUE");
    w->fused_epilogue = fe && fe[0] == '1' && w->g_fwd && w->g_bwd && !L.d.return_sequence && !L.d.stateful;
    // =2: the forward gates in the GEMM epilogue, the backward as separate kernels (keeps the split-K dH product)
    w->fused_fwd = w->fused_epilogue || (fe && fe[0] == '2' && w->g_fwd && !L.d.return_sequence && !L.d.stateful);
    if (!x3 && !w->fused_fwd && !w->xw && w->g_fwd)
      w->persist = lstm_persist_create((int)B, (int)T, (int)U, (int)F, w->h, w->wr_cat, e->device);
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
  if (w->fast_math)
    launch_k(lstm_gates_fwd_kernel<FI, true>, ew_grid(n), 256, 0, s, z, x_t, w->wk_cat, w->b_cat, xw_t, c_prev, c_out, h_out, w->B, w->U, first,
             act, ract, has_bias);
  else
    launch_k(lstm_gates_fwd_kernel<FI, false>, ew_grid(n), 256, 0, s, z, x_t, w->wk_cat, w->b_cat, xw_t, c_prev, c_out, h_out, w->B, w->U, first,
             act, ract, has_bias);
  SB_LAUNCHED();
}

}  // namespace

void lstm_plan(sb_engine* e) {
  for (auto& L : e->L)
    if (L.d.kind == SB_LAYER_LSTM) ws_of(e, L);  // allocate outside any stream capture
}

void lstm_forward(sb_engine* e, LayerPlan& L, int, bool store_state) {
  LstmWs* w = ws_of(e, L);
  cudaStream_t s = e->stream;
  const int B = w->B, T = w->T, F = w->F, U = w->U;
  const long long BU = (long long)B * U;
  const int act = L.d.activation, ract = L.d.recurrent_activation;
  // pack the per-gate tensors (the optimizer updated them since the last step)
  launch_k(lstm_pack_kernel, ew_grid((long long)U * 4 * U), 256, 0, s, gate_ptrs(e, L, 1, false), w->wr_cat, U, U, w->fast_math ? 1 : 0);
  SB_LAUNCHED();
  launch_k(lstm_pack_kernel, ew_grid((long long)F * 4 * U), 256, 0, s, gate_ptrs(e, L, 0, false), w->wk_cat, F, U, 0);
  SB_LAUNCHED();
  if (L.d.use_bias) {
    launch_k(lstm_pack_kernel, ew_grid(4LL * U), 256, 0, s, gate_ptrs(e, L, 2, false), w->b_cat, 1, U, 0);
    SB_LAUNCHED();
  }
  launch_k(lstm_transpose_kernel, ew_grid((long long)B * T * F), 256, 0, s, e->acts[L.buf_in], w->x_tm, B, T, F);
  SB_LAUNCHED();
  if (w->xw) gemm_nn(w->x_tm, w->wk_cat, w->xw, T * B, 4 * U, F, e->ws, e->ws_floats, s);
  // stateful: the state parameters seed the first step (layer/RNN.scala:66-67).  h_{-1} goes into slot 0 of the hidden states
  // (so H_prev_all stays one matrix for dWr), c_{-1} into its own buffer: the arena copies are overwritten by this very step.
  const bool stateful = L.d.stateful != 0;
  float* st_c = stateful ? e->arena + e->params[L.p_aux[12]].offset : nullptr;
  float* st_h = stateful ? e->arena + e->params[L.p_aux[13]].offset : nullptr;
  if (stateful) {
    SB_CUDA(cudaMemcpyAsync(w->h, st_h, (size_t)BU * 4, cudaMemcpyDeviceToDevice, s));
    SB_CUDA(cudaMemcpyAsync(w->c_init, st_c, (size_t)BU * 4, cudaMemcpyDeviceToDevice, s));
  }
  if (w->persist) {  // the whole time loop in one launch
    lstm_persist_forward(w->persist, w->x_tm, w->wk_cat, w->b_cat, stateful ? w->c_init : nullptr, w->acts, w->c, w->h, act, ract,
                         L.d.use_bias, stateful ? 0 : 1, s);
  }
  for (int t = 0; t < (w->persist ? 0 : T); ++t) {
    float* z = w->acts + (long long)t * B * 4 * U;
    const float* h_prev = w->h + (long long)t * BU;
    float* h_t = w->h + (long long)(t + 1) * BU;
    float* c_t = w->c + (long long)t * BU;
    const float* c_prev = t ? w->c + (long long)(t - 1) * BU : (stateful ? w->c_init : nullptr);
    if (t > 0 && w->fused_fwd && !w->xw) {
      // Z_t = H_{t-1} . Wr_cat with the gate math in the GEMM epilogue: activations -> acts[t], c_t, h_t (no separate kernel)
      LstmEpi ep{};
      ep.mode = 1; ep.F = F; ep.U = U; ep.act = act; ep.ract = ract; ep.has_bias = L.d.use_bias; ep.fast = w->fast_math ? 1 : 0;
      ep.x_t = w->x_tm + (long long)t * B * F; ep.wk_cat = w->wk_cat; ep.b_cat = w->b_cat;
      ep.c_prev = c_prev; ep.c_out = c_t; ep.h_out = h_t;
      tc_gemm_run_lstm(w->g_fwd, t, 0, t, ep, s);
      continue;
    }
    if (t > 0 || stateful) {  // stateless: h_{-1} = 0, no product at t = 0
      if (w->g_fwd) tc_gemm_run(w->g_fwd, t, 0, t, s);
      else gemm_nn(h_prev, w->wr_cat, z, B, 4 * U, U, e->ws, e->ws_floats, s);
    }
    const float* x_t = w->x_tm + (long long)t * B * F;
    const float* xw_t = w->xw ? w->xw + (long long)t * B * 4 * U : nullptr;
    const int first = t == 0 && !stateful;
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
  // a train step hands the last cell state on with the parameters (`nextParams = nextWeights ++ nextState`, Optimizer.scala:141-144)
  if (stateful && store_state) {
    SB_CUDA(cudaMemcpyAsync(st_h, w->h + (long long)T * BU, (size_t)BU * 4, cudaMemcpyDeviceToDevice, s));
    SB_CUDA(cudaMemcpyAsync(st_c, w->c + (long long)(T - 1) * BU, (size_t)BU * 4, cudaMemcpyDeviceToDevice, s));
  }
  if (L.d.return_sequence) {  // (batch, time, units) <- the time-major hidden states h_0 .. h_{T-1}
    launch_k(lstm_transpose_kernel, ew_grid((long long)T * BU), 256, 0, s, w->h + BU, e->acts[L.buf_out], T, B, U);
    SB_LAUNCHED();
  } else {  // the last hidden state
    SB_CUDA(cudaMemcpyAsync(e->acts[L.buf_out], w->h + (long long)T * BU, (size_t)BU * 4, cudaMemcpyDeviceToDevice, s));
  }
}

void lstm_backward(sb_engine* e, LayerPlan& L, int) {
  LstmWs* w = ws_of(e, L);
  cudaStream_t s = e->stream;
  const int B = w->B, T = w->T, F = w->F, U = w->U;
  const long long BU = (long long)B * U;
  const int act = L.d.activation, ract = L.d.recurrent_activation;
  const float* dh = e->dacts[L.buf_out];  // dLoss/dh_{T-1}
  const bool seq = L.d.return_sequence != 0;
  if (seq) {  // dLoss/d(output)[B, T, U] -> time-major; step t adds its slice to the recurrent gradient
    launch_k(lstm_transpose_kernel, ew_grid((long long)T * BU), 256, 0, s, e->dacts[L.buf_out], w->dout_tm, B, T, U);
    SB_LAUNCHED();
    dh = w->dout_tm + (long long)(T - 1) * BU;
  }
  int dh_parts = 1;
  for (int t = T - 1; t >= 0; --t) {
    float* a = w->acts + (long long)t * B * 4 * U;
    const float* c_t = w->c + (long long)t * BU;
    const bool stateful = L.d.stateful != 0;
    const float* c_prev = t ? w->c + (long long)(t - 1) * BU : (stateful ? w->c_init : nullptr);
    if (t == T - 1 || !w->fused_epilogue) {  // fused mode: the gate backward of step t < T-1 ran in the epilogue of step t+1
      const float* dh_seq = (seq && t < T - 1) ? w->dout_tm + (long long)t * BU : nullptr;
      if (w->fast_math)
        launch_k(lstm_gates_bwd_kernel<true>, ew_grid(BU), 256, 0, s, a, c_t, c_prev, dh, dh_seq, w->dc, B, U, t == 0 && !stateful, t == T - 1,
                 act, ract, dh_parts);
      else
        launch_k(lstm_gates_bwd_kernel<false>, ew_grid(BU), 256, 0, s, a, c_t, c_prev, dh, dh_seq, w->dc, B, U, t == 0 && !stateful, t == T - 1,
                 act, ract, dh_parts);
      SB_LAUNCHED();
    }
    if (t > 0) {
      if (w->fused_epilogue) {
        // dH_{t-1} = dZ_t . Wr_cat^T, consumed in the epilogue: gate backward of step t-1 (acts[t-1] -> dZ_{t-1}, dc -> dc_{t-2})
        LstmEpi ep{};
        ep.mode = 2; ep.F = F; ep.U = U; ep.act = act; ep.ract = ract; ep.has_bias = L.d.use_bias; ep.fast = w->fast_math ? 1 : 0;
        ep.acts_dz = w->acts + (long long)(t - 1) * B * 4 * U;
        ep.c_cur = w->c + (long long)(t - 1) * BU;
        ep.c_prev = t - 1 > 0 ? w->c + (long long)(t - 2) * BU : nullptr;
        ep.dc_io = w->dc;
        tc_gemm_run_lstm(w->g_bwd, t, 0, 0, ep, s);
      } else {
        dh = w->dh;
        dh_parts = 1;
        if (w->g_bwd && tc_gemm_is_split(w->g_bwd) && w->bwd_fold) dh_parts = tc_gemm_run_partials(w->g_bwd, t, 0, s, &dh);  // summed by the gate kernel
        else if (w->g_bwd) tc_gemm_run(w->g_bwd, t, 0, 0, s);  // dH_{t-1} = dZ_t . Wr_cat^T
        else gemm_nt(a, w->wr_cat, w->dh, B, U, 4 * U, e->ws, e->ws_floats, s);
      }
    }
  }
  // weight gradients over the whole sequence (acts now holds dZ_all[T*B, 4U]; h slots 0..T-1 are H_prev_all[T*B, U])
  if (w->g_dwr) tc_gemm_run(w->g_dwr, 0, 0, 0, s);
  else gemm_tn(w->h, w->acts, w->dwr_cat, U, 4 * U, T * B, e->ws, e->ws_floats, s);
  bool fused_kb = false;
  if (w->dwkb_cat) {  // db and dWk in one pass over dZ_all
    const long long R = (long long)T * B;
    switch (F) {
      case 1: fused_kb = launch_dwk_db<1>(w->acts, w->x_tm, w->dwkb_cat, R, 4 * U, e->ws, e->ws_floats, s); break;
      case 2: fused_kb = launch_dwk_db<2>(w->acts, w->x_tm, w->dwkb_cat, R, 4 * U, e->ws, e->ws_floats, s); break;
      case 3: fused_kb = launch_dwk_db<3>(w->acts, w->x_tm, w->dwkb_cat, R, 4 * U, e->ws, e->ws_floats, s); break;
      case 4: fused_kb = launch_dwk_db<4>(w->acts, w->x_tm, w->dwkb_cat, R, 4 * U, e->ws, e->ws_floats, s); break;
      case 5: fused_kb = launch_dwk_db<5>(w->acts, w->x_tm, w->dwkb_cat, R, 4 * U, e->ws, e->ws_floats, s); break;
      case 6: fused_kb = launch_dwk_db<6>(w->acts, w->x_tm, w->dwkb_cat, R, 4 * U, e->ws, e->ws_floats, s); break;
      case 7: fused_kb = launch_dwk_db<7>(w->acts, w->x_tm, w->dwkb_cat, R, 4 * U, e->ws, e->ws_floats, s); break;
      case 8: fused_kb = launch_dwk_db<8>(w->acts, w->x_tm, w->dwkb_cat, R, 4 * U, e->ws, e->ws_floats, s); break;
      default: break;
    }
  }
  const float* dwk_src = fused_kb ? w->dwkb_cat + 4 * U : w->dwk_cat;
  const float* db_src = fused_kb ? w->dwkb_cat : w->db_cat;
  if (!fused_kb) gemm_tn(w->x_tm, w->acts, w->dwk_cat, F, 4 * U, T * B, e->ws, e->ws_floats, s);
  launch_k(lstm_unpack_kernel, ew_grid((long long)U * 4 * U), 256, 0, s, w->dwr_cat, gate_ptrs(e, L, 1, true), U, U);
  SB_LAUNCHED();
  launch_k(lstm_unpack_kernel, ew_grid((long long)F * 4 * U), 256, 0, s, dwk_src, gate_ptrs(e, L, 0, true), F, U);
  SB_LAUNCHED();
  if (L.d.use_bias) {
    if (!fused_kb) col_sum(w->acts, w->db_cat, (long long)T * B, 4 * U, e->ws, e->ws_floats, s);
    launch_k(lstm_unpack_kernel, ew_grid(4LL * U), 256, 0, s, db_src, gate_ptrs(e, L, 2, true), 1, U);
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
    if (!L.lstm || L.d.kind != SB_LAYER_LSTM) continue;
    LstmWs* w = (LstmWs*)L.lstm;
    for (float* p : {w->wr_cat, w->wk_cat, w->b_cat, w->x_tm, w->acts, w->c, w->h, w->xw, w->dh, w->dc, w->dwr_cat, w->dwk_cat,
                     w->db_cat, w->dwkb_cat, w->dx_tm, w->dout_tm, w->c_init, w->bwd_ws})
      cudaFree(p);
    tc_gemm_destroy(w->g_fwd); tc_gemm_destroy(w->g_bwd); tc_gemm_destroy(w->g_dwr);
    lstm_persist_destroy(w->persist);
    delete w;
    L.lstm = nullptr;
  }
}

}  // namespace sb

namespace sb {
void trace_set_lstm(unsigned long long* p) { sb_trace_set_local(p); }  // SB_TRACE timeline buffer of this translation unit
}  // namespace sb
