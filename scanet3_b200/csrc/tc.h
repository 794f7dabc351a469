// Internal: tensor-core (tcgen05) fast paths, the LSTM macro-op and host-side initialisers.
#pragma once
#include <stdint.h>
typedef struct CUstream_st* cudaStream_t;  // as in driver_types.h (host_init.cpp is built without the CUDA headers)
struct sb_engine;
namespace sb {
struct LayerPlan;
struct ReduceBatch;  // kernels.cuh
// decide per layer whether a tcgen05 kernel applies (precision != FP32 and the shape fits); build TMA maps
void tc_plan_layers(sb_engine* e);
void tc_free_layers(sb_engine* e);
void tc_params_changed(sb_engine* e);
// each returns false when the layer has no tensor-core path (the caller then runs the fp32 kernel)
bool tc_dense_fwd(sb_engine* e, LayerPlan& L, int li);
bool tc_dense_wgrad(sb_engine* e, LayerPlan& L, int li);
bool tc_dense_dgrad(sb_engine* e, LayerPlan& L, int li);
// skinny classifier head Z[M<=128, N<=16] = X[M,K] . W[K,N] as split-K partials[chunks][M][N] on tcgen05 (TF32 mode, K % 32 == 0)
bool tc_head_fwd(sb_engine* e, LayerPlan& L, const float* x, const float* w, float* partial, size_t partial_floats, int* chunks_out);
// its backward: dW[K,N] = X^T . G and (dx != null) dX[M,K] = G . W^T in one kernel (TF32 mode, M <= 128, K % 128 == 0)
bool tc_head_bwd(sb_engine* e, LayerPlan& L, const float* x, const float* g, const float* w, float* dw, float* dx, cudaStream_t st);
bool tc_conv_fwd(sb_engine* e, LayerPlan& L, int li);
bool tc_conv_wgrad(sb_engine* e, LayerPlan& L, int li, cudaStream_t side = nullptr, ReduceBatch* defer = nullptr);  // s: side stream (null = the engine's)
bool tc_conv_dgrad(sb_engine* e, LayerPlan& L, int li);

// ---- tc_gemm.cu: tcgen05 TF32 GEMM plans (3-D tensor maps: the outer coordinate picks a slice, e.g. an LSTM time step) ----
struct TcGemm;
// C[M,N] = op(A).op(B).  a_mn_major: A stored [K][M] (else [M][K]); b_mn_major: B stored [K][N] (else [N][K]).
// Returns null when the shape / alignment does not fit (caller falls back to the fp32 kernels).
TcGemm* tc_gemm_create(int a_mn_major, int b_mn_major, const float* A, int a_slices, const float* B, int b_slices, float* C,
                       int c_slices, int M, int N, int K, int allow_split, float* ws, size_t ws_floats, int device, int x3 = 0);
void tc_gemm_set_epilogue(TcGemm* g, const float* bias, int act, float thr);
bool tc_gemm_is_split(const TcGemm* g);
void tc_gemm_run(TcGemm* g, int za, int zb, int zc, cudaStream_t s);
// split-K plans (tc_gemm_is_split): launch without the reduction; returns the number of partials [splits][M*N] left in the workspace
int tc_gemm_run_partials(TcGemm* g, int za, int zb, cudaStream_t s, const float** partial);
// LSTM pointwise work fused into the GEMM epilogue (columns are unit-interleaved: col = 4*unit + gate, csrc/lstm.cu):
//  mode 1 (forward, C = Z_t = H_{t-1}.Wr_cat): z += x_t.Wk_cat + b ; activations written to C in place of z ; c_t, h_t stored
//  mode 2 (backward, C = dH_{t-1} = dZ_t.Wr_cat^T, not stored): the gate backward of step t-1 -> dZ_{t-1} in place of acts_{t-1}
struct LstmEpi {
  int mode;
  int F, U, act, ract, has_bias;
  int fast;             // approximate exp / reciprocal in the activations (TF32 mode; the same forms as lstm.cu actf_t)
  const float* x_t;     // [B, F]   (mode 1)
  const float* wk_cat;  // [F, 4U]
  const float* b_cat;   // [4U]
  const float* c_prev;  // [B, U]   mode 1: c_{t-1} ; mode 2: c_{t-2} (null when step t-1 is the first)
  float* c_out;         // [B, U]   mode 1: c_t
  float* h_out;         // [B, U]   mode 1: h_t
  float* acts_dz;       // [B, 4U]  mode 2: acts_{t-1} in, dZ_{t-1} out
  const float* c_cur;   // [B, U]   mode 2: c_{t-1}
  float* dc_io;         // [B, U]   mode 2: dc_t in, dc_{t-1} out
};
void tc_gemm_run_lstm(TcGemm* g, int za, int zb, int zc, const LstmEpi& epi, cudaStream_t s);
void tc_gemm_destroy(TcGemm* g);

// halo fprop with the Pool2D(2x2, stride 1, VALID) that follows folded into its epilogue (tc.cu); false: not available
bool tc_conv_enable_pool_fusion(sb_engine* e, LayerPlan& L, float* pooled, unsigned char* map);

// persistent LSTM time loop (lstm_persist.cu): one launch for all T forward steps; create returns null when the shape is not covered
void* lstm_persist_create(int B, int T, int U, int F, float* h, const float* wr_cat, int device);
void lstm_persist_destroy(void* plan);
void lstm_persist_forward(void* plan, const float* x_tm, const float* wk_cat, const float* b_cat, const float* c_init, float* acts,
                          float* c, float* h, int act, int ract, int has_bias, int zero_state, cudaStream_t s);

void lstm_plan(sb_engine* e);
void lstm_forward(sb_engine* e, LayerPlan& L, int li, bool store_state);  // store_state: a train step of a stateful RNN
void lstm_backward(sb_engine* e, LayerPlan& L, int li);
void lstm_free(sb_engine* e);
// RNN(SimpleRNNCell >> Bias >> Activate) (rnn.cu)
void rnn_plan(sb_engine* e);
void rnn_forward(sb_engine* e, LayerPlan& L, int li, bool store_state);
void rnn_backward(sb_engine* e, LayerPlan& L, int li);
void rnn_free(sb_engine* e);

// Orthogonal initializer (models/Initializer.scala:202-219): QR of a TF-Philox normal matrix on the host
void orthogonal_host(float* out, int64_t rows, int64_t cols, unsigned long long seed, float gain);
}  // namespace sb
