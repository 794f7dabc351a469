// Internal: tensor-core (tcgen05) fast paths, the LSTM macro-op and host-side initialisers.
#pragma once
#include <stdint.h>
typedef struct CUstream_st* cudaStream_t;  // as in driver_types.h (host_init.cpp is built without the CUDA headers)
struct sb_engine;
namespace sb {
struct LayerPlan;
// decide per layer whether a tcgen05 kernel applies (precision != FP32 and the shape fits); build TMA maps
void tc_plan_layers(sb_engine* e);
void tc_free_layers(sb_engine* e);
void tc_params_changed(sb_engine* e);
// each returns false when the layer has no tensor-core path (the caller then runs the fp32 kernel)
bool tc_dense_fwd(sb_engine* e, LayerPlan& L, int li);
bool tc_dense_wgrad(sb_engine* e, LayerPlan& L, int li);
bool tc_dense_dgrad(sb_engine* e, LayerPlan& L, int li);
bool tc_conv_fwd(sb_engine* e, LayerPlan& L, int li);
bool tc_conv_wgrad(sb_engine* e, LayerPlan& L, int li);
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
void tc_gemm_destroy(TcGemm* g);

void lstm_plan(sb_engine* e);
void lstm_forward(sb_engine* e, LayerPlan& L, int li);
void lstm_backward(sb_engine* e, LayerPlan& L, int li);
void lstm_free(sb_engine* e);

// Orthogonal initializer (models/Initializer.scala:202-219): QR of a TF-Philox normal matrix on the host
void orthogonal_host(float* out, int64_t rows, int64_t cols, unsigned long long seed, float gain);
}  // namespace sb
