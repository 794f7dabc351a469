// Internal: tensor-core (tcgen05) fast paths, the LSTM macro-op and host-side initialisers.
#pragma once
#include <stdint.h>
struct sb_engine;
namespace sb {
struct LayerPlan;
// decide per layer whether a tcgen05 kernel applies (precision != FP32 and the shape fits); build TMA maps
void tc_plan_layers(sb_engine* e);
void tc_free_layers(sb_engine* e);
void tc_params_changed(sb_engine* e);
// each returns false when the layer has no tensor-core path (the caller then runs the fp32 kernel)
bool tc_dense_fwd(sb_engine* e, LayerPlan& L, int li);
bool tc_dense_bwd(sb_engine* e, LayerPlan& L, int li);
bool tc_conv_fwd(sb_engine* e, LayerPlan& L, int li);
bool tc_conv_wgrad(sb_engine* e, LayerPlan& L, int li);
bool tc_conv_dgrad(sb_engine* e, LayerPlan& L, int li);

void lstm_plan(sb_engine* e);
void lstm_forward(sb_engine* e, LayerPlan& L, int li);
void lstm_backward(sb_engine* e, LayerPlan& L, int li);
void lstm_free(sb_engine* e);

// Orthogonal initializer (models/Initializer.scala:202-219): QR of a TF-Philox normal matrix on the host
void orthogonal_host(float* out, int64_t rows, int64_t cols, unsigned long long seed, float gain);
}  // namespace sb
