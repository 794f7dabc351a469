// Kernel launchers of the scanet-b200 engine (sm_100a).  Every launcher enqueues on `s` and never syncs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Device-resident step bookkeeping: lets a whole epoch run without a host sync
// (reference: `iter = globalIter + localIter + 1`, optimizers/Optimizer.scala:143).
struct StepState {
  long long iter;       // last completed GLOBAL step t-1; the prologue kernel advances it
  long long batch_idx;  // next batch of the resident shard; advanced by the optimizer kernel
  float bc1;            // 1 - beta1^t   (`boost`, math/alg/AllKernels.scala:546-549)
  float bc2;            // 1 - beta2^t
  float loss;           // loss of the latest forward
  float pad;            // running sum of the L1/L2 penalties (reg_penalty_grad)
  // bias corrections of step `bc_iter`, computed one step ahead by a spare thread of the optimizer kernel (two double-precision
  // pow() calls, ~2 us of dependent latency that used to sit on the prologue kernel's critical path); the prologue takes them when
  // bc_iter matches the step it starts and recomputes otherwise (the host moved `iter`, another optimizer, first step)
  long long bc_iter;
  float bc1_next, bc2_next;
  long long avg_total_iters;  // sum of all ranks' iteration counts at the latest device-synchronised averaging (group.cu)
  int err;              // sticky device-side error (1: cross-GPU averaging barrier timed out, group.cu); read by sb_sync / sb_read_loss
  int pad2;
};

struct ConvGeom {
  int N, H, W, Cin, Cout, KH, KW, SH, SW, PT, PL, OH, OW;
};

struct PoolGeom {
  int N, H, W, C, KH, KW, SH, SW, PT, PL, OH, OW;
};

struct OptDesc {
  int kind;  // sb_optimizer
  float rate, beta1, beta2, epsilon;
  int nesterov;
  int minimizing;
};

// ---- step prologue: gather the batch out of the resident shard + advance iter / bias corrections ----
// prev_loss_slot (may be null): receives st->loss, i.e. the loss of the step BEFORE this one (steps chained in one graph)
void step_prologue(StepState* st, const float* ds_x, const float* ds_y, float* bx, float* by, long long x_per_batch,
                   long long y_per_batch, float beta1, float beta2, int batch_from_state, cudaStream_t s,
                   float* prev_loss_slot = nullptr);

// *slot = st->loss (last step of a host-fed group); by_cursor: slot[st->batch_idx - 1] = st->loss
void store_loss(const StepState* st, float* slot, cudaStream_t s, int by_cursor = 0);

// ---- fp32 FFMA contractions (parity mode; also the shapes tcgen05 does not fit) ----
// C[M,N] = A[M,K] . B[K,N]
void gemm_nn(const float* A, const float* B, float* C, int M, int N, int K, float* ws, size_t ws_floats, cudaStream_t s);
// C[M,N] = A[M,K] . B[N,K]^T            (dX = G . W^T)
void gemm_nt(const float* A, const float* B, float* C, int M, int N, int K, float* ws, size_t ws_floats, cudaStream_t s);
// C[M,N] = A[K,M]^T . B[K,N]            (dW = X^T . G), deterministic split-K through `ws`
void gemm_tn(const float* A, const float* B, float* C, int M, int N, int K, float* ws, size_t ws_floats, cudaStream_t s);
void conv2d_fwd(const float* x, const float* w, float* y, const ConvGeom& g, float* ws, size_t ws_floats, cudaStream_t s);
void conv2d_dgrad(const float* dy, const float* w, float* dx, const ConvGeom& g, float* ws, size_t ws_floats, cudaStream_t s);
void conv2d_wgrad(const float* x, const float* dy, float* dw, const ConvGeom& g, float* ws, size_t ws_floats, cudaStream_t s);

// ---- bandwidth-bound kernels ----
void bias_add(float* x, const float* b, long long rows, int C, cudaStream_t s);
void col_sum(const float* g, float* out, long long rows, int C, float* ws, size_t ws_floats, cudaStream_t s);
void act_fwd(float* x, int act, float thr, long long n, cudaStream_t s);
void act_bwd(float* g, const float* y, int act, float thr, long long n, cudaStream_t s);
void softmax_fwd(float* x, int rows, int C, cudaStream_t s);
void softmax_bwd(float* g, const float* p, int rows, int C, cudaStream_t s);
// loss value -> st->loss ; dpred (may be null) = dLoss/dpred.  rows = y.shape[0].
void loss_fwd_bwd(int loss, const float* pred, const float* y, float* dpred, StepState* st, long long rows, long long cols,
                  cudaStream_t s);
// fused Softmax + CategoricalCrossentropy head: p = softmax(z) in place, loss, dz
void softmax_cce_fwd_bwd(float* z_inout_p, const float* y, float* dz, StepState* st, int rows, int C, cudaStream_t s);
void pool_fwd(const float* x, float* y, const PoolGeom& g, int reduce, cudaStream_t s);
// max: gather form with the reference's first-max tie rule (math/nn/KernelsSpec.scala:277-288)
void pool_bwd(const float* x, const float* dy, float* dx, const PoolGeom& g, int reduce, cudaStream_t s);
// pool_bwd fused with the gradient of a ReLU that produced x in place:  dx = [x > thr] * poolgrad
void pool_bwd_relu(const float* x, const float* dy, float* dx, const PoolGeom& g, float thr, cudaStream_t s);

// batch norm over the last axis (C), rows = product of the other axes
void bn_fwd(const float* x, float* y, const float* gamma, const float* beta, float* mov_mean, float* mov_var,
            float* save_mean, float* save_var, long long rows, int C, float eps, float momentum, int fused, int bn_mode,
            int update_moving, float* ws, size_t ws_floats, cudaStream_t s);
// bn_pool.cu: one-pass batch statistics, and BatchNorm >> Pool2D(2x2, stride 2, VALID) as one forward / one backward (false: shape not covered)
bool bn_lean_ok(int C);
bool bn_pool_ok(const PoolGeom& g);
bool bn_stats(const float* x, long long rows, int C, float* save_mean, float* save_var, float* mov_mean, float* mov_var, float momentum,
              int fused, int bn_mode, int update_moving, float* ws, size_t ws_floats, cudaStream_t s);
bool bn_apply_pool_fwd(const float* x, float* y_pooled, unsigned char* map, const float* gamma, const float* beta, const float* mean,
                       const float* var, float eps, int fused, const PoolGeom& g, cudaStream_t s);
bool bn_pool_bwd(const float* x, const float* dy_pooled, const unsigned char* map, float* dx, const float* gamma, const float* save_mean,
                 const float* save_var, float* dgamma, float* dbeta, const PoolGeom& g, float eps, int fused, int bn_mode, int relu, float thr,
                 float* ws, size_t ws_floats, cudaStream_t s);
void bn_fwd_infer(const float* x, float* y, const float* gamma, const float* beta, const float* mov_mean, const float* mov_var,
                  long long rows, int C, float eps, int fused, cudaStream_t s);
void bn_bwd(const float* x, const float* dy, float* dx, const float* gamma, const float* save_mean, const float* save_var,
            float* dgamma, float* dbeta, long long rows, int C, float eps, int fused, int bn_mode, float* ws,
            size_t ws_floats, cudaStream_t s);

// L1/L2 penalty of one weight tensor (models/Regularization.scala:16-44; models/Model.scala:97):
//   st->loss += lambda * sum(|w| or w^2) / 2  -- penalties are summed in tensor order into a scratch first (first/last flags),
//   g (may be null) += dPenalty/dw  (L2: lambda*w ; L1: |lambda|/2, the reference's Abs gradient)
void reg_penalty_grad(const float* w, float* g, long long n, int kind, float lambda, StepState* st, int first, int last,
                      cudaStream_t s);

// estimators (estimators/package.scala:18-137) over one forwarded batch: pred,y [rows,C]; acc[4] (doubles) accumulates
//   kind 0 accuracy : acc[0] += rows where rint(pred) == y on every output
//   kind 1/2 squared / absolute error : acc[0] += sum (pred-y)^2 | |pred-y|
//   kind 3 R2 (C == 1): acc[0] += (pred-y)^2, acc[2] += pred, acc[3] += pred^2 ; acc[1] += rows always
void estimator_accumulate(const float* pred, const float* y, long long rows, int C, int kind, double* acc, cudaStream_t s);

// fused multi-tensor optimizer over the flat arena: w,g,state regions of n floats each
void optimizer_step(const OptDesc& o, float* w, const float* g, float* s0, float* s1, float* s2, long long n,
                    StepState* st, int advance_batch, cudaStream_t s);

// ---- init / utility ----
void fill_const(float* p, float v, long long n, cudaStream_t s);
void device_delay(long long ns, cudaStream_t s);  // profiling aid (not counted as a launch)
void scale_inplace(float* p, float w, long long n, cudaStream_t s);
// TF RandomUniform / RandomStandardNormal (Philox-4x32-10), out = sample*scale + shift
void philox_fill(float* p, long long n, unsigned long long seed, int normal, float scale, int has_scale, float shift,
                 int has_shift, cudaStream_t s);


}  // namespace sb

namespace sb {
// ---- kernels_fast.cu: vectorised / fused variants (return false when the shape is not covered) ----
// map != null (2x2 stride-1 pools): also records one winner byte per output and channel -- index of the first maximum | (maximum
// > thr) << 2 when `relu` (always set otherwise) -- for the map-driven backward kernels
void pool_max_fwd_v4(const float* x, float* y, const PoolGeom& g, cudaStream_t s, unsigned char* map = nullptr, int relu = 0,
                     float thr = 0.f);
// Pool2D(2x2, stride 1) backward fused with the input gradient of the classifier head Dense(N <= 16) that consumes the flattened
// pool output: dx[b,h,w,c] = sum over the windows won by (h,w) of (G[b,:] . W[f(window,c),:]).  false: shape not covered
bool pool_bwd_head_dgrad(const unsigned char* map, const float* gmat, const float* w, float* dx, const PoolGeom& g, int N, int relu,
                         cudaStream_t s);
// map != null (2x2 stride-1 pools whose forward recorded winner bytes): routed by the map, x is not read
void pool_max_bwd_v4(const float* x, const float* dy, float* dx, const PoolGeom& g, int relu, float thr, cudaStream_t s,
                     const unsigned char* map = nullptr);
bool pool22_map_ok(const PoolGeom& g);  // the winner-map kernels cover this pool geometry
// 2x2 stride-1 max-pool backward (+ReLU mask) that also accumulates the filter gradient of the small-K Conv2D feeding the pool
// (conv >> [in-place ReLU] >> pool); dx is written only when store_dx.  false: shape not covered
struct ReduceBatch;
bool pool_bwd_conv_wgrad(const float* x, const float* dy, float* dx, const PoolGeom& g, int relu, float thr, const float* conv_in,
                         const ConvGeom& cg, float* dw, int store_dx, float* ws, size_t ws_floats, cudaStream_t s,
                         const unsigned char* map = nullptr, ReduceBatch* defer = nullptr);
void conv_smallk_fwd(const float* x, const float* w, float* y, const ConvGeom& g, int relu, float thr, cudaStream_t s);
// conv (small K, stride 1) >> [ReLU] >> max-pool 2x2 stride 1 VALID: y = the (activated) conv output, pooled = the pool output.
// false: shape not covered
// map != null: one winner byte per pool output and channel (index of the first maximum | ReLU bit) is stored INSTEAD of y -- all the
// backward pass needs from the conv activations (pool_max_bwd_v4 / pool_bwd_conv_wgrad with the same map)
bool conv_smallk_pool_fwd(const float* x, const float* w, float* y, float* pooled, const ConvGeom& g, int relu, float thr,
                          unsigned char* map, cudaStream_t s);
bool conv_smallk_wgrad(const float* x, const float* dy, float* dw, const ConvGeom& g, float* ws, size_t ws_floats, cudaStream_t s);
bool dense_skinny_fwd(const float* x, const float* w, float* z, int M, int N, int K, float* ws, size_t ws_floats, cudaStream_t s);
// split-K partials only (partial[chunks][M][N]); the consumer (head kernel or partial_reduce) sums them in chunk order
bool dense_head_fwd(const float* x, const float* w, float* partial, int* chunks_out, int M, int N, int K, size_t ws_floats,
                    cudaStream_t s);
// wgrad and (dx != null) dgrad in one pass
bool dense_head_bwd(const float* x, const float* g, const float* w, float* dw, float* dx, int M, int N, int K, float* ws,
                    size_t ws_floats, cudaStream_t s);
// z (or the sum of `chunks` split-K partials when partial != null) + bias -> softmax -> CCE loss, dz, dbias
// scratch: kHeadScratchFloats floats, zero-initialised once (per-CTA partials + the completion ticket)
constexpr size_t kHeadScratchFloats = 1024 * 40 + 32;
// deferred_ctas != null: the kernel leaves its per-CTA loss / bias-gradient partials in `scratch` and returns their count; the
// caller runs head_loss_combine (any stream ordered after this kernel) to form st->loss and dbias -- same summation, same bits
bool head_bias_softmax_cce(float* z, const float* partial, int chunks, const float* bias, const float* y, float* dz, float* dbias,
                           StepState* st, int rows, int C, float* scratch, cudaStream_t s, int* deferred_ctas = nullptr);
void head_loss_combine(const float* scratch, int n_ctas, int C, int rows, StepState* st, float* dbias, cudaStream_t s);
// g <- act'(y)*g in place and out[c] = column sums of the result, one pass (act = 0: plain column sum); false if C % 4 != 0
bool act_bwd_col_sum(float* g, const float* y, int act, float thr, float* out, long long rows, int C, float* ws, size_t ws_floats,
                     cudaStream_t s);
// out[i] = sum_c partial[c*n + i], fixed order
void partial_reduce(const float* partial, float* out, int n, int parts, cudaStream_t s);
// Reductions that are all due at the same point of the step (the filter-gradient partials of several layers, just before the
// optimizer) collected into ONE launch: each is summed in the order partial_reduce uses for > 48 parts (32 lane groups), so
// only reductions with > 48 parts are accepted (same bits as the separate launch).
struct ReduceBatch {
  static constexpr int kMax = 4;
  const float* partial[kMax];
  float* out[kMax];
  int n[kMax], parts[kMax], blocks[kMax];
  int count = 0;
  bool add(const float* p, float* o, int n_, int parts_) {
    if (count >= kMax || parts_ <= 48) return false;
    partial[count] = p; out[count] = o; n[count] = n_; parts[count] = parts_; blocks[count] = 0;
    ++count;
    return true;
  }
};
void partial_reduce_multi(ReduceBatch& rb, cudaStream_t s);  // no-op when empty; leaves the batch empty
}  // namespace sb
