/* scanet_b200.h -- C ABI of the B200-native train engine for scanet3's data-parallel hot path.
 *
 * The reference (pashashiz/scanet3) has no FFI seam of its own for this path: its only native
 * boundary is TensorFlow-Java's generic JNI behind `Session.runner.feed().fetch().run()`
 * (src/main/scala/scanet/core/Session.scala:83-94).  The natural seam is the trio of compiled
 * closures the optimizer obtains and calls (src/main/scala/scanet/optimizers/Optimizer.scala):
 *
 *   backprop : (x, y, weights, optParams, state, iter) => (nextWeights, nextOpt, nextState)   :132,142-143,169-194
 *   loss     : (x, y, params) => (loss, state)                                                :133-134,148,163-167
 *   agg      : (leftParams, rightParams) => averagedParams  (model + optimizer state)        :196-230
 *
 * plus what surrounds them in `optimizeOnPartition` (:106-161): epoch-0 initialisation, the
 * TensorIterator batch loop (optimizers/Iterators.scala:39-92) and the global step counter.
 * Each entry point below names the reference lines it replaces.  The JVM-side binding a
 * maintainer would add (Panama FFM) is shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every tensor is fp32, C-contiguous row-major, conv
 * tensors NHWC, filters HWIO -- the layout of the reference's `Tensor` direct buffers
 * (core/TensorCodec.scala:31-41).  Every function returns an `sb_status`; the message of the last
 * failure on the calling thread is `sb_last_error()`.  There is NO CPU fallback: without a CUDA
 * device every compute entry point fails with SB_ERR_CUDA.
 *
 * Threading: one `sb_engine` = one worker = one GPU (one Spark partition task,
 * Optimizer.scala:75-85).  An engine is single-owner (no internal locking); different engines may be
 * driven concurrently from different threads.  `sb_group_*` is called by one coordinator thread.
 */
#ifndef SCANET_B200_H
#define SCANET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_ABI_VERSION 4

typedef enum {
  SB_OK = 0,
  SB_ERR_INVALID_ARG = 1, /* reference: IllegalArgumentException via require(...) (core/Require.scala:5-6) */
  SB_ERR_UNSUPPORTED = 2, /* a reference feature outside the accelerated path (no CPU fallback)          */
  SB_ERR_CUDA = 3,        /* reference: TF_Status -> RuntimeException (native/package.scala:6-9)          */
  SB_ERR_NCCL = 4,
  SB_ERR_STATE = 5,       /* call order violated (e.g. train before params are initialised)               */
  SB_ERR_TIMEOUT = 6      /* a peer rank never reached the device-synchronised averaging barrier           */
} sb_status;

/* ---- model descriptor: the flattened leaf-layer list of `Composed` (layer/Composed.scala:89-99) ---- */
typedef enum {
  SB_LAYER_DENSE = 1,     /* layer/Dense.scala:42-71   params: "weights" [in,out]                        */
  SB_LAYER_BIAS = 2,      /* layer/Bias.scala:22-43    params: "weights" [features]                      */
  SB_LAYER_ACTIVATE = 3,  /* layer/Activate.scala:12   no params                                         */
  SB_LAYER_FLATTEN = 4,   /* layer/Flatten.scala:10    no params                                         */
  SB_LAYER_CONV2D = 5,    /* layer/Conv2D.scala:70-125 params: "weights" [KH,KW,Cin,filters]             */
  SB_LAYER_POOL2D = 6,    /* layer/Pool2D.scala:28-57  no params; always Max (the layer ignores `reduce`) */
  SB_LAYER_BATCHNORM = 7, /* layer/BatchNorm.scala:63-174 params: beta,gamma (+state moving_mean/variance) */
  SB_LAYER_LSTM = 8,      /* layer/RNN.scala:208-339   RNN(LSTMCell): per gate kernel/recurrent/bias       */
  SB_LAYER_SIMPLE_RNN = 9 /* layer/RNN.scala:96-206    RNN(SimpleRNNCell >> Bias >> act): "0/kernel_weights" [F,U],
                             "0/recurrent_weights" [U,U], "1/weights" [U]; the state fed back is the PRE-bias,
                             PRE-activation sum (:190-196), as in the reference                                   */
} sb_layer_kind;

typedef enum { SB_ACT_IDENTITY = 0, SB_ACT_SIGMOID = 1, SB_ACT_TANH = 2, SB_ACT_SOFTMAX = 3, SB_ACT_RELU = 4 } sb_activation;
/* math/nn/AllKernels.scala:39-103.  EXPLICIT (Conv2D only; Pool2D refuses it like the reference, :270-277) takes the four
 * pad_* fields of the layer: `Explicit(height = (top, bottom), width = (left, right))`. */
typedef enum { SB_PAD_VALID = 0, SB_PAD_SAME = 1, SB_PAD_EXPLICIT = 2 } sb_padding;
typedef enum { SB_POOL_MAX = 0, SB_POOL_AVG = 1 } sb_pool_reduce;
/* models/Regularization.scala:16-44: penalty lambda*sum|w|/2 (L1) or lambda*sum(w^2)/2 (L2) added to the loss
 * (models/Model.scala:97).  Gradients follow the reference's autodiff: L2 -> lambda*w; L1 -> |lambda|/2 per element
 * (the reference's Abs.localGrad is |g|, math/alg/AllKernels.scala:330-336 -- reproduced). */
typedef enum { SB_REG_ZERO = 0, SB_REG_L1 = 1, SB_REG_L2 = 2 } sb_regularization;

/* models/Initializer.scala:28-220.  Random kinds are bit-compatible with TF's seeded Philox ops
 * (math/rand/AllKernels.scala:12-43); an unseeded random initializer is rejected. */
typedef enum {
  SB_INIT_ZEROS = 0,
  SB_INIT_ONES = 1,
  SB_INIT_CONST = 2,          /* a = value                     */
  SB_INIT_RANDOM_UNIFORM = 3, /* a = min, b = max              */
  SB_INIT_RANDOM_NORMAL = 4,  /* a = mean, b = std             */
  SB_INIT_GLOROT_UNIFORM = 5,
  SB_INIT_HE_UNIFORM = 6,
  SB_INIT_LECUN_UNIFORM = 7,
  SB_INIT_ORTHOGONAL = 8,     /* a = gain (0 => none); QR on the host, values uploaded */
  SB_INIT_TRUNCATED_NORMAL = 9, /* a = mean, b = std (RandomNormalTruncated) */
  SB_INIT_GLOROT_NORMAL = 10, /* truncated normal, std = sqrt(2/(fan_in+fan_out)) / 0.8796... */
  SB_INIT_HE_NORMAL = 11,     /* std = sqrt(2/fan_in) / 0.8796...  */
  SB_INIT_LECUN_NORMAL = 12   /* std = sqrt(1/fan_in) / 0.8796...  */
} sb_init_kind;

typedef struct {
  int32_t kind;     /* sb_init_kind */
  int32_t has_seed; /* `seed: Option[Long]` */
  int64_t seed;
  float a, b;
} sb_initializer;

typedef struct {
  int32_t kind;                 /* sb_layer_kind */
  int32_t units;                /* Dense outputs | Bias features | Conv2D filters | LSTM units */
  int32_t activation;           /* Activate: sb_activation; LSTM: cell activation (default Tanh) */
  int32_t recurrent_activation; /* LSTM (default Sigmoid) */
  float relu_threshold;         /* ReLU(threshold) (models/Activation.scala:80-83) */
  int32_t window_h, window_w;   /* Conv2D kernel | Pool2D window */
  int32_t stride_h, stride_w;
  int32_t padding;              /* sb_padding */
  int32_t pool_reduce;          /* sb_pool_reduce, recorded but ignored like the reference layer unless pool_honor_reduce */
  int32_t use_bias;             /* LSTM, SimpleRNN */
  int32_t trainable;
  float bn_momentum, bn_epsilon;
  int32_t bn_mode;              /* 0 = reference (bug-for-bug fused wiring), 1 = correct */
  int32_t bn_fused;             /* -1 = None (auto), 0 = Some(false), 1 = Some(true) */
  sb_initializer init[4];       /* [0] kernel | gamma, [1] recurrent | beta, [2] bias | moving_mean,
                                   [3] forget-bias | moving_variance */
  int32_t reg_kind;             /* sb_regularization of this layer's "weights" (Dense kernel, Bias) */
  float reg_lambda;
  int32_t pad_top, pad_bottom, pad_left, pad_right; /* Conv2D with SB_PAD_EXPLICIT (math/nn/AllKernels.scala:67-102) */
  int32_t return_sequence;      /* LSTM, SimpleRNN: output (batch, time, units) instead of the last step (layer/RNN.scala:21-27, 69-71) */
  int32_t pool_honor_reduce;    /* Pool2D: 1 = apply `pool_reduce` (Avg pooling works); 0 = the reference layer's behaviour, which never
                                   forwards `reduce` and therefore always max-pools (layer/Pool2D.scala:36-42) */
  int32_t stateful;             /* LSTM, SimpleRNN: the state tensors ("cell_state", "hidden_state" | "0/state", [batch, units],
                                   Zeros, aggregation None) are parameters: they seed the first step and every train step stores
                                   the last cell state back (layer/RNN.scala:32-35, 66-72; Optimizer.scala:141-144) */
} sb_layer_desc;

typedef struct {
  int32_t n_layers;
  const sb_layer_desc* layers;
  int32_t input_rank;  /* per-sample feature shape (`ds.shapes._1`, optimizers/SparkExt.scala:36-44) */
  int64_t input_dims[4];
  int32_t label_rank;
  int64_t label_dims[4];
} sb_model_desc;

typedef enum { SB_LOSS_MSE = 0, SB_LOSS_BCE = 1, SB_LOSS_CCE = 2 } sb_loss; /* models/Loss.scala:18-57 */

/* optimizers/{SGD,Adam,AdaGrad,RMSProp,AdaDelta,Adamax,Nadam,AMSGrad}.scala */
typedef enum {
  SB_OPT_SGD = 0, SB_OPT_ADAM = 1, SB_OPT_ADAGRAD = 2, SB_OPT_RMSPROP = 3,
  SB_OPT_ADADELTA = 4, SB_OPT_ADAMAX = 5, SB_OPT_NADAM = 6, SB_OPT_AMSGRAD = 7
} sb_optimizer;

/* Arithmetic of the contractions (MatMul / Conv2D fwd, dgrad, wgrad). */
typedef enum {
  SB_PREC_FP32 = 0,  /* fp32 FFMA kernels: parity mode */
  SB_PREC_TF32 = 1,  /* tcgen05 kind::tf32, fp32 accumulate in TMEM (throughput mode) where the shape fits */
  SB_PREC_3XTF32 = 2 /* MatMul: split-operand 3xTF32 on tcgen05 (fp32-class accuracy); Conv2D: the fp32 FFMA kernels */
} sb_precision;

typedef struct {
  int32_t batch;     /* `.batch(n)` (Optimizer.scala:296-297) */
  int32_t loss;      /* sb_loss */
  int32_t optimizer; /* sb_optimizer */
  float rate;        /* SGD/Adam/AdaGrad/RMSProp/AdaDelta/Adamax/AMSGrad `rate` (Nadam has none) */
  float beta1;       /* Adam/Adamax/Nadam/AMSGrad beta1 | SGD momentum | RMSProp/AdaDelta rho */
  float beta2;
  float epsilon;
  float init_acc;    /* Adamax/Nadam initial accumulator (Const(initAcc)) */
  int32_t nesterov;  /* SGD */
  int32_t minimizing; /* Optimizer.minimize (1): w - delta ; maximize (0): w + delta (Optimizer.scala:186,326-338) */
  int32_t precision; /* sb_precision */
  int32_t use_graph; /* 1: capture the step into a CUDA graph (default), 0: eager launches */
} sb_train_desc;

typedef struct sb_engine sb_engine;
typedef struct sb_group sb_group;

/* thread-local message of the last failure */
const char* sb_last_error(void);
int sb_abi_version(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int sb_device_count(void);

/* Builds the flat parameter arena, streams, kernels' workspaces and the captured step graphs.
 * Replaces `model.params`, `alg.params`, `compileBackprop`, `compileLoss`
 * (Optimizer.scala:116-134,163-194).  `device` < 0 builds a PLAN-ONLY engine (no CUDA calls): it
 * answers sb_param_* / sb_output_dims so host code can be exercised without a GPU; every compute
 * call on it fails with SB_ERR_CUDA. */
int sb_engine_create(const sb_model_desc* model, const sb_train_desc* train, int device, sb_engine** out);
void sb_engine_destroy(sb_engine* e);

/* ---- parameters by the reference's path strings (core/Params.scala:3-81; Optimizer.scala:117-122) ---- */
typedef enum { SB_ROLE_WEIGHT = 0, SB_ROLE_STATE = 1, SB_ROLE_OPT = 2 } sb_param_role;
typedef enum { SB_AGG_NONE = 0, SB_AGG_AVG = 1 } sb_aggregation;

int sb_param_count(const sb_engine* e);
/* path: caller buffer of `path_cap` bytes (NUL-terminated on return); dims: up to 4 entries */
int sb_param_info(const sb_engine* e, int index, char* path, size_t path_cap, int64_t* dims, int32_t* rank,
                  int32_t* role, int32_t* aggregation, int64_t* arena_offset);
int sb_output_dims(const sb_engine* e, int64_t* dims, int32_t* rank); /* model.outputShape incl. batch */

/* H2D / D2H of one tensor (`initParams`, broadcast in, `StepResult` out, `Tensor` read-back:
 * Optimizer.scala:72-73,124-126,149-155; core/Tensor.scala:13-32). n = element count (checked). */
int sb_params_set(sb_engine* e, const char* path, const float* host, size_t n);
int sb_params_get(sb_engine* e, const char* path, float* host, size_t n);
/* run the declared initializers on the device (Optimizer.scala:123-131; optimizer state too) */
int sb_init_params(sb_engine* e);
/* only the optimizer state (when model params came from `initParams`) */
int sb_init_opt_params(sb_engine* e);

/* ---- parameter wire format + checkpoint / resume ----
 * One tensor record is byte-for-byte the reference's Kryo `TensorSerializer.write` (optimizers/KryoSerializers.scala:33-49,
 * Kryo 4.0.2: fixed-width big-endian ints): int32 type code (DT_FLOAT = 1) | int32 rank | int64 dims[rank] | int32 payload
 * bytes | row-major little-endian fp32 payload (core/Tensor.scala:46-52) -- `TensorSerializer.read` (:51-64) rebuilds the
 * `Tensor[Float]`.  buf == NULL in an encode call only reports the size in *written. */
int sb_wire_encode(const int64_t* dims, int32_t rank, const float* data, void* buf, size_t cap, size_t* written);
/* data == NULL: parse the header only (rank, dims, *consumed = record size) */
int sb_wire_decode(const void* buf, size_t n, int64_t* dims, int32_t dims_cap, int32_t* rank, float* data, size_t data_cap,
                   size_t* consumed);
int sb_param_to_wire(sb_engine* e, const char* path, void* buf, size_t cap, size_t* written);
int sb_param_from_wire(sb_engine* e, const char* path, const void* buf, size_t n, size_t* consumed);
/* Every tensor of the arena (weights, layer state, optimizer state) keyed by the reference's path strings, plus the global
 * iteration and epoch (`Step`, Optimizer.scala:17-24), in one file: "SB3CKPT1" | int64 BE iter | int64 BE epoch | int32 BE n | n x { int32 BE path bytes | path | tensor record }.
 * The reference has no checkpointing (README.md:112); a resumed run continues bit-identically (tests/test_gpu_parity.py). */
int sb_checkpoint_save(sb_engine* e, const char* file, int64_t global_iter, int64_t epoch);
int sb_checkpoint_load(sb_engine* e, const char* file, int64_t* global_iter, int64_t* epoch);

/* ---- the train path ---- */
/* shard resident in HBM (the partition iterator + TensorIterator, Iterators.scala:39-83).
 * x: [n_rows, prod(input_dims)], y: [n_rows, prod(label_dims)]; host memory (pinned or pageable). */
int sb_dataset_upload(sb_engine* e, const float* x, const float* y, int64_t n_rows);
#define SB_ITER_FROM_DEVICE (-1)
/* one epoch of `optimizeOnPartition` (Optimizer.scala:135-157): all FULL batches in row order
 * (partial tail dropped), iter = global_iter + j + 1, then the forward-only loss on the last batch
 * with the updated params.  No host sync inside the epoch.  Fewer rows than one batch ->
 * SB_ERR_INVALID_ARG (the reference throws NoSuchElementException, Iterators.scala:79-81).
 * global_iter may be SB_ITER_FROM_DEVICE (see sb_train_steps_async). */
int sb_train_epoch(sb_engine* e, int64_t global_iter, int64_t* iterations_out, float* last_loss_out);
/* asynchronous variant: enqueue `n_steps` batches starting at batch `first_batch` of the shard.
 * global_iter == SB_ITER_FROM_DEVICE: the device's own step counter continues (after sb_group_average_device it already
 * holds globalIter + the SUM of all ranks' iteration counts, Optimizer.scala:88,223). */
int sb_train_steps_async(sb_engine* e, int64_t global_iter, int64_t first_batch, int64_t n_steps);
/* Capture + instantiate + upload the CUDA graphs of the train path now (the reference's analogue: `compileBackprop` /
 * `compileLoss` are compiled once per shape signature and cached, Optimizer.scala:132-134, TF.scala:401-417) instead of
 * lazily inside the first epoch.  Nothing is executed.  SB_WARM_RESIDENT needs an uploaded shard. */
enum { SB_WARM_RESIDENT = 1, SB_WARM_LOSS = 2, SB_WARM_HOSTFED = 4, SB_WARM_ALL = 7 };
int sb_warm(sb_engine* e, int what);
/* single `backprop` call on HOST buffers x[batch,...], y[batch,...] with iter = `iter` (1-based global
 * step): H2D copy, forward, loss, backward, update.  `loss_out` (optional) receives the PRE-update loss
 * of this batch after a D2H read (one sync). */
int sb_train_step(sb_engine* e, const float* x, const float* y, int64_t iter, float* loss_out);
/* asynchronous `backprop` on HOST buffers: enqueues the H2D copies of x,y (pinned memory for real overlap), the
 * step, and a D2H read of the PRE-update loss into `loss_out_host` (pinned; may be NULL).  No sync: call sb_sync
 * before reading the loss or reusing x / y. */
int sb_train_step_async(sb_engine* e, const float* x, const float* y, int64_t iter, float* loss_out_host);
/* The `optimize` loop over HOST batches (Optimizer.scala:135-147): n_batches consecutive batches x[n_batches*batch,...],
 * y[n_batches*batch,...] (pinned memory for real overlap) stream through a double-buffered staging area in groups of up to 8;
 * a group is one H2D copy pair on the copy stream and one CUDA graph of that many steps on the compute stream, so the copy of
 * group g+1 overlaps the steps of group g.  Batch j runs with iter = global_iter + j + 1.  losses_out_host (optional, pinned,
 * n_batches floats) receives the PRE-update loss of every batch.  No sync: call sb_sync before reading them or reusing x / y. */
int sb_train_batches_async(sb_engine* e, const float* x, const float* y, int64_t n_batches, int64_t global_iter,
                           float* losses_out_host);
/* the `loss` closure on HOST buffers (training-mode forward, state discarded) */
int sb_eval_loss(sb_engine* e, const float* x, const float* y, float* loss_out);
/* forward only (`TrainedModel.result`, models/Model.scala:154-163): out[batch, ...] on the host */
int sb_predict(sb_engine* e, const float* x, float* out, size_t n_out);
/* Estimators on the device (estimators/package.scala:18-137; `RecordAccuracy`, optimizers/Effect.scala:52-64): forwards
 * n_rows rows in inference mode, batch by batch (the tail batch is partial, `remaining = Partial`), and reduces on the GPU.
 * x == NULL: the rows of the resident shard (sb_dataset_upload), n_rows <= 0 meaning all of them; otherwise HOST buffers.
 *   out[0]: ACCURACY -> positives (rows with round(prediction) == label on every output, :30-33);
 *           SQUARED_ERROR / ABS_ERROR -> sum of (p-y)^2 / |p-y| over all outputs (:53-63; MSE/MAE = out[0]/out[1]);
 *           R2 -> sum (p-y)^2 (:127-131)
 *   out[1]: rows evaluated;  out[2], out[3] (R2 only, labels of shape (1)): sum p, sum p^2 -- the reference centres
 *           SST on the mean of the PREDICTIONS (`map(_._2)`, :124), so R2 = 1 - out[0] / (out[3] - out[2]^2/out[1]). */
typedef enum { SB_EST_ACCURACY = 0, SB_EST_SQUARED_ERROR = 1, SB_EST_ABS_ERROR = 2, SB_EST_R2 = 3 } sb_estimator;
int sb_estimate(sb_engine* e, int estimator, const float* x, const float* y, int64_t n_rows, double out[4]);
/* gradients of the trainable weights for one HOST batch without updating anything
 * (`LossModel.grad`, models/Model.scala:107-111) -- parity-test hook. */
int sb_compute_grads(sb_engine* e, const float* x, const float* y, float* loss_out);
int sb_grad_get(sb_engine* e, const char* weight_path, float* host, size_t n);
int sb_sync(sb_engine* e);
/* loss of the latest forward (or the combined loss after sb_group_average_device): one stream sync + 4-byte D2H */
int sb_read_loss(sb_engine* e, float* loss_out);

/* ---- raw access for host plumbing (torch.distributed over NCCL, timing on the right stream) ---- */
/* device pointer + float count of the averaged arena region [model params | optimizer state] */
int sb_arena(sb_engine* e, void** device_ptr, int64_t* n_floats);
/* cudaStream_t the engine launches on (for CUDA-event timing by the caller) */
int sb_stream(sb_engine* e, void** stream);
/* scale the whole arena by `w` on the engine's stream (pre-multiply of a weighted all-reduce) */
int sb_arena_scale(sb_engine* e, float w);
/* re-initialise tensors whose ParamDef.aggregation == None (Optimizer.scala:209) */
int sb_reset_unaggregated(sb_engine* e);
/* cudaIpcMemHandle_t (64 bytes) of the arena allocation, for peer mapping from another process */
int sb_arena_ipc_handle(sb_engine* e, void* handle64, int64_t* byte_offset);

/* ---- epoch-end averaging across the k workers of one box: `treeReduce(aggregateParams)`
 * (Optimizer.scala:86,196-230; models/Aggregation.scala:16-19) ---- */
typedef enum {
  SB_AVG_MEAN = 0,        /* arithmetic mean (== the reference for k <= 2)                       */
  SB_AVG_SPARK_CHAIN = 1, /* chain of pairwise (l+r)/2 in Spark's treeReduce shape, canonical order  */
  SB_AVG_WEIGHTS = 2      /* caller-supplied convex weights                                      */
} sb_avg_mode;
typedef enum {
  SB_COLL_P2P = 0,  /* one fused kernel per GPU: peer loads over NVLink, weighted sum in rank order
                       (bit-identical to the chain), state reset; needs peer access               */
  SB_COLL_NCCL = 1  /* pre-scale + ncclAllReduce(sum) (libnccl loaded with dlopen)               */
} sb_collective;

int sb_group_create(sb_engine** engines, int k, int avg_mode, const float* weights_or_null, int collective,
                    sb_group** out);
/* averages the arenas in place on every engine, resets aggregation==None tensors, combines
 * loss (same weights) and iterations (plain sum, Optimizer.scala:223).  In/out arrays of length k;
 * on return every entry holds the combined value. */
int sb_group_average(sb_group* g, float* loss_inout, int64_t* iterations_inout);
void sb_group_destroy(sb_group* g);
/* One process per GPU (torchrun / one JVM executor per GPU): `handles` = k x 64 bytes, entry q = rank q's
 * sb_arena_ipc_handle.  sb_group_average_local launches THIS rank's slice of the fused P2P reduce on the
 * engine stream; the caller places a cross-process barrier before it (all epochs done) and after it
 * (all slices stored), then calls sb_reset_unaggregated and combines loss/iterations itself. */
/* sb_group_create_ipc also zeroes this rank's barrier flags: put a cross-process barrier before it (no peer still runs an
 * averaging kernel of an earlier group) and after it (every rank zeroed before anyone raises a flag). */
int sb_group_create_ipc(sb_engine* self, int rank, int k, const void* handles, int avg_mode,
                        const float* weights_or_null, sb_group** out);
int sb_group_average_local(sb_group* g);
/* Device-synchronised variant for IPC groups (one process per GPU, every rank on its OWN GPU): ONE kernel per rank does
 * a flag barrier over NVLink peer memory (all epochs done), the fused reduce-scatter + all-gather of the arena, the
 * combination of the k last-batch losses (same weights and order; result replaces this engine's loss, read it with
 * sb_read_loss) and a second flag barrier (all slices stored).  No host synchronisation: the call only enqueues.
 * `epoch` must increase on every call, by the same amount on every rank (a repeated epoch is rejected: stale flags would
 * open the barrier).  `my_iterations` = batches this rank ran in the epoch; the kernel adds the OTHER ranks' counts to this
 * engine's device step counter, so the next sb_train_steps_async(SB_ITER_FROM_DEVICE, ...) continues at
 * globalIter + sum of all counts (Optimizer.scala:88,223) without a host exchange.  A peer that does not arrive within
 * SB_P2P_TIMEOUT_MS (env, default 600000) leaves this rank's arena un-averaged and the next sb_sync / sb_read_loss returns
 * SB_ERR_TIMEOUT; the CUDA context stays usable (sb_group_average_local + host barriers is the fallback path). */
int sb_group_average_device(sb_group* g, uint64_t epoch, int64_t my_iterations);
/* result of the latest sb_group_average_device: one stream sync, the combined loss ((lossL+lossR)/2 chain, Optimizer.scala:221)
 * and the summed iteration count (Optimizer.scala:223).  SB_ERR_TIMEOUT if a peer never reached the barrier. */
int sb_group_read_result(sb_group* g, float* loss_out, int64_t* total_iterations_out);
/* effective weights the group applies (length k) */
int sb_group_weights(const sb_group* g, float* weights_out);
/* the Spark 3.3.1 treeReduce(depth=2) weight vector for k partitions in ascending completion order */
int sb_spark_chain_weights(int k, float* weights_out);

/* ---- per-kernel-class device timing (CUDA events on the engine stream, eager mode) ---- */
int sb_profile_enable(sb_engine* e, int on);
int sb_profile_count(const sb_engine* e);
int sb_profile_get(const sb_engine* e, int index, char* name, size_t name_cap, double* total_ms, int64_t* launches,
                   double* algorithmic_bytes, double* algorithmic_flops);
int sb_profile_reset(sb_engine* e);
/* kernels launched by this engine since creation (graph replays count their kernel nodes) */
int64_t sb_launch_count(const sb_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* SCANET_B200_H */
