"""ctypes binding of `libscanet_b200.so` (declared in include/scanet_b200.h).

The product path: there is no Python/NumPy fallback -- if the shared library is missing or no CUDA device
is present, the calls fail loudly with `NativeError`."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB_LIB") or os.path.join(_HERE, "libscanet_b200.so")  # SB_LIB: a developer variant (`make trace`)

SB_OK, SB_ERR_INVALID_ARG, SB_ERR_UNSUPPORTED, SB_ERR_CUDA, SB_ERR_NCCL, SB_ERR_STATE, SB_ERR_TIMEOUT = range(7)
ITER_FROM_DEVICE = -1
WARM_RESIDENT, WARM_LOSS, WARM_HOSTFED, WARM_ALL = 1, 2, 4, 7

# enums (mirror include/scanet_b200.h)
(LAYER_DENSE, LAYER_BIAS, LAYER_ACTIVATE, LAYER_FLATTEN, LAYER_CONV2D, LAYER_POOL2D, LAYER_BATCHNORM, LAYER_LSTM,
 LAYER_SIMPLE_RNN) = range(1, 10)
ACT_IDENTITY, ACT_SIGMOID, ACT_TANH, ACT_SOFTMAX, ACT_RELU = range(5)
PAD_VALID, PAD_SAME, PAD_EXPLICIT = 0, 1, 2
POOL_MAX, POOL_AVG = 0, 1
(INIT_ZEROS, INIT_ONES, INIT_CONST, INIT_RANDOM_UNIFORM, INIT_RANDOM_NORMAL, INIT_GLOROT_UNIFORM, INIT_HE_UNIFORM,
 INIT_LECUN_UNIFORM, INIT_ORTHOGONAL, INIT_TRUNCATED_NORMAL, INIT_GLOROT_NORMAL, INIT_HE_NORMAL, INIT_LECUN_NORMAL) = range(13)
LOSS_MSE, LOSS_BCE, LOSS_CCE = range(3)
OPT_SGD, OPT_ADAM, OPT_ADAGRAD, OPT_RMSPROP, OPT_ADADELTA, OPT_ADAMAX, OPT_NADAM, OPT_AMSGRAD = range(8)
PREC_FP32, PREC_TF32, PREC_3XTF32 = range(3)
ROLE_WEIGHT, ROLE_STATE, ROLE_OPT = range(3)
AGG_NONE, AGG_AVG = 0, 1
AVG_MEAN, AVG_SPARK_CHAIN, AVG_WEIGHTS = range(3)
COLL_P2P, COLL_NCCL = 0, 1
EST_ACCURACY, EST_SQUARED_ERROR, EST_ABS_ERROR, EST_R2 = range(4)


class NativeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[sb_status {code}] {message}")
        self.code = code
        self.message = message


class IllegalArgument(NativeError, ValueError):
    """SB_ERR_INVALID_ARG: what the reference raises as IllegalArgumentException (core/Require.scala:5-6)."""


class Unsupported(NativeError, NotImplementedError):
    pass


class sb_initializer(C.Structure):
    _fields_ = [("kind", C.c_int32), ("has_seed", C.c_int32), ("seed", C.c_int64), ("a", C.c_float), ("b", C.c_float)]


class sb_layer_desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("units", C.c_int32), ("activation", C.c_int32), ("recurrent_activation", C.c_int32),
                ("relu_threshold", C.c_float), ("window_h", C.c_int32), ("window_w", C.c_int32), ("stride_h", C.c_int32),
                ("stride_w", C.c_int32), ("padding", C.c_int32), ("pool_reduce", C.c_int32), ("use_bias", C.c_int32),
                ("trainable", C.c_int32), ("bn_momentum", C.c_float), ("bn_epsilon", C.c_float), ("bn_mode", C.c_int32),
                ("bn_fused", C.c_int32), ("init", sb_initializer * 4), ("reg_kind", C.c_int32), ("reg_lambda", C.c_float),
                ("pad_top", C.c_int32), ("pad_bottom", C.c_int32), ("pad_left", C.c_int32), ("pad_right", C.c_int32),
                ("return_sequence", C.c_int32), ("pool_honor_reduce", C.c_int32), ("stateful", C.c_int32)]


class sb_model_desc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("layers", C.POINTER(sb_layer_desc)), ("input_rank", C.c_int32),
                ("input_dims", C.c_int64 * 4), ("label_rank", C.c_int32), ("label_dims", C.c_int64 * 4)]


class sb_train_desc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("loss", C.c_int32), ("optimizer", C.c_int32), ("rate", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float), ("init_acc", C.c_float),
                ("nesterov", C.c_int32), ("minimizing", C.c_int32), ("precision", C.c_int32), ("use_graph", C.c_int32)]


_P = C.c_void_p
_FP = C.POINTER(C.c_float)

# name -> (restype, argtypes); every symbol include/scanet_b200.h declares (tests check the export list)
PROTOTYPES = {
    "sb_last_error": (C.c_char_p, []),
    "sb_abi_version": (C.c_int, []),
    "sb_device_count": (C.c_int, []),
    "sb_engine_create": (C.c_int, [C.POINTER(sb_model_desc), C.POINTER(sb_train_desc), C.c_int, C.POINTER(_P)]),
    "sb_engine_destroy": (None, [_P]),
    "sb_param_count": (C.c_int, [_P]),
    "sb_param_info": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "sb_output_dims": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "sb_params_set": (C.c_int, [_P, C.c_char_p, _P, C.c_size_t]),
    "sb_params_get": (C.c_int, [_P, C.c_char_p, _P, C.c_size_t]),
    "sb_init_params": (C.c_int, [_P]),
    "sb_init_opt_params": (C.c_int, [_P]),
    "sb_dataset_upload": (C.c_int, [_P, _P, _P, C.c_int64]),
    "sb_train_epoch": (C.c_int, [_P, C.c_int64, C.POINTER(C.c_int64), _FP]),
    "sb_train_steps_async": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64]),
    "sb_warm": (C.c_int, [_P, C.c_int]),
    "sb_train_step": (C.c_int, [_P, _P, _P, C.c_int64, _FP]),
    "sb_train_step_async": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "sb_train_batches_async": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, _P]),
    "sb_eval_loss": (C.c_int, [_P, _P, _P, _FP]),
    "sb_predict": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "sb_wire_encode": (C.c_int, [C.POINTER(C.c_int64), C.c_int32, _P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sb_wire_decode": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_int32), _P, C.c_size_t,
                                 C.POINTER(C.c_size_t)]),
    "sb_param_to_wire": (C.c_int, [_P, C.c_char_p, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sb_param_from_wire": (C.c_int, [_P, C.c_char_p, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sb_checkpoint_save": (C.c_int, [_P, C.c_char_p, C.c_int64, C.c_int64]),
    "sb_checkpoint_load": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "sb_estimate": (C.c_int, [_P, C.c_int, _P, _P, C.c_int64, C.POINTER(C.c_double)]),
    "sb_compute_grads": (C.c_int, [_P, _P, _P, _FP]),
    "sb_grad_get": (C.c_int, [_P, C.c_char_p, _P, C.c_size_t]),
    "sb_sync": (C.c_int, [_P]),
    "sb_read_loss": (C.c_int, [_P, _FP]),
    "sb_arena": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "sb_stream": (C.c_int, [_P, C.POINTER(_P)]),
    "sb_arena_scale": (C.c_int, [_P, C.c_float]),
    "sb_reset_unaggregated": (C.c_int, [_P]),
    "sb_arena_ipc_handle": (C.c_int, [_P, _P, C.POINTER(C.c_int64)]),
    "sb_group_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _FP, C.c_int, C.POINTER(_P)]),
    "sb_group_average": (C.c_int, [_P, _FP, C.POINTER(C.c_int64)]),
    "sb_group_destroy": (None, [_P]),
    "sb_group_create_ipc": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _FP, C.POINTER(_P)]),
    "sb_group_average_local": (C.c_int, [_P]),
    "sb_group_average_device": (C.c_int, [_P, C.c_uint64, C.c_int64]),
    "sb_group_read_result": (C.c_int, [_P, _FP, C.POINTER(C.c_int64)]),
    "sb_group_weights": (C.c_int, [_P, _FP]),
    "sb_spark_chain_weights": (C.c_int, [C.c_int, _FP]),
    "sb_profile_enable": (C.c_int, [_P, C.c_int]),
    "sb_profile_count": (C.c_int, [_P]),
    "sb_profile_get": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "sb_profile_reset": (C.c_int, [_P]),
    "sb_launch_count": (C.c_int64, [_P]),
}

_lib = None


def lib():
    """The loaded shared library; raises if it has not been built (run `python -c 'import __graft_entry__ as g; g.build()'`
    or `make`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(SB_ERR_CUDA, f"{LIB_PATH} is missing: build it with `make` (there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code):
    if code == SB_OK:
        return
    msg = lib().sb_last_error().decode("utf-8", "replace")
    if code == SB_ERR_INVALID_ARG:
        raise IllegalArgument(code, msg)
    if code == SB_ERR_UNSUPPORTED:
        raise Unsupported(code, msg)
    raise NativeError(code, msg)
