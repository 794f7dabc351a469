"""Thin object wrapper over the C ABI (one `Engine` = one worker = one GPU)."""
import ctypes as C
import os
from collections import OrderedDict
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from .dsl import Algorithm, Layer, Loss


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class ParamInfo:
    __slots__ = ("path", "shape", "role", "aggregation", "offset")

    def __init__(self, path, shape, role, aggregation, offset):
        self.path, self.shape, self.role, self.aggregation, self.offset = path, shape, role, aggregation, offset

    @property
    def numel(self):
        return int(np.prod(self.shape)) if len(self.shape) else 1


class Engine:
    """Builds the device plan for `model` trained with `loss`/`alg` at batch size `batch`.

    `input_shape` / `label_shape` are per-sample shapes (`ds.shapes`, optimizers/SparkExt.scala:36-44).
    `device=-1` gives a plan-only engine (param enumeration without CUDA); everything else needs a GPU."""

    def __init__(self, model: Layer, loss: Loss, alg: Algorithm, batch: int, input_shape: Sequence[int],
                 label_shape: Sequence[int], device: int = 0, precision: int = N.PREC_FP32, use_graph: bool = True,
                 minimizing: bool = True):
        self._h = None
        lib = N.lib()
        leaves = model.leaves()
        arr = (N.sb_layer_desc * len(leaves))(*[l.to_native() for l in leaves])
        md = N.sb_model_desc()
        md.n_layers = len(leaves)
        md.layers = arr
        md.input_rank = len(input_shape)
        md.label_rank = len(label_shape)
        for i, d in enumerate(input_shape):
            md.input_dims[i] = int(d)
        for i, d in enumerate(label_shape):
            md.label_dims[i] = int(d)
        td = N.sb_train_desc(int(batch), loss.code, alg.code, alg.rate, alg.beta1, alg.beta2, alg.epsilon, alg.initAcc,
                             int(alg.nesterov), int(minimizing), int(precision), int(use_graph))
        h = C.c_void_p()
        N.check(lib.sb_engine_create(C.byref(md), C.byref(td), int(device), C.byref(h)))
        self._h = h
        self._lib = lib
        self.model, self.loss, self.alg = model, loss, alg
        self.batch, self.device = int(batch), int(device)
        self.input_shape, self.label_shape = tuple(int(d) for d in input_shape), tuple(int(d) for d in label_shape)
        # a single (non-composed) layer keeps un-prefixed paths like the reference (`Params(Weights -> w)`,
        # Conv2DLayerSpec.scala:35); the C ABI always numbers leaves, so strip / add the "0/" here
        from .dsl import Composed
        self._single = not isinstance(model, Composed)
        self.params: "OrderedDict[str, ParamInfo]" = OrderedDict()
        for i in range(lib.sb_param_count(h)):
            path = C.create_string_buffer(256)
            dims = (C.c_int64 * 4)()
            rank, role, agg, off = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
            N.check(lib.sb_param_info(h, i, path, 256, dims, C.byref(rank), C.byref(role), C.byref(agg), C.byref(off)))
            p = path.value.decode()
            if self._single:
                p = p.split("/", 1)[1]
            self.params[p] = ParamInfo(p, tuple(dims[j] for j in range(rank.value)), role.value, agg.value, off.value)
        dims = (C.c_int64 * 5)()
        rank = C.c_int32()
        N.check(lib.sb_output_dims(h, dims, C.byref(rank)))
        self.output_shape = tuple(dims[j] for j in range(rank.value))

    # -- lifetime ------------------------------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            self._lib.sb_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._h

    def _np(self, path: str) -> bytes:
        return (("0/" + path) if self._single else path).encode()

    # -- params --------------------------------------------------------------------------------------------
    def model_param_paths(self):
        return [p for p, i in self.params.items() if i.role != N.ROLE_OPT]

    def weight_paths(self):
        return [p for p, i in self.params.items() if i.role == N.ROLE_WEIGHT]

    def opt_param_paths(self):
        return [p for p, i in self.params.items() if i.role == N.ROLE_OPT]

    def set_param(self, path: str, value):
        v = _f32(value)
        N.check(self._lib.sb_params_set(self._h, self._np(path), v.ctypes.data, v.size))

    def get_param(self, path: str) -> np.ndarray:
        info = self.params[path]
        out = np.empty(info.shape, np.float32)
        N.check(self._lib.sb_params_get(self._h, self._np(path), out.ctypes.data, out.size))
        return out

    def set_params(self, values: Dict[str, np.ndarray]):
        for k, v in values.items():
            self.set_param(k, v)

    def get_params(self, paths=None) -> "OrderedDict[str, np.ndarray]":
        return OrderedDict((p, self.get_param(p)) for p in (paths if paths is not None else self.params))

    def get_model_params(self):
        return self.get_params(self.model_param_paths())

    def get_opt_params(self):
        return self.get_params(self.opt_param_paths())

    def init_params(self):
        N.check(self._lib.sb_init_params(self._h))

    def init_opt_params(self):
        N.check(self._lib.sb_init_opt_params(self._h))

    # -- train path ------------------------------------------------------------------------------------------
    def upload_dataset(self, x, y):
        x, y = _f32(x), _f32(y)
        n = x.shape[0]
        if y.shape[0] != n:
            raise ValueError("features and labels must have the same number of rows")
        N.check(self._lib.sb_dataset_upload(self._h, x.ctypes.data, y.ctypes.data, n))
        self.dataset_rows = n

    def train_epoch(self, global_iter: int) -> Tuple[int, float]:
        it, loss = C.c_int64(), C.c_float()
        N.check(self._lib.sb_train_epoch(self._h, int(global_iter), C.byref(it), C.byref(loss)))
        return it.value, loss.value

    def warm(self, what: int = N.WARM_ALL):
        """`sb_warm`: capture + instantiate + upload the CUDA graphs of the train path now instead of inside the first epoch."""
        if not getattr(self, "dataset_rows", 0):
            what &= ~N.WARM_RESIDENT
        N.check(self._lib.sb_warm(self._h, int(what)))

    def train_steps_async(self, global_iter: int, first_batch: int, n_steps: int):
        N.check(self._lib.sb_train_steps_async(self._h, int(global_iter), int(first_batch), int(n_steps)))

    def train_step(self, x, y, it: int, want_loss: bool = True) -> Optional[float]:
        x, y = _f32(x), _f32(y)
        loss = C.c_float()
        N.check(self._lib.sb_train_step(self._h, x.ctypes.data, y.ctypes.data, int(it), C.byref(loss) if want_loss else None))
        return loss.value if want_loss else None

    def train_step_async_ptr(self, x_ptr: int, y_ptr: int, it: int, loss_ptr: int = 0):
        """pointer form for pinned host buffers (end-to-end timing)"""
        N.check(self._lib.sb_train_step_async(self._h, x_ptr, y_ptr, int(it), loss_ptr or None))

    def train_batches_async_ptr(self, x_ptr: int, y_ptr: int, n_batches: int, global_iter: int, losses_ptr: int = 0):
        """`sb_train_batches_async`: n consecutive HOST batches (pinned) in groups of up to 8 steps per graph; no sync."""
        N.check(self._lib.sb_train_batches_async(self._h, x_ptr, y_ptr, int(n_batches), int(global_iter), losses_ptr or None))

    def train_batches(self, x, y, global_iter: int) -> np.ndarray:
        """synchronous convenience form: -> the PRE-update loss of every full batch of (x, y)"""
        x, y = _f32(x), _f32(y)
        n = x.shape[0] // self.batch
        losses = np.empty(n, np.float32)
        N.check(self._lib.sb_train_batches_async(self._h, x.ctypes.data, y.ctypes.data, n, int(global_iter), losses.ctypes.data))
        self.sync()
        return losses

    def eval_loss(self, x, y) -> float:
        x, y = _f32(x), _f32(y)
        loss = C.c_float()
        N.check(self._lib.sb_eval_loss(self._h, x.ctypes.data, y.ctypes.data, C.byref(loss)))
        return loss.value

    # ---- wire format / checkpoint (optimizers/KryoSerializers.scala:27-63 layout per tensor) ----
    def param_to_wire(self, path: str) -> bytes:
        n = C.c_size_t()
        N.check(self._lib.sb_param_to_wire(self._h, self._np(path), None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        N.check(self._lib.sb_param_to_wire(self._h, self._np(path), buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    def param_from_wire(self, path: str, record: bytes) -> int:
        used = C.c_size_t()
        N.check(self._lib.sb_param_from_wire(self._h, self._np(path), record, len(record), C.byref(used)))
        return used.value

    def save_checkpoint(self, file: str, global_iter: int, epoch: int = 0):
        N.check(self._lib.sb_checkpoint_save(self._h, os.fsencode(file), int(global_iter), int(epoch)))

    def load_checkpoint(self, file: str) -> Tuple[int, int]:
        """-> (global_iter, epoch) ; every tensor of the arena is restored"""
        it, ep = C.c_int64(), C.c_int64()
        N.check(self._lib.sb_checkpoint_load(self._h, os.fsencode(file), C.byref(it), C.byref(ep)))
        return it.value, ep.value

    def estimate(self, kind: int, x=None, y=None, rows: int = 0) -> Tuple[float, float, float, float]:
        """Device-side estimator sums (`sb_estimate`): host buffers, or the resident shard when x is None."""
        out = (C.c_double * 4)()
        if x is None:
            N.check(self._lib.sb_estimate(self._h, int(kind), None, None, int(rows), out))
        else:
            x, y = _f32(x), _f32(y)
            N.check(self._lib.sb_estimate(self._h, int(kind), x.ctypes.data, y.ctypes.data, int(x.shape[0]), out))
        return tuple(out)

    def predict(self, x) -> np.ndarray:
        x = _f32(x)
        out = np.empty(self.output_shape, np.float32)
        N.check(self._lib.sb_predict(self._h, x.ctypes.data, out.ctypes.data, out.size))
        return out

    def compute_grads(self, x, y) -> Tuple[float, "OrderedDict[str, np.ndarray]"]:
        x, y = _f32(x), _f32(y)
        loss = C.c_float()
        N.check(self._lib.sb_compute_grads(self._h, x.ctypes.data, y.ctypes.data, C.byref(loss)))
        grads = OrderedDict()
        for p in self.weight_paths():
            g = np.empty(self.params[p].shape, np.float32)
            N.check(self._lib.sb_grad_get(self._h, self._np(p), g.ctypes.data, g.size))
            grads[p] = g
        return loss.value, grads

    def read_loss(self) -> float:
        out = C.c_float()
        N.check(self._lib.sb_read_loss(self._h, C.byref(out)))
        return float(out.value)

    def sync(self):
        N.check(self._lib.sb_sync(self._h))

    # -- plumbing --------------------------------------------------------------------------------------------
    def arena(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_int64()
        N.check(self._lib.sb_arena(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stream(self) -> int:
        s = C.c_void_p()
        N.check(self._lib.sb_stream(self._h, C.byref(s)))
        return s.value or 0

    def arena_scale(self, w: float):
        N.check(self._lib.sb_arena_scale(self._h, float(w)))

    def reset_unaggregated(self):
        N.check(self._lib.sb_reset_unaggregated(self._h))

    def arena_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        off = C.c_int64()
        N.check(self._lib.sb_arena_ipc_handle(self._h, buf, C.byref(off)))
        return buf.raw

    def arena_tensor(self):
        """torch view of the arena (for torch.distributed collectives over NCCL)."""
        import torch

        ptr, n = self.arena()

        class _Holder:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}

        t = torch.as_tensor(_Holder(), device=f"cuda:{self.device}")
        t._sb_keepalive = self
        return t

    # -- profiling -------------------------------------------------------------------------------------------
    def profile_enable(self, on: bool):
        N.check(self._lib.sb_profile_enable(self._h, int(on)))

    def profile_reset(self):
        N.check(self._lib.sb_profile_reset(self._h))

    def profile(self):
        out = []
        for i in range(self._lib.sb_profile_count(self._h)):
            name = C.create_string_buffer(128)
            ms, bytes_, flops = C.c_double(), C.c_double(), C.c_double()
            launches = C.c_int64()
            N.check(self._lib.sb_profile_get(self._h, i, name, 128, C.byref(ms), C.byref(launches), C.byref(bytes_), C.byref(flops)))
            out.append({"name": name.value.decode(), "total_ms": ms.value, "launches": launches.value,
                        "bytes": bytes_.value, "flops": flops.value})
        return out

    def launch_count(self) -> int:
        return int(self._lib.sb_launch_count(self._h))


def wire_encode(t) -> bytes:
    """One `TensorSerializer.write` record of an fp32 array (KryoSerializers.scala:33-49)."""
    t = np.array(t, dtype=np.float32, order="C")  # keeps rank 0 (ascontiguousarray would promote a scalar to rank 1)
    dims = (C.c_int64 * max(1, t.ndim))(*t.shape)
    n = C.c_size_t()
    lib = N.lib()
    N.check(lib.sb_wire_encode(dims, t.ndim, None, None, 0, C.byref(n)))
    buf = C.create_string_buffer(n.value)
    N.check(lib.sb_wire_encode(dims, t.ndim, t.ctypes.data, buf, n.value, C.byref(n)))
    return buf.raw[:n.value]


def wire_decode(record: bytes):
    """-> (array, bytes consumed) ; the inverse of `wire_encode` (`TensorSerializer.read`, KryoSerializers.scala:51-64)."""
    lib = N.lib()
    dims = (C.c_int64 * 8)()
    rank, used = C.c_int32(), C.c_size_t()
    N.check(lib.sb_wire_decode(record, len(record), dims, 8, C.byref(rank), None, 0, C.byref(used)))
    shape = tuple(dims[i] for i in range(rank.value))
    out = np.empty(shape, np.float32)
    N.check(lib.sb_wire_decode(record, len(record), dims, 8, C.byref(rank), out.ctypes.data, out.size, C.byref(used)))
    return out, used.value


def spark_chain_weights(k: int):
    w = (C.c_float * k)()
    N.check(N.lib().sb_spark_chain_weights(k, w))
    return [float(v) for v in w]


class Group:
    """k engines of ONE process averaged in place (single-process multi-GPU: a JVM with one thread per GPU)."""

    def __init__(self, engines, avg_mode=N.AVG_SPARK_CHAIN, weights=None, collective=N.COLL_P2P):
        self.engines = list(engines)
        k = len(self.engines)
        hs = (C.c_void_p * k)(*[e.handle for e in self.engines])
        w = (C.c_float * k)(*weights) if weights is not None else None
        g = C.c_void_p()
        N.check(N.lib().sb_group_create(hs, k, int(avg_mode), w, int(collective), C.byref(g)))
        self._g = g

    def weights(self):
        k = len(self.engines)
        w = (C.c_float * k)()
        N.check(N.lib().sb_group_weights(self._g, w))
        return [float(v) for v in w]

    def average(self, losses, iterations):
        k = len(self.engines)
        l = (C.c_float * k)(*losses)
        it = (C.c_int64 * k)(*iterations)
        N.check(N.lib().sb_group_average(self._g, l, it))
        return float(l[0]), int(it[0])

    def close(self):
        if self._g is not None:
            N.lib().sb_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
