"""The per-worker train loop, the epoch loop and the epoch-end averaging, restated from
`src/main/scala/scanet/optimizers/Optimizer.scala:59-230` and `optimizers/Iterators.scala:39-92`.
Test infrastructure only (also the timed CPU "port" baseline of bench.py)."""
import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .layers import (Layer, model_backward, model_forward, model_params, model_penalty, model_penalty_grads)

__all__ = ["tensor_iterator", "LossModel", "StepResult", "init_params", "opt_param_defs", "train_step",
           "eval_loss", "optimize_on_partition", "aggregate_pair", "spark_chain_weights", "spark_tree_reduce",
           "weighted_average", "run_training", "shard_rows"]


# ----------------------------------------------------------------------------------------------
# TensorIterator (optimizers/Iterators.scala:39-83)
# ----------------------------------------------------------------------------------------------
def tensor_iterator(x_rows: np.ndarray, y_rows: np.ndarray, batch: int, remaining: str = "skip"):
    """rows -> (x[B,...], y[B,...]); `skip` (training default) drops the partial tail batch,
    `partial` trims it, `pad_zeros` pads it."""
    n = x_rows.shape[0]
    full = n // batch
    for j in range(full):
        yield x_rows[j * batch:(j + 1) * batch], y_rows[j * batch:(j + 1) * batch]
    rest = n - full * batch
    if rest and remaining == "partial":
        yield x_rows[full * batch:], y_rows[full * batch:]
    elif rest and remaining == "pad_zeros":
        xp = np.zeros((batch,) + x_rows.shape[1:], x_rows.dtype)
        yp = np.zeros((batch,) + y_rows.shape[1:], y_rows.dtype)
        xp[:rest], yp[:rest] = x_rows[full * batch:], y_rows[full * batch:]
        yield xp, yp


# ----------------------------------------------------------------------------------------------
# LossModel (models/Model.scala:84-127)
# ----------------------------------------------------------------------------------------------
@dataclass
class LossModel:
    model: Layer
    loss: object

    def loss_stateful(self, x, y, params):
        """loss = L(model(x; params), y) + penalty(params)  (Model.scala:92-99)"""
        out, caches, next_state = model_forward(self.model, x, params, training=True)
        val = self.loss.value(out, y) + model_penalty(self.model, params, x.dtype)
        return val, out, caches, next_state

    def grad_stateful(self, x, y, params):
        """-> (loss, grads of trainable weights, next_state)  (Model.scala:113-119)"""
        val, out, caches, next_state = self.loss_stateful(x, y, params)
        g = self.loss.grad(out, y)
        grads = model_backward(self.model, g, caches, params)
        for k, v in model_penalty_grads(self.model, params).items():
            grads[k] = grads[k] + v
        return val, grads, next_state


@dataclass
class StepResult:
    """Optimizer.scala:26-35"""
    iterations: int
    model_params: "OrderedDict[str, np.ndarray]"
    opt_params: "OrderedDict[str, np.ndarray]"
    loss: float
    step_losses: List[float] = field(default_factory=list)  # oracle extra: pre-update loss of every step


def opt_param_defs(model_defs, alg):
    """`<weightPath>/<stateName>` for every TRAINABLE param (Optimizer.scala:118-121)."""
    out = OrderedDict()
    for path, d in model_defs.items():
        if d.trainable:
            for name, init in alg.state_defs().items():
                out[f"{path}/{name}"] = (tuple(d.shape), init)
    return out


def init_params(model, alg, in_shape_batched, dtype=np.float32, init: Optional[Dict[str, np.ndarray]] = None):
    """Epoch-0 initialisation (Optimizer.scala:123-131)."""
    defs = model_params(model, in_shape_batched)
    if init is not None:
        mp = OrderedDict((k, np.array(init[k], dtype=dtype).reshape(defs[k].shape)) for k in defs)
    else:
        mp = OrderedDict((k, d.initialize().astype(dtype)) for k, d in defs.items())
    op = OrderedDict((k, i.build(s).astype(dtype)) for k, (s, i) in opt_param_defs(defs, alg).items())
    return defs, mp, op


def train_step(lm: LossModel, alg, x, y, model_params_, opt_params, t: int, minimizing=True):
    """One `backprop` call (Optimizer.scala:169-194): grads, per-weight update, next state.
    Returns (next_model_params, next_opt_params, pre-update loss)."""
    val, grads, next_state = lm.grad_stateful(x, y, model_params_)
    nxt = OrderedDict(model_params_)
    nopt = OrderedDict(opt_params)
    for path, g in grads.items():
        w = model_params_[path]
        st = {k[len(path) + 1:]: v for k, v in opt_params.items() if k.startswith(path + "/") and
              "/" not in k[len(path) + 1:]}
        delta, st2 = alg.build(g.reshape(w.shape).astype(w.dtype), st, t)
        nxt[path] = (w - delta if minimizing else w + delta).astype(w.dtype)
        for k, v in st2.items():
            nopt[f"{path}/{k}"] = v.astype(w.dtype)
    for k, v in next_state.items():  # nextParams = nextWeights ++ nextState (Optimizer.scala:144)
        nxt[k] = v.astype(x.dtype)
    return nxt, nopt, val


def eval_loss(lm: LossModel, x, y, params):
    """`compileLoss` closure (Optimizer.scala:163-167): training-mode forward, state discarded."""
    return lm.loss_stateful(x, y, params)[0]


def optimize_on_partition(lm: LossModel, alg, x_rows, y_rows, batch: int, global_iter: int, model_params_,
                          opt_params, minimizing=True) -> StepResult:
    """Optimizer.scala:106-161: sequential full batches in row order, `iter = globalIter+j+1`, then a
    forward-only loss on the LAST batch with the UPDATED params."""
    n_batches = x_rows.shape[0] // batch
    if n_batches == 0:
        raise IndexError("shard smaller than one batch (NoSuchElementException, Iterators.scala:79-81)")
    mp, op = model_params_, opt_params
    losses = []
    x = y = None
    for j, (x, y) in enumerate(tensor_iterator(x_rows, y_rows, batch, "skip")):
        mp, op, val = train_step(lm, alg, x, y, mp, op, global_iter + j + 1, minimizing)
        losses.append(float(val))
    final = eval_loss(lm, x, y, mp)
    return StepResult(n_batches, mp, op, float(final), losses)


# ----------------------------------------------------------------------------------------------
# Averaging (Optimizer.scala:196-230; models/Aggregation.scala:16-19) and Spark's reduce shape
# ----------------------------------------------------------------------------------------------
def aggregate_pair(left: StepResult, right: StepResult, model_defs=None) -> StepResult:
    """One node of the reduce: `(l + r) / 2` on every tensor with aggregation=Avg, re-initialise
    tensors with aggregation=None, `(lossL+lossR)/2`, `iterL+iterR`."""
    def agg(l, r, defs):
        out = OrderedDict()
        for k in l:
            d = defs.get(k) if defs is not None else None
            if d is not None and d.aggregation is None:
                out[k] = d.initialize().astype(l[k].dtype)
            else:
                out[k] = ((l[k] + r[k]) / l[k].dtype.type(2)).astype(l[k].dtype)
        return out
    t = np.float32
    loss = float((t(left.loss) + t(right.loss)) / t(2))
    return StepResult(left.iterations + right.iterations, agg(left.model_params, right.model_params, model_defs),
                      agg(left.opt_params, right.opt_params, None), loss)


def _chain_weights(n):
    """weights of ((a0 (+) a1) (+) a2) ... (+) a_{n-1} with (+) = pairwise average."""
    if n == 1:
        return [1.0]
    w = [0.5 ** (n - 1)] + [0.5 ** (n - i) for i in range(1, n)]
    return w


def spark_chain_weights(k: int) -> List[float]:
    """Effective convex weights of Spark 3.3.1 `RDD.treeReduce(f, depth=2)` under the declared canonical
    completion order (ascending partition index) -- SURVEY 3.3 (restated from the pinned Spark version,
    `build.sbt:16`; Spark sources are absent from /root/reference, the k>2 rows are PARITY UNPINNED).
      scale = max(ceil(k^(1/2)), 2); while k > scale + ceil(k/scale): one level keyed i % (k/scale)."""
    groups = [[(i, 1.0)] for i in range(k)]  # per-partition reduceLeft already applied (1 element each)
    parts = k
    scale = max(int(math.ceil(parts ** 0.5)), 2)
    while parts > scale + math.ceil(parts / scale):
        parts //= scale
        merged: List[List[Tuple[int, float]]] = [[] for _ in range(parts)]
        buckets: List[List[List[Tuple[int, float]]]] = [[] for _ in range(parts)]
        for i, g in enumerate(groups):
            buckets[i % parts].append(g)
        for b, lst in enumerate(buckets):
            cw = _chain_weights(len(lst))
            for w, g in zip(cw, lst):
                merged[b].extend((idx, w * gw) for idx, gw in g)
        groups = merged
    cw = _chain_weights(len(groups))
    out = [0.0] * k
    for w, g in zip(cw, groups):
        for idx, gw in g:
            out[idx] += w * gw
    return out


def spark_tree_reduce(results: Sequence[StepResult], model_defs=None) -> StepResult:
    """Apply `aggregate_pair` in the same structure `spark_chain_weights` describes."""
    groups = [[r] for r in results]
    parts = len(results)
    scale = max(int(math.ceil(parts ** 0.5)), 2)

    def chain(lst):
        acc = lst[0]
        for r in lst[1:]:
            acc = aggregate_pair(acc, r, model_defs)
        return acc
    level = [chain(g) for g in groups]
    while parts > scale + math.ceil(parts / scale):
        parts //= scale
        buckets = [[] for _ in range(parts)]
        for i, r in enumerate(level):
            buckets[i % parts].append(r)
        level = [chain(b) for b in buckets]
    return chain(level)


def weighted_average(results: Sequence[StepResult], weights: Sequence[float], model_defs=None) -> StepResult:
    """Weighted all-reduce form: sum_r w_r * theta_r in rank order (what the engine computes);
    tensors with aggregation=None are re-initialised when k > 1; loss uses the same weights;
    iterations are a plain sum (Optimizer.scala:223)."""
    k = len(results)
    if k == 1:
        return results[0]

    def agg(dicts, defs):
        out = OrderedDict()
        for key in dicts[0]:
            d = defs.get(key) if defs is not None else None
            if d is not None and d.aggregation is None:
                out[key] = d.initialize().astype(dicts[0][key].dtype)
                continue
            t = dicts[0][key].dtype.type
            acc = dicts[0][key] * t(weights[0])
            for r in range(1, k):
                acc = acc + dicts[r][key] * t(weights[r])
            out[key] = acc
        return out
    loss = np.float32(0)
    for r in range(k):
        loss = loss + np.float32(results[r].loss) * np.float32(weights[r])
    return StepResult(sum(r.iterations for r in results), agg([r.model_params for r in results], model_defs),
                      agg([r.opt_params for r in results], None), float(loss))


def shard_rows(n_rows: int, k: int) -> List[Tuple[int, int]]:
    """contiguous shard i = rows [i*N/k, (i+1)*N/k)  (SURVEY 8d)."""
    return [(i * n_rows // k, (i + 1) * n_rows // k) for i in range(k)]


# ----------------------------------------------------------------------------------------------
# Epoch loop (Optimizer.scala:59-104)
# ----------------------------------------------------------------------------------------------
def run_training(model: Layer, loss, alg, x_rows, y_rows, batch: int, epochs: int, k: int = 1,
                 avg_mode: str = "spark_chain", init: Optional[Dict[str, np.ndarray]] = None, dtype=np.float32,
                 minimizing=True, on_epoch: Optional[Callable] = None):
    """`.stopAfter(n.epochs)`: returns (model_params, opt_params, history[(epoch, global_iter, loss)])."""
    lm = LossModel(model, loss)
    in_shape = (batch,) + tuple(x_rows.shape[1:])
    defs = model_params(model, in_shape)
    shards = shard_rows(x_rows.shape[0], k)
    x_rows = x_rows.astype(dtype, copy=False)
    y_rows = y_rows.astype(dtype, copy=False)
    global_iter = 0
    mp = op = None
    hist = []
    for epoch in range(1, epochs + 1):
        results = []
        for (a, b) in shards:
            if global_iter == 0:  # every partition initialises independently (Optimizer.scala:123-131)
                _, mp0, op0 = init_params(model, alg, in_shape, dtype, init)
            else:
                mp0, op0 = mp, op
            results.append(optimize_on_partition(lm, alg, x_rows[a:b], y_rows[a:b], batch, global_iter, mp0, op0,
                                                 minimizing))
        if k == 1:
            red = results[0]
        elif avg_mode == "spark_chain":
            red = spark_tree_reduce(results, defs)
        elif avg_mode == "mean":
            red = weighted_average(results, [1.0 / k] * k, defs)
        else:
            raise ValueError(avg_mode)
        mp, op = red.model_params, red.opt_params
        global_iter += red.iterations
        hist.append((epoch, global_iter, red.loss))
        if on_epoch is not None:
            on_epoch(epoch, global_iter, red)
    return mp, op, hist
