"""Host runtime: the epoch loop of `Optimizer.run` and the phantom-typed builder, mirrored over the C ABI.

Reference: /root/reference/src/main/scala/scanet/optimizers/Optimizer.scala:59-104 (run), :267-338 (Builder),
optimizers/{Condition,Effect}.scala, optimizers/SparkExt.scala:61-68 (`ds.train(model)`),
models/Model.scala:154-187 (TrainedModel), estimators/package.scala:18-43 (accuracy).

    trained = (Dataset(x, y).train(Dense(50, Sigmoid) >> Dense(10, Softmax))
               .loss(CategoricalCrossentropy).using(Adam(0.01)).batch(1000)
               .each(epochs(1), RecordLoss()).stopAfter(epochs(25)).run())

What differs from the reference is only what executes underneath: `broadcast` becomes "params stay on the
device", `mapPartitions(optimizeOnPartition)` becomes k engines running `sb_train_epoch`, and
`treeReduce(aggregateParams)` becomes one fused averaging pass over NVLink."""
import time
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field, replace
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _native as N
from .dsl import Algorithm, Layer, Loss
from .engine import Engine, Group


# ---- dataset (optimizers/SparkExt.scala:11-47) -----------------------------------------------------------------
class Dataset:
    """`Dataset[Record[A]]` with shape metadata: rows of (features, labels)."""

    def __init__(self, features, labels, partitions: int = 1):
        self.features = np.ascontiguousarray(features, dtype=np.float32)
        self.labels = np.ascontiguousarray(labels, dtype=np.float32)
        if self.features.shape[0] != self.labels.shape[0]:
            raise ValueError("features and labels must have the same number of rows")
        self.partitions = int(partitions)

    @property
    def shapes(self):
        return tuple(self.features.shape[1:]), tuple(self.labels.shape[1:])

    def __len__(self):
        return self.features.shape[0]

    def repartition(self, k: int) -> "Dataset":
        return Dataset(self.features, self.labels, k)

    def coalesce(self, k: int) -> "Dataset":
        return self.repartition(k)

    def shard(self, i: int, k: int):
        """contiguous shard i = rows [i*N/k, (i+1)*N/k)"""
        n = len(self)
        a, b = i * n // k, (i + 1) * n // k
        return self.features[a:b], self.labels[a:b]

    def train(self, model: Layer) -> "Builder":
        return Builder(_Config(model=model, dataset=self))


# ---- step bookkeeping (Optimizer.scala:19-41) --------------------------------------------------------------------
@dataclass(frozen=True)
class Step:
    batch: int
    epoch: int = 0
    iter: int = 0

    def __str__(self):
        return f"{self.epoch}:{self.iter}"


@dataclass
class StepResult:
    iterations: int
    loss: float
    fetch_params: Callable[[], "OrderedDict[str, np.ndarray]"]
    _cache: Optional[OrderedDict] = None

    @property
    def modelParams(self):
        if self._cache is None:
            self._cache = self.fetch_params()
        return self._cache


@dataclass
class StepContext:
    step: Step
    result: StepResult
    lossModel: "LossModel"
    time: int  # ms
    engine: Engine = None


# ---- conditions (optimizers/Condition.scala:8-27) and effects (optimizers/Effect.scala:9-65) ---------------------
def always(ctx):
    return True


def never(ctx):
    return False


def epochs(number: int):
    return lambda ctx: ctx.step.epoch % number == 0


def iterations(number: int):
    return lambda ctx: ctx.step.iter % number == 0


class RecordIteration:
    def __call__(self, events: List[str], ctx: StepContext):
        events.append(f"{ctx.step.epoch}/{ctx.step.iter} {ctx.time / 1000.0}s")


class RecordLoss:
    def __init__(self, console: bool = True, tensorboard: bool = False):
        if tensorboard:
            raise N.Unsupported(N.SB_ERR_UNSUPPORTED, "TensorBoard writing stays on the JVM side (out of scope)")
        self.console = console
        self.history: List[float] = []

    def __call__(self, events, ctx):
        self.history.append(ctx.result.loss)
        if self.console:
            events.append(f"loss: {ctx.result.loss}")


class RecordAccuracy:
    def __init__(self, ds: Dataset, console: bool = True):
        self.ds, self.console = ds, console
        self.history: List[float] = []

    def __call__(self, events, ctx):
        trained = TrainedModel(ctx.lossModel, ctx.result.modelParams, ctx.engine)
        a = accuracy(trained, self.ds)
        self.history.append(a)
        if self.console:
            events.append(f"accuracy: {a}")


class SaveCheckpoint:
    """Effect: write the averaged arena (weights, layer state, optimizer state) + `Step` to `file` (atomic rename).  The
    reference has no checkpointing (README.md:112); tensors use its Kryo `TensorSerializer` layout (KryoSerializers.scala:27-63)."""

    def __init__(self, file: str):
        self.file = file

    def __call__(self, events, ctx):
        ctx.engine.save_checkpoint(self.file, ctx.step.iter, ctx.step.epoch)


# ---- models (models/Model.scala:84-187) ------------------------------------------------------------------------------
@dataclass(frozen=True)
class LossModel:
    model: Layer
    lossF: Loss

    def __repr__(self):
        return f"{self.lossF!r}({self.model!r})"


class TrainedModel:
    """`TrainedModel(lossModel.freeze, params)`; `result` / `loss` run on the engine that trained it."""

    def __init__(self, lossModel: LossModel, params: "OrderedDict[str, np.ndarray]", engine: Engine):
        self.lossModel, self.params, self._engine = lossModel, params, engine

    def result(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        e = self._engine
        b = e.batch
        outs = []
        for a in range(0, x.shape[0], b):
            chunk = x[a:a + b]
            n = chunk.shape[0]
            if n < b:  # pad the tail batch; rows are independent in inference mode
                chunk = np.concatenate([chunk, np.zeros((b - n,) + chunk.shape[1:], np.float32)])
            outs.append(e.predict(chunk)[:n])
        return np.concatenate(outs) if outs else np.zeros((0,) + tuple(e.output_shape[1:]), np.float32)

    def loss(self, x, y) -> float:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.shape[0] != self._engine.batch:
            raise ValueError(f"loss needs exactly one batch of {self._engine.batch} rows (the compiled shape)")
        return self._engine.eval_loss(x, y)

    def outputShape(self):
        return self._engine.output_shape


# ---- estimators (estimators/package.scala:18-137): forward + reduction on the device (`sb_estimate`) ---------------------
# `batch` is accepted for signature parity; the engine forwards in batches of its compiled size (rows are independent).
def accuracy(model: TrainedModel, ds: Dataset, batch: int = 1000) -> float:
    """estimators/package.scala:18-43: a row is positive when round(prediction) == label on every output."""
    pos, total, _, _ = model._engine.estimate(N.EST_ACCURACY, ds.features, ds.labels)
    return float(np.float32(pos) / np.float32(total))


def MSE(model: TrainedModel, ds: Dataset, batch: int = 1000) -> float:
    """estimators/package.scala:51-57: sum of squared errors over all outputs / rows."""
    s, total, _, _ = model._engine.estimate(N.EST_SQUARED_ERROR, ds.features, ds.labels)
    return float(np.float32(s) / np.float32(total))


def RMSE(model: TrainedModel, ds: Dataset, batch: int = 1000) -> float:
    """estimators/package.scala:45-49."""
    return float(np.float32(np.sqrt(MSE(model, ds, batch))))


def MAE(model: TrainedModel, ds: Dataset, batch: int = 1000) -> float:
    """estimators/package.scala:59-65."""
    s, total, _, _ = model._engine.estimate(N.EST_ABS_ERROR, ds.features, ds.labels)
    return float(np.float32(s) / np.float32(total))


def R2Score(model: TrainedModel, ds: Dataset, batch: int = 1000) -> float:
    """estimators/package.scala:99-137; SST is centred on the mean of the PREDICTIONS, as the reference does (:124)."""
    if tuple(np.shape(ds.labels)[1:]) not in ((1,), ()):
        raise ValueError("requirement failed: labels should have shape (1)")
    ssr, n, sp, sp2 = model._engine.estimate(N.EST_R2, ds.features, ds.labels)
    sst = sp2 - sp * sp / n
    return float(np.float32(1.0 - ssr / sst))


# ---- builder (Optimizer.scala:267-338) ---------------------------------------------------------------------------
@dataclass(frozen=True)
class _Config:
    model: Layer = None
    dataset: Dataset = None
    loss: Loss = None
    alg: Algorithm = None
    initParams: Optional[Callable[[], Dict[str, np.ndarray]]] = None
    batchSize: int = 10000  # Optimizer.scala:321
    minimizing: bool = True
    stop: Callable = None
    doOnEach: tuple = field(default_factory=lambda: ((always, RecordIteration()),))
    # engine knobs (not in the reference)
    precision: int = N.PREC_FP32
    avg_mode: int = N.AVG_SPARK_CHAIN
    collective: int = N.COLL_P2P
    devices: Optional[tuple] = None
    use_graph: bool = True
    quiet: bool = False
    resume: Optional[str] = None


class Builder:
    def __init__(self, cfg: _Config):
        self._c = cfg

    def _with(self, **kw) -> "Builder":
        return Builder(replace(self._c, **kw))

    def loss(self, loss: Loss):
        return self._with(loss=loss)

    def using(self, alg: Algorithm):
        return self._with(alg=alg)

    def initParams(self, params):
        return self._with(initParams=(params if callable(params) else (lambda: params)))

    def on(self, dataset: Dataset):
        return self._with(dataset=dataset)

    def stopWhen(self, condition):
        return self._with(stop=condition)

    def stopAfter(self, condition):
        return self.stopWhen(condition)

    def epochs(self, number: int):
        return self.stopWhen(epochs(number))

    def iterations(self, number: int):
        return self.stopWhen(iterations(number))

    def batch(self, size: int):
        return self._with(batchSize=int(size))

    def eachIter(self, effect):
        return self.each(always, effect)

    def each(self, when, effect):
        return self._with(doOnEach=self._c.doOnEach + ((when, effect),))

    # engine knobs
    def precision(self, p: int):
        return self._with(precision=int(p))

    def averaging(self, avg_mode: int, collective: int = N.COLL_P2P):
        return self._with(avg_mode=int(avg_mode), collective=int(collective))

    def devices(self, devs: Sequence[int]):
        return self._with(devices=tuple(devs))

    def graph(self, on: bool):
        return self._with(use_graph=bool(on))

    def resumeFrom(self, file: str):
        """continue from a `SaveCheckpoint` file: params, optimizer state, iteration and epoch counters"""
        return self._with(resume=str(file))

    def quiet(self, on: bool = True):
        return self._with(quiet=bool(on))

    def build(self) -> "Optimizer":
        c = self._c
        missing = [n for n, v in (("loss", c.loss), ("using", c.alg), ("on", c.dataset), ("stopAfter", c.stop)) if v is None]
        if missing:  # the reference refuses to compile run() unless the builder state is Complete
            raise TypeError("builder is incomplete, missing: " + ", ".join(missing))
        return Optimizer(c)

    def run(self) -> TrainedModel:
        return self.build().run()


def minimize(model: Layer) -> Builder:
    """Optimizer.minimize (Optimizer.scala:312-324)"""
    return Builder(_Config(model=model))


def maximize(model: Layer) -> Builder:
    """Optimizer.maximize (Optimizer.scala:326-338): the update becomes `w + delta` (Optimizer.scala:186)"""
    return Builder(_Config(model=model, minimizing=False))


class Optimizer:
    """One worker (engine) per dataset partition, all on the GPUs of this process."""

    def __init__(self, cfg: _Config):
        self.c = cfg
        self.lossModel = LossModel(cfg.model, cfg.loss)

    def run(self) -> TrainedModel:
        c = self.c
        ds = c.dataset
        k = max(1, ds.partitions)
        devices = list(c.devices) if c.devices is not None else list(range(k))
        if len(devices) != k:
            raise ValueError(f"{k} partitions need {k} devices, got {devices}")
        in_shape, lab_shape = ds.shapes
        if not c.quiet:
            try:  # Optimizer.scala:65 (`println(s"Training model:\n${model.describe[A](batchSize +: shapes._1)}")`)
                print(f"Training model:\n{c.model.describe((c.batchSize,) + tuple(in_shape))}")
            except Exception as ex:  # a description must never stop a training run
                print(f"Training model: {c.model!r} (no description: {ex})")
        engines = [Engine(c.model, c.loss, c.alg, c.batchSize, in_shape, lab_shape, device=d, precision=c.precision,
                          use_graph=c.use_graph, minimizing=c.minimizing) for d in devices]
        group = pool = None
        handed_over = False  # engines[0] leaves with the TrainedModel on success; every other path closes all k engines
        try:
            group = Group(engines, c.avg_mode, collective=c.collective) if k > 1 else None
            pool = ThreadPoolExecutor(max_workers=k) if k > 1 else None
            step = Step(c.batchSize)
            for i, e in enumerate(engines):
                x, y = ds.shard(i, k)
                e.upload_dataset(x, y)
                if c.resume is not None:  # every worker restarts from the same averaged arena
                    it, ep = e.load_checkpoint(c.resume)
                    step = Step(c.batchSize, ep, it)
                    continue
                # epoch 0: every partition initialises (Optimizer.scala:123-131); seeded => identical starts
                init = c.initParams() if c.initParams is not None else None
                if init is not None:
                    for p in e.model_param_paths():
                        if p not in init:
                            raise ValueError(f"missing {p} param")
                        e.set_param(p, init[p])
                    e.init_opt_params()
                else:
                    e.init_params()
            while True:
                start = time.time()
                if pool is None:
                    results = [engines[0].train_epoch(step.iter)]
                else:
                    results = list(pool.map(lambda e: e.train_epoch(step.iter), engines))
                if group is not None:
                    loss, iters = group.average([r[1] for r in results], [r[0] for r in results])
                else:
                    iters, loss = results[0]
                elapsed = int((time.time() - start) * 1000)
                step = Step(c.batchSize, step.epoch + 1, step.iter + iters)
                result = StepResult(iters, loss, engines[0].get_model_params)
                ctx = StepContext(step, result, self.lossModel, elapsed, engines[0])
                events: List[str] = []
                for when, effect in c.doOnEach:
                    if when(ctx):
                        effect(events, ctx)
                if events and not c.quiet:
                    print(", ".join(events))
                if c.stop(ctx):
                    params = result.modelParams
                    break
            handed_over = True
            return TrainedModel(self.lossModel, params, engines[0])
        finally:
            if pool is not None:
                pool.shutdown()
            if group is not None:
                group.close()
            for e in (engines[1:] if handed_over else engines):  # device arenas, streams and graphs
                e.close()
