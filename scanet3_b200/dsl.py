"""Host-side mirror of the reference's model DSL for the train path (same names, argument meaning, defaults and
param-path conventions), lowering to the C-ABI descriptors of include/scanet_b200.h.

Reference (under /root/reference/src/main/scala/scanet/): models/layer/{Layer,Composed,Dense,Bias,Activate,Flatten,
Conv2D,Pool2D,BatchNorm,RNN}.scala, models/{Activation,Loss,Initializer}.scala, optimizers/{SGD..AMSGrad}.scala.
This module contains NO arithmetic: it only describes models; all compute happens in libscanet_b200.so."""
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

from . import _native as N


# ---- initializers (models/Initializer.scala:28-220) ---------------------------------------------------------
@dataclass(frozen=True)
class Initializer:
    kind: int
    seed: Optional[int] = None
    a: float = 0.0
    b: float = 0.0
    name: str = ""

    def to_native(self) -> N.sb_initializer:
        # seed = None: TF draws a fresh seed per `initialize.eval`, and every partition initialises independently on the first
        # epoch (Optimizer.scala:123-131) -- so the host runtime draws one per lowering (= per engine / partition).  The C ABI
        # itself only takes explicit seeds; bit parity with the reference is claimed for explicitly seeded initializers only.
        seed = self.seed
        if seed is None and self.kind not in (N.INIT_ZEROS, N.INIT_ONES, N.INIT_CONST):
            import secrets
            seed = secrets.randbits(31)
        return N.sb_initializer(self.kind, 0 if seed is None else 1, 0 if seed is None else int(seed),
                                float(self.a), float(self.b))

    def __repr__(self):
        return self.name


Zeros = Initializer(N.INIT_ZEROS, name="Zeros")
Ones = Initializer(N.INIT_ONES, name="Ones")


def Const(value: float):
    return Initializer(N.INIT_CONST, a=value, name=f"Const({value})")


def RandomUniform(min: float = -0.05, max: float = 0.05, seed: Optional[int] = None):
    return Initializer(N.INIT_RANDOM_UNIFORM, seed, min, max, f"RandomUniform(min={min},max={max})")


def RandomNormal(mean: float = 0.0, std: float = 0.05, seed: Optional[int] = None):
    return Initializer(N.INIT_RANDOM_NORMAL, seed, mean, std, f"RandomNormal(mean={mean},std={std})")


def GlorotUniform(seed: Optional[int] = None):
    return Initializer(N.INIT_GLOROT_UNIFORM, seed, name="GlorotUniform")


def HeUniform(seed: Optional[int] = None):
    return Initializer(N.INIT_HE_UNIFORM, seed, name="HeUniform")


def LecunUniform(seed: Optional[int] = None):
    return Initializer(N.INIT_LECUN_UNIFORM, seed, name="LecunUniform")


def RandomNormalTruncated(mean: float = 0.0, std: float = 0.05, seed: Optional[int] = None):
    """models/Initializer.scala:62-70 (TF TruncatedNormal)"""
    return Initializer(N.INIT_TRUNCATED_NORMAL, seed, mean, std, name="RandomNormalTruncated")


def GlorotNormal(seed: Optional[int] = None):
    return Initializer(N.INIT_GLOROT_NORMAL, seed, name="GlorotNormal")


def HeNormal(seed: Optional[int] = None):
    return Initializer(N.INIT_HE_NORMAL, seed, name="HeNormal")


def LecunNormal(seed: Optional[int] = None):
    return Initializer(N.INIT_LECUN_NORMAL, seed, name="LecunNormal")


def Orthogonal(gain: Optional[float] = None, seed: Optional[int] = None):
    return Initializer(N.INIT_ORTHOGONAL, seed, 0.0 if gain is None else gain, name="Orthogonal")


# ---- activations (models/Activation.scala:26-84) ------------------------------------------------------------
@dataclass(frozen=True)
class Activation:
    code: int
    name: str
    threshold: float = 0.0

    @property
    def identity(self):
        return self.code == N.ACT_IDENTITY

    @property
    def layer(self):
        return Activate(self)

    def __repr__(self):
        return self.name


Identity = Activation(N.ACT_IDENTITY, "Identity")
Sigmoid = Activation(N.ACT_SIGMOID, "Sigmoid")
Tanh = Activation(N.ACT_TANH, "Tanh")
Softmax = Activation(N.ACT_SOFTMAX, "Softmax")


def ReLU(threshold: float = 0.0):
    return Activation(N.ACT_RELU, f"ReLU({float(threshold)})", float(threshold))  # case class toString: ReLU(0.0)


# ---- losses (models/Loss.scala:10-58) -----------------------------------------------------------------------
@dataclass(frozen=True)
class Loss:
    code: int
    name: str

    def __repr__(self):
        return self.name


MeanSquaredError = Loss(N.LOSS_MSE, "MeanSquaredError")
BinaryCrossentropy = Loss(N.LOSS_BCE, "BinaryCrossentropy")
CategoricalCrossentropy = Loss(N.LOSS_CCE, "CategoricalCrossentropy")


# ---- regularisation (models/Regularization.scala:16-44) ------------------------------------------------------
class _Zero:
    code, lam = 0, 0.0

    def __repr__(self):
        return "Zero"


Zero = _Zero()


class L1:
    """penalty lambda * sum|w| / 2"""
    code = 1

    def __init__(self, lam: float = 0.01):
        self.lam = float(lam)

    def __repr__(self):
        return f"L1({self.lam})"


class L2:
    """penalty lambda * sum(w^2) / 2"""
    code = 2

    def __init__(self, lam: float = 0.01):
        self.lam = float(lam)

    def __repr__(self):
        return f"L2({self.lam})"


def _check_reg(reg):
    if not isinstance(reg, (_Zero, L1, L2)):
        raise N.IllegalArgument(N.SB_ERR_INVALID_ARG, f"unknown regularisation {reg!r}")


# ---- layers --------------------------------------------------------------------------------------------------
# ---- model description: `Model.describe` (models/Model.scala:62-81), printed by `Optimizer.run` (Optimizer.scala:65) ----
def _power(shape) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


def _shape_str(shape) -> str:
    """core/Shape.scala toString: `(3, 3, 1, 32)`, `(10)`, `()`"""
    return "(" + ", ".join(str(int(d)) for d in shape) + ")"


def _group_concat(items) -> str:
    """layer/LayerInfo.scala:10-37: a list made of one repeated group prints as `[group]xN`"""
    items = list(items)
    elements, times, size = items, 1, 1
    while len(items) >= size / 2:
        if len(items) % size == 0 and len({tuple(items[i:i + size]) for i in range(0, len(items), size)}) == 1:
            elements, times = items[:size], len(items) // size
            break
        size += 1
    value = ", ".join(str(e) for e in elements)
    return f"[{value}]x{times}" if times > 1 else value


def _format_size(nbytes: int) -> str:
    """utils/Bytes.scala:4-8"""
    if nbytes < 1024:
        return f"{nbytes} B"
    z = (nbytes.bit_length() - 1) // 10
    return "%.1f %sB" % (nbytes / float(1 << (z * 10)), " KMGTPE"[z])


def _tabulate(table) -> str:
    """utils/Tabulator.scala:5-30 with a header row"""
    widths = [max(len(str(row[c])) for row in table) for c in range(len(table[0]))]
    sep = "+" + "+".join("-" * w for w in widths) + "+"
    rows = ["|" + "|".join(("" if w == 0 else str(cell).ljust(w)) for cell, w in zip(row, widths)) + "|" for row in table]
    return "\n".join([sep, rows[0], sep] + rows[1:] + [sep])


def _window_out(size: int, window: int, stride: int, padding) -> int:
    """output extent of a conv / pool axis (math/nn/AllKernels.scala:12-103); `padding` = Valid | Same | (before, after)"""
    if padding == Same:
        return -(-size // stride)
    before, after = (0, 0) if padding == Valid else padding
    return (size + before + after - window) // stride + 1


@dataclass(frozen=True)
class LayerInfo:
    """layer/LayerInfo.scala:8-50"""
    name: str
    weights: Tuple[Tuple[int, ...], ...]
    state: Tuple[Tuple[int, ...], ...]
    output: Tuple[int, ...]

    @property
    def weightsTotal(self) -> int:
        return sum(_power(w) for w in self.weights)

    @property
    def stateTotal(self) -> int:
        return sum(_power(w) for w in self.state)

    def toRow(self) -> List[str]:
        return [self.name, _group_concat(_shape_str(w) for w in self.weights), _group_concat(_power(w) for w in self.weights),
                _group_concat(_power(w) for w in self.state), "(" + ",".join(["_"] + [str(d) for d in self.output[1:]]) + ")"]


class Layer:
    """`layer/Layer.scala:24-26`: `>>` composes; compositions flatten (`Composed.scala:89-99`)."""

    def leaves(self) -> List["Layer"]:
        return [self]

    # -- shapes and description (host logic only; the engine plans the same shapes in C++) --
    def outputShape(self, input: Sequence[int]) -> Tuple[int, ...]:
        return tuple(input)

    def paramDefs(self, input: Sequence[int]) -> List[Tuple[Tuple[int, ...], bool]]:
        """(shape, trainable) of the layer's parameters in the order of the reference's `params(input)`"""
        return []

    def info(self, input: Sequence[int]) -> List["LayerInfo"]:
        """models/Model.scala:62-69: non-trainable parameters are listed as state"""
        defs = self.paramDefs(input)
        return [LayerInfo(repr(self), tuple(sh for sh, tr in defs if tr), tuple(sh for sh, tr in defs if not tr),
                          self.outputShape(input))]

    def describe(self, input: Sequence[int], itemsize: int = 4) -> str:
        """models/Model.scala:71-81 (`itemsize` = bytes of the element type: 4 for Float).  Rows of Dense / Bias / Activate /
        Flatten / Conv2D / Pool2D layers are pinned by the reference's golden table (models/CNNSpec.scala:39-63); BatchNorm and RNN
        rows follow the same rules but the reference has no golden for them (an LSTM's 12 parameters live in a hash map there)."""
        infos = self.info(input)
        rows = [LayerInfo("Input", (), (), tuple(input)).toRow()] + [i.toRow() for i in infos]
        table = _tabulate([["name", "weights", "weights params", "state params", "output"]] + rows)
        wt, st = sum(i.weightsTotal for i in infos), sum(i.stateTotal for i in infos)
        return f"{table}\nTotal weight params: {wt} ({_format_size(wt * itemsize)}), state params: {st} ({_format_size(st * itemsize)})"

    def __rshift__(self, other: "Layer") -> "Layer":
        return Composed(self.leaves() + other.leaves())

    def andThen(self, other):
        return self >> other

    def to_native(self) -> N.sb_layer_desc:
        raise NotImplementedError

    def _desc(self, kind, **kw):
        d = N.sb_layer_desc()
        d.kind = kind
        d.trainable = 1
        d.window_h = d.window_w = d.stride_h = d.stride_w = 1
        d.bn_fused = -1
        for k, v in kw.items():
            setattr(d, k, v)
        return d


class Composed(Layer):
    def __init__(self, layers: Sequence[Layer]):
        flat: List[Layer] = []
        for l in layers:
            flat.extend(l.leaves())
        if not flat:
            raise ValueError("requirement failed: at least 1 layer is required")
        self.layers = flat

    def leaves(self):
        return list(self.layers)

    def outputShape(self, input):
        for l in self.layers:
            input = l.outputShape(input)
        return tuple(input)

    def paramDefs(self, input):
        out = []
        for l in self.layers:
            out += l.paramDefs(input)
            input = l.outputShape(input)
        return out

    def info(self, input):  # layer/Composed.scala:71-77
        out = []
        for l in self.layers:
            out += l.info(input)
            input = l.outputShape(input)
        return out

    def __repr__(self):
        return " >> ".join(repr(l) for l in self.layers)


class _DenseKernel(Layer):
    def __init__(self, outputs, initializer, trainable, reg=Zero):
        self.outputs, self.initializer, self.trainable, self.reg = outputs, initializer, trainable, reg

    def to_native(self):
        d = self._desc(N.LAYER_DENSE, units=self.outputs, trainable=int(self.trainable), reg_kind=self.reg.code,
                       reg_lambda=self.reg.lam)
        d.init[0] = self.initializer.to_native()
        return d

    def outputShape(self, input):  # layer/Dense.scala:66
        return (input[0], self.outputs)

    def paramDefs(self, input):  # layer/Dense.scala:50-55
        return [((input[1], self.outputs), bool(self.trainable))]

    def __repr__(self):
        return f"Dense({self.outputs})"


class Bias(Layer):
    def __init__(self, features, reg=Zero, initializer=Zeros, trainable=True):
        _check_reg(reg)
        self.features, self.initializer, self.trainable, self.reg = features, initializer, trainable, reg

    def to_native(self):
        d = self._desc(N.LAYER_BIAS, units=self.features, trainable=int(self.trainable), reg_kind=self.reg.code,
                       reg_lambda=self.reg.lam)
        d.init[2] = self.initializer.to_native()
        return d

    def paramDefs(self, input):  # layer/Bias.scala:27-30
        return [((self.features,), bool(self.trainable))]

    def __repr__(self):
        return f"Bias({self.features})"


class Activate(Layer):
    def __init__(self, activation: Activation):
        self.activation = activation

    def to_native(self):
        return self._desc(N.LAYER_ACTIVATE, activation=self.activation.code, relu_threshold=self.activation.threshold)

    def __repr__(self):
        return repr(self.activation)


class _Flatten(Layer):
    def to_native(self):
        return self._desc(N.LAYER_FLATTEN)

    def outputShape(self, input):  # layer/Flatten.scala:20
        return (input[0], _power(input[1:]))

    def __repr__(self):
        return "Flatten"


Flatten = _Flatten()


def Dense(outputs: int, activation: Activation = Identity, reg=Zero, bias: bool = True,
          kernelInitializer: Initializer = GlorotUniform(), biasInitializer: Initializer = Zeros,
          trainable: bool = True) -> Layer:
    """layer/Dense.scala:30-41 -> Dense >> Bias >> activation.layer"""
    _check_reg(reg)
    layers: List[Layer] = [_DenseKernel(outputs, kernelInitializer, trainable, reg)]
    if bias:
        layers.append(Bias(outputs, reg, biasInitializer))
    if not activation.identity:
        layers.append(activation.layer)
    return layers[0] if len(layers) == 1 else Composed(layers)


Valid, Same = "Valid", "Same"
_PAD = {Valid: N.PAD_VALID, Same: N.PAD_SAME}


class Explicit:
    """`Explicit(height = (pad_top, pad_bottom), width = (pad_left, pad_right))` (math/nn/AllKernels.scala:67-102); Conv2D only."""

    def __init__(self, height: Tuple[int, int], width: Tuple[int, int]):
        self.height, self.width = (int(height[0]), int(height[1])), (int(width[0]), int(width[1]))

    def __repr__(self):
        return f"Explicit({self.height},{self.width})"

    def __eq__(self, other):
        return isinstance(other, Explicit) and (self.height, self.width) == (other.height, other.width)

    def __hash__(self):
        return hash((self.height, self.width))


def _pad_code(padding):
    if isinstance(padding, Explicit):
        return N.PAD_EXPLICIT
    if padding not in _PAD:
        raise ValueError(f"unknown padding {padding!r}")
    return _PAD[padding]


class _Conv2DKernel(Layer):
    def __init__(self, filters, kernel, strides, padding, initializer, trainable):
        self.filters, self.kernel, self.strides, self.padding = filters, tuple(kernel), tuple(strides), padding
        self.initializer, self.trainable = initializer, trainable

    def to_native(self):
        d = self._desc(N.LAYER_CONV2D, units=self.filters, window_h=self.kernel[0], window_w=self.kernel[1],
                       stride_h=self.strides[0], stride_w=self.strides[1], padding=_pad_code(self.padding),
                       trainable=int(self.trainable))
        if isinstance(self.padding, Explicit):
            (d.pad_top, d.pad_bottom), (d.pad_left, d.pad_right) = self.padding.height, self.padding.width
        d.init[0] = self.initializer.to_native()
        return d

    def _pads(self):
        if isinstance(self.padding, Explicit):
            return self.padding.height, self.padding.width
        return self.padding, self.padding

    def outputShape(self, input):  # layer/Conv2D.scala:109-122
        if len(input) != 4:
            raise ValueError(f"requirement failed: Conv2D input should have a shape (NHWC) or (NCHW) but was {_shape_str(input)}")
        ph, pw = self._pads()
        return (input[0], _window_out(input[1], self.kernel[0], self.strides[0], ph),
                _window_out(input[2], self.kernel[1], self.strides[1], pw), self.filters)

    def paramDefs(self, input):  # layer/Conv2D.scala:84-91: HWIO
        return [((self.kernel[0], self.kernel[1], input[3], self.filters), bool(self.trainable))]

    def __repr__(self):
        return f"Conv2D({self.filters},({self.kernel[0]},{self.kernel[1]}),({self.strides[0]},{self.strides[1]}),{self.padding},NHWC)"


def Conv2D(filters: int, kernel: Tuple[int, int] = (3, 3), strides: Tuple[int, int] = (1, 1), padding=Valid,
           format: str = "NHWC", activation: Activation = Identity, bias: bool = False,
           kernelInitializer: Initializer = GlorotUniform(), biasInitializer: Initializer = Zeros, biasReg=Zero,
           trainable: bool = True) -> Layer:
    """layer/Conv2D.scala:54-68 -> Conv2D [>> Bias] [>> activation]; NOTE bias defaults to False."""
    if format != "NHWC":
        raise N.Unsupported(N.SB_ERR_UNSUPPORTED, "NCHW is outside the accelerated path (SURVEY 8b)")
    layers: List[Layer] = [_Conv2DKernel(filters, kernel, strides, padding, kernelInitializer, trainable)]
    if bias:
        layers.append(Bias(filters, biasReg, biasInitializer))
    if not activation.identity:
        layers.append(activation.layer)
    return layers[0] if len(layers) == 1 else Composed(layers)


class Pool2D(Layer):
    """layer/Pool2D.scala:28-57: window (2,2), strides (1,1), Valid, Max.  `reduce` is recorded but the layer's
    build never forwards it, so pooling is always Max."""

    def __init__(self, window=(2, 2), strides=(1, 1), padding=Valid, format="NHWC", reduce="Max", honorReduce=False):
        self.honorReduce = bool(honorReduce)  # not in the reference: True makes `reduce="Avg"` average-pool
        if format != "NHWC":
            raise N.Unsupported(N.SB_ERR_UNSUPPORTED, "NCHW is outside the accelerated path (SURVEY 8b)")
        if isinstance(padding, Explicit) or padding not in _PAD:
            raise ValueError("requirement failed: only [Same, Valid] paddings are supported")
        self.window, self.strides, self.padding, self.reduce = tuple(window), tuple(strides), padding, reduce

    def to_native(self):
        return self._desc(N.LAYER_POOL2D, window_h=self.window[0], window_w=self.window[1], stride_h=self.strides[0],
                          stride_w=self.strides[1], padding=_PAD[self.padding],
                          pool_reduce=N.POOL_AVG if self.reduce == "Avg" else N.POOL_MAX,
                          pool_honor_reduce=int(self.honorReduce))

    def outputShape(self, input):  # layer/Pool2D.scala:44-56
        if len(input) != 4:
            raise ValueError(f"requirement failed: Pool2D input should have a shape (NHWC) or (NCHW) but was {_shape_str(input)}")
        return (input[0], _window_out(input[1], self.window[0], self.strides[0], self.padding),
                _window_out(input[2], self.window[1], self.strides[1], self.padding), input[3])

    def __repr__(self):
        return f"Pool2D(({self.window[0]},{self.window[1]}),({self.strides[0]},{self.strides[1]}),{self.padding},NHWC,{self.reduce})"


class BatchNorm(Layer):
    """layer/BatchNorm.scala:63-96.  `bnMode`: "reference" reproduces the fused path's output wiring bug for bug
    (SURVEY App. B-Q5), "correct" is the TF op as intended."""

    def __init__(self, axes=(-1,), momentum=0.99, epsilon=1e-3, betaInitializer=Zeros, gammaInitializer=Ones,
                 movingMeanInitializer=Zeros, movingVarianceInitializer=Ones, fused: Optional[bool] = None,
                 trainable=True, bnMode="reference"):
        if tuple(axes) != (-1,):
            raise N.Unsupported(N.SB_ERR_UNSUPPORTED, "BatchNorm axes other than [-1] are outside the accelerated path")
        self.momentum, self.epsilon, self.fused, self.trainable, self.bnMode = momentum, epsilon, fused, trainable, bnMode
        self.inits = (gammaInitializer, betaInitializer, movingMeanInitializer, movingVarianceInitializer)

    def to_native(self):
        d = self._desc(N.LAYER_BATCHNORM, bn_momentum=self.momentum, bn_epsilon=self.epsilon,
                       bn_mode=0 if self.bnMode == "reference" else 1,
                       bn_fused=-1 if self.fused is None else int(self.fused), trainable=int(self.trainable))
        for i, init in enumerate(self.inits):
            d.init[i] = init.to_native()
        return d

    def paramDefs(self, input):  # layer/BatchNorm.scala:78-95: beta, gamma, moving_mean, moving_variance; axes [-1] keep the last dim
        shape = tuple([1] * (len(input) - 1) + [input[-1]])
        return [(shape, bool(self.trainable)), (shape, bool(self.trainable)), (shape, False), (shape, False)]

    def __repr__(self):
        return "BatchNorm(List(-1))"


class RNN(Layer):
    """layer/RNN.scala:27-94 around an LSTMCell or a SimpleRNNCell (>> Bias >> activation)."""

    def __init__(self, cell, returnSequence=False, stateful=False):
        if not isinstance(cell, (LSTMCell, SimpleRNNCell)):
            raise N.Unsupported(N.SB_ERR_UNSUPPORTED, "only LSTMCell and SimpleRNNCell are on the accelerated path")
        self.cell, self.returnSequence, self.stateful = cell, bool(returnSequence), bool(stateful)

    def to_native(self):
        c = self.cell
        if isinstance(c, SimpleRNNCell):
            d = self._desc(N.LAYER_SIMPLE_RNN, units=c.units, activation=c.activation.code, use_bias=int(c.bias),
                           return_sequence=int(self.returnSequence), stateful=int(self.stateful),
                           relu_threshold=float(getattr(c.activation, "threshold", 0.0)))
            for i, init in enumerate((c.kernelInitializer, c.recurrentInitializer, c.biasInitializer)):
                d.init[i] = init.to_native()
            return d
        d = self._desc(N.LAYER_LSTM, units=c.units, activation=c.activation.code,
                       recurrent_activation=c.recurrentActivation.code, use_bias=int(c.bias),
                       return_sequence=int(self.returnSequence), stateful=int(self.stateful))
        for i, init in enumerate((c.kernelInitializer, c.recurrentInitializer, c.biasInitializer, c.biasForgetInitializer)):
            d.init[i] = init.to_native()
        return d

    def outputShape(self, input):  # layer/RNN.scala:80-86
        if len(input) != 3:
            raise ValueError("requirement failed: RNN requires input to have Shape(batch, time, features)")
        return (input[0], input[1], self.cell.units) if self.returnSequence else (input[0], self.cell.units)

    def paramDefs(self, input):  # layer/RNN.scala:32-39, 182-188, 293-305 (weights first here; the reference's map order is unpinned)
        c, f, u = self.cell, input[2], self.cell.units
        per_cell = [((f, u), True), ((u, u), True)] + ([((u,), True)] if c.bias else [])
        weights = per_cell * (4 if isinstance(c, LSTMCell) else 1)
        n_state = (2 if isinstance(c, LSTMCell) else 1) if self.stateful else 0
        return weights + [((input[0], u), False)] * n_state

    def __repr__(self):
        return f"RNN({self.cell!r})"


@dataclass
class LSTMCell:
    units: int
    activation: Activation = Tanh
    recurrentActivation: Activation = Sigmoid
    bias: bool = True
    kernelInitializer: Initializer = GlorotUniform()
    recurrentInitializer: Initializer = Orthogonal()
    biasInitializer: Initializer = Zeros
    biasForgetInitializer: Initializer = Ones

    def __repr__(self):
        return f"LSTMCell({self.units},{self.activation!r},{self.recurrentActivation!r},{str(self.bias).lower()})"


@dataclass
class SimpleRNNCell:
    """layer/RNN.scala:140-206: `SimpleRNNCell >> Bias >> activation`; the state fed back is the pre-bias, pre-activation sum."""
    units: int
    activation: Activation = Tanh
    bias: bool = True
    kernelInitializer: Initializer = GlorotUniform()
    recurrentInitializer: Initializer = Orthogonal()
    biasInitializer: Initializer = Zeros

    def __repr__(self):
        s = f"SimpleRNNCell({self.units})"
        if self.bias:
            s += f" >> Bias({self.units})"
        if not self.activation.identity:
            s += f" >> {self.activation!r}"
        return s


def SimpleRNN(units, activation=Tanh, bias=True, kernelInitializer=GlorotUniform(), recurrentInitializer=Orthogonal(),
              biasInitializer=Zeros, returnSequence=False, stateful=False) -> RNN:
    """layer/RNN.scala:96-138"""
    return RNN(SimpleRNNCell(units, activation, bias, kernelInitializer, recurrentInitializer, biasInitializer), returnSequence,
               stateful)


def LSTM(units, activation=Tanh, recurrentActivation=Sigmoid, bias=True, kernelInitializer=GlorotUniform(),
         recurrentInitializer=Orthogonal(), biasInitializer=Zeros, biasForgetInitializer=Ones, returnSequence=False,
         stateful=False) -> RNN:
    """layer/RNN.scala:208-239"""
    return RNN(LSTMCell(units, activation, recurrentActivation, bias, kernelInitializer, recurrentInitializer,
                        biasInitializer, biasForgetInitializer), returnSequence, stateful)


# ---- optimizers (optimizers/*.scala; defaults as in the case classes) ----------------------------------------
@dataclass(frozen=True)
class Algorithm:
    code: int
    rate: float = 0.0
    beta1: float = 0.0
    beta2: float = 0.0
    epsilon: float = 0.0
    initAcc: float = 0.0
    nesterov: bool = False
    name: str = ""

    def __repr__(self):
        return self.name


def SGD(rate=0.01, momentum=0.0, nesterov=False):
    return Algorithm(N.OPT_SGD, rate=rate, beta1=momentum, nesterov=nesterov, name=f"SGD({rate},{momentum},{nesterov})")


def Adam(rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-7):
    return Algorithm(N.OPT_ADAM, rate, beta1, beta2, epsilon, name=f"Adam({rate},{beta1},{beta2},{epsilon})")


def AdaGrad(rate=0.001, epsilon=1e-7):
    return Algorithm(N.OPT_ADAGRAD, rate=rate, epsilon=epsilon, name=f"AdaGrad({rate},{epsilon})")


def RMSProp(rate=0.001, rho=0.9, epsilon=1e-7):
    return Algorithm(N.OPT_RMSPROP, rate=rate, beta1=rho, epsilon=epsilon, name=f"RMSProp({rate},{rho},{epsilon})")


def AdaDelta(rate=1.0, rho=0.9, epsilon=1e-7):
    return Algorithm(N.OPT_ADADELTA, rate=rate, beta1=rho, epsilon=epsilon, name=f"AdaDelta({rate},{rho},{epsilon})")


def Adamax(rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-7, initAcc=0.001):
    return Algorithm(N.OPT_ADAMAX, rate, beta1, beta2, epsilon, initAcc, name=f"Adamax({rate},{beta1},{beta2},{epsilon},{initAcc})")


def Nadam(beta1=0.9, beta2=0.999, epsilon=1e-7, initAcc=0.001):
    return Algorithm(N.OPT_NADAM, 0.0, beta1, beta2, epsilon, initAcc, name=f"Nadam({beta1},{beta2},{epsilon},{initAcc})")


def AMSGrad(rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-7):
    return Algorithm(N.OPT_AMSGRAD, rate, beta1, beta2, epsilon, name=f"AMSGrad({rate},{beta1},{beta2},{epsilon})")


def LinearRegression(bias=True, kernelInitializer=Zeros, biasInitializer=Zeros):
    """models/LinearRegression.scala:20-31"""
    return Dense(1, Identity, bias=bias, kernelInitializer=kernelInitializer, biasInitializer=biasInitializer)


def LogisticRegression(bias=True, kernelInitializer=Zeros, biasInitializer=Zeros):
    """models/LogisticRegression.scala:19-30"""
    return Dense(1, Sigmoid, bias=bias, kernelInitializer=kernelInitializer, biasInitializer=biasInitializer)
