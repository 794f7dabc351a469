"""Layer DSL + forward/backward math restated from the reference (test infrastructure only).

Follows `src/main/scala/scanet/models/layer/*.scala`, `models/Activation.scala`,
`math/nn/AllKernels.scala` (conv / pool / batch-norm semantics and padding arithmetic) and the
local gradients in `math/alg/AllKernels.scala`.  Each class cites the lines it restates.
All tensors are C-contiguous row-major, conv tensors NHWC, filters HWIO.
"""
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

from . import initializers as I

__all__ = ["ParamDef", "Layer", "Composed", "Dense", "DenseKernel", "Bias", "Activate", "Flatten", "Conv2D",
           "Conv2DKernel", "Pool2D", "BatchNorm", "RNN", "SimpleRNNCell", "SimpleRNN", "LSTMCell", "LSTM",
           "Identity", "Sigmoid", "Tanh", "Softmax", "ReLU", "Zero", "L1", "L2", "conv2d", "conv2d_wgrad",
           "conv2d_dgrad", "pool2d", "pool2d_grad", "flatten_layers", "model_params", "model_forward",
           "model_backward", "model_penalty", "model_penalty_grads", "split_weights_state"]


# ----------------------------------------------------------------------------------------------
# ParamDef (models/ParamDef.scala:5-12) and regularisation (models/Regularization.scala:16-45)
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ParamDef:
    shape: Tuple[int, ...]
    initializer: I.Initializer = I.Zeros
    aggregation: Optional[str] = None  # "avg" | "sum" | None
    trainable: bool = False

    def initialize(self):
        return self.initializer.build(tuple(self.shape))


@dataclass(frozen=True)
class _Zero:
    def penalty(self, w):
        return w.dtype.type(0)

    def grad(self, w):
        return None


Zero = _Zero()


@dataclass(frozen=True)
class L2:
    lam: float = 0.01

    def penalty(self, w):  # lambda * sum(w^2) / 2
        t = w.dtype.type
        return t(self.lam) * np.sum(w * w, dtype=w.dtype) / t(2)

    def grad(self, w):
        return w.dtype.type(self.lam) * w


@dataclass(frozen=True)
class L1:
    lam: float = 0.01

    def penalty(self, w):
        t = w.dtype.type
        return t(self.lam) * np.sum(np.abs(w), dtype=w.dtype) / t(2)

    def grad(self, w):
        # reference quirk: Abs.localGrad = |g| (alg/AllKernels.scala:330-336)
        return np.full_like(w, abs(self.lam) / 2)


# ----------------------------------------------------------------------------------------------
# Activations (models/Activation.scala:26-84; grads alg/AllKernels.scala:288-304,391-437)
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class _Act:
    name: str
    threshold: float = 0.0

    @property
    def identity(self):
        return self.name == "Identity"

    def fwd(self, x):
        t = x.dtype.type
        if self.name == "Identity":
            return x
        if self.name == "Sigmoid":
            return (t(1) / (t(1) + np.exp(-x))).astype(x.dtype)
        if self.name == "Tanh":
            return np.tanh(x).astype(x.dtype)
        if self.name == "Softmax":
            # un-stabilised: e / reshape(sum_axis1 e)  (Activation.scala:63-69)
            e = np.exp(x)
            s = np.sum(e, axis=1, dtype=x.dtype)
            return (e / s.reshape(s.shape + (1,))).astype(x.dtype)
        if self.name == "ReLU":
            return np.maximum(t(self.threshold), x)
        raise ValueError(self.name)

    def bwd(self, g, x, y):
        t = g.dtype.type
        if self.name == "Identity":
            return g
        if self.name == "Sigmoid":  # (s * (1 - s)) * g   (alg:397-405)
            return (y * (t(1) - y)) * g
        if self.name == "Tanh":  # (1 - tanh^2) * g       (alg:414-420)
            return (t(1) - y * y) * g
        if self.name == "Softmax":
            # autodiff through Exp, Sum(axis=1), Reshape, Div:  dz = p * (g - sum_j g_j p_j)
            dot = np.sum(g * y, axis=1, keepdims=True, dtype=g.dtype)
            return y * (g - dot)
        if self.name == "ReLU":  # Max grad: [thr < x] * g, strict  (alg:297-303)
            return np.where(x > t(self.threshold), g, t(0)).astype(g.dtype)
        raise ValueError(self.name)

    def __repr__(self):
        return self.name


Identity = _Act("Identity")
Sigmoid = _Act("Sigmoid")
Tanh = _Act("Tanh")
Softmax = _Act("Softmax")


def ReLU(threshold: float = 0.0):
    return _Act("ReLU", float(threshold))


# ----------------------------------------------------------------------------------------------
# Layer base + composition (layer/Layer.scala:24-26, layer/Composed.scala:14-103)
# ----------------------------------------------------------------------------------------------
class Layer:
    trainable = True

    def params(self, in_shape) -> "OrderedDict[str, ParamDef]":
        return OrderedDict()

    def out_shape(self, in_shape):
        raise NotImplementedError

    def forward(self, x, p: Dict[str, np.ndarray], training: bool):
        """-> (y, cache, next_state: dict)"""
        raise NotImplementedError

    def backward(self, g, cache, p, need_dx: bool):
        """-> (dx | None, grads: dict)"""
        raise NotImplementedError

    def penalty(self, p, dtype):
        return dtype.type(0)

    def penalty_grads(self, p):
        return {}

    def leaves(self) -> List["Layer"]:
        return [self]

    def __rshift__(self, other: "Layer") -> "Layer":
        return Composed(self.leaves() + other.leaves())


class Composed(Layer):
    """`>>` flattens nested compositions; each leaf's params are prefixed with its index."""

    def __init__(self, layers: Sequence[Layer]):
        flat: List[Layer] = []
        for l in layers:
            flat.extend(l.leaves())
        assert flat, "at least 1 layer is required"
        self.layers = flat

    def leaves(self):
        return list(self.layers)

    def params(self, in_shape):
        return model_params(self, in_shape)

    def out_shape(self, in_shape):
        s = tuple(in_shape)
        for l in self.layers:
            s = l.out_shape(s)
        return s

    def __repr__(self):
        return " >> ".join(repr(l) for l in self.layers)


def flatten_layers(model: Layer) -> List[Layer]:
    return model.leaves()


def model_params(model: Layer, in_shape) -> "OrderedDict[str, ParamDef]":
    """All params (weights and state) keyed by the reference's path strings (Composed.scala:18-27).
    A single (non-composed) layer keeps un-prefixed paths, like `Params(Weights -> filters)` in
    `Conv2DLayerSpec.scala:35`."""
    leaves = model.leaves()
    out: "OrderedDict[str, ParamDef]" = OrderedDict()
    s = tuple(in_shape)
    single = not isinstance(model, Composed)
    for i, l in enumerate(leaves):
        for k, d in l.params(s).items():
            out[k if single else f"{i}/{k}"] = d
        s = l.out_shape(s)
    return out


def _layer_params(model, params, i):
    if not isinstance(model, Composed):
        return dict(params)
    pre = f"{i}/"
    return {k[len(pre):]: v for k, v in params.items() if k.startswith(pre)}


def model_forward(model: Layer, x, params: Dict[str, np.ndarray], training: bool = True):
    """-> (out, caches, next_state) ; next_state keyed by full paths (Composed.scala:42-53)."""
    leaves = model.leaves()
    single = not isinstance(model, Composed)
    caches, next_state = [], OrderedDict()
    h = x
    for i, l in enumerate(leaves):
        h, c, st = l.forward(h, _layer_params(model, params, i), training)
        caches.append(c)
        for k, v in st.items():
            next_state[k if single else f"{i}/{k}"] = v
    return h, caches, next_state


def model_backward(model: Layer, g, caches, params):
    """Reverse pass; gradients only for trainable weights; no gradient w.r.t. the model input
    (GradCalc walks only paths to requested leaves, grad/package.scala:17-40)."""
    leaves = model.leaves()
    single = not isinstance(model, Composed)
    grads: Dict[str, np.ndarray] = {}
    for i in range(len(leaves) - 1, -1, -1):
        l = leaves[i]
        need_dx = i > 0
        if g is None:
            if any(isinstance(leaves[j], (DenseKernel, Bias, Conv2DKernel, BatchNorm, RNN)) for j in range(i + 1)):
                raise NotImplementedError("gradient through an RNN's sequence input is outside the hot path")
            break
        g, gr = l.backward(g, caches[i], _layer_params(model, params, i), need_dx)
        for k, v in gr.items():
            grads[k if single else f"{i}/{k}"] = v
    return grads


def model_penalty(model: Layer, params, dtype):
    total = dtype.type(0)
    for i, l in enumerate(model.leaves()):
        total = total + l.penalty(_layer_params(model, params, i), dtype)
    return total


def model_penalty_grads(model: Layer, params):
    single = not isinstance(model, Composed)
    out = {}
    for i, l in enumerate(model.leaves()):
        for k, v in l.penalty_grads(_layer_params(model, params, i)).items():
            if v is not None:
                out[k if single else f"{i}/{k}"] = v
    return out


def split_weights_state(defs: "OrderedDict[str, ParamDef]"):
    w = [k for k, d in defs.items() if d.trainable]
    s = [k for k, d in defs.items() if not d.trainable]
    return w, s


# ----------------------------------------------------------------------------------------------
# Dense / Bias / Activate / Flatten
# ----------------------------------------------------------------------------------------------
class DenseKernel(Layer):
    """layer/Dense.scala:42-71: Z = X . W, X[B,in], W[in,out]; grads alg/AllKernels.scala:370-389."""

    def __init__(self, outputs, reg=Zero, initializer=I.GlorotUniform(), trainable=True):
        self.outputs, self.reg, self.initializer, self.trainable = outputs, reg, initializer, trainable

    def params(self, in_shape):
        return OrderedDict(weights=ParamDef((in_shape[1], self.outputs), self.initializer, "avg", self.trainable))

    def out_shape(self, in_shape):
        return (in_shape[0], self.outputs)

    def forward(self, x, p, training):
        return x @ p["weights"], x, {}

    def backward(self, g, x, p, need_dx):
        dw = x.T @ g
        dx = g @ p["weights"].T if need_dx else None
        return dx, ({"weights": dw} if self.trainable else {})

    def penalty(self, p, dtype):
        return self.reg.penalty(p["weights"])

    def penalty_grads(self, p):
        return {"weights": self.reg.grad(p["weights"])}

    def __repr__(self):
        return f"Dense({self.outputs})"


class Bias(Layer):
    """layer/Bias.scala:22-43; db = sum over all-but-last axes (alg/AllKernels.scala:12-33)."""

    def __init__(self, features, reg=Zero, initializer=I.Zeros, trainable=True):
        self.features, self.reg, self.initializer, self.trainable = features, reg, initializer, trainable

    def params(self, in_shape):
        return OrderedDict(weights=ParamDef((self.features,), self.initializer, "avg", self.trainable))

    def out_shape(self, in_shape):
        return tuple(in_shape)

    def forward(self, x, p, training):
        return x + p["weights"], None, {}

    def backward(self, g, cache, p, need_dx):
        db = np.sum(g.reshape(-1, g.shape[-1]), axis=0, dtype=g.dtype)
        return g, ({"weights": db} if self.trainable else {})

    def penalty(self, p, dtype):
        return self.reg.penalty(p["weights"]) if self.trainable else dtype.type(0)

    def penalty_grads(self, p):
        return {"weights": self.reg.grad(p["weights"])} if self.trainable else {}

    def __repr__(self):
        return f"Bias({self.features})"


class Activate(Layer):
    """layer/Activate.scala:12-21."""
    trainable = False

    def __init__(self, activation):
        self.activation = activation

    def out_shape(self, in_shape):
        return tuple(in_shape)

    def forward(self, x, p, training):
        y = self.activation.fwd(x)
        return y, (x, y), {}

    def backward(self, g, cache, p, need_dx):
        x, y = cache
        return self.activation.bwd(g, x, y), {}

    def __repr__(self):
        return repr(self.activation)


class _Flatten(Layer):
    """layer/Flatten.scala:10-21."""
    trainable = False

    def out_shape(self, in_shape):
        return (in_shape[0], int(np.prod(in_shape[1:])))

    def forward(self, x, p, training):
        assert x.ndim >= 2
        return x.reshape(x.shape[0], -1), x.shape, {}

    def backward(self, g, shape, p, need_dx):
        return g.reshape(shape), {}

    def __repr__(self):
        return "Flatten"


Flatten = _Flatten()


def Dense(outputs, activation=Identity, reg=Zero, bias=True, kernel_initializer=I.GlorotUniform(),
          bias_initializer=I.Zeros, trainable=True) -> Layer:
    """layer/Dense.scala:30-41: Dense >> Bias >> activation.layer"""
    layers: List[Layer] = [DenseKernel(outputs, reg, kernel_initializer, trainable)]
    if bias:
        layers.append(Bias(outputs, reg, bias_initializer))
    if not activation.identity:
        layers.append(Activate(activation))
    return layers[0] if len(layers) == 1 else Composed(layers)


# ----------------------------------------------------------------------------------------------
# Conv2D (math/nn/AllKernels.scala:12-103 padding, :143-261 ops; layer/Conv2D.scala:54-125)
# ----------------------------------------------------------------------------------------------
def _pad_amounts(padding, in_size, k, s, axis):
    """TF padding arithmetic. -> (out_size, pad_before, pad_after)"""
    if padding == "valid":
        return -(-(in_size - k + 1) // s), 0, 0
    if padding == "same":
        out = -(-in_size // s)
        total = max((out - 1) * s + k - in_size, 0)
        return out, total // 2, total - total // 2
    (pt, pb), (pl, pr) = padding
    before, after = (pt, pb) if axis == 0 else (pl, pr)
    return (in_size + before + after - k) // s + 1, before, after


def _conv_geometry(x_shape, kh, kw, strides, padding):
    n, h, w, c = x_shape
    sh, sw = strides
    oh, pt, pb = _pad_amounts(padding, h, kh, sh, 0)
    ow, pl, pr = _pad_amounts(padding, w, kw, sw, 1)
    return oh, ow, (pt, pb), (pl, pr)


def _im2col(xp, kh, kw, sh, sw, oh, ow):
    # xp: padded [N,HP,WP,C] -> [N*OH*OW, KH*KW*C] with (kh,kw,c) order = HWIO filter flattening
    win = sliding_window_view(xp, (kh, kw), axis=(1, 2))  # [N,H',W',C,kh,kw]
    win = win[:, ::sh, ::sw][:, :oh, :ow]
    cols = win.transpose(0, 1, 2, 4, 5, 3)
    n = xp.shape[0]
    return np.ascontiguousarray(cols).reshape(n * oh * ow, kh * kw * xp.shape[3])


def conv2d(x, w, strides=(1, 1), padding="valid"):
    """TF Conv2D, NHWC cross-correlation with HWIO filters (nn/AllKernels.scala:143-192)."""
    kh, kw, ci, co = w.shape
    oh, ow, ph, pw = _conv_geometry(x.shape, kh, kw, strides, padding)
    xp = np.pad(x, ((0, 0), ph, pw, (0, 0)))
    cols = _im2col(xp, kh, kw, strides[0], strides[1], oh, ow)
    y = cols @ w.reshape(kh * kw * ci, co)
    return y.reshape(x.shape[0], oh, ow, co)


def conv2d_wgrad(x, w_shape, g, strides=(1, 1), padding="valid"):
    """TF Conv2DBackpropFilter (nn/AllKernels.scala:211-235)."""
    kh, kw, ci, co = w_shape
    oh, ow, ph, pw = _conv_geometry(x.shape, kh, kw, strides, padding)
    xp = np.pad(x, ((0, 0), ph, pw, (0, 0)))
    cols = _im2col(xp, kh, kw, strides[0], strides[1], oh, ow)
    return (cols.T @ g.reshape(-1, co)).reshape(kh, kw, ci, co)


def conv2d_dgrad(x_shape, w, g, strides=(1, 1), padding="valid"):
    """TF Conv2DBackpropInput (nn/AllKernels.scala:237-261)."""
    kh, kw, ci, co = w.shape
    n, h, wd, _ = x_shape
    sh, sw = strides
    oh, ow, ph, pw = _conv_geometry(x_shape, kh, kw, strides, padding)
    dxp = np.zeros((n, h + ph[0] + ph[1], wd + pw[0] + pw[1], ci), g.dtype)
    g2 = g.reshape(-1, co)
    for a in range(kh):
        for b in range(kw):
            contrib = (g2 @ w[a, b].T).reshape(n, oh, ow, ci)
            dxp[:, a:a + (oh - 1) * sh + 1:sh, b:b + (ow - 1) * sw + 1:sw, :] += contrib
    return np.ascontiguousarray(dxp[:, ph[0]:ph[0] + h, pw[0]:pw[0] + wd, :])


class Conv2DKernel(Layer):
    def __init__(self, filters, kernel=(3, 3), strides=(1, 1), padding="valid", initializer=I.GlorotUniform(),
                 trainable=True):
        self.filters, self.kernel, self.strides, self.padding = filters, tuple(kernel), tuple(strides), padding
        self.initializer, self.trainable = initializer, trainable

    def params(self, in_shape):
        assert len(in_shape) == 4, f"Conv2D input should have a shape (NHWC) but was {in_shape}"
        return OrderedDict(weights=ParamDef((self.kernel[0], self.kernel[1], in_shape[3], self.filters),
                                            self.initializer, "avg", self.trainable))

    def out_shape(self, in_shape):
        oh, ow, _, _ = _conv_geometry(in_shape, self.kernel[0], self.kernel[1], self.strides, self.padding)
        return (in_shape[0], oh, ow, self.filters)

    def forward(self, x, p, training):
        return conv2d(x, p["weights"], self.strides, self.padding), x, {}

    def backward(self, g, x, p, need_dx):
        w = p["weights"]
        grads = {"weights": conv2d_wgrad(x, w.shape, g, self.strides, self.padding)} if self.trainable else {}
        dx = conv2d_dgrad(x.shape, w, g, self.strides, self.padding) if need_dx else None
        return dx, grads

    def __repr__(self):
        pad = self.padding.capitalize() if isinstance(self.padding, str) else f"Explicit{self.padding}"
        return f"Conv2D({self.filters},({self.kernel[0]},{self.kernel[1]}),({self.strides[0]},{self.strides[1]}),{pad},NHWC)"


def Conv2D(filters, kernel=(3, 3), strides=(1, 1), padding="valid", activation=Identity, bias=False,
           kernel_initializer=I.GlorotUniform(), bias_initializer=I.Zeros, trainable=True) -> Layer:
    """layer/Conv2D.scala:54-68: Conv2D [>> Bias] [>> activation]; bias defaults to FALSE."""
    layers: List[Layer] = [Conv2DKernel(filters, kernel, strides, padding, kernel_initializer, trainable)]
    if bias:
        layers.append(Bias(filters, Zero, bias_initializer))
    if not activation.identity:
        layers.append(Activate(activation))
    return layers[0] if len(layers) == 1 else Composed(layers)


# ----------------------------------------------------------------------------------------------
# Pool2D (nn/AllKernels.scala:270-397; layer/Pool2D.scala:28-57)
# ----------------------------------------------------------------------------------------------
def _pool_windows(x, window, strides, padding, fill):
    kh, kw = window
    oh, ow, ph, pw = _conv_geometry(x.shape, kh, kw, strides, padding)
    xp = np.pad(x, ((0, 0), ph, pw, (0, 0)), constant_values=fill)
    win = sliding_window_view(xp, (kh, kw), axis=(1, 2))[:, ::strides[0], ::strides[1]][:, :oh, :ow]
    return win.reshape(win.shape[:4] + (kh * kw,)), oh, ow, ph, pw  # [N,OH,OW,C,kh*kw] row-major (kh,kw)


def pool2d(x, window=(2, 2), strides=(1, 1), padding="valid", reduce="max"):
    if reduce == "max":
        win, *_ = _pool_windows(x, window, strides, padding, -np.inf)
        return win.max(axis=-1).astype(x.dtype)
    win, *_ = _pool_windows(x, window, strides, padding, 0)
    ones, *_ = _pool_windows(np.ones_like(x), window, strides, padding, 0)
    return (win.sum(axis=-1, dtype=x.dtype) / ones.sum(axis=-1, dtype=x.dtype)).astype(x.dtype)


def pool2d_grad(x, g, window=(2, 2), strides=(1, 1), padding="valid", reduce="max"):
    """MaxPoolGradV2 routes each window's gradient to the FIRST maximal element in row-major (kh,kw)
    order and sums over overlapping windows (golden: nn/KernelsSpec.scala:277-288)."""
    kh, kw = window
    n, h, w, c = x.shape
    sh, sw = strides
    if reduce == "max":
        win, oh, ow, ph, pw = _pool_windows(x, window, strides, padding, -np.inf)
        arg = win.argmax(axis=-1)  # first occurrence
        a, b = arg // kw, arg % kw
        ni, oi, oj, ci = np.meshgrid(np.arange(n), np.arange(oh), np.arange(ow), np.arange(c), indexing="ij")
        dxp = np.zeros((n, h + ph[0] + ph[1], w + pw[0] + pw[1], c), g.dtype)
        np.add.at(dxp, (ni, oi * sh + a, oj * sw + b, ci), g)
        return np.ascontiguousarray(dxp[:, ph[0]:ph[0] + h, pw[0]:pw[0] + w, :])
    ones, oh, ow, ph, pw = _pool_windows(np.ones_like(x), window, strides, padding, 0)
    per = g / ones.sum(axis=-1, dtype=g.dtype)
    dxp = np.zeros((n, h + ph[0] + ph[1], w + pw[0] + pw[1], c), g.dtype)
    for a in range(kh):
        for b in range(kw):
            dxp[:, a:a + (oh - 1) * sh + 1:sh, b:b + (ow - 1) * sw + 1:sw, :] += per
    return np.ascontiguousarray(dxp[:, ph[0]:ph[0] + h, pw[0]:pw[0] + w, :])


class Pool2D(Layer):
    """Defaults window 2x2, stride **1**, VALID, Max.  The layer's `build` does not forward `reduce`
    (layer/Pool2D.scala:36-42) so it is ALWAYS Max; `honor_reduce=True` is the
    flag to fix that."""
    trainable = False

    def __init__(self, window=(2, 2), strides=(1, 1), padding="valid", reduce="max", honor_reduce=False):
        self.window, self.strides, self.padding = tuple(window), tuple(strides), padding
        self.reduce = reduce
        self.effective = reduce if honor_reduce else "max"

    def out_shape(self, in_shape):
        oh, ow, _, _ = _conv_geometry(in_shape, self.window[0], self.window[1], self.strides, self.padding)
        return (in_shape[0], oh, ow, in_shape[3])

    def forward(self, x, p, training):
        return pool2d(x, self.window, self.strides, self.padding, self.effective), x, {}

    def backward(self, g, x, p, need_dx):
        return pool2d_grad(x, g, self.window, self.strides, self.padding, self.effective), {}

    def __repr__(self):
        return (f"Pool2D(({self.window[0]},{self.window[1]}),({self.strides[0]},{self.strides[1]}),"
                f"{self.padding.capitalize()},NHWC,{self.reduce.capitalize()})")


# ----------------------------------------------------------------------------------------------
# BatchNorm (layer/BatchNorm.scala:63-174; nn/AllKernels.scala:399-545, 663-684)
# ----------------------------------------------------------------------------------------------
class BatchNorm(Layer):
    """`bn_mode` applies to the fused (rank-4, channel-axis) path only:
       "reference": bug-for-bug wiring of FusedBatchNormV3 outputs -- moving
                    variance tracks the batch MEAN and the grad kernel receives
                    mean_in = Bessel-corrected variance, var_in = batch mean.  PARITY UNPINNED.
       "correct":   the TF op as intended."""

    def __init__(self, axes=(-1,), momentum=0.99, epsilon=1e-3, beta_initializer=I.Zeros, gamma_initializer=I.Ones,
                 moving_mean_initializer=I.Zeros, moving_variance_initializer=I.Ones, fused=None, trainable=True,
                 bn_mode="reference"):
        self.axes, self.momentum, self.epsilon = tuple(axes), momentum, epsilon
        self.inits = (beta_initializer, gamma_initializer, moving_mean_initializer, moving_variance_initializer)
        self.fused, self.trainable, self.bn_mode = fused, trainable, bn_mode

    def _norm_axes(self, rank):
        return tuple(a % rank for a in self.axes)

    def _param_shape(self, in_shape):
        ax = self._norm_axes(len(in_shape))
        return tuple(d if i in ax else 1 for i, d in enumerate(in_shape))

    def _use_fused(self, in_shape):
        ax = self._norm_axes(len(in_shape))
        can = len(in_shape) == 4 and ax == (3,)
        if len(in_shape) == 4 and ax == (1,):
            raise NotImplementedError("NCHW is unsupported by the oracle")
        if self.fused is True and not can:
            raise ValueError(f"Cannot use fused implementation with input shape {in_shape}")
        return can and self.fused in (None, True)

    def params(self, in_shape):
        s = self._param_shape(in_shape)
        b, g, mm, mv = self.inits
        return OrderedDict(beta=ParamDef(s, b, "avg", self.trainable), gamma=ParamDef(s, g, "avg", self.trainable),
                           moving_mean=ParamDef(s, mm, "avg", False), moving_variance=ParamDef(s, mv, "avg", False))

    def out_shape(self, in_shape):
        return tuple(in_shape)

    def forward(self, x, p, training):
        t = x.dtype.type
        training = training and self.trainable
        eps, mom = t(self.epsilon), t(self.momentum)
        red = tuple(i for i in range(x.ndim) if i not in self._norm_axes(x.ndim))
        gamma, beta, mmean, mvar = p["gamma"], p["beta"], p["moving_mean"], p["moving_variance"]
        fused = self._use_fused(x.shape)
        if training:
            m = np.mean(x, axis=red, keepdims=True, dtype=x.dtype)
            v = np.mean((x - m) ** 2, axis=red, keepdims=True, dtype=x.dtype)  # biased (alg:505-513)
        else:
            m, v = mmean, mvar
        if fused:
            y = (x - m) * ((t(1) / np.sqrt(v + eps)) * gamma) + beta
        else:
            inv = (t(1) / np.sqrt(v + eps)) * gamma  # rsqrt(v+eps)*gamma   (nn:679-681)
            y = ((x * inv) - (m * inv)) + beta
        state = {}
        if training:
            r = t(int(np.prod([x.shape[i] for i in red])))
            batch_mean = m
            if fused and self.bn_mode == "reference":
                batch_var = m  # TakeOut index 1 used twice (nn/AllKernels.scala:444-445)
            elif fused:
                batch_var = v * (r / (r - t(1)))  # TF returns the Bessel-corrected variance
            else:
                batch_var = v
            one_m = t(1) - mom
            state["moving_mean"] = (mom * mmean) + (one_m * batch_mean)  # decayingAvg (alg:541-544)
            state["moving_variance"] = (mom * mvar) + (one_m * batch_var)
        return y.astype(x.dtype), (x, m, v, red, fused, training), state

    def backward(self, g, cache, p, need_dx):
        x, m, v, red, fused, training = cache
        t = x.dtype.type
        eps = t(self.epsilon)
        gamma = p["gamma"]
        r = t(int(np.prod([x.shape[i] for i in red])))
        if not training:  # frozen: plain affine map
            inv = (t(1) / np.sqrt(v + eps)) * gamma
            return g * inv, {}
        if fused and self.bn_mode == "reference":
            mean_in = v * (r / (r - t(1)))  # reserveSpace1 := TF output 2 (batch_variance, Bessel)
            var_in = m                      # reserveSpace2 := TF output 3 (saved mean)
        else:
            mean_in, var_in = m, v
        xc = x - mean_in
        istd = t(1) / np.sqrt(var_in + eps)
        dgamma = np.sum(g * xc * istd, axis=red, keepdims=True, dtype=x.dtype)
        dbeta = np.sum(g, axis=red, keepdims=True, dtype=x.dtype)
        dx = None
        if need_dx:
            mg = np.mean(g, axis=red, keepdims=True, dtype=x.dtype)
            mgx = np.mean(g * xc, axis=red, keepdims=True, dtype=x.dtype)
            dx = gamma * istd * (g - mg - xc * mgx / (var_in + eps))
        return dx, {"gamma": dgamma.astype(x.dtype), "beta": dbeta.astype(x.dtype)}

    def __repr__(self):
        return f"BatchNorm({list(self.axes)})"


# ----------------------------------------------------------------------------------------------
# RNN / SimpleRNNCell / LSTMCell (layer/RNN.scala:27-355)
# ----------------------------------------------------------------------------------------------
class SimpleRNNCell(Layer):
    """h = act((x.Wk + state.Wr) + b); NOTE the state fed back is the PRE-bias, PRE-activation sum
    because `SimpleRNNCell.build` returns `Params(State -> result)` before Bias/Activate are composed
    (layer/RNN.scala:190-196) -- a reference quirk reproduced here."""

    def __init__(self, units, activation=Tanh, bias=True, kernel_initializer=I.GlorotUniform(),
                 recurrent_initializer=I.Orthogonal(), bias_initializer=I.Zeros, trainable=True):
        self.units, self.activation, self.bias = units, activation, bias
        self.inits = (kernel_initializer, recurrent_initializer, bias_initializer)
        self.trainable = trainable

    def params(self, in_shape):  # in_shape = (batch, features)
        k, r, b = self.inits
        out = OrderedDict()
        out["0/kernel_weights"] = ParamDef((in_shape[1], self.units), k, "avg", self.trainable)
        out["0/recurrent_weights"] = ParamDef((self.units, self.units), r, "avg", self.trainable)
        out["0/state"] = ParamDef((in_shape[0], self.units), I.Zeros, None, False)
        if self.bias:
            out["1/weights"] = ParamDef((self.units,), b, "avg", True)
        return out

    def out_shape(self, in_shape):
        return (in_shape[0], self.units)

    def step(self, x, state, p):
        z = (x @ p["0/kernel_weights"]) + (state @ p["0/recurrent_weights"])
        zb = z + p["1/weights"] if self.bias else z
        h = self.activation.fwd(zb)
        return h, z, (x, state, zb, h)

    def step_bwd(self, dh, dstate_next, cache, p):
        x, state, zb, h = cache
        dzb = self.activation.bwd(dh, zb, h)
        dz = dzb + dstate_next  # state_next = z (pre-bias)
        g = {"0/kernel_weights": x.T @ dz, "0/recurrent_weights": state.T @ dz}
        if self.bias:
            g["1/weights"] = np.sum(dzb, axis=0, dtype=dh.dtype)
        self._last_dx = dz @ p["0/kernel_weights"].T  # MatMul grad w.r.t. the step input (a stacked RNN below needs it)
        return dz @ p["0/recurrent_weights"].T, g

    def __repr__(self):
        s = f"SimpleRNNCell({self.units})"
        if self.bias:
            s += f" >> Bias({self.units})"
        if not self.activation.identity:
            s += f" >> {self.activation!r}"
        return s


class LSTMCell(Layer):
    """layer/RNN.scala:257-339: per gate G in (forget,input,gate,output):
    z_G = (x.Wk_G + h_prev.Wr_G) + b_G ; c = c_prev*f + i*g ; h = o*act(c)."""
    GATES = ("forget", "input", "gate", "output")

    def __init__(self, units, activation=Tanh, recurrent_activation=Sigmoid, bias=True,
                 kernel_initializer=I.GlorotUniform(), recurrent_initializer=I.Orthogonal(),
                 bias_initializer=I.Zeros, bias_forget_initializer=I.Ones, trainable=True):
        self.units, self.activation, self.recurrent_activation, self.bias = units, activation, recurrent_activation, bias
        self.inits = (kernel_initializer, recurrent_initializer, bias_initializer, bias_forget_initializer)
        self.trainable = trainable

    def params(self, in_shape):
        k, r, b, bf = self.inits
        out = OrderedDict()
        for gname in self.GATES:
            out[f"{gname}/0/kernel_weights"] = ParamDef((in_shape[1], self.units), k, "avg", self.trainable)
            out[f"{gname}/0/recurrent_weights"] = ParamDef((self.units, self.units), r, "avg", self.trainable)
            if self.bias:
                out[f"{gname}/1/weights"] = ParamDef((self.units,), bf if gname == "forget" else b, "avg", True)
        out["cell_state"] = ParamDef((in_shape[0], self.units), I.Zeros, None, False)
        out["hidden_state"] = ParamDef((in_shape[0], self.units), I.Zeros, None, False)
        return out

    def out_shape(self, in_shape):
        return (in_shape[0], self.units)

    def _act(self, gname):
        return self.activation if gname == "gate" else self.recurrent_activation

    def step(self, x, c_prev, h_prev, p):
        zs, acts = {}, {}
        for gname in self.GATES:
            z = (x @ p[f"{gname}/0/kernel_weights"]) + (h_prev @ p[f"{gname}/0/recurrent_weights"])
            if self.bias:
                z = z + p[f"{gname}/1/weights"]
            zs[gname] = z
            acts[gname] = self._act(gname).fwd(z)
        f, i, g, o = (acts[n] for n in self.GATES)
        c = c_prev * f + i * g
        ac = self.activation.fwd(c)
        h = o * ac
        return h, c, (x, c_prev, h_prev, zs, acts, c, ac)

    def step_bwd(self, dh, dc_next, cache, p):
        x, c_prev, h_prev, zs, acts, c, ac = cache
        f, i, g, o = (acts[n] for n in self.GATES)
        do = dh * ac
        dc = self.activation.bwd(dh * o, c, ac) + dc_next
        dacts = {"forget": dc * c_prev, "input": dc * g, "gate": dc * i, "output": do}
        dc_prev = dc * f
        dh_prev = np.zeros_like(h_prev)
        dx = np.zeros_like(x)
        grads = {}
        for gname in self.GATES:
            dz = self._act(gname).bwd(dacts[gname], zs[gname], acts[gname])
            grads[f"{gname}/0/kernel_weights"] = x.T @ dz
            grads[f"{gname}/0/recurrent_weights"] = h_prev.T @ dz
            if self.bias:
                grads[f"{gname}/1/weights"] = np.sum(dz, axis=0, dtype=dh.dtype)
            dh_prev = dh_prev + dz @ p[f"{gname}/0/recurrent_weights"].T
            dx = dx + dz @ p[f"{gname}/0/kernel_weights"].T  # MatMul grad w.r.t. the step input (alg/AllKernels.scala:370-389)
        self._last_dx = dx
        return dh_prev, dc_prev, grads

    def __repr__(self):
        return f"LSTMCell({self.units},{self.activation!r},{self.recurrent_activation!r},{str(self.bias).lower()})"


class RNN(Layer):
    """layer/RNN.scala:27-94: x[B,T,F]; statically unrolled; stateless => zero initial state and the
    state params are not part of `model.params`; output = last step (or the sequence)."""

    def __init__(self, cell, return_sequence=False, stateful=False):
        self.cell, self.return_sequence, self.stateful = cell, return_sequence, stateful
        self.trainable = cell.trainable

    def _drop_time(self, s):
        return (s[0],) + tuple(s[2:])

    def params(self, in_shape):
        allp = self.cell.params(self._drop_time(in_shape))
        if self.stateful:  # `if (stateful) state ++ weights else weights` (RNN.scala:32-35)
            return allp
        return OrderedDict((k, d) for k, d in allp.items() if not k.split("/")[-1].endswith("state"))

    def _state_keys(self):
        return ("cell_state", "hidden_state") if isinstance(self.cell, LSTMCell) else ("0/state",)

    def out_shape(self, in_shape):
        co = self.cell.out_shape(self._drop_time(in_shape))
        return (co[0], in_shape[1], co[1]) if self.return_sequence else co

    def forward(self, x, p, training):
        assert x.ndim == 3, "RNN requires input to have Shape(batch, time, features)"
        b, T, _ = x.shape
        u = self.cell.units
        h = np.zeros((b, u), x.dtype)
        c = np.zeros((b, u), x.dtype)
        caches, outs = [], []
        lstm = isinstance(self.cell, LSTMCell)
        if self.stateful:  # the state params ARE the initial state (RNN.scala:66-67); zeros otherwise
            if lstm:
                c, h = p["cell_state"].astype(x.dtype), p["hidden_state"].astype(x.dtype)
            else:
                h = p["0/state"].astype(x.dtype)
        for t in range(T):
            xt = x[:, t, :]
            if lstm:
                out, c, cache = self.cell.step(xt, c, h, p)
                h = out
            else:
                out, h, cache = self.cell.step(xt, h, p)  # h := pre-activation state (quirk)
            caches.append(cache)
            outs.append(out)
        y = np.stack(outs, axis=1) if self.return_sequence else outs[-1]
        # the last cell state is the layer's output state (RNN.scala:68-72); only a stateful RNN declares it, so only then does it
        # travel on with the params (the stateless case is SURVEY App. B-Q13)
        state = {}
        if self.stateful:
            state = {"cell_state": c, "hidden_state": h} if lstm else {"0/state": h}
        return y, (caches, x.shape), state

    def backward(self, g, cache, p, need_dx):
        caches, xshape = cache
        b, T, _ = xshape
        u = self.cell.units
        lstm = isinstance(self.cell, LSTMCell)
        grads: Dict[str, np.ndarray] = {}
        dstate = np.zeros((b, u), g.dtype)  # dL/d(recurrent input of step t+1)
        dc = np.zeros((b, u), g.dtype)
        dx = np.zeros(xshape, g.dtype) if need_dx else None
        for t in range(T - 1, -1, -1):
            dout = g[:, t, :] if self.return_sequence else (g if t == T - 1 else np.zeros((b, u), g.dtype))
            if lstm:
                dstate, dc, gr = self.cell.step_bwd(dout + dstate, dc, caches[t], p)
                if dx is not None:
                    dx[:, t, :] = self.cell._last_dx  # a stacked RNN: the layer below needs dL/d(sequence input)
            else:
                dstate, gr = self.cell.step_bwd(dout, dstate, caches[t], p)
                if dx is not None:
                    dx[:, t, :] = self.cell._last_dx
            for k, v in gr.items():
                grads[k] = grads[k] + v if k in grads else v
        return dx, grads

    def __repr__(self):
        return f"RNN({self.cell!r})"


def SimpleRNN(units, activation=Tanh, bias=True, return_sequence=False, **kw):
    return RNN(SimpleRNNCell(units, activation, bias, **kw), return_sequence)


def LSTM(units, activation=Tanh, recurrent_activation=Sigmoid, bias=True, return_sequence=False, **kw):
    return RNN(LSTMCell(units, activation, recurrent_activation, bias, **kw), return_sequence)
