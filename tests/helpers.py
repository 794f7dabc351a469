"""Builds the SAME model twice from one spec: in the CPU oracle's DSL and in the product's DSL, so parity tests
drive both with identical inputs.  Spec items:
  ("dense", units, act) | ("conv", filters, act, kernel, strides, padding, bias) | ("pool", window, strides, padding)
  ("flatten",) | ("bn", momentum, bn_mode) | ("lstm", units) | ("act", act) | ("bias", features)"""
import numpy as np

import oracle as O
import scanet3_b200 as S

_OACT = {"identity": O.Identity, "sigmoid": O.Sigmoid, "tanh": O.Tanh, "softmax": O.Softmax, "relu": O.ReLU()}
_SACT = {"identity": S.Identity, "sigmoid": S.Sigmoid, "tanh": S.Tanh, "softmax": S.Softmax, "relu": S.ReLU()}
OLOSS = {"mse": O.MeanSquaredError, "bce": O.BinaryCrossentropy, "cce": O.CategoricalCrossentropy}
SLOSS = {"mse": S.MeanSquaredError, "bce": S.BinaryCrossentropy, "cce": S.CategoricalCrossentropy}


def build_models(spec, seed=1):
    ol, sl = [], []
    for it in spec:
        kind = it[0]
        if kind == "dense":
            _, units, act = it
            ol.append(O.Dense(units, _OACT[act], kernel_initializer=O.GlorotUniform(seed)))
            sl.append(S.Dense(units, _SACT[act], kernelInitializer=S.GlorotUniform(seed)))
        elif kind == "conv":
            _, filters, act, kernel, strides, padding, bias = it
            # padding: "Valid" | "Same" | ((top, bottom), (left, right)) for Explicit
            opad = padding.lower() if isinstance(padding, str) else padding
            spad = padding if isinstance(padding, str) else S.Explicit(*padding)
            ol.append(O.Conv2D(filters, kernel, strides, opad, _OACT[act], bias, O.GlorotUniform(seed)))
            sl.append(S.Conv2D(filters, kernel, strides, spad, activation=_SACT[act], bias=bias,
                               kernelInitializer=S.GlorotUniform(seed)))
        elif kind == "pool":
            window, strides, padding = it[1], it[2], it[3]  # ("pool", window, strides, padding[, reduce honoured])
            if len(it) > 4:
                ol.append(O.Pool2D(window, strides, padding.lower(), it[4].lower(), honor_reduce=True))
                sl.append(S.Pool2D(window, strides, padding, reduce=it[4], honorReduce=True))
            else:
                ol.append(O.Pool2D(window, strides, padding.lower()))
                sl.append(S.Pool2D(window, strides, padding))
        elif kind == "flatten":
            ol.append(O.Flatten)
            sl.append(S.Flatten)
        elif kind == "bn":
            _, momentum, mode = it
            ol.append(O.BatchNorm(momentum=momentum, bn_mode=mode))
            sl.append(S.BatchNorm(momentum=momentum, bnMode=mode))
        elif kind == "lstm":
            # ("lstm", units[, return_sequence[, stateful]])
            units, seq, stateful = it[1], (len(it) > 2 and bool(it[2])), (len(it) > 3 and bool(it[3]))
            ocell = O.LSTMCell(units, kernel_initializer=O.GlorotUniform(seed), recurrent_initializer=O.Orthogonal(seed=seed))
            ol.append(O.RNN(ocell, seq, stateful))
            sl.append(S.LSTM(units, kernelInitializer=S.GlorotUniform(seed), recurrentInitializer=S.Orthogonal(seed=seed),
                             returnSequence=seq, stateful=stateful))
        elif kind == "rnn":
            # ("rnn", units, act[, return_sequence[, bias[, stateful]]])
            units, act = it[1], it[2]
            seq = len(it) > 3 and bool(it[3])
            bias = it[4] if len(it) > 4 else True
            stateful = len(it) > 5 and bool(it[5])
            ocell = O.SimpleRNNCell(units, _OACT[act], bias, kernel_initializer=O.GlorotUniform(seed),
                                    recurrent_initializer=O.Orthogonal(seed=seed))
            ol.append(O.RNN(ocell, seq, stateful))
            sl.append(S.SimpleRNN(units, _SACT[act], bias, kernelInitializer=S.GlorotUniform(seed),
                                  recurrentInitializer=S.Orthogonal(seed=seed), returnSequence=seq, stateful=stateful))
        elif kind == "act":
            ol.append(O.Activate(_OACT[it[1]]))
            sl.append(S.Activate(_SACT[it[1]]))
        elif kind == "bias":
            ol.append(O.Bias(it[1]))
            sl.append(S.Bias(it[1]))
        else:
            raise ValueError(kind)
    om = ol[0] if len(ol) == 1 else O.Composed(ol)
    sm = sl[0] if len(sl) == 1 else S.Composed(sl)
    return om, sm


OALG = {
    "sgd": lambda **k: O.SGD(**k), "adam": lambda **k: O.Adam(**k), "adagrad": lambda **k: O.AdaGrad(**k),
    "rmsprop": lambda **k: O.RMSProp(**k), "adadelta": lambda **k: O.AdaDelta(**k), "adamax": lambda **k: O.Adamax(**k),
    "nadam": lambda **k: O.Nadam(**k), "amsgrad": lambda **k: O.AMSGrad(**k),
}
SALG = {
    "sgd": lambda **k: S.SGD(**k), "adam": lambda **k: S.Adam(**k), "adagrad": lambda **k: S.AdaGrad(**k),
    "rmsprop": lambda **k: S.RMSProp(**k), "adadelta": lambda **k: S.AdaDelta(**k),
    "adamax": lambda **k: S.Adamax(**{("initAcc" if a == "init_acc" else a): v for a, v in k.items()}),
    "nadam": lambda **k: S.Nadam(**{("initAcc" if a == "init_acc" else a): v for a, v in k.items()}),
    "amsgrad": lambda **k: S.AMSGrad(**k),
}


def synth_classification(n, in_shape, classes, seed):
    """features U[1/256,1], one-hot labels with class = row % classes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.uniform(1.0 / 256, 1.0, size=(n,) + tuple(in_shape)).astype(np.float32)
    y = np.zeros((n, classes), np.float32)
    y[np.arange(n), np.arange(n) % classes] = 1.0
    return x, y


def assert_params_close(got, want, rtol, atol, what=""):
    assert set(got.keys()) == set(want.keys()), (sorted(got.keys()), sorted(want.keys()))
    for k in want:
        np.testing.assert_allclose(got[k].reshape(-1), np.asarray(want[k]).reshape(-1), rtol=rtol, atol=atol,
                                   err_msg=f"{what} param {k}")
