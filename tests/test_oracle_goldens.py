"""Pins the CPU oracle against every golden vector the reference's own tests hold for the hot path
(SURVEY.md Appendix D).  Each test cites the reference spec (file:line under
/root/reference/src/test/scala/scanet/) whose literals it re-encodes.  CPU only."""
import os

import numpy as np
import pytest

import oracle as O

f32 = np.float32


def rnd(x, k):
    """`roundAt(k)` restated: round(x * 10^k) / 10^k in the tensor's own precision, TF Round = half-to-even
    (math/alg/AllKernels.scala:526-530); literals are compared as fp32 like `Tensor.matrix(0.574442f)`."""
    x = np.asarray(x)
    t = np.float64 if x.dtype == np.float64 and x.ndim and x.size > 30 else np.float32
    p = t(10.0) ** t(k)
    return np.round(np.asarray(x, t) * p) / p


def close_rounded(x, golden, k):
    """golden was produced by `roundAt(k)` on TF-CPU fp32 output: accept half a unit in the last printed digit
    plus 2 fp32 ulp (TF's Eigen exp/tanh differ from NumPy's at the 1-ulp level)."""
    x = np.asarray(x, np.float64)
    g = np.asarray(golden, np.float64)
    return bool(np.all(np.abs(x - g) <= 0.5 * 10.0 ** (-k) + 2.4e-7 * np.maximum(np.abs(g), 1e-3)))


X4 = np.array([[0, 0, 1], [0, 1, 1], [1, 0, 1], [1, 1, 1]], f32)
W34 = np.array([[1, 0.5, 1, 0.1], [0.1, 1, 1, 1], [1, 0, 0.2, 0.3]], f32)
IMG = np.array([[2, 1, 2, 0, 1], [1, 3, 2, 2, 3], [1, 1, 3, 3, 0], [2, 2, 0, 1, 1], [0, 0, 3, 1, 2]], np.float64)
FIL = np.array([[2, 3], [0, 1]], np.float64)


# ---- models/layer/DenseLayerSpec.scala:17-78 ---------------------------------------------------
def test_dense_forward():
    model = O.Dense(4, O.Sigmoid)
    y, _, _ = O.model_forward(model, X4, {"0/weights": W34, "1/weights": np.zeros(4, f32)})
    exp = [[0.731059, 0.500000, 0.549834, 0.574442], [0.750260, 0.731059, 0.768525, 0.785835],
           [0.880797, 0.622459, 0.768525, 0.598688], [0.890903, 0.817574, 0.900250, 0.802184]]
    assert close_rounded(y, exp, 6)


def test_dense_penalty_l2():
    model = O.Dense(4, O.Sigmoid, reg=O.L2(1.0))
    w = np.array([[0, 0], [1, 0.5], [0.1, 1], [1, 0]], f32)
    pen = O.model_penalty(model, {"0/weights": w, "1/weights": np.zeros(2, f32)}, np.dtype(f32))
    assert f32(pen) == f32(1.63)


def test_dense_bce_gradient():
    lm = O.LossModel(O.Dense(4, O.Sigmoid), O.BinaryCrossentropy)
    y = np.array([[0.7310586, 0.5, 0.549834, 0.5744425], [0.7502601, 0.7310586, 0.76852477, 0.7858349],
                  [0.8807971, 0.6224593, 0.76852477, 0.59868765], [0.89090323, 0.81757444, 0.9002496, 0.80218387]], f32)
    _, g, _ = lm.grad_stateful(X4, y, {"0/weights": np.zeros((3, 4), f32), "1/weights": np.zeros(4, f32)})
    wg = np.array([[-0.048231270, -0.027502108, -0.041798398, -0.025054470],
                   [-0.040072710, -0.034289565, -0.041798398, -0.036751173],
                   [-0.078313690, -0.041943270, -0.061695820, -0.047571808]], f32)
    bg = np.array([-0.078313690, -0.041943270, -0.061695820, -0.047571808], f32)
    np.testing.assert_allclose(g["0/weights"], wg, rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(g["1/weights"], bg, rtol=2e-6, atol=1e-8)


# ---- models/layer/ComposedLayerSpec.scala:17-120 -----------------------------------------------
def test_composed_forward_and_param_index_convention():
    model = O.Dense(4, O.Sigmoid) >> O.Dense(1, O.Sigmoid)
    assert list(O.model_params(model, (4, 3)).keys()) == ["0/weights", "1/weights", "3/weights", "4/weights"]
    p = {"0/weights": W34, "1/weights": np.zeros(4, f32), "3/weights": np.array([[0.1], [0.5], [1], [0]], f32),
         "4/weights": np.zeros(1, f32)}
    y, _, _ = O.model_forward(model, X4, p)
    assert close_rounded(y, [[0.705357], [0.770136], [0.762753], [0.801886]], 6)


def test_composed_penalty():
    model = O.Dense(4, O.Sigmoid, reg=O.L2(1.0)) >> O.Dense(1, O.Sigmoid, reg=O.L2(1.0))
    p = {"0/weights": np.array([[1, 0.1, 1], [0.5, 1, 0], [1, 1, 0.2], [0.1, 1, 0.3]], f32),
         "1/weights": np.zeros(4, f32), "3/weights": np.array([[0.1, 0.5, 1, 0]], f32), "4/weights": np.zeros(1, f32)}
    assert abs(float(O.model_penalty(model, p, np.dtype(f32))) - 3.83) < 1e-6


def _mlp3():
    x = np.array([[0, 0, 1], [0, 1, 1]], f32)
    y = np.array([[1], [0]], f32)
    p = {"0/weights": np.array([[1, 0.5, 1], [0.1, 1, 1], [1, 0, 0.2]], f32), "1/weights": np.zeros(3, f32),
         "3/weights": np.array([[0.1], [0.5], [1]], f32), "4/weights": np.zeros(1, f32)}
    return x, y, p


def test_composed_mse_loss():
    x, y, p = _mlp3()
    lm = O.LossModel(O.Dense(3, O.Sigmoid) >> O.Dense(1, O.Sigmoid), O.MeanSquaredError)
    assert rnd(O.eval_loss(lm, x, y, p), 6) == rnd(0.339962, 6)


def test_composed_mse_loss_with_penalty():
    x, y, p = _mlp3()
    lm = O.LossModel(O.Dense(3, O.Sigmoid, O.L2(1.0)) >> O.Dense(1, O.Sigmoid, O.L2(1.0)), O.MeanSquaredError)
    assert abs(float(O.eval_loss(lm, x, y, p)) - 3.6199622) < 5e-7


# ---- models/RegressionSpec.scala:13-105 ---------------------------------------------------------
def test_linear_regression_loss_result_grad():
    lin = O.Dense(1, O.Identity, kernel_initializer=O.Zeros)
    assert repr(lin) == "Dense(1) >> Bias(1)"
    x = np.array([[1, 2], [2, 4]], f32)
    y = np.array([[6], [12]], f32)
    p = {"0/weights": np.array([[2], [3]], f32), "1/weights": np.array([1], f32)}
    lm = O.LossModel(lin, O.MeanSquaredError)
    assert float(O.eval_loss(lm, x, y, p)) == 17.0
    assert float(O.eval_loss(lm, x.astype(np.float64), y.astype(np.float64),
                             {k: v.astype(np.float64) for k, v in p.items()})) == 17.0
    out, _, _ = O.model_forward(lin, x, p)
    assert np.array_equal(out, np.array([[9], [17]], f32))
    _, g, _ = lm.grad_stateful(x, y, {"0/weights": np.zeros((2, 1), f32), "1/weights": np.zeros(1, f32)})
    assert np.array_equal(g["0/weights"], np.array([[-30], [-60]], f32))
    assert np.array_equal(g["1/weights"], np.array([-18], f32))


def test_logistic_regression_loss_result_grad():
    log = O.Dense(1, O.Sigmoid, kernel_initializer=O.Zeros)
    assert repr(log) == "Dense(1) >> Bias(1) >> Sigmoid"
    x = np.array([[0.34, 0.78], [0.6, 0.86]], f32)
    y = np.array([[0.402], [0.47800002]], f32)
    p = {"0/weights": np.array([[0.2], [0.3]], f32), "1/weights": np.array([0.1], f32)}
    lm = O.LossModel(log, O.BinaryCrossentropy)
    assert abs(float(O.eval_loss(lm, x, y, p)) - 0.7422824) < 1e-6
    out, _, _ = O.model_forward(log, x, p)
    assert np.array_equal(rnd(out, 6), rnd([[0.599168], [0.617276]], 6))
    _, g, _ = lm.grad_stateful(x, y, p)
    assert np.array_equal(rnd(g["0/weights"], 6), rnd([[0.075301], [0.136784]], 6))
    assert np.array_equal(rnd(g["1/weights"], 6), rnd([0.168222], 6))


# ---- Bias / Activate / Flatten layer specs ------------------------------------------------------
def test_bias_activate_flatten_layers():
    y, _, _ = O.model_forward(O.Bias(3), X4, {"weights": np.array([1, 2, 3], f32)})
    assert np.array_equal(y, np.array([[1, 2, 4], [1, 3, 4], [2, 2, 4], [2, 3, 4]], f32))  # BiasLayerSpec:13-28
    y, _, _ = O.model_forward(O.Activate(O.Sigmoid), X4, {})  # ActivateLayerSpec:13-30
    np.testing.assert_allclose(y, np.where(X4 > 0, f32(0.7310586), f32(0.5)), rtol=1e-7)
    x = np.ones((10, 5, 5), f32)
    y, _, _ = O.model_forward(O.Flatten, x, {})  # FlattenLayerSpec:10-16
    assert y.shape == (10, 25)
    assert repr(O.Bias(3)) == "Bias(3)" and repr(O.Activate(O.Sigmoid)) == "Sigmoid"


# ---- math/nn/KernelsSpec.scala:45-188 (conv) ----------------------------------------------------
def _img():
    return IMG.reshape(1, 5, 5, 1), FIL.reshape(2, 2, 1, 1)


@pytest.mark.parametrize("padding,strides,expected", [
    ("same", (1, 1), [[10, 10, 6, 6, 2], [12, 15, 13, 13, 6], [7, 11, 16, 7, 0], [10, 7, 4, 7, 2], [0, 9, 9, 8, 4]]),
    ("same", (2, 2), [[10, 6, 2], [7, 16, 0], [0, 9, 4]]),
    ("valid", (1, 1), [[10, 10, 6, 6], [12, 15, 13, 13], [7, 11, 16, 7], [10, 7, 4, 7]]),
    ("valid", (2, 2), [[10, 6], [7, 16]]),
    (((1, 0), (1, 0)), (2, 2), [[2, 2, 1], [4, 15, 13], [6, 7, 7]]),
])
def test_conv2d_forward(padding, strides, expected):
    x, w = _img()
    y = O.conv2d(x, w, strides, padding)
    e = np.array(expected, np.float64)
    assert y.shape == (1,) + e.shape + (1,)
    assert np.array_equal(y.reshape(e.shape), e)


def test_conv2d_filter_and_input_gradients():
    x, w = _img()
    g = np.ones((1, 4, 4, 1))
    assert np.array_equal(O.conv2d_wgrad(x, w.shape, g).reshape(2, 2), [[26, 25], [25, 27]])  # :146-157
    exp = [[2, 5, 5, 5, 3], [2, 6, 6, 6, 4], [2, 6, 6, 6, 4], [2, 6, 6, 6, 4], [0, 1, 1, 1, 1]]
    assert np.array_equal(O.conv2d_dgrad(x.shape, w, g).reshape(5, 5), exp)  # :174-188


def test_conv2d_layer():  # models/layer/Conv2DLayerSpec.scala:13-36
    model = O.Conv2D(filters=1, kernel=(2, 2))
    assert repr(model) == "Conv2D(1,(2,2),(1,1),Valid,NHWC)"
    x, w = _img()
    y, _, _ = O.model_forward(model, x, {"weights": w})
    assert np.array_equal(y.reshape(4, 4), [[10, 10, 6, 6], [12, 15, 13, 13], [7, 11, 16, 7], [10, 7, 4, 7]])


# ---- math/nn/KernelsSpec.scala:200-305 (pool) ---------------------------------------------------
@pytest.mark.parametrize("padding,strides,reduce,expected", [
    ("same", (1, 1), "max", [[3, 3, 2, 3, 3], [3, 3, 3, 3, 3], [2, 3, 3, 3, 1], [2, 3, 3, 2, 2], [0, 3, 3, 2, 2]]),
    ("same", (1, 1), "avg", [[1.75, 2.0, 1.5, 1.5, 2.0], [1.5, 2.25, 2.5, 2.0, 1.5], [1.5, 1.5, 1.75, 1.25, 0.5],
                             [1.0, 1.25, 1.25, 1.25, 1.5], [0.0, 1.5, 2.0, 1.5, 2.0]]),
    ("same", (2, 2), "max", [[3, 2, 3], [2, 3, 1], [0, 3, 2]]),
    ("valid", (1, 1), "max", [[3, 3, 2, 3], [3, 3, 3, 3], [2, 3, 3, 3], [2, 3, 3, 2]]),
    ("valid", (2, 2), "max", [[3, 2], [2, 3]]),
])
def test_pool2d_forward(padding, strides, reduce, expected):
    x, _ = _img()
    y = O.pool2d(x, (2, 2), strides, padding, reduce)
    e = np.array(expected, np.float64)
    assert np.array_equal(y.reshape(e.shape), e)


def test_pool2d_gradients_first_max_rule():
    x, _ = _img()
    g = np.ones((1, 5, 5, 1))
    exp_max = [[0, 0, 1, 0, 0], [0, 4, 0, 0, 4], [0, 0, 3, 1, 0], [2, 0, 0, 0, 1], [1, 0, 4, 0, 4]]
    assert np.array_equal(O.pool2d_grad(x, g, (2, 2), (1, 1), "same", "max").reshape(5, 5), exp_max)  # :277-288
    exp_avg = [[0.25, 0.5, 0.5, 0.5, 0.75], [0.5, 1, 1, 1, 1.5], [0.5, 1, 1, 1, 1.5], [0.5, 1, 1, 1, 1.5],
               [0.75, 1.5, 1.5, 1.5, 2.25]]
    assert np.array_equal(O.pool2d_grad(x, g, (2, 2), (1, 1), "same", "avg").reshape(5, 5), exp_avg)  # :294-305


def test_pool2d_layer_ignores_reduce():  # Pool2DLayerSpec.scala:12-30 + SURVEY App. B-Q7
    model = O.Pool2D(window=(2, 2))
    assert repr(model) == "Pool2D((2,2),(1,1),Valid,NHWC,Max)"
    x, _ = _img()
    y, _, _ = O.model_forward(model, x, {})
    assert np.array_equal(y.reshape(4, 4), [[3, 3, 2, 3], [3, 3, 3, 3], [2, 3, 3, 3], [2, 3, 3, 2]])
    y2, _, _ = O.model_forward(O.Pool2D(reduce="avg"), x, {})
    assert np.array_equal(y2, y)


# ---- models/layer/BatchNormSpec.scala:30-65 -----------------------------------------------------
BN_IN = np.array([[0, 0, 10], [0, 8, 4], [5, 0, 13], [20, 3, 8]], f32)


def test_batchnorm_generic_training_and_inference():
    model = O.Dense(4) >> O.BatchNorm(momentum=0.5)
    base = {"0/weights": W34, "1/weights": np.zeros(4, f32)}
    bn = {"2/beta": np.zeros((1, 4), f32), "2/gamma": np.ones((1, 4), f32), "2/moving_mean": np.zeros((1, 4), f32),
          "2/moving_variance": np.ones((1, 4), f32)}
    assert list(O.model_params(model, (4, 3)).keys()) == ["0/weights", "1/weights", "2/beta", "2/gamma",
                                                          "2/moving_mean", "2/moving_variance"]
    out, _, st = O.model_forward(model, BN_IN, {**base, **bn}, training=True)
    exp = [[-0.595, -1.168, -1.042, -1.231], [-1.181, 0.422, -0.232, 1.313], [0.307, -0.671, -0.375, -0.656],
           [1.469, 1.417, 1.649, 0.574]]
    assert np.array_equal(rnd(out, 3), rnd(exp, 3))
    assert np.array_equal(rnd(st["2/moving_mean"], 3), rnd([[7.638, 2.938, 5.375, 3.0]], 3))
    assert np.array_equal(rnd(st["2/moving_variance"], 3), rnd([[39.828, 13.148, 35.764, 3.47]], 3))
    frozen = O.Dense(4) >> O.BatchNorm(momentum=0.5, trainable=False)
    bn2 = dict(bn)
    bn2["2/moving_mean"] = np.full((1, 4), 5, f32)
    bn2["2/moving_variance"] = np.full((1, 4), 20, f32)
    out, _, st = O.model_forward(frozen, BN_IN, {**base, **bn2}, training=True)
    exp = [[1.118, -1.118, -0.671, -0.447], [-0.045, 0.671, 0.85, 0.939], [2.907, -0.559, 0.581, -0.134],
           [5.21, 1.789, 4.383, 0.537]]
    assert np.array_equal(rnd(out, 3), rnd(exp, 3)) and not st


# ---- models/layer/RNNLayerSpec.scala:17-79 ------------------------------------------------------
def test_simple_rnn_forward():
    layer = O.RNN(O.SimpleRNNCell(units=2, activation=O.Identity))
    x = np.array([1, 2, 3], f32).reshape(1, 3, 1)
    p = {"0/kernel_weights": np.array([[-0.5302748, -1.2049599]], f32),
         "0/recurrent_weights": np.array([[-0.77547526, -0.6313778], [-0.6313778, 0.7754754]], f32),
         "1/weights": np.zeros(2, f32)}
    y, _, _ = O.model_forward(layer, x, p)
    assert np.array_equal(rnd(y, 6), rnd([[0.222901, -6.019066]], 6))
    assert repr(O.RNN(O.SimpleRNNCell(units=2))) == "RNN(SimpleRNNCell(2) >> Bias(2) >> Tanh)"


LSTM_GOLD = {
    "forget": ([[0.17034012, -0.39808878]], [[0.4032409, -0.37492862], [-0.0409357, 0.22285524]], [1, 1]),
    "input": ([[0.5514972, -0.6295431]], [[-0.50807524, -0.07556351], [0.02593641, 0.44693834]], [0, 0]),
    "gate": ([[0.39550006, 0.03499722]], [[-0.11996187, 0.5238996], [0.34965965, -0.09114401]], [0, 0]),
    "output": ([[-0.02363521, 0.1830315]], [[-0.37248448, 0.07327155], [-0.70414686, -0.3490578]], [0, 0]),
}


def lstm_gold_params(dtype=f32):
    p = {}
    for g, (k, r, b) in LSTM_GOLD.items():
        p[f"{g}/0/kernel_weights"] = np.array(k, dtype)
        p[f"{g}/0/recurrent_weights"] = np.array(r, dtype)
        p[f"{g}/1/weights"] = np.array(b, dtype)
    return p


def test_lstm_forward_and_param_layout():
    layer = O.RNN(O.LSTMCell(units=2))
    assert repr(layer) == "RNN(LSTMCell(2,Tanh,Sigmoid,true))"
    x = np.array([1, 2, 3], f32).reshape(1, 3, 1)
    p = lstm_gold_params()
    assert set(O.model_params(layer, (1, 3, 1)).keys()) == set(p.keys())
    y, _, _ = O.model_forward(layer, x, p)
    assert np.array_equal(rnd(y, 6), rnd([[0.382158, 0.029766]], 6))


# ---- models/ActivationSpec.scala:11-46 ----------------------------------------------------------
def test_activations():
    assert abs(float(O.Tanh.fwd(np.array(0.5, f32))) - 0.46211717) < 6e-8
    x = np.array([[1.3, 5.1, 2.2, 0.7, 1.1], [2.1, 2.2, 0.1, 3.2, 1.1]], f32)
    exp = [[0.02019, 0.90254, 0.04966, 0.01108, 0.01653], [0.17817, 0.19691, 0.02411, 0.53526, 0.06555]]
    assert np.array_equal(rnd(O.Softmax.fwd(x), 5), rnd(exp, 5))
    relu = O.ReLU()
    assert [float(relu.fwd(np.array(v, f32))) for v in (5, 0, -5)] == [5, 0, 0]
    # strict gradient at the threshold (alg/AllKernels.scala:297-303)
    assert float(relu.bwd(np.array(1, f32), np.array(0, f32), np.array(0, f32))) == 0


# ---- models/InitializerSpec.scala:23-66 (TF Philox restated) ------------------------------------
def test_seeded_initializers():
    assert np.array_equal(rnd(O.RandomUniform(-1, 1, seed=1).build((3,)), 2), rnd([0.63, -0.84, 0.78], 2))
    assert np.array_equal(rnd(O.RandomNormal(std=1, seed=1).build((3,)), 2), rnd([0.31, 0.57, 0.42], 2))
    assert np.array_equal(rnd(O.RandomNormalTruncated(std=1, seed=1).build((3,)), 2), rnd([0.31, 0.57, 0.42], 2))
    assert np.array_equal(rnd(O.GlorotUniform(seed=1).build((3,)), 2), rnd([0.63, -0.84, 0.78], 2))
    assert np.array_equal(rnd(O.GlorotNormal(seed=2).build((3,)), 2), rnd([-0.25, -0.31, -0.34], 2))
    ort = O.Orthogonal(seed=1).build((2, 3))
    # Householder QR in fp32 (Eigen in TF vs LAPACK here) differs in the last bits: 1.5e-6 absolute
    np.testing.assert_allclose(ort, [[0.363699, 0.498735, -0.786757], [-0.189386, -0.787368, -0.586672]], atol=1.5e-6)
    assert np.array_equal(O.Zeros.build((3,)), [0, 0, 0]) and np.array_equal(O.Ones.build((3,)), [1, 1, 1])


def test_fans_oddity():  # Initializer.scala:16-25
    assert O.fans((3, 3, 32, 64)) == (2048 * 3, 2048 * 3)
    assert O.fans((784, 50)) == (784, 50) and O.fans((10,)) == (10, 10) and O.fans(()) == (1, 1)


# ---- models/RegularizationSpec.scala:11-27 -------------------------------------------------------
def test_regularizers():
    w = np.array([-1, 2, 3], f32)
    assert float(O.Zero.penalty(w)) == 0
    assert abs(float(O.L1(0.1).penalty(w)) - 0.3) < 1e-7
    assert abs(float(O.L2(0.1).penalty(w)) - 0.7) < 1e-7


# ---- math/alg/KernelsSpec.scala:445-461, 556-585 -------------------------------------------------
def test_alg_helpers_and_matmul_grads():
    assert float(O.decaying_avg(np.float64(10), np.float64(5), np.float64(0.9))) == 9.5
    assert float(O.sqrt_zero_safe(f32(8), f32(1))) == 3.0
    assert float(O.boost(np.array(1.5, f32), f32(0.5), 2)) == 2.0
    assert float(O.boost(np.array(1.5, f32), f32(0.5), 1000000)) == 1.5
    a = np.array([[1, 2, 3], [4, 5, 6]], f32)
    x = np.array([[5, 10], [15, 20], [25, 30]], f32)
    d = O.DenseKernel(2)
    _, g = d.backward(np.ones((2, 2), f32), a, {"weights": x}, need_dx=False)
    assert np.array_equal(g["weights"], [[5, 5], [7, 7], [9, 9]])
    dx, _ = d.backward(np.ones((2, 2), f32), a, {"weights": x}, need_dx=True)
    assert np.array_equal(dx, [[15, 35, 55], [15, 35, 55]])


# ---- optimizers/IteratorsSpec.scala:11-74 --------------------------------------------------------
def test_tensor_iterator():
    x = np.array([[1, 2], [4, 5]], f32)
    y = np.array([[3], [6]], f32)
    it = list(O.tensor_iterator(x, y, 2))
    assert len(it) == 1 and np.array_equal(it[0][0], x) and np.array_equal(it[0][1], y)
    x4 = np.array([[1, 2], [4, 5], [7, 8], [10, 11]], f32)
    y4 = np.array([[3], [6], [9], [12]], f32)
    it = list(O.tensor_iterator(x4, y4, 2))
    assert len(it) == 2 and np.array_equal(it[1][0], [[7, 8], [10, 11]]) and np.array_equal(it[1][1], [[9], [12]])
    it = list(O.tensor_iterator(x, y, 3, "pad_zeros"))
    assert np.array_equal(it[0][0], [[1, 2], [4, 5], [0, 0]]) and np.array_equal(it[0][1], [[3], [6], [0]])
    it = list(O.tensor_iterator(x, y, 3, "partial"))
    assert np.array_equal(it[0][0], x)
    assert list(O.tensor_iterator(x, y, 3, "skip")) == []


# ---- optimizers/*Spec.scala convergence thresholds on the reference's deterministic fixtures ----
def _linear(fixtures_dir):
    d = np.loadtxt(os.path.join(fixtures_dir, "linear_function_1.scv"), delimiter=",", dtype=np.float32)
    assert d.shape == (97, 2)
    return d[:, :1].copy(), d[:, 1:].copy()


@pytest.mark.parametrize("alg,epochs,threshold", [
    (O.SGD(), 100, 11.0), (O.SGD(momentum=0.1), 100, 11.0), (O.SGD(momentum=0.1, nesterov=True), 100, 11.0),
    (O.Adam(rate=0.1), 100, 12.0), (O.AdaGrad(rate=1.0), 100, 9.0), (O.RMSProp(rate=0.06), 100, 9.4),
    (O.AdaDelta(), 100, 50.0), (O.Adamax(), 70, 9.0), (O.Nadam(), 50, 9.0), (O.AMSGrad(rate=0.1), 100, 9.0),
])
def test_optimizer_convergence_thresholds(fixtures_dir, alg, epochs, threshold):
    x, y = _linear(fixtures_dir)
    model = O.Dense(1, O.Identity, kernel_initializer=O.Zeros)  # LinearRegression() (models/LinearRegression.scala:20-31)
    mp, _, hist = O.run_training(model, O.MeanSquaredError, alg, x, y, batch=97, epochs=epochs)
    loss = float(O.eval_loss(O.LossModel(model, O.MeanSquaredError), x, y, mp))
    assert loss <= threshold, (alg, loss)
    assert hist[-1][1] == epochs  # one iteration per epoch, global iter advances by the sum


def test_adam_logistic_regression(fixtures_dir):  # AdamSpec.scala:44-64
    d = np.loadtxt(os.path.join(fixtures_dir, "logistic_regression_1.scv"), delimiter=",", dtype=np.float64)
    x = (d[:, :2].astype(f32) / f32(100)).astype(f32)
    y = d[:, 2:].astype(f32)
    model = O.Dense(1, O.Sigmoid, kernel_initializer=O.Zeros)
    mp, _, _ = O.run_training(model, O.BinaryCrossentropy, O.Adam(0.1), x, y, batch=100, epochs=100)
    lm = O.LossModel(model, O.BinaryCrossentropy)
    assert float(O.eval_loss(lm, x, y, mp)) <= 0.4
    pred, _, _ = O.model_forward(model, x, mp)
    assert np.mean((np.round(pred) == y).astype(np.float64)) >= 0.9
    probe, _, _ = O.model_forward(model, np.array([[0.3462, 0.7802], [0.6018, 0.8630]], f32), mp)
    assert np.array_equal(np.round(probe), [[0], [1]])


@pytest.mark.parametrize("seed", [1, 2])
def test_mlp_minimises_logistic_regression(fixtures_dir, seed):  # models/ANNSpec.scala:21-39 (unseeded GlorotUniform there)
    d = np.loadtxt(os.path.join(fixtures_dir, "logistic_regression_1.scv"), delimiter=",", dtype=np.float64)
    x = (d[:, :2].astype(f32) / f32(100)).astype(f32)
    y = d[:, 2:].astype(f32)
    model = (O.Dense(4, O.Sigmoid, kernel_initializer=O.GlorotUniform(seed))
             >> O.Dense(1, O.Sigmoid, kernel_initializer=O.GlorotUniform(seed + 10)))
    mp, _, _ = O.run_training(model, O.BinaryCrossentropy, O.Adam(0.2), x, y, batch=100, epochs=55)
    assert float(O.eval_loss(O.LossModel(model, O.BinaryCrossentropy), x, y, mp)) <= 0.5
    pred, _, _ = O.model_forward(model, x, mp)
    assert np.mean((np.round(pred) == y).astype(np.float64)) >= 0.89
    probe, _, _ = O.model_forward(model, np.array([[0.3462, 0.7802], [0.6018, 0.8630]], f32), mp)
    assert np.array_equal(np.round(probe), [[0], [1]])


def test_sgd_minimises_x_squared():  # SGDSpec.scala:16-31 with models/Math.scala:11-25
    w = np.array(5.0, f32)
    alg = O.SGD(rate=0.1)
    st = {"velocity": np.zeros((), f32)}
    for t in range(1, 101):
        delta, st = alg.build(f32(2) * w, st, t)
        w = w - delta
    assert float(w * w) <= 0.1


# ---- the timed CPU baseline of bench.py is a torch-CPU port of the cfg2 step: pin it against the NumPy oracle ----
def test_torch_port_matches_numpy_oracle():
    torch = pytest.importorskip("torch")
    from oracle.torch_port import TorchLeNetPort
    om = (O.Conv2D(32, activation=O.ReLU(), kernel_initializer=O.GlorotUniform(1)) >> O.Pool2D()
          >> O.Conv2D(64, activation=O.ReLU(), kernel_initializer=O.GlorotUniform(1)) >> O.Pool2D() >> O.Flatten
          >> O.Dense(10, O.Softmax, kernel_initializer=O.GlorotUniform(1)))
    alg = O.Adam(0.001)
    batch = 20
    _, mp, op = O.init_params(om, alg, (batch, 28, 28, 1))
    lm = O.LossModel(om, O.CategoricalCrossentropy)
    rng = np.random.Generator(np.random.PCG64(5))
    x = rng.uniform(1 / 256, 1.0, size=(3 * batch, 28, 28, 1)).astype(f32)
    y = np.zeros((3 * batch, 10), f32)
    y[np.arange(3 * batch), np.arange(3 * batch) % 10] = 1
    port = TorchLeNetPort(mp, rate=0.001)
    for t in range(1, 4):
        xb, yb = x[(t - 1) * batch:t * batch], y[(t - 1) * batch:t * batch]
        mp, op, ol = O.train_step(lm, alg, xb, yb, mp, op, t)
        tl = port.train_step(xb, yb, t)
        assert abs(tl - float(ol)) <= 2e-5 * abs(float(ol)), (t, tl, float(ol))
    got = port.params_numpy()
    for k, v in mp.items():  # a sign flip of a near-zero gradient moves a weight by one Adam step (1e-3)
        assert float(np.max(np.abs(got[k] - v))) <= 3e-3 and float(np.mean(np.abs(got[k] - v))) <= 5e-5, k


def test_estimators_hand_computed():  # estimators/package.scala:18-137 on y = 2x with one label off by one
    om = O.Dense(1, O.Identity, kernel_initializer=O.Zeros)
    defs = O.model_params(om, (2, 1))
    params = {k: (np.full(d.shape, 2.0, np.float32) if len(d.shape) == 2 else np.zeros(d.shape, np.float32))
              for k, d in defs.items()}
    x = np.array([[1], [2], [3]], np.float32)
    y = np.array([[2], [4], [7]], np.float32)  # predictions 2, 4, 6
    E = O.estimators
    assert E.accuracy(om, params, x, y, batch=2) == pytest.approx(2 / 3, rel=1e-6)   # tail batch is partial, not dropped
    assert E.mse(om, params, x, y, batch=2) == pytest.approx(1 / 3, rel=1e-6)
    assert E.mae(om, params, x, y, batch=2) == pytest.approx(1 / 3, rel=1e-6)
    assert E.rmse(om, params, x, y, batch=2) == pytest.approx(np.sqrt(1 / 3), rel=1e-6)
    # SST is centred on the mean PREDICTION (4): sst = 8, ssr = 1
    assert E.r2_score(om, params, x, y, batch=2) == pytest.approx(0.875, rel=1e-6)
    with pytest.raises(ValueError):
        E.r2_score(om, params, x, np.concatenate([y, y], axis=1))
