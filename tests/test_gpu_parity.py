"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs and against
the reference's golden vectors.  Tolerances are stated per test.  Run on the B200 box with `-m gpu`.

fp32 (SB_PREC_FP32) tolerance rationale: every contraction accumulates in fp32 with FMA in a different association
order than NumPy/BLAS (and than TF-CPU's Eigen), transcendental functions differ at the 1-2 ulp level; so values are
compared with rtol 2e-5 / atol 2e-6 after one op chain and rtol 1e-3 / atol 2e-5 after tens of optimizer steps."""
import os

import numpy as np
import pytest

import oracle as O
import scanet3_b200 as S
from helpers import OALG, OLOSS, SALG, SLOSS, assert_params_close, build_models, synth_classification

pytestmark = pytest.mark.gpu
f32 = np.float32


def oracle_init(omodel, alg, in_shape_batched, seed_params=None):
    defs, mp, op = O.init_params(omodel, alg, in_shape_batched, np.float32, seed_params)
    return defs, mp, op


def make_engine(smodel, loss, alg, batch, in_shape, lab_shape, **kw):
    return S.Engine(smodel, loss, alg, batch, in_shape, lab_shape, device=0, **kw)


# ---------------------------------------------------------------------------------------------------------------
# initialisers: device Philox is bit-compatible with the restated TF Philox (InitializerSpec.scala:23-42)
# ---------------------------------------------------------------------------------------------------------------
def test_device_philox_matches_tf_goldens_and_oracle():
    e = make_engine(S.Dense(3, kernelInitializer=S.RandomUniform(-1, 1, seed=1), bias=False), S.MeanSquaredError, S.SGD(), 1,
                    (1,), (3,))
    e.init_params()
    w = e.get_param("weights").reshape(-1)
    assert np.array_equal(np.round(w * f32(100)) / f32(100), np.array([0.63, -0.84, 0.78], f32))
    assert np.array_equal(w, O.RandomUniform(-1, 1, seed=1).build((3,)))
    e.close()
    om, sm = build_models([("dense", 50, "sigmoid"), ("dense", 10, "softmax")], seed=7)
    e = make_engine(sm, S.CategoricalCrossentropy, S.Adamax(), 4, (784,), (10,))
    e.init_params()
    _, mp, op = oracle_init(om, O.Adamax(), (4, 784))
    got = e.get_model_params()
    for k, v in mp.items():
        assert np.array_equal(got[k], v), k  # GlorotUniform: bit exact
    gop = e.get_opt_params()
    for k, v in op.items():
        assert np.array_equal(gop[k], v), k  # Const(initAcc)
    e.close()
    e = make_engine(S.Dense(64, kernelInitializer=S.RandomNormal(0.0, 1.0, seed=5), bias=False), S.MeanSquaredError, S.SGD(),
                    1, (33,), (64,))
    e.init_params()
    np.testing.assert_allclose(e.get_param("weights"), O.RandomNormal(0.0, 1.0, seed=5).build((33, 64)), rtol=0, atol=2e-6)
    e.close()


# ---------------------------------------------------------------------------------------------------------------
# the reference's golden vectors, evaluated through the engine
# ---------------------------------------------------------------------------------------------------------------
X4 = np.array([[0, 0, 1], [0, 1, 1], [1, 0, 1], [1, 1, 1]], f32)
W34 = np.array([[1, 0.5, 1, 0.1], [0.1, 1, 1, 1], [1, 0, 0.2, 0.3]], f32)


def test_dense_layer_goldens():  # models/layer/DenseLayerSpec.scala:17-41, 55-78
    e = make_engine(S.Dense(4, S.Sigmoid), S.BinaryCrossentropy, S.SGD(), 4, (3,), (4,))
    e.set_params({"0/weights": W34, "1/weights": np.zeros(4, f32)})
    exp = np.array([[0.731059, 0.500000, 0.549834, 0.574442], [0.750260, 0.731059, 0.768525, 0.785835],
                    [0.880797, 0.622459, 0.768525, 0.598688], [0.890903, 0.817574, 0.900250, 0.802184]], f32)
    np.testing.assert_allclose(e.predict(X4), exp, atol=7e-7)
    y = np.array([[0.7310586, 0.5, 0.549834, 0.5744425], [0.7502601, 0.7310586, 0.76852477, 0.7858349],
                  [0.8807971, 0.6224593, 0.76852477, 0.59868765], [0.89090323, 0.81757444, 0.9002496, 0.80218387]], f32)
    e.set_params({"0/weights": np.zeros((3, 4), f32), "1/weights": np.zeros(4, f32)})
    _, g = e.compute_grads(X4, y)
    wg = np.array([[-0.048231270, -0.027502108, -0.041798398, -0.025054470],
                   [-0.040072710, -0.034289565, -0.041798398, -0.036751173],
                   [-0.078313690, -0.041943270, -0.061695820, -0.047571808]], f32)
    np.testing.assert_allclose(g["0/weights"], wg, rtol=3e-6, atol=1e-8)
    np.testing.assert_allclose(g["1/weights"], wg[2], rtol=3e-6, atol=1e-8)
    e.close()


def test_regression_goldens():  # models/RegressionSpec.scala:13-61, 67-105
    e = make_engine(S.LinearRegression(), S.MeanSquaredError, S.SGD(), 2, (2,), (1,))
    x = np.array([[1, 2], [2, 4]], f32)
    y = np.array([[6], [12]], f32)
    e.set_params({"0/weights": np.array([[2], [3]], f32), "1/weights": np.array([1], f32)})
    assert e.eval_loss(x, y) == 17.0
    assert np.array_equal(e.predict(x), np.array([[9], [17]], f32))
    e.set_params({"0/weights": np.zeros((2, 1), f32), "1/weights": np.zeros(1, f32)})
    _, g = e.compute_grads(x, y)
    assert np.array_equal(g["0/weights"], np.array([[-30], [-60]], f32)) and np.array_equal(g["1/weights"], np.array([-18], f32))
    e.close()
    e = make_engine(S.LogisticRegression(), S.BinaryCrossentropy, S.SGD(), 2, (2,), (1,))
    x = np.array([[0.34, 0.78], [0.6, 0.86]], f32)
    y = np.array([[0.402], [0.47800002]], f32)
    e.set_params({"0/weights": np.array([[0.2], [0.3]], f32), "1/weights": np.array([0.1], f32)})
    assert abs(e.eval_loss(x, y) - 0.7422824) < 1e-6
    np.testing.assert_allclose(e.predict(x), [[0.599168], [0.617276]], atol=7e-7)
    _, g = e.compute_grads(x, y)
    np.testing.assert_allclose(g["0/weights"], [[0.075301], [0.136784]], atol=7e-7)
    np.testing.assert_allclose(g["1/weights"], [0.168222], atol=7e-7)
    e.close()


def test_composed_mse_golden():  # models/layer/ComposedLayerSpec.scala:17-48, 71-95
    e = make_engine(S.Dense(3, S.Sigmoid) >> S.Dense(1, S.Sigmoid), S.MeanSquaredError, S.SGD(), 2, (3,), (1,))
    assert e.model_param_paths() == ["0/weights", "1/weights", "3/weights", "4/weights"]
    e.set_params({"0/weights": np.array([[1, 0.5, 1], [0.1, 1, 1], [1, 0, 0.2]], f32), "1/weights": np.zeros(3, f32),
                  "3/weights": np.array([[0.1], [0.5], [1]], f32), "4/weights": np.zeros(1, f32)})
    assert abs(e.eval_loss(np.array([[0, 0, 1], [0, 1, 1]], f32), np.array([[1], [0]], f32)) - 0.339962) < 1e-6
    e.close()


IMG = np.array([[2, 1, 2, 0, 1], [1, 3, 2, 2, 3], [1, 1, 3, 3, 0], [2, 2, 0, 1, 1], [0, 0, 3, 1, 2]], f32).reshape(1, 5, 5, 1)
FIL = np.array([[2, 3], [0, 1]], f32).reshape(2, 2, 1, 1)


@pytest.mark.parametrize("padding,strides,expected", [
    ("Same", (1, 1), [[10, 10, 6, 6, 2], [12, 15, 13, 13, 6], [7, 11, 16, 7, 0], [10, 7, 4, 7, 2], [0, 9, 9, 8, 4]]),
    ("Same", (2, 2), [[10, 6, 2], [7, 16, 0], [0, 9, 4]]),
    ("Valid", (1, 1), [[10, 10, 6, 6], [12, 15, 13, 13], [7, 11, 16, 7], [10, 7, 4, 7]]),
    ("Valid", (2, 2), [[10, 6], [7, 16]]),
    (S.Explicit(height=(1, 0), width=(1, 0)), (2, 2), [[2, 2, 1], [4, 15, 13], [6, 7, 7]]),  # KernelsSpec.scala:105-119
])
def test_conv2d_goldens(padding, strides, expected):  # math/nn/KernelsSpec.scala:45-119; Conv2DLayerSpec.scala:13-36
    exp = np.array(expected, f32)
    e = make_engine(S.Conv2D(1, (2, 2), strides, padding), S.MeanSquaredError, S.SGD(), 1, (5, 5, 1), exp.shape + (1,))
    e.set_param("weights", FIL)
    assert np.array_equal(e.predict(IMG).reshape(exp.shape), exp)
    e.close()


def test_conv2d_gradient_goldens():  # math/nn/KernelsSpec.scala:146-157 (wgrad), 174-188 (dgrad via a second conv in front)
    # d(sum(conv))/dW through MSE: with y = conv - n/2 the MSE gradient w.r.t. the output is exactly 1 everywhere
    e = make_engine(S.Conv2D(1, (2, 2)), S.MeanSquaredError, S.SGD(), 1, (5, 5, 1), (4, 4, 1))
    e.set_param("weights", FIL)
    out = e.predict(IMG)
    _, g = e.compute_grads(IMG, out - f32(8.0))  # dL/dout = 2*(8)/16 = 1
    assert np.array_equal(g["weights"].reshape(2, 2), np.array([[26, 25], [25, 27]], f32))
    e.close()
    # dgrad: put a 1x1 identity conv in front so the 2x2 conv's input gradient becomes the first layer's weight gradient
    # contracted with the image; check it against the oracle instead of a literal
    om, sm = build_models([("conv", 1, "identity", (1, 1), (1, 1), "Valid", False), ("conv", 1, "identity", (2, 2), (1, 1), "Valid", False)])
    e = make_engine(sm, S.MeanSquaredError, S.SGD(), 1, (5, 5, 1), (4, 4, 1))
    p = {"0/weights": np.ones((1, 1, 1, 1), f32), "1/weights": FIL}
    e.set_params(p)
    y = np.zeros((1, 4, 4, 1), f32)
    _, g = e.compute_grads(IMG, y)
    _, og, _ = O.LossModel(om, O.MeanSquaredError).grad_stateful(IMG, y, p)
    np.testing.assert_allclose(g["0/weights"], og["0/weights"], rtol=2e-6)
    np.testing.assert_allclose(g["1/weights"], og["1/weights"], rtol=2e-6)
    e.close()


@pytest.mark.parametrize("padding,strides,expected", [
    ("Same", (1, 1), [[3, 3, 2, 3, 3], [3, 3, 3, 3, 3], [2, 3, 3, 3, 1], [2, 3, 3, 2, 2], [0, 3, 3, 2, 2]]),
    ("Same", (2, 2), [[3, 2, 3], [2, 3, 1], [0, 3, 2]]),
    ("Valid", (1, 1), [[3, 3, 2, 3], [3, 3, 3, 3], [2, 3, 3, 3], [2, 3, 3, 2]]),
    ("Valid", (2, 2), [[3, 2], [2, 3]]),
])
def test_maxpool_goldens(padding, strides, expected):  # math/nn/KernelsSpec.scala:200-256; Pool2DLayerSpec.scala:12-30
    exp = np.array(expected, f32)
    # a 1x1 conv with weight 1 in front gives the pool a trainable predecessor (the engine needs one parameter)
    e = make_engine(S.Conv2D(1, (1, 1)) >> S.Pool2D((2, 2), strides, padding), S.MeanSquaredError, S.SGD(), 1, (5, 5, 1),
                    exp.shape + (1,))
    e.set_param("0/weights", np.ones((1, 1, 1, 1), f32))
    assert np.array_equal(e.predict(IMG).reshape(exp.shape), exp)
    e.close()


def test_maxpool_gradient_first_max_rule():  # math/nn/KernelsSpec.scala:277-288
    # dL/d(conv weight) = sum(image * poolgrad): compare the routed gradient through a 1x1 conv with the golden map
    e = make_engine(S.Conv2D(1, (1, 1)) >> S.Pool2D((2, 2), (1, 1), "Same"), S.MeanSquaredError, S.SGD(), 1, (5, 5, 1), (5, 5, 1))
    e.set_param("0/weights", np.ones((1, 1, 1, 1), f32))
    out = e.predict(IMG)
    _, g = e.compute_grads(IMG, out - f32(12.5))  # dL/dpool = 2*12.5/25 = 1
    golden = np.array([[0, 0, 1, 0, 0], [0, 4, 0, 0, 4], [0, 0, 3, 1, 0], [2, 0, 0, 0, 1], [1, 0, 4, 0, 4]], f32)
    assert float(g["0/weights"].reshape(())) == float(np.sum(golden * IMG.reshape(5, 5)))
    e.close()


def test_batchnorm_generic_golden():  # models/layer/BatchNormSpec.scala:30-49
    x = np.array([[0, 0, 10], [0, 8, 4], [5, 0, 13], [20, 3, 8]], f32)
    e = make_engine(S.Dense(4) >> S.BatchNorm(momentum=0.5), S.MeanSquaredError, S.SGD(rate=0.0), 4, (3,), (4,))
    assert e.model_param_paths() == ["0/weights", "1/weights", "2/beta", "2/gamma", "2/moving_mean", "2/moving_variance"]
    e.set_params({"0/weights": W34, "1/weights": np.zeros(4, f32), "2/beta": np.zeros((1, 4), f32), "2/gamma": np.ones((1, 4), f32),
                  "2/moving_mean": np.zeros((1, 4), f32), "2/moving_variance": np.ones((1, 4), f32)})
    e.init_opt_params()
    e.train_step(x, np.zeros((4, 4), f32), 1)  # rate 0: only the moving statistics change
    np.testing.assert_allclose(e.get_param("2/moving_mean"), [[7.638, 2.938, 5.375, 3.0]], atol=6e-4)
    np.testing.assert_allclose(e.get_param("2/moving_variance"), [[39.828, 13.148, 35.764, 3.47]], atol=6e-4)
    e.close()


# ---------------------------------------------------------------------------------------------------------------
# trajectories: per-step loss and parameters after n steps vs the oracle
# ---------------------------------------------------------------------------------------------------------------
def run_trajectory(spec, loss, alg_name, alg_kw, batch, in_shape, classes, steps, seed=3, rtol=1e-3, atol=2e-5, loss_rtol=2e-5,
                   precision=S.PREC_FP32, use_graph=True):
    om, sm = build_models(spec)
    oalg, salg = OALG[alg_name](**alg_kw), SALG[alg_name](**alg_kw)
    x, y = synth_classification(batch * steps, in_shape, classes, seed)
    e = make_engine(sm, SLOSS[loss], salg, batch, in_shape, (classes,), precision=precision, use_graph=use_graph)
    e.init_params()
    _, mp, op = oracle_init(om, oalg, (batch,) + tuple(in_shape))
    assert_params_close(e.get_model_params(), mp, 0, 0, "init")
    lm = O.LossModel(om, OLOSS[loss])
    for t in range(1, steps + 1):
        xb, yb = x[(t - 1) * batch:t * batch], y[(t - 1) * batch:t * batch]
        gl = e.train_step(xb, yb, t)
        mp, op, ol = O.train_step(lm, oalg, xb, yb, mp, op, t)
        assert abs(gl - float(ol)) <= loss_rtol * abs(float(ol)) + 1e-6, (t, gl, float(ol))
    assert_params_close(e.get_model_params(), mp, rtol, atol, f"after {steps} steps")
    assert_params_close(e.get_opt_params(), op, rtol * 5, atol, f"opt state after {steps} steps")
    e.close()


MLP = [("dense", 50, "sigmoid"), ("dense", 10, "softmax")]


@pytest.mark.parametrize("alg,kw", [
    ("sgd", dict(rate=0.05)), ("sgd", dict(rate=0.05, momentum=0.1, nesterov=True)), ("adam", dict(rate=0.01)),
    ("adagrad", dict(rate=0.05)), ("rmsprop", dict(rate=0.005)), ("adadelta", dict()), ("adamax", dict(rate=0.01)),
    ("nadam", dict()), ("amsgrad", dict(rate=0.01)),
])
def test_mlp_trajectory_all_optimizers(alg, kw):  # BASELINE config 1 shape (784 -> 50 -> 10), smaller batch
    run_trajectory(MLP, "cce", alg, kw, batch=200, in_shape=(784,), classes=10, steps=12)


def test_mlp_full_batch_config1():  # BASELINE config 1 exactly: batch 1000
    run_trajectory(MLP, "cce", "adam", dict(rate=0.01), batch=1000, in_shape=(784,), classes=10, steps=5)


LENET = [("conv", 32, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"),
         ("conv", 64, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"), ("flatten",),
         ("dense", 10, "softmax")]


def test_lenet_small_trajectory():
    run_trajectory(LENET, "cce", "adam", dict(rate=0.001), batch=8, in_shape=(12, 12, 1), classes=10, steps=8, rtol=2e-3,
                   atol=3e-5, loss_rtol=5e-5)


def test_lenet_config2_full_size():  # BASELINE config 2: 28x28x1, batch 100 (fp32 contractions)
    run_trajectory(LENET, "cce", "adam", dict(rate=0.001), batch=100, in_shape=(28, 28, 1), classes=10, steps=3, rtol=2e-3,
                   atol=3e-5, loss_rtol=5e-5)


def test_strided_same_conv_with_bias_and_sigmoid_trajectory():
    spec = [("conv", 8, "sigmoid", (3, 3), (2, 2), "Same", True), ("pool", (2, 2), (2, 2), "Same"),
            ("conv", 16, "tanh", (2, 2), (1, 1), "Same", True), ("flatten",), ("dense", 4, "sigmoid")]
    run_trajectory(spec, "mse", "sgd", dict(rate=0.1, momentum=0.5), batch=6, in_shape=(9, 11, 3), classes=4, steps=6,
                   rtol=2e-3, atol=3e-5, loss_rtol=5e-5)


@pytest.mark.parametrize("mode", ["reference", "correct"])
def test_batchnorm_cnn_trajectory(mode):  # BASELINE config 5 pattern (fused rank-4 BN), small
    spec = [("conv", 8, "relu", (3, 3), (1, 1), "Valid", False), ("bn", 0.99, mode), ("pool", (2, 2), (2, 2), "Valid"),
            ("flatten",), ("dense", 16, "identity"), ("bn", 0.9, mode), ("dense", 10, "softmax")]
    run_trajectory(spec, "cce", "adam", dict(rate=0.001), batch=16, in_shape=(10, 10, 3), classes=10, steps=6, rtol=3e-3,
                   atol=5e-5, loss_rtol=1e-4)


BN_POOL = [("conv", 16, "relu", (3, 3), (1, 1), "Valid", False), ("bn", 0.99, "correct"), ("pool", (2, 2), (2, 2), "Valid"),
           ("conv", 32, "relu", (3, 3), (1, 1), "Same", True), ("bn", 0.9, "correct"), ("pool", (2, 2), (2, 2), "Valid"),
           ("flatten",), ("dense", 10, "softmax")]


@pytest.mark.parametrize("kernel,cin,cout", [((5, 5), 1, 32), ((3, 3), 3, 128), ((3, 3), 1, 64), ((3, 3), 3, 24)])
def test_first_layer_filter_gradient_kernels(kernel, cin, cout):
    """Conv2D with KH*KW*Cin <= 32 as the FIRST layer and no 2x2/1 pool behind it: the filter gradient runs in
    conv_smallk_wgrad_tile_kernel (Cout >= 32: pixel tiles + im2col rows in shared memory) or the register-accumulator kernel
    (Cout = 24), the forward in the persistent small-K kernel -- trajectories against the oracle in fp32."""
    spec = [("conv", cout, "tanh", kernel, (1, 1), "Valid", False), ("flatten",), ("dense", 6, "softmax")]
    run_trajectory(spec, "cce", "adam", dict(rate=0.003), batch=9, in_shape=(11, 13, cin), classes=6, steps=5, rtol=2e-3, atol=3e-5,
                   loss_rtol=5e-5)


def test_batchnorm_pool_block_falls_back_when_the_layout_does_not_fit():
    """24 channels: 256 is not a multiple of C/4 = 6, so bn_pool.cu's kernels do not apply and the two-pass BatchNorm, the generic
    pool and the ReLU backward run as separate kernels -- same trajectory against the oracle."""
    spec = [("conv", 24, "relu", (3, 3), (1, 1), "Valid", False), ("bn", 0.9, "correct"), ("pool", (2, 2), (2, 2), "Valid"),
            ("flatten",), ("dense", 10, "softmax")]
    run_trajectory(spec, "cce", "adam", dict(rate=0.001), batch=10, in_shape=(9, 9, 3), classes=10, steps=5, rtol=3e-3, atol=5e-5,
                   loss_rtol=1e-4)


@pytest.mark.parametrize("mode", ["reference", "correct"])
def test_batchnorm_pool_block_with_odd_extents(mode):
    """Conv2D(ReLU) >> BatchNorm >> Pool2D(2x2, stride 2) runs as one forward and one backward (csrc/bn_pool.cu).  13x11 images:
    the first block pools 11x9 -> 5x4 and the second 5x4 -> 2x2, so both have a last row / column outside every window (their dy is
    zero, their BatchNorm dx is not)."""
    spec = [(it[0], it[1], mode) if it[0] == "bn" else it for it in BN_POOL]
    run_trajectory(spec, "cce", "adam", dict(rate=0.001), batch=12, in_shape=(13, 11, 3), classes=10, steps=5, rtol=3e-3, atol=5e-5,
                   loss_rtol=1e-4)


def test_batchnorm_pool_block_matches_the_separate_kernels(monkeypatch):
    """The same model with SB_NO_BN_POOL=1 SB_BN_TWO_PASS=1 (two-pass moments, BatchNorm / Pool2D / ReLU backward as separate
    kernels): loss and gradients agree to fp32 rounding of the sums (the one-pass moments merge partial (n, mean, M2) triples, so the
    bits differ), and so does predict(), which normalises with the moving statistics."""
    _, sm = build_models(BN_POOL)
    x, y = synth_classification(24, (13, 11, 3), 10, 5)
    outs = []
    for off in ("1", "0"):
        monkeypatch.setenv("SB_NO_BN_POOL", off)
        monkeypatch.setenv("SB_BN_TWO_PASS", off)
        e = make_engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), 24, (13, 11, 3), (10,))
        e.init_params()
        loss, grads = e.compute_grads(x, y)
        l1 = e.train_step(x, y, 1)
        outs.append((loss, l1, grads, e.get_params(), e.predict(x), e.eval_loss(x, y)))
        e.close()
    a, b = outs
    assert abs(a[0] - b[0]) <= 2e-6 * abs(a[0]) and abs(a[1] - b[1]) <= 2e-6 * abs(a[1]) and abs(a[5] - b[5]) <= 2e-6 * abs(a[5])
    for k in a[2]:
        scale = float(np.max(np.abs(a[2][k]))) + 1e-12
        assert float(np.max(np.abs(a[2][k] - b[2][k]))) <= 2e-4 * scale, k
    for k in a[3]:
        if "moving" in k:
            np.testing.assert_allclose(b[3][k], a[3][k], rtol=1e-5, atol=1e-7, err_msg=k)
    np.testing.assert_allclose(b[4], a[4], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("alg", ["adam", "amsgrad"])
def test_early_optimizer_on_the_side_branch_is_bit_identical(alg, monkeypatch):
    """Default; SB_EARLY_OPT=0 turns it off: conv2's side branch (filter gradient + reduction) also runs the optimizer update of every parameter from conv2
    on, once conv2's dgrad has read the weights; the tail of the step only updates conv1.  Same arithmetic on the same gradients:
    losses, parameters and optimizer state must be the same bits (TF32 mode: the tensor-core conv path is the one that forks)."""
    _, sm = build_models(LENET)
    x, y = synth_classification(24 * 4, (14, 14, 1), 10, 9)
    outs = []
    for on in ("0", "1"):
        monkeypatch.setenv("SB_EARLY_OPT", on)
        e = make_engine(sm, S.CategoricalCrossentropy, SALG[alg](rate=0.01), 24, (14, 14, 1), (10,), precision=S.PREC_TF32)
        e.init_params()
        losses = [e.train_step(x[i * 24:(i + 1) * 24], y[i * 24:(i + 1) * 24], i + 1) for i in range(4)]
        outs.append((losses, {k: v.tobytes() for k, v in e.get_params().items()}))
        e.close()
    assert outs[0][0] == outs[1][0]
    assert outs[0][1] == outs[1][1]


def test_graph_and_eager_are_bit_identical():
    om, sm = build_models(MLP)
    x, y = synth_classification(400, (784,), 10, 11)
    outs = []
    for use_graph in (True, False):
        e = make_engine(sm, S.CategoricalCrossentropy, S.Adam(0.01), 100, (784,), (10,), use_graph=use_graph)
        e.init_params()
        losses = [e.train_step(x[i * 100:(i + 1) * 100], y[i * 100:(i + 1) * 100], i + 1) for i in range(4)]
        outs.append((losses, e.get_params()))
        e.close()
    assert outs[0][0] == outs[1][0]
    for k in outs[0][1]:
        assert np.array_equal(outs[0][1][k], outs[1][1][k]), k


# ---------------------------------------------------------------------------------------------------------------
# the per-worker epoch (optimizeOnPartition, Optimizer.scala:106-161) with the shard resident in HBM
# ---------------------------------------------------------------------------------------------------------------
def test_resident_epoch_matches_optimize_on_partition():
    om, sm = build_models(MLP)
    batch, rows = 64, 64 * 5 + 17  # partial tail batch of 17 rows is dropped (Iterators.scala:46-47,90)
    x, y = synth_classification(rows, (784,), 10, 5)
    oalg = O.Adam(rate=0.01)
    e = make_engine(sm, S.CategoricalCrossentropy, S.Adam(0.01), batch, (784,), (10,))
    e.init_params()
    e.upload_dataset(x, y)
    _, mp, op = oracle_init(om, oalg, (batch, 784))
    lm = O.LossModel(om, O.CategoricalCrossentropy)
    global_iter = 0
    for epoch in range(3):
        iters, loss = e.train_epoch(global_iter)
        res = O.optimize_on_partition(lm, oalg, x, y, batch, global_iter, mp, op)
        assert iters == res.iterations == 5
        assert abs(loss - res.loss) <= 3e-5 * abs(res.loss) + 1e-6  # forward-only loss on the last batch, updated params
        mp, op = res.model_params, res.opt_params
        global_iter += iters  # the step counter is global (Optimizer.scala:88,143)
    assert_params_close(e.get_model_params(), mp, 1e-3, 2e-5, "after 3 epochs")
    assert_params_close(e.get_opt_params(), op, 5e-3, 2e-5, "opt after 3 epochs")
    # a global_iter offset (what a second worker contributes) changes the bias correction: check it is honoured
    iters, _ = e.train_epoch(1000)
    res = O.optimize_on_partition(lm, oalg, x, y, batch, 1000, mp, op)
    assert_params_close(e.get_model_params(), res.model_params, 1e-3, 2e-5, "epoch at global_iter=1000")
    e.close()


def test_shard_smaller_than_one_batch_is_rejected():  # Iterators.scala:79-81 NoSuchElementException
    _, sm = build_models(MLP)
    e = make_engine(sm, S.CategoricalCrossentropy, S.Adam(), 64, (784,), (10,))
    e.init_params()
    x, y = synth_classification(10, (784,), 10, 1)
    e.upload_dataset(x, y)
    with pytest.raises(S.IllegalArgument):
        e.train_epoch(0)
    e.close()


def test_state_errors_and_unsupported():
    _, sm = build_models(MLP)
    e = make_engine(sm, S.CategoricalCrossentropy, S.Adam(), 8, (784,), (10,))
    x, y = synth_classification(8, (784,), 10, 1)
    with pytest.raises(S.NativeError):
        e.train_step(x, y, 1)  # params not initialised
    with pytest.raises(S.IllegalArgument):
        e.set_param("0/weights", np.zeros(3, f32))  # wrong size
    with pytest.raises(S.IllegalArgument, match="missing 9/nope param"):  # core/Params.scala:22-23
        S.native.check(e._lib.sb_params_get(e.handle, b"9/nope", np.zeros(1, f32).ctypes.data, 1))
    e.close()
    # unseeded random initializers: the host runtime draws a seed per engine, as every partition initialises independently in the
    # reference (Optimizer.scala:123-131); the C ABI itself refuses a random initializer without a seed
    ws = []
    for _ in range(2):
        e = S.Engine(S.Dense(3, kernelInitializer=S.GlorotUniform()), S.MeanSquaredError, S.SGD(), 2, (2,), (3,))
        e.init_params()
        ws.append(e.get_param("0/weights"))
        e.close()
    lim = np.sqrt(6.0 / 5.0)
    assert np.all(np.abs(ws[0]) <= lim) and np.any(ws[0] != 0) and not np.array_equal(ws[0], ws[1])
    leaves = S.Dense(3, kernelInitializer=S.GlorotUniform(7)).leaves()
    arr = (S.native.sb_layer_desc * len(leaves))(*[l.to_native() for l in leaves])
    arr[0].init[0].has_seed = 0
    md = S.native.sb_model_desc(); md.n_layers = len(leaves); md.layers = arr; md.input_rank = 1; md.label_rank = 1
    md.input_dims[0] = 2; md.label_dims[0] = 3
    td = S.native.sb_train_desc(2, S.native.LOSS_MSE, S.native.OPT_SGD, 0.01, 0.9, 0.999, 1e-7, 0.0, 0, 1, 0, 1)
    import ctypes as C
    h = C.c_void_p()
    S.native.check(e._lib.sb_engine_create(C.byref(md), C.byref(td), 0, C.byref(h)))
    with pytest.raises(S.Unsupported):
        S.native.check(e._lib.sb_init_params(h))
    e._lib.sb_engine_destroy(h)


# ---------------------------------------------------------------------------------------------------------------
# the reference's optimizer tests, re-run through the drop-in API (optimizers/*Spec.scala)
# ---------------------------------------------------------------------------------------------------------------
def _linear(fixtures_dir):
    d = np.loadtxt(os.path.join(fixtures_dir, "linear_function_1.scv"), delimiter=",", dtype=np.float32)
    return S.Dataset(d[:, :1], d[:, 1:])


@pytest.mark.parametrize("alg,n_epochs,threshold", [
    (S.SGD(), 100, 11.0), (S.SGD(momentum=0.1), 100, 11.0), (S.SGD(momentum=0.1, nesterov=True), 100, 11.0),
    (S.Adam(rate=0.1), 100, 12.0), (S.AdaGrad(rate=1.0), 100, 9.0), (S.RMSProp(rate=0.06), 100, 9.4),
    (S.AdaDelta(), 100, 50.0), (S.Adamax(), 70, 9.0), (S.Nadam(), 50, 9.0), (S.AMSGrad(rate=0.1), 100, 9.0),
])
def test_optimizer_specs_linear_regression(fixtures_dir, alg, n_epochs, threshold):
    ds = _linear(fixtures_dir)
    rec = S.RecordLoss(console=False)
    trained = (ds.train(S.LinearRegression()).loss(S.MeanSquaredError).using(alg).batch(97)
               .each(S.epochs(1), rec).stopAfter(S.epochs(n_epochs)).quiet().run())
    loss = trained.loss(ds.features[:97], ds.labels[:97])
    assert loss <= threshold
    assert len(rec.history) == n_epochs
    # and the whole run agrees with the oracle's run of the same training
    d = np.loadtxt(os.path.join(fixtures_dir, "linear_function_1.scv"), delimiter=",", dtype=np.float32)
    name = {0: "sgd", 1: "adam", 2: "adagrad", 3: "rmsprop", 4: "adadelta", 5: "adamax", 6: "nadam", 7: "amsgrad"}[alg.code]
    kw = {"sgd": dict(rate=alg.rate, momentum=alg.beta1, nesterov=alg.nesterov), "adam": dict(rate=alg.rate),
          "adagrad": dict(rate=alg.rate), "rmsprop": dict(rate=alg.rate), "adadelta": dict(), "adamax": dict(), "nadam": dict(),
          "amsgrad": dict(rate=alg.rate)}[name]
    om = O.Dense(1, O.Identity, kernel_initializer=O.Zeros)
    mp, _, hist = O.run_training(om, O.MeanSquaredError, OALG[name](**kw), d[:, :1], d[:, 1:], batch=97, epochs=n_epochs)
    np.testing.assert_allclose(rec.history[-1], hist[-1][2], rtol=2e-3)
    assert_params_close(trained.params, mp, 2e-3, 1e-4, "final params")


def test_adam_logistic_regression_spec(fixtures_dir):  # optimizers/AdamSpec.scala:44-64
    d = np.loadtxt(os.path.join(fixtures_dir, "logistic_regression_1.scv"), delimiter=",", dtype=np.float64)
    ds = S.Dataset((d[:, :2].astype(f32) / f32(100)), d[:, 2:].astype(f32))
    trained = (ds.train(S.LogisticRegression()).loss(S.BinaryCrossentropy).using(S.Adam(0.1)).batch(100)
               .each(S.epochs(1), S.RecordLoss(console=False)).stopAfter(S.epochs(100)).quiet().run())
    assert trained.loss(ds.features, ds.labels) <= 0.4
    assert S.accuracy(trained, ds) >= 0.9
    probe = trained.result(np.array([[0.3462, 0.7802], [0.6018, 0.8630]], f32))
    assert np.array_equal(np.round(probe), [[0], [1]])


# ---------------------------------------------------------------------------------------------------------------
# regularisation penalties (models/Regularization.scala:16-44; loss + penalty, models/Model.scala:97)
# ---------------------------------------------------------------------------------------------------------------
def test_l2_penalty_goldens():  # models/layer/ComposedLayerSpec.scala:50-69, 97-120
    model = S.Dense(3, S.Sigmoid, S.L2(1.0)) >> S.Dense(1, S.Sigmoid, S.L2(1.0))
    e = make_engine(model, S.MeanSquaredError, S.SGD(), 2, (3,), (1,))
    e.set_params({"0/weights": np.array([[1, 0.5, 1], [0.1, 1, 1], [1, 0, 0.2]], f32), "1/weights": np.zeros(3, f32),
                  "3/weights": np.array([[0.1], [0.5], [1]], f32), "4/weights": np.zeros(1, f32)})
    x, y = np.array([[0, 0, 1], [0, 1, 1]], f32), np.array([[1], [0]], f32)
    assert abs(e.eval_loss(x, y) - 3.6199622) < 1e-6  # 0.339962 (MSE) + 3.28 (penalty)
    e.close()


@pytest.mark.parametrize("reg", ["l1", "l2"])
def test_regularised_mlp_trajectory(reg):
    oreg, sreg = (O.L1(0.02), S.L1(0.02)) if reg == "l1" else (O.L2(0.05), S.L2(0.05))
    om = O.Dense(24, O.Sigmoid, reg=oreg, kernel_initializer=O.GlorotUniform(1)) >> O.Dense(10, O.Softmax, reg=oreg,
                                                                                             kernel_initializer=O.GlorotUniform(1))
    sm = S.Dense(24, S.Sigmoid, sreg, kernelInitializer=S.GlorotUniform(1)) >> S.Dense(10, S.Softmax, sreg,
                                                                                      kernelInitializer=S.GlorotUniform(1))
    batch, steps = 32, 8
    x, y = synth_classification(batch * steps, (40,), 10, 3)
    e = make_engine(sm, S.CategoricalCrossentropy, S.Adam(0.01), batch, (40,), (10,))
    e.init_params()
    oalg = O.Adam(0.01)
    _, mp, op = oracle_init(om, oalg, (batch, 40))
    lm = O.LossModel(om, O.CategoricalCrossentropy)
    for t in range(1, steps + 1):
        xb, yb = x[(t - 1) * batch:t * batch], y[(t - 1) * batch:t * batch]
        gl = e.train_step(xb, yb, t)
        mp, op, ol = O.train_step(lm, oalg, xb, yb, mp, op, t)
        assert abs(gl - float(ol)) <= 2e-5 * abs(float(ol)) + 1e-6, (t, gl, float(ol))  # loss includes the penalty
    assert_params_close(e.get_model_params(), mp, 1e-3, 2e-5, reg)
    e.close()


def test_device_truncated_normal_initializers():  # models/InitializerSpec.scala:44-66 goldens + the oracle's restated TF stream
    e = make_engine(S.Dense(3, kernelInitializer=S.RandomNormalTruncated(0.0, 1.0, seed=1), bias=False), S.MeanSquaredError, S.SGD(),
                    1, (1,), (3,))
    e.init_params()
    w = e.get_param("weights").reshape(-1)
    assert np.array_equal(np.round(w * f32(100)) / f32(100), np.array([0.31, 0.57, 0.42], f32))
    e.close()
    for sinit, oinit, shape in [(S.GlorotNormal(2), O.GlorotNormal(seed=2), (37, 50)), (S.HeNormal(3), O.HeNormal(seed=3), (64, 9)),
                                (S.LecunNormal(4), O.LecunNormal(seed=4), (5, 130)),
                                (S.RandomNormalTruncated(0.5, 0.2, seed=9), O.RandomNormalTruncated(0.5, 0.2, seed=9), (33, 7))]:
        e = make_engine(S.Dense(shape[1], kernelInitializer=sinit, bias=False), S.MeanSquaredError, S.SGD(), 1, (shape[0],), (shape[1],))
        e.init_params()
        got, want = e.get_param("weights"), oinit.build(shape)
        assert np.max(np.abs(got)) < 2.0 * np.max(np.abs(want)) + 1.0
        np.testing.assert_allclose(got, want, rtol=0, atol=3e-6)  # same accepted samples; libm vs CUDA sincos/log at the ulp level
        e.close()


# ---------------------------------------------------------------------------------------------------------------
# geometry sweep of the vectorised / fused kernels (strip max-pool, small-K conv, skinny head, fused bias+activation backward):
# ragged widths, single rows/columns, channel counts that fall back to the generic kernels, batch 1 and batch > 128
# ---------------------------------------------------------------------------------------------------------------
GEOMETRIES = [
    # (id, spec, batch, in_shape, classes)
    ("pool_w2_c4", [("conv", 4, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"), ("flatten",),
                    ("dense", 3, "softmax")], 2, (4, 4, 1), 3),
    ("pool_odd_c12", [("conv", 12, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"), ("flatten",),
                      ("dense", 10, "softmax")], 3, (9, 13, 3), 10),
    ("pool_c6_fallback", [("conv", 6, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"), ("flatten",),
                          ("dense", 10, "softmax")], 3, (7, 8, 1), 10),
    ("pool_3x3_s2_same", [("conv", 8, "tanh", (3, 3), (1, 1), "Same", False), ("pool", (3, 3), (2, 2), "Same"), ("flatten",),
                          ("dense", 10, "softmax")], 2, (9, 7, 3), 10),
    ("conv5x5_same", [("conv", 16, "relu", (5, 5), (1, 1), "Same", False), ("pool", (2, 2), (1, 1), "Valid"), ("flatten",),
                      ("dense", 10, "softmax")], 2, (8, 9, 1), 10),
    ("conv_strided_smallk", [("conv", 8, "sigmoid", (3, 3), (2, 2), "Valid", True), ("flatten",), ("dense", 10, "softmax")], 3,
     (11, 9, 3), 10),
    ("head_k2048_n16", [("dense", 2048, "tanh"), ("dense", 16, "softmax")], 5, (40,), 16),
    ("head_batch130", [("dense", 512, "sigmoid"), ("dense", 10, "softmax")], 130, (24,), 10),
    ("head_batch1_n3", [("dense", 1024, "tanh"), ("dense", 3, "softmax")], 1, (8,), 3),
    ("head_k_not_mult4_fallback", [("dense", 1026, "tanh"), ("dense", 10, "softmax")], 4, (8,), 10),
    ("bias_act_c6_fallback", [("dense", 6, "relu"), ("dense", 10, "softmax")], 9, (5,), 10),
]


@pytest.mark.parametrize("gid,spec,batch,in_shape,classes", GEOMETRIES, ids=[g[0] for g in GEOMETRIES])
def test_fast_kernel_geometries_match_the_oracle(gid, spec, batch, in_shape, classes):
    om, sm = build_models(spec)
    x, y = synth_classification(batch, in_shape, classes, 77)
    x = (x - 0.5).astype(f32)  # both signs: ReLU masks and max-pool ties are exercised
    e = make_engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), batch, in_shape, (classes,))
    e.init_params()
    _, mp, _ = oracle_init(om, O.Adam(0.001), (batch,) + tuple(in_shape))
    ol, og, _ = O.LossModel(om, O.CategoricalCrossentropy).grad_stateful(x, y, mp)
    gl, gg = e.compute_grads(x, y)
    assert abs(gl - float(ol)) <= 2e-5 * abs(float(ol)) + 1e-7
    for k, v in og.items():
        np.testing.assert_allclose(gg[k], v, rtol=2e-4, atol=2e-6, err_msg=k)
    np.testing.assert_allclose(e.predict(x), O.model_forward(om, x, mp)[0], rtol=2e-5, atol=2e-6)
    e.close()


# ---------------------------------------------------------------------------------------------------------------
# estimators on the device (estimators/package.scala:18-137)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows", [64, 150, 37])  # whole batches, a partial tail (`remaining = Partial`), less than one batch
def test_device_estimators_match_the_oracle_classifier(rows):
    spec = [("dense", 16, "sigmoid"), ("dense", 4, "softmax")]
    om, sm = build_models(spec)
    x, y = synth_classification(rows, (12,), 4, seed=5)
    e = make_engine(sm, S.CategoricalCrossentropy, S.Adam(0.05), 32, (12,), (4,))
    _, mp, _ = oracle_init(om, O.Adam(rate=0.05), (32, 12))
    e.set_params(mp)
    e.init_opt_params()
    # train a little on full batches so that predictions are not all on one side of 0.5
    xt, yt = synth_classification(32 * 20, (12,), 4, seed=6)
    for t in range(20):
        e.train_step(xt[t * 32:(t + 1) * 32], yt[t * 32:(t + 1) * 32], t + 1)
    params = e.get_model_params()
    trained = S.TrainedModel(S.LossModel(sm, S.CategoricalCrossentropy), params, e)
    ds = S.Dataset(x, y)
    # counts are exact as long as no prediction sits within rounding noise of .5; the oracle runs on the ENGINE's weights
    assert S.accuracy(trained, ds) == O.estimators.accuracy(om, params, x, y, batch=32)
    np.testing.assert_allclose(S.MSE(trained, ds), O.estimators.mse(om, params, x, y, batch=32), rtol=2e-5)
    np.testing.assert_allclose(S.RMSE(trained, ds), O.estimators.rmse(om, params, x, y, batch=32), rtol=2e-5)
    np.testing.assert_allclose(S.MAE(trained, ds), O.estimators.mae(om, params, x, y, batch=32), rtol=2e-5)
    # resident-shard variant gives the same sums as the host-buffer variant
    if rows >= 32:
        e.upload_dataset(x, y)
        assert e.estimate(S.native.EST_ACCURACY) == e.estimate(S.native.EST_ACCURACY, x, y)
        a, b = e.estimate(S.native.EST_SQUARED_ERROR), e.estimate(S.native.EST_SQUARED_ERROR, x, y)
        np.testing.assert_allclose(a, b, rtol=1e-12)
    e.close()


def test_device_r2_score_matches_the_oracle(fixtures_dir):  # estimators/package.scala:99-137, labels of shape (1)
    ds = _linear(fixtures_dir)
    trained = (ds.train(S.LinearRegression()).loss(S.MeanSquaredError).using(S.Adam(rate=0.1)).batch(97)
               .stopAfter(S.epochs(30)).quiet().run())
    om = O.Dense(1, O.Identity, kernel_initializer=O.Zeros)
    want = O.estimators.r2_score(om, trained.params, ds.features, ds.labels, batch=97)
    np.testing.assert_allclose(S.R2Score(trained, ds), want, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(S.MSE(trained, ds), O.estimators.mse(om, trained.params, ds.features, ds.labels, batch=97),
                               rtol=2e-5)
    with pytest.raises(ValueError):  # `require(ds.labelsShape == Shape(1))`
        S.R2Score(trained, S.Dataset(ds.features, np.concatenate([ds.labels, ds.labels], axis=1)))


# ---------------------------------------------------------------------------------------------------------------
# wire format + checkpoint / resume (optimizers/KryoSerializers.scala:27-63)
# ---------------------------------------------------------------------------------------------------------------
def test_param_wire_records_round_trip_through_the_device():
    from oracle import wire as W
    spec = [("conv", 4, "relu", (3, 3), (1, 1), "Valid", True), ("flatten",), ("dense", 3, "softmax")]
    _, sm = build_models(spec)
    e = make_engine(sm, S.CategoricalCrossentropy, S.Adam(0.01), 8, (6, 6, 2), (3,))
    e.init_params()
    for path, t in e.get_params().items():
        rec = e.param_to_wire(path)
        assert rec == W.tensor_write(t), path       # TensorSerializer.write layout, byte for byte
        e.set_param(path, np.zeros_like(t))
        assert e.param_from_wire(path, rec) == len(rec)
        assert np.array_equal(e.get_param(path), t)
    w = next(iter(e.get_params()))
    with pytest.raises(S.native.IllegalArgument):   # a record of another shape is refused
        e.param_from_wire(w, W.tensor_write(np.zeros((2, 2), np.float32)))
    e.close()


@pytest.mark.parametrize("k", [1, 2])
def test_checkpoint_resume_continues_bit_identically(tmp_path, k):
    spec = [("dense", 24, "tanh"), ("bn", 0.9, 0), ("dense", 5, "softmax")]
    x, y = synth_classification(64 * 4 * k, (10,), 5, seed=11)

    def builder():
        _, sm = build_models(spec)
        return (S.Dataset(x, y, partitions=k).train(sm).loss(S.CategoricalCrossentropy).using(S.Adam(0.02)).batch(64)
                .devices([0] * k).quiet())

    rec_full = S.RecordLoss(console=False)
    full = builder().each(S.epochs(1), rec_full).stopAfter(S.epochs(6)).run()
    ck = str(tmp_path / "model.sb3")
    rec_a, rec_b = S.RecordLoss(console=False), S.RecordLoss(console=False)
    builder().each(S.epochs(1), rec_a).each(S.epochs(3), S.SaveCheckpoint(ck)).stopAfter(S.epochs(4)).run()
    # the file was last written at epoch 3; a fresh process-equivalent resumes there and runs to epoch 6
    resumed = builder().resumeFrom(ck).each(S.epochs(1), rec_b).stopAfter(S.epochs(6)).run()
    assert rec_a.history[:3] + rec_b.history == rec_full.history
    for p, t in full.params.items():
        assert np.array_equal(resumed.params[p], t), p
    with pytest.raises(S.native.IllegalArgument):   # a checkpoint of another model is refused
        _, other = build_models([("dense", 5, "softmax")])
        (S.Dataset(x, y).train(other).loss(S.CategoricalCrossentropy).using(S.Adam(0.02)).batch(64).resumeFrom(ck)
         .stopAfter(S.epochs(1)).quiet().run())


# ---------------------------------------------------------------------------------------------------------------
# host-fed multi-batch call (the `optimize` loop over host batches, Optimizer.scala:135-147)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_batches", [19, 8, 3])  # two graph groups + 3 single steps / exactly one group / singles only
def test_train_batches_matches_step_by_step(n_batches):
    spec = [("conv", 4, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"), ("flatten",),
            ("dense", 5, "softmax")]
    om, sm = build_models(spec)
    B = 16
    x, y = synth_classification(B * n_batches + 5, (8, 8, 1), 5, seed=13)  # 5 trailing rows: the partial batch is dropped
    _, mp, _ = oracle_init(om, O.Adam(rate=0.01), (B, 8, 8, 1))
    ea = make_engine(sm, S.CategoricalCrossentropy, S.Adam(0.01), B, (8, 8, 1), (5,))
    eb = make_engine(sm, S.CategoricalCrossentropy, S.Adam(0.01), B, (8, 8, 1), (5,))
    for e in (ea, eb):
        e.set_params(mp)
        e.init_opt_params()
    it0 = 7  # a resumed run: global iteration does not start at 0
    want = [ea.train_step(x[j * B:(j + 1) * B], y[j * B:(j + 1) * B], it0 + j + 1) for j in range(n_batches)]
    got = eb.train_batches(x, y, it0)
    assert got.shape == (n_batches,)
    assert np.array_equal(got, np.asarray(want, np.float32))            # same kernels, same order: bit-identical
    pa, pb = ea.get_params(), eb.get_params()
    for k in pa:
        assert np.array_equal(pa[k], pb[k]), k
    # and the trajectory is the oracle's
    lm = O.LossModel(om, O.CategoricalCrossentropy)
    p, op = dict(mp), O.init_params(om, O.Adam(rate=0.01), (B, 8, 8, 1), np.float32, mp)[2]
    for j in range(min(n_batches, 4)):
        p, op, loss = O.train_step(lm, O.Adam(rate=0.01), x[j * B:(j + 1) * B], y[j * B:(j + 1) * B], p, op, it0 + j + 1)
        np.testing.assert_allclose(got[j], loss, rtol=1e-4)
    ea.close(); eb.close()


# ---------------------------------------------------------------------------------------------------------------
# Explicit padding (math/nn/AllKernels.scala:67-102)
# ---------------------------------------------------------------------------------------------------------------
def test_explicit_padding_trajectory_and_errors():
    # asymmetric pads, stride 2 on the second conv (pads chosen to fit), bias + sigmoid: forward, wgrad, dgrad and updates
    spec = [("conv", 4, "relu", (3, 3), (1, 1), ((2, 1), (0, 3)), False), ("pool", (2, 2), (1, 1), "Valid"),
            ("conv", 8, "sigmoid", (2, 2), (2, 2), ((1, 0), (1, 1)), True), ("flatten",), ("dense", 5, "softmax")]
    run_trajectory(spec, "cce", "adam", dict(rate=0.01), batch=6, in_shape=(7, 6, 2), classes=5, steps=8)
    with pytest.raises(S.IllegalArgument):  # pads that do not fit the stride: the reference's ceil and TF's floor would disagree
        make_engine(S.Conv2D(2, (2, 2), (2, 2), S.Explicit((1, 0), (0, 0))), S.MeanSquaredError, S.SGD(), 1, (4, 4, 1), (3, 2, 2))
    with pytest.raises(ValueError):         # Pool2D refuses Explicit like the reference (KernelsSpec.scala:258-268)
        S.Pool2D(padding=S.Explicit((1, 0), (1, 0)))


# ---------------------------------------------------------------------------------------------------------------
# average pooling when the descriptor asks to honour `reduce` (math/nn/KernelsSpec.scala:213-225, 294-305)
# ---------------------------------------------------------------------------------------------------------------
def test_avg_pool_goldens_and_trajectory():
    exp = np.array([[1.75, 2.0, 1.5, 1.5, 2.0], [1.5, 2.25, 2.5, 2.0, 1.5], [1.5, 1.5, 1.75, 1.25, 0.5],
                    [1.0, 1.25, 1.25, 1.25, 1.5], [0.0, 1.5, 2.0, 1.5, 2.0]], f32)
    e = make_engine(S.Pool2D((2, 2), (1, 1), "Same", reduce="Avg", honorReduce=True), S.MeanSquaredError, S.SGD(), 1, (5, 5, 1),
                    (5, 5, 1))
    e.init_params()
    assert np.array_equal(e.predict(IMG).reshape(5, 5), exp)
    e.close()
    e = make_engine(S.Pool2D((2, 2), (1, 1), "Same", reduce="Avg"), S.MeanSquaredError, S.SGD(), 1, (5, 5, 1), (5, 5, 1))
    e.init_params()
    assert e.predict(IMG).reshape(5, 5)[0, 0] == 3.0          # the reference layer: `reduce` ignored, max-pools
    e.close()
    spec = [("conv", 4, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (2, 2), "Same", "Avg"),
            ("conv", 6, "tanh", (2, 2), (1, 1), "Valid", True), ("pool", (2, 2), (1, 1), "Valid", "Avg"), ("flatten",),
            ("dense", 5, "softmax")]
    run_trajectory(spec, "cce", "adam", dict(rate=0.01), batch=6, in_shape=(9, 8, 2), classes=5, steps=8)


@pytest.mark.parametrize("prec", [S.PREC_FP32, S.PREC_TF32], ids=["fp32", "tf32"])
def test_winner_map_path_is_bit_identical_to_the_activation_path(prec, monkeypatch):
    """conv >> ReLU >> Pool2D(2x2, stride 1) blocks: the forward records one winner byte per pool output instead of the conv
    activations and the backward routes dy with it (kernels_fast.cu win4 / pool22_bwd_map_kernel; the tensor-core conv's pooled
    epilogue in TF32 mode).  Loss, every gradient and the updated parameters must be the SAME BITS as with SB_NO_POOL_MAP=1."""
    _, sm = build_models(LENET)
    x, y = synth_classification(12, (14, 14, 1), 10, 21)
    outs = []
    # activation path | winner maps, head dgrad unfused | winner maps + the head's input gradient formed inside pool2's backward
    # (conv1's map is the default; the map of the standalone pool kernel and the head fusion are opt-in: SB_POOL_MAP_ALL=1)
    monkeypatch.setenv("SB_TC_FOLD", "0")  # the KW-folded conv kernel sums the taps of a kernel row in its epilogue: other bits
    for off, nofuse in (("1", "1"), ("0", "1"), ("0", "0")):
        monkeypatch.setenv("SB_NO_POOL_MAP", off)
        monkeypatch.setenv("SB_POOL_MAP_ALL", "1")
        monkeypatch.setenv("SB_TC_POOL_FUSION", "1")  # TF32: conv2's fprop pools in its own epilogue (opt-in as well)
        monkeypatch.setenv("SB_NO_HEAD_DGRAD_FUSION", nofuse)
        e = make_engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), 12, (14, 14, 1), (10,), precision=prec)
        e.init_params()
        loss, grads = e.compute_grads(x, y)
        l1 = e.train_step(x, y, 1)
        l2 = e.train_step(x, y, 2)
        outs.append((loss, l1, l2, {k: v.tobytes() for k, v in grads.items()}, {k: v.tobytes() for k, v in e.get_params().items()}))
        e.close()
    for o in outs[1:]:
        assert outs[0][:3] == o[:3]
        for k in outs[0][3]:
            assert outs[0][3][k] == o[3][k], f"gradient {k} differs"
        assert outs[0][4] == o[4]
