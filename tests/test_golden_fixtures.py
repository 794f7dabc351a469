"""The committed golden fixtures (tests/golden/reference_vectors.json: the reference's own known-answer vectors, see
tests/golden/transcribe.py) against the CPU oracle (runs everywhere) and against the CUDA engine through the C ABI
(`-m gpu`).  Each case names the reference spec it was transcribed from."""
import json
import os

import numpy as np
import pytest

import oracle as O
import scanet3_b200 as S
from test_oracle_goldens import close_rounded, rnd

f32 = np.float32
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))["cases"]


def by_kind(kind):
    return [pytest.param(c, id=c["ref"].split("/")[-1]) for c in CASES if c["kind"] == kind]


def nhwc(a, dtype=np.float64):
    a = np.asarray(a, dtype)
    return a.reshape((1,) + a.shape + (1,))


def hwio(a, dtype=np.float64):
    a = np.asarray(a, dtype)
    return a.reshape(a.shape + (1, 1))


def lstm_params(c):
    p = {}
    for g, (k, r, b) in c["gates"].items():
        p[f"{g}/0/kernel_weights"], p[f"{g}/0/recurrent_weights"], p[f"{g}/1/weights"] = (np.array(v, f32) for v in (k, r, b))
    return p


# ---------------------------------------------------------------------------------------------------------------
# oracle (CPU)
# ---------------------------------------------------------------------------------------------------------------
def test_fixture_inventory():
    kinds = {c["kind"] for c in CASES}
    assert {"conv2d_fwd", "conv2d_wgrad", "conv2d_dgrad", "pool2d_fwd", "pool2d_grad", "dense_sigmoid_fwd",
            "dense_sigmoid_bce_grad", "linear_regression", "logistic_regression", "simple_rnn_fwd", "lstm_fwd",
            "batchnorm_generic_train", "random_uniform_seeded"} <= kinds
    assert all(c["ref"].endswith(tuple("0123456789")) for c in CASES)  # every case cites file:line


@pytest.mark.parametrize("c", by_kind("conv2d_fwd"))
def test_oracle_conv2d_fwd(c):
    y = O.conv2d(nhwc(c["x"]), hwio(c["w"]), tuple(c["strides"]), c["padding"])
    assert np.array_equal(y[0, :, :, 0], np.array(c["expected"], np.float64))


@pytest.mark.parametrize("c", by_kind("pool2d_fwd"))
def test_oracle_pool2d_fwd(c):
    y = O.pool2d(nhwc(c["x"]), tuple(c["window"]), tuple(c["strides"]), c["padding"], c["reduce"])
    assert np.array_equal(y[0, :, :, 0], np.array(c["expected"], np.float64))


@pytest.mark.parametrize("c", by_kind("pool2d_grad"))
def test_oracle_pool2d_grad(c):
    x = nhwc(c["x"])
    g = O.pool2d_grad(x, np.ones_like(x), tuple(c["window"]), tuple(c["strides"]), c["padding"], c["reduce"])
    assert np.array_equal(g[0, :, :, 0], np.array(c["expected"], np.float64))


def test_oracle_conv2d_gradients():
    (cw,), (cd,) = [[c for c in CASES if c["kind"] == k] for k in ("conv2d_wgrad", "conv2d_dgrad")]
    x = nhwc(cw["x"])
    g = np.ones((1, 4, 4, 1))
    assert np.array_equal(O.conv2d_wgrad(x, (2, 2, 1, 1), g).reshape(2, 2), np.array(cw["expected"], np.float64))
    assert np.array_equal(O.conv2d_dgrad((1, 5, 5, 1), hwio(cd["w"]), g).reshape(5, 5), np.array(cd["expected"], np.float64))


@pytest.mark.parametrize("c", by_kind("lstm_fwd"))
def test_oracle_lstm_fwd(c):
    y, _, _ = O.model_forward(O.RNN(O.LSTMCell(units=2)), np.array(c["x"], f32).reshape(1, -1, 1), lstm_params(c))
    assert np.array_equal(rnd(y, c["round"]), rnd(c["expected"], c["round"]))


@pytest.mark.parametrize("c", by_kind("dense_sigmoid_fwd"))
def test_oracle_dense_fwd(c):
    y, _, _ = O.model_forward(O.Dense(4, O.Sigmoid), np.array(c["x"], f32),
                              {"0/weights": np.array(c["w"], f32), "1/weights": np.array(c["b"], f32)})
    assert close_rounded(y, c["expected"], c["round"])


@pytest.mark.parametrize("c", by_kind("random_uniform_seeded"))
def test_oracle_seeded_uniform(c):
    v = O.RandomUniform(c["min"], c["max"], seed=c["seed"]).build(tuple(c["shape"]))
    assert np.array_equal(rnd(v, c["round"]), rnd(c["expected"], c["round"]))


# ---------------------------------------------------------------------------------------------------------------
# CUDA engine through the C ABI (fp32 parity mode)
# ---------------------------------------------------------------------------------------------------------------
def _engine(model, loss, batch, in_shape, lab_shape):
    return S.Engine(model, loss, S.SGD(), batch, in_shape, lab_shape, device=0)


@pytest.mark.gpu
@pytest.mark.parametrize("c", [p for p in by_kind("conv2d_fwd")])
def test_engine_conv2d_fwd(c):
    exp = np.array(c["expected"], f32)
    e = _engine(S.Conv2D(1, (2, 2), tuple(c["strides"]), c["padding"].capitalize()), S.MeanSquaredError, 1, (5, 5, 1), exp.shape + (1,))
    e.set_param("weights", hwio(c["w"], f32))
    assert np.array_equal(e.predict(nhwc(c["x"], f32))[0, :, :, 0], exp)
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("c", [p for p in by_kind("pool2d_fwd") if p.values[0]["reduce"] == "max"])
def test_engine_maxpool_fwd(c):
    exp = np.array(c["expected"], f32)
    model = S.Conv2D(1, (1, 1)) >> S.Pool2D(tuple(c["window"]), tuple(c["strides"]), c["padding"].capitalize())
    e = _engine(model, S.MeanSquaredError, 1, (5, 5, 1), exp.shape + (1,))
    e.set_param("0/weights", np.ones((1, 1, 1, 1), f32))  # identity 1x1 conv in front: the engine needs a computing layer
    assert np.array_equal(e.predict(nhwc(c["x"], f32))[0, :, :, 0], exp)
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("c", by_kind("lstm_fwd"))
def test_engine_lstm_fwd(c):
    e = _engine(S.LSTM(2), S.MeanSquaredError, 1, (3, 1), (2,))
    e.set_params(lstm_params(c))
    np.testing.assert_allclose(e.predict(np.array(c["x"], f32).reshape(1, 3, 1)), np.array(c["expected"], f32), atol=6e-7)
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("c", by_kind("dense_sigmoid_bce_grad"))
def test_engine_dense_bce_grad(c):
    e = _engine(S.Dense(4, S.Sigmoid), S.BinaryCrossentropy, 4, (3,), (4,))
    e.set_params({"0/weights": np.zeros((3, 4), f32), "1/weights": np.zeros(4, f32)})
    _, g = e.compute_grads(np.array(c["x"], f32), np.array(c["y"], f32))
    np.testing.assert_allclose(g["0/weights"], np.array(c["expected_w"], f32), rtol=3e-6, atol=1e-8)
    np.testing.assert_allclose(g["1/weights"], np.array(c["expected_b"], f32), rtol=3e-6, atol=1e-8)
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("c", by_kind("logistic_regression"))
def test_engine_logistic_regression(c):
    e = _engine(S.Dense(1, S.Sigmoid), S.BinaryCrossentropy, 2, (2,), (1,))
    e.set_params({"0/weights": np.array(c["w"], f32), "1/weights": np.array(c["b"], f32)})
    x, y = np.array(c["x"], f32), np.array(c["y"], f32)
    assert abs(e.eval_loss(x, y) - c["loss"]) < 1e-6
    assert np.array_equal(rnd(e.predict(x), 6), rnd(c["result"], 6))
    _, g = e.compute_grads(x, y)
    assert np.array_equal(rnd(g["0/weights"], 6), rnd(c["grad_w"], 6))
    assert np.array_equal(rnd(g["1/weights"], 6), rnd(c["grad_b"], 6))
    e.close()
