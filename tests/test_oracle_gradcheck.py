"""The reference has no test for LSTM gradients, conv/pool inside a trained model, or BatchNorm gradients
(SURVEY 8c "parity unpinned").  The oracle's hand-derived backward passes are therefore checked here against
central finite differences of its own (golden-pinned) forward in float64.  CPU only."""
import numpy as np
import pytest

import oracle as O


def fd_check(model, loss, in_shape, classes, seed=0, eps=1e-6, tol=2e-6, n_probe=12):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.1, 1.0, size=in_shape)
    y = np.zeros((in_shape[0], classes))
    if classes > 1:
        y[np.arange(in_shape[0]), np.arange(in_shape[0]) % classes] = 1.0
    else:
        y[:] = rng.uniform(size=y.shape)
    defs = O.model_params(model, in_shape)
    params = {}
    for k, d in defs.items():
        if d.initializer in (O.Zeros, O.Ones):
            params[k] = d.initialize().astype(np.float64) + (0.1 * rng.normal(size=d.shape) if d.trainable else 0)
        else:
            params[k] = 0.5 * rng.normal(size=d.shape)
    lm = O.LossModel(model, loss)
    _, grads, _ = lm.grad_stateful(x, y, params)
    for k in grads:
        flat = params[k].reshape(-1)
        idx = rng.choice(flat.size, size=min(n_probe, flat.size), replace=False)
        for i in idx:
            old = flat[i]
            flat[i] = old + eps
            lp = float(lm.loss_stateful(x, y, params)[0])
            flat[i] = old - eps
            lmn = float(lm.loss_stateful(x, y, params)[0])
            flat[i] = old
            num = (lp - lmn) / (2 * eps)
            ana = grads[k].reshape(-1)[i]
            assert abs(num - ana) <= tol * max(1.0, abs(num)), (k, i, num, ana)


def test_lstm_bptt_gradients():
    fd_check(O.LSTM(5, kernel_initializer=O.GlorotUniform(1), recurrent_initializer=O.Orthogonal(seed=1)) >> O.Dense(1, O.Tanh),
             O.MeanSquaredError, (3, 6, 2), 1)


def test_stacked_lstm_gradients_through_the_returned_sequence():  # returnSequence = true feeds a second LSTM (RNN.scala:24-25, 70)
    model = (O.LSTM(4, return_sequence=True, kernel_initializer=O.GlorotUniform(1), recurrent_initializer=O.Orthogonal(seed=1))
             >> O.LSTM(3, kernel_initializer=O.GlorotUniform(2), recurrent_initializer=O.Orthogonal(seed=2)) >> O.Dense(1, O.Tanh))
    fd_check(model, O.MeanSquaredError, (3, 5, 2), 1)


def test_simple_rnn_gradients_with_the_pre_activation_state_quirk():
    fd_check(O.SimpleRNN(4, kernel_initializer=O.GlorotUniform(1), recurrent_initializer=O.Orthogonal(seed=1)) >> O.Dense(1),
             O.MeanSquaredError, (3, 5, 2), 1)


def test_stacked_simple_rnn_gradients_through_the_returned_sequence():
    model = (O.SimpleRNN(4, return_sequence=True, kernel_initializer=O.GlorotUniform(1), recurrent_initializer=O.Orthogonal(seed=1))
             >> O.SimpleRNN(3, activation=O.Sigmoid, kernel_initializer=O.GlorotUniform(2), recurrent_initializer=O.Orthogonal(seed=2))
             >> O.Dense(1))
    fd_check(model, O.MeanSquaredError, (3, 5, 2), 1)


def test_lenet_style_cnn_gradients():
    model = (O.Conv2D(3, activation=O.ReLU()) >> O.Pool2D() >> O.Conv2D(4, (2, 2), (2, 2), "same", O.Sigmoid, bias=True)
             >> O.Pool2D((2, 2), (2, 2), "same") >> O.Flatten >> O.Dense(3, O.Softmax))
    fd_check(model, O.CategoricalCrossentropy, (2, 9, 8, 2), 3)


def test_batchnorm_gradients_generic_and_fused_correct():
    fd_check(O.Dense(4) >> O.BatchNorm(momentum=0.5) >> O.Dense(2, O.Sigmoid), O.BinaryCrossentropy, (6, 3), 2)
    model = O.Conv2D(3, activation=O.ReLU()) >> O.BatchNorm(bn_mode="correct") >> O.Flatten >> O.Dense(2, O.Softmax)
    fd_check(model, O.CategoricalCrossentropy, (4, 5, 5, 2), 2)


def test_fused_softmax_cce_formula_equals_autodiff():
    rng = np.random.default_rng(0)
    z = rng.normal(size=(7, 10))
    y = np.eye(10)[np.arange(7) % 10]
    p = O.Softmax.fwd(z)
    g = O.CategoricalCrossentropy.grad(p, y)
    dz_autodiff = O.Softmax.bwd(g, z, p)
    eps = np.float64(np.float32(1e-8))
    r = y / (p + eps)
    dz_fused = -(p / 7) * (r - np.sum(r * p, axis=1, keepdims=True))
    np.testing.assert_allclose(dz_fused, dz_autodiff, rtol=0, atol=1e-16)


@pytest.mark.parametrize("k", [1, 2, 4, 8])
def test_multi_partition_run_in_both_averaging_modes(k):
    rng = np.random.default_rng(1)
    x = rng.uniform(size=(64 * k + 5, 6)).astype(np.float32)
    y = np.eye(3, dtype=np.float32)[np.arange(x.shape[0]) % 3]
    model = O.Dense(5, O.Sigmoid, kernel_initializer=O.GlorotUniform(1)) >> O.Dense(3, O.Softmax, kernel_initializer=O.GlorotUniform(2))
    for mode in ("spark_chain", "mean"):
        mp, op, hist = O.run_training(model, O.CategoricalCrossentropy, O.Adam(0.01), x, y, batch=16, epochs=2, k=k, avg_mode=mode)
        # global step counter advances by the SUM of all partitions' iterations (Optimizer.scala:88,223)
        assert hist[-1][1] == 2 * sum(((b - a) // 16) for a, b in O.shard_rows(x.shape[0], k))
        assert all(np.all(np.isfinite(v)) for v in mp.values())
    if k <= 2:  # the chain IS the mean for k <= 2
        a = O.run_training(model, O.CategoricalCrossentropy, O.Adam(0.01), x, y, batch=16, epochs=2, k=k, avg_mode="spark_chain")[0]
        b = O.run_training(model, O.CategoricalCrossentropy, O.Adam(0.01), x, y, batch=16, epochs=2, k=k, avg_mode="mean")[0]
        for key in a:
            assert np.array_equal(a[key], b[key])
