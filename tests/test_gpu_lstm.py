"""RNN(LSTMCell) on the device against the reference's forward golden (models/layer/RNNLayerSpec.scala:44-79) and against the
oracle's BPTT (the reference has no LSTM-gradient test; the oracle's BPTT is pinned by finite differences in
tests/test_oracle_gradcheck.py).  fp32 contractions; tolerances as in test_gpu_parity.py: one op chain rtol 2e-5-class,
a 64-step recurrence accumulates ~T of them."""
import numpy as np
import pytest

import oracle as O
import scanet3_b200 as S
from helpers import assert_params_close, build_models
from test_oracle_goldens import LSTM_GOLD, lstm_gold_params

pytestmark = pytest.mark.gpu
f32 = np.float32


def synth_series(n, t, f, seed):
    """min-max scaled series windows like the sunspot data prep (test/Datasets.scala:80-85): features and targets in [0,1]"""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.uniform(0.0, 1.0, size=(n, t, f)).astype(f32)
    y = rng.uniform(0.0, 1.0, size=(n, 1)).astype(f32)
    return x, y


def test_lstm_forward_golden_through_the_engine():
    e = S.Engine(S.LSTM(2), S.MeanSquaredError, S.SGD(), 1, (3, 1), (2,), device=0)
    assert set(e.model_param_paths()) == set(lstm_gold_params().keys())
    e.set_params(lstm_gold_params())
    y = e.predict(np.array([1, 2, 3], f32).reshape(1, 3, 1))
    np.testing.assert_allclose(y, np.array([[0.382158, 0.029766]], f32), atol=6e-7)
    e.close()


@pytest.mark.parametrize("units,t,f,batch", [(8, 5, 3, 6), (16, 7, 1, 9), (12, 4, 11, 5), (32, 1, 2, 4)],
                         ids=["f3", "f1", "f11_gemm_projection", "single_step"])
def test_lstm_gradients_match_oracle_bptt(units, t, f, batch):
    om, sm = build_models([("lstm", units), ("dense", 1, "tanh")])
    x, y = synth_series(batch, t, f, 5)
    e = S.Engine(sm, S.MeanSquaredError, S.Adam(), batch, (t, f), (1,), device=0)
    e.init_params()
    _, mp, _ = O.init_params(om, O.Adam(), (batch, t, f))
    assert_params_close(e.get_model_params(), mp, 0, 2e-6, "init")  # Orthogonal: host QR, GlorotUniform: bit exact
    e.set_params(mp)
    ol, og, _ = O.LossModel(om, O.MeanSquaredError).grad_stateful(x, y, mp)
    gl, gg = e.compute_grads(x, y)
    assert abs(gl - float(ol)) <= 2e-5 * abs(float(ol)) + 1e-7
    for k, v in og.items():
        np.testing.assert_allclose(gg[k], v, rtol=2e-4, atol=2e-6, err_msg=k)
    np.testing.assert_allclose(e.predict(x), O.model_forward(om, x, mp)[0], rtol=2e-5, atol=2e-6)
    e.close()


def run_series_trajectory(units, t, f, batch, steps, rate, rtol, atol, loss_rtol, use_graph=True):
    om, sm = build_models([("lstm", units), ("dense", 1, "tanh")])
    x, y = synth_series(batch * steps, t, f, 11)
    e = S.Engine(sm, S.MeanSquaredError, S.Adam(rate), batch, (t, f), (1,), device=0, use_graph=use_graph)
    e.init_params()
    oalg = O.Adam(rate)
    _, mp, op = O.init_params(om, oalg, (batch, t, f))
    e.set_params(mp)
    e.init_opt_params()
    lm = O.LossModel(om, O.MeanSquaredError)
    for s in range(1, steps + 1):
        xb, yb = x[(s - 1) * batch:s * batch], y[(s - 1) * batch:s * batch]
        gl = e.train_step(xb, yb, s)
        mp, op, ol = O.train_step(lm, oalg, xb, yb, mp, op, s)
        assert abs(gl - float(ol)) <= loss_rtol * abs(float(ol)) + 1e-7, (s, gl, float(ol))
    assert_params_close(e.get_model_params(), mp, rtol, atol, f"after {steps} steps")
    e.close()


def test_lstm_small_trajectory_config3_pattern():  # LSTM >> Dense(1, Tanh), MSE, Adam (BASELINE config 3), small
    run_series_trajectory(units=16, t=10, f=1, batch=32, steps=8, rate=1e-3, rtol=2e-3, atol=3e-5, loss_rtol=1e-4)


def test_lstm_config3_full_size_two_steps():  # BASELINE config 3 exactly: LSTM(256), seq len 64, batch 1024, F = 1
    run_series_trajectory(units=256, t=64, f=1, batch=1024, steps=2, rate=1e-3, rtol=5e-3, atol=1e-4, loss_rtol=5e-4)


def test_lstm_resident_epoch_and_graph_determinism():
    _, sm = build_models([("lstm", 16), ("dense", 1, "tanh")])
    x, y = synth_series(32 * 3, 6, 2, 2)
    res = []
    for use_graph in (True, False):
        e = S.Engine(sm, S.MeanSquaredError, S.Adam(1e-3), 32, (6, 2), (1,), device=0, use_graph=use_graph)
        e.init_params()
        e.upload_dataset(x, y)
        it, loss = e.train_epoch(0)
        res.append((it, loss, e.get_model_params()))
        e.close()
    assert res[0][:2] == res[1][:2]
    for k in res[0][2]:
        assert np.array_equal(res[0][2][k], res[1][2][k]), k


# ---------------------------------------------------------------------------------------------------------------
# returnSequence = true (layer/RNN.scala:21-27, 69-71): the sequence output, and LSTMs stacked on it
# ---------------------------------------------------------------------------------------------------------------
SEQ_MODELS = {
    "sequence_into_dense": ([("lstm", 6, True), ("flatten",), ("dense", 1, "tanh")], 5, 3, 7),
    "stacked": ([("lstm", 8, True), ("lstm", 5), ("dense", 1, "tanh")], 6, 2, 9),
    "stacked_wide_input_projection": ([("lstm", 12, True), ("lstm", 4, True), ("lstm", 3), ("dense", 1, "tanh")], 4, 3, 5),
}


@pytest.mark.parametrize("name", list(SEQ_MODELS))
def test_lstm_return_sequence_gradients_match_oracle(name):
    spec, t, f, batch = SEQ_MODELS[name]
    om, sm = build_models(spec)
    x, y = synth_series(batch, t, f, 17)
    e = S.Engine(sm, S.MeanSquaredError, S.Adam(), batch, (t, f), (1,), device=0)
    _, mp, _ = O.init_params(om, O.Adam(), (batch, t, f))
    assert set(e.model_param_paths()) == set(mp.keys())
    e.set_params(mp)
    e.init_opt_params()
    ol, og, _ = O.LossModel(om, O.MeanSquaredError).grad_stateful(x, y, mp)
    gl, gg = e.compute_grads(x, y)
    assert abs(gl - float(ol)) <= 2e-5 * abs(float(ol)) + 1e-7
    for k, v in og.items():
        np.testing.assert_allclose(gg[k], v, rtol=3e-4, atol=3e-6, err_msg=k)
    np.testing.assert_allclose(e.predict(x), O.model_forward(om, x, mp)[0], rtol=2e-5, atol=2e-6)
    e.close()


def test_lstm_return_sequence_output_shape_and_values():
    om, sm = build_models([("lstm", 4, True)])
    x = synth_series(3, 5, 2, 23)[0]
    e = S.Engine(sm, S.MeanSquaredError, S.SGD(), 3, (5, 2), (5, 4), device=0)
    assert tuple(e.output_shape) == (3, 5, 4)                      # (batch, time, units)
    _, mp, _ = O.init_params(om, O.SGD(), (3, 5, 2))
    e.set_params(mp)
    np.testing.assert_allclose(e.predict(x), O.model_forward(om, x, mp)[0], rtol=2e-5, atol=2e-6)
    e.close()


def test_stacked_lstm_trajectory():
    spec = [("lstm", 8, True), ("lstm", 6), ("dense", 1, "tanh")]
    om, sm = build_models(spec)
    batch, t, f, steps, rate = 16, 6, 2, 8, 2e-3
    x, y = synth_series(batch * steps, t, f, 29)
    e = S.Engine(sm, S.MeanSquaredError, S.Adam(rate), batch, (t, f), (1,), device=0)
    oalg = O.Adam(rate)
    _, mp, op = O.init_params(om, oalg, (batch, t, f))
    e.set_params(mp)
    e.init_opt_params()
    lm = O.LossModel(om, O.MeanSquaredError)
    for s_ in range(1, steps + 1):
        xb, yb = x[(s_ - 1) * batch:s_ * batch], y[(s_ - 1) * batch:s_ * batch]
        gl = e.train_step(xb, yb, s_)
        mp, op, ol = O.train_step(lm, oalg, xb, yb, mp, op, s_)
        assert abs(gl - float(ol)) <= 1e-4 * abs(float(ol)) + 1e-7, (s_, gl, float(ol))
    assert_params_close(e.get_model_params(), mp, 2e-3, 3e-5, f"after {steps} steps")
    e.close()


# ---------------------------------------------------------------------------------------------------------------
# RNN(SimpleRNNCell >> Bias >> act) (layer/RNN.scala:96-206).  The state fed back is the pre-bias,
# pre-activation sum (reference quirk); forward golden models/layer/RNNLayerSpec.scala:17-34
# ---------------------------------------------------------------------------------------------------------------
def test_simple_rnn_forward_golden_through_the_engine():
    e = S.Engine(S.SimpleRNN(2, activation=S.Identity), S.MeanSquaredError, S.SGD(), 1, (3, 1), (2,), device=0)
    assert set(e.model_param_paths()) == {"0/kernel_weights", "0/recurrent_weights", "1/weights"}
    e.set_params({"0/kernel_weights": np.array([[-0.5302748, -1.2049599]], f32),
                  "0/recurrent_weights": np.array([[-0.77547526, -0.6313778], [-0.6313778, 0.7754754]], f32),
                  "1/weights": np.zeros(2, f32)})
    y = e.predict(np.array([1, 2, 3], f32).reshape(1, 3, 1))
    np.testing.assert_allclose(y, np.array([[0.222901, -6.019066]], f32), atol=2e-6)
    e.close()
    assert repr(S.SimpleRNN(2)) == "RNN(SimpleRNNCell(2) >> Bias(2) >> Tanh)"


RNN_MODELS = {
    "last_step_tanh": ([("rnn", 8, "tanh"), ("dense", 1, "tanh")], 6, 3, 7),
    "sequence_into_dense_sigmoid": ([("rnn", 5, "sigmoid", True), ("flatten",), ("dense", 1, "identity")], 4, 2, 5),
    "stacked_no_bias_relu": ([("rnn", 6, "tanh", True), ("rnn", 4, "relu", False, False), ("dense", 1, "tanh")], 5, 2, 9),
    "under_an_lstm": ([("rnn", 6, "tanh", True), ("lstm", 4), ("dense", 1, "tanh")], 5, 3, 6),
}


@pytest.mark.parametrize("name", list(RNN_MODELS))
def test_simple_rnn_gradients_match_oracle_bptt(name):
    spec, t, f, batch = RNN_MODELS[name]
    om, sm = build_models(spec)
    x, y = synth_series(batch, t, f, 31)
    e = S.Engine(sm, S.MeanSquaredError, S.Adam(), batch, (t, f), (1,), device=0)
    _, mp, _ = O.init_params(om, O.Adam(), (batch, t, f))
    assert set(e.model_param_paths()) == set(mp.keys())
    e.set_params(mp)
    e.init_opt_params()
    ol, og, _ = O.LossModel(om, O.MeanSquaredError).grad_stateful(x, y, mp)
    gl, gg = e.compute_grads(x, y)
    assert abs(gl - float(ol)) <= 2e-5 * abs(float(ol)) + 1e-7
    for k, v in og.items():
        np.testing.assert_allclose(gg[k], v, rtol=3e-4, atol=3e-6, err_msg=k)
    np.testing.assert_allclose(e.predict(x), O.model_forward(om, x, mp)[0], rtol=2e-5, atol=2e-6)
    e.close()


def test_simple_rnn_trajectory_and_device_init():
    spec = [("rnn", 8, "tanh"), ("dense", 1, "tanh")]
    om, sm = build_models(spec)
    batch, t, f, steps, rate = 16, 6, 2, 8, 2e-3
    x, y = synth_series(batch * steps, t, f, 37)
    e = S.Engine(sm, S.MeanSquaredError, S.Adam(rate), batch, (t, f), (1,), device=0)
    e.init_params()
    oalg = O.Adam(rate)
    _, mp, op = O.init_params(om, oalg, (batch, t, f))
    assert_params_close(e.get_model_params(), mp, 0, 2e-6, "init")  # GlorotUniform bit exact, Orthogonal by host QR
    e.set_params(mp)
    e.init_opt_params()
    lm = O.LossModel(om, O.MeanSquaredError)
    for s_ in range(1, steps + 1):
        xb, yb = x[(s_ - 1) * batch:s_ * batch], y[(s_ - 1) * batch:s_ * batch]
        gl = e.train_step(xb, yb, s_)
        mp, op, ol = O.train_step(lm, oalg, xb, yb, mp, op, s_)
        assert abs(gl - float(ol)) <= 1e-4 * abs(float(ol)) + 1e-7, (s_, gl, float(ol))
    assert_params_close(e.get_model_params(), mp, 2e-3, 3e-5, f"after {steps} steps")
    e.close()


# ---------------------------------------------------------------------------------------------------------------
# stateful = true (layer/RNN.scala:32-35, 66-72): the state tensors are parameters -- they seed the first step, every train step
# stores the last cell state back (Optimizer.scala:141-144), the `loss` closure discards it (:163-167)
# ---------------------------------------------------------------------------------------------------------------
STATEFUL = {
    "lstm": ([("lstm", 6, False, True), ("dense", 1, "tanh")], {"0/cell_state", "0/hidden_state"}),
    "lstm_sequence_stack": ([("lstm", 5, True, True), ("lstm", 4, False, True), ("dense", 1, "tanh")],
                            {"0/cell_state", "0/hidden_state", "1/cell_state", "1/hidden_state"}),
    "simple_rnn": ([("rnn", 6, "tanh", False, True, True), ("dense", 1, "tanh")], {"0/0/state"}),
}


@pytest.mark.parametrize("name", list(STATEFUL))
def test_stateful_rnn_trajectory_carries_the_state(name):
    spec, state_paths = STATEFUL[name]
    om, sm = build_models(spec)
    batch, t, f, steps, rate = 8, 5, 2, 6, 5e-3
    x, y = synth_series(batch * steps, t, f, 41)
    e = S.Engine(sm, S.MeanSquaredError, S.Adam(rate), batch, (t, f), (1,), device=0)
    oalg = O.Adam(rate)
    _, mp, op = O.init_params(om, oalg, (batch, t, f))
    assert set(e.model_param_paths()) == set(mp.keys()) and state_paths <= set(mp.keys())
    e.set_params(mp)
    e.init_opt_params()
    lm = O.LossModel(om, O.MeanSquaredError)
    for s_ in range(1, steps + 1):
        xb, yb = x[(s_ - 1) * batch:s_ * batch], y[(s_ - 1) * batch:s_ * batch]
        gl = e.train_step(xb, yb, s_)
        mp, op, ol = O.train_step(lm, oalg, xb, yb, mp, op, s_)
        assert abs(gl - float(ol)) <= 2e-4 * abs(float(ol)) + 1e-7, (s_, gl, float(ol))
    got = e.get_model_params()
    for k in state_paths:
        assert np.abs(got[k]).max() > 1e-3                        # the state moved away from its Zeros initialiser ...
    assert_params_close(got, mp, 3e-3, 5e-5, f"after {steps} steps")  # ... exactly as the oracle's did
    # the `loss` closure and inference read the state but do not store it
    before = {k: got[k].copy() for k in state_paths}
    l1 = e.eval_loss(x[:batch], y[:batch])
    np.testing.assert_allclose(l1, float(O.eval_loss(lm, x[:batch], y[:batch], mp)), rtol=2e-4)
    e.predict(x[:batch])
    after = e.get_model_params()
    for k in state_paths:
        assert np.array_equal(after[k], before[k]), k
    e.close()


def test_stateful_state_is_reinitialised_by_the_averaging():  # aggregation None => `paramDef.initializer.build` (Optimizer.scala:209)
    _, sm = build_models([("lstm", 4, False, True), ("dense", 1, "tanh")])
    x, y = synth_series(8 * 4, 4, 2, 43)
    trained = (S.Dataset(x, y, partitions=2).train(sm).loss(S.MeanSquaredError).using(S.Adam(1e-3)).batch(8).devices([0, 0])
               .stopAfter(S.epochs(2)).quiet().run())
    assert np.all(trained.params["0/cell_state"] == 0) and np.all(trained.params["0/hidden_state"] == 0)
    single = (S.Dataset(x, y).train(sm).loss(S.MeanSquaredError).using(S.Adam(1e-3)).batch(8).stopAfter(S.epochs(2)).quiet().run())
    assert np.abs(single.params["0/hidden_state"]).max() > 0      # one partition: nothing is reduced, the state lives on
