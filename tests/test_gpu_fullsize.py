"""BASELINE.json configs 3, 4 and 5 at their FULL sizes through the C ABI against the oracle (these were only
covered at "pattern" size).  Two train steps each: the first pins forward + gradients + the optimizer from the shared initial
state, the second pins that the updated state feeds the next step; cost is dominated by the NumPy oracle (seconds).

Tolerances.  fp32 mode (FFMA contractions): everything is judged against the oracle run in FLOAT64, with the float32 oracle's own
distance from that truth as the yardstick.  Loss: rtol 1e-4 on the first step, then within 4x the float32 oracle's drift.
Gradients: the engine's relative L2 error per tensor must stay within 8x the float32 oracle's (+ 2e-4: sequential fp32 column sums
of 8192 rows against NumPy's pairwise ones): reductions
of up to 8192 x 4096 terms run in a different order than NumPy's, a few hundred of the 33 M ReLU pre-activations per layer sit
within fp32 rounding of zero (their masks flip and each flip changes single terms of the gradient sums by 100 %), and the
BatchNorm backward subtracts nearly equal sums -- measured 5e-3 of max on a bias gradient of config 4 and 1e-2 relative L2 on the
last conv's filter gradient of config 5 between two float32 implementations; parameters after Adam steps: an Adam step moves every weight by about `rate`
whatever the gradient's magnitude, so a gradient whose sign is within rounding noise of zero moves its weight the other way --
the parameters are therefore checked by max |err| <= 2 * steps * rate and by the FRACTION of elements off by more than
(rtol 2e-3, atol 3e-5), which must stay below twice the float32 oracle's own fraction + 3e-3
(measured: 6 of the 4096 elements of a config-4 bias).  TF32 mode: loss rtol 3e-3, gradient direction cosine > 0.995 and norm within
5 % (ReLU masks flip for pre-activations within the tf32 error of zero, tests/test_gpu_tf32.py), parameters max |err| <= 3 * steps
* rate.  3xTF32 mode: the fp32-mode bounds.  LSTM (64 recurrent steps): per-tensor |err| <= tol * max|ref|, and tol * max(max|ref|, sqrt(loss)) for the
output layer's (Dense(1) >> Bias) gradients, which are cancelling sums of 1024 terms of size ~|p - y| ~ sqrt(loss), so its natural scale is sqrt(loss), not
its own (small) value; 2e-3 per tensor in 3xTF32.  Plain TF32 through 64 recurrent steps pins the direction of the whole gradient
(cosine > 0.98 -- measured 0.991 --, norm ratio in (0.8, 1.05) -- measured 0.86: kind::tf32 truncates, so 64 steps of BPTT shrink the gradient) and a coarse 0.3 x largest-entry bound per tensor (the test explains why)."""
import numpy as np
import pytest

import oracle as O
import scanet3_b200 as S
from helpers import build_models, synth_classification

pytestmark = pytest.mark.gpu
f32 = np.float32


def rel_to_max(got, ref):
    return float(np.max(np.abs(got.astype(np.float64) - np.asarray(ref, np.float64))) / (np.max(np.abs(ref)) + 1e-30))


def rel_l2(a, t):
    a, t = np.asarray(a, np.float64).ravel(), np.asarray(t, np.float64).ravel()
    return float(np.linalg.norm(a - t) / (np.linalg.norm(t) + 1e-300))


def check_two_steps(om, sm, oloss, sloss, x, y, batch, in_shape, out_shape, precision, rate=0.001):
    """The oracle runs twice: in float64 (the truth) and in float32 (how far a correct float32 implementation may drift from it --
    on config 5 the float32 oracle's second-step loss is 9.5e-4 off the float64 one and 32 % of conv1's weights sit outside
    (rtol 2e-3, atol 3e-5) after two Adam steps, while the engine's loss agrees with the float64 run to 1e-6)."""
    steps = 2
    e = S.Engine(sm, sloss, S.Adam(rate), batch, in_shape, out_shape, device=0, precision=precision)
    oalg = O.Adam(rate)
    _, mp, op = O.init_params(om, oalg, (batch,) + tuple(in_shape))
    e.init_params()
    lm = O.LossModel(om, oloss)
    fp32_class = precision != S.PREC_TF32
    f64 = lambda d: {k: np.asarray(v, np.float64) for k, v in d.items()}
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    # ---- gradients from the shared initial state ----
    gl, gg = e.compute_grads(x[:batch], y[:batch])
    ol, og, _ = lm.grad_stateful(x[:batch], y[:batch], mp)
    tl, tg, _ = lm.grad_stateful(x64[:batch], y64[:batch], f64(mp))
    assert abs(gl - float(tl)) <= (1e-4 if fp32_class else 3e-3) * abs(float(tl)), (gl, float(tl))
    for k in tg:
        if fp32_class:
            err_engine, err_oracle = rel_l2(gg[k], tg[k]), rel_l2(og[k], tg[k])
            assert err_engine <= 8 * err_oracle + 2e-4, (k, err_engine, err_oracle)
        else:
            a, b = gg[k].astype(np.float64).ravel(), np.asarray(tg[k], np.float64).ravel()
            cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))
            assert cos > 0.995, (k, cos)
            assert abs(np.linalg.norm(a) / (np.linalg.norm(b) + 1e-300) - 1) < 0.05, k
    # ---- two optimizer steps ----
    mp32, op32, mp64, op64 = mp, op, f64(mp), f64(op)
    for t in range(1, steps + 1):
        sl = slice((t - 1) * batch, t * batch)
        loss = e.train_step(x[sl], y[sl], t)
        mp32, op32, l32 = O.train_step(lm, oalg, x[sl], y[sl], mp32, op32, t)
        mp64, op64, l64 = O.train_step(lm, oalg, x64[sl], y64[sl], mp64, op64, t)
        drift = abs(float(l32) - float(l64))
        bound = max(4 * drift, 1e-4 * abs(float(l64))) if fp32_class else 3e-3 * abs(float(l64))
        assert abs(loss - float(l64)) <= bound, (t, loss, float(l64), float(l32))
    got = e.get_model_params()
    for k, v in mp64.items():
        err = np.abs(got[k].astype(np.float64) - v)
        assert float(err.max()) <= (2 if fp32_class else 3) * steps * rate + 1e-6, (k, float(err.max()))
        if fp32_class:
            off = float((err > (3e-5 + 2e-3 * np.abs(v))).mean())
            off_oracle = float((np.abs(np.asarray(mp32[k], np.float64) - v) > (3e-5 + 2e-3 * np.abs(v))).mean())
            assert off <= 2 * off_oracle + 3e-3, (k, off, off_oracle)
    assert e.launch_count() > 0
    e.close()


@pytest.mark.parametrize("precision", [S.PREC_FP32, S.PREC_TF32, S.PREC_3XTF32], ids=["fp32", "tf32", "3xtf32"])
def test_config4_wide_mlp_full_size(precision):
    """BASELINE configs[3]: Dense(4096, ReLU) x 4 >> Dense(10, Softmax), batch 8192, 784 inputs (53.6 M parameters)"""
    spec = [("dense", 4096, "relu")] * 4 + [("dense", 10, "softmax")]
    om, sm = build_models(spec)
    batch = 8192
    x, y = synth_classification(batch * 2, (784,), 10, 1238)
    x = (x - 0.5).astype(f32)
    check_two_steps(om, sm, O.CategoricalCrossentropy, S.CategoricalCrossentropy, x, y, batch, (784,), (10,), precision)


CIFAR = lambda mode: [("conv", 64, "relu", (3, 3), (1, 1), "Valid", False), ("bn", 0.99, mode), ("pool", (2, 2), (2, 2), "Valid"),
                      ("conv", 128, "relu", (3, 3), (1, 1), "Valid", False), ("bn", 0.99, mode), ("pool", (2, 2), (2, 2), "Valid"),
                      ("conv", 256, "relu", (3, 3), (1, 1), "Valid", False), ("bn", 0.99, mode), ("pool", (2, 2), (2, 2), "Valid"),
                      ("flatten",), ("dense", 10, "softmax")]


@pytest.mark.parametrize("mode", ["reference", "correct"])
@pytest.mark.parametrize("precision", [S.PREC_FP32, S.PREC_TF32], ids=["fp32", "tf32"])
def test_config5_cifar_batchnorm_cnn_full_size(precision, mode):
    """BASELINE configs[4] (SURVEY 8d topology): Conv2D 64/128/256 + fused rank-4 BatchNorm + Pool2D(s2) + Dense, batch 256,
    32x32x3, both `bn_mode`s (the reference's fused-output wiring quirk, SURVEY App. B-Q5, and the TF op as intended)"""
    om, sm = build_models(CIFAR(mode))
    batch = 256
    x, y = synth_classification(batch * 2, (32, 32, 3), 10, 1239)
    check_two_steps(om, sm, O.CategoricalCrossentropy, S.CategoricalCrossentropy, x, y, batch, (32, 32, 3), (10,), precision)


def _tf32_trunc(a):
    """what kind::tf32 keeps of an fp32 operand: the low 13 mantissa bits are ignored"""
    return (np.ascontiguousarray(a, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def cfg5_tf32_emulation_report(batch, setattr_):
    """engine (TF32 mode) vs the oracle whose Conv2D contractions truncate their operands to tf32 where the engine runs tcgen05 kernels"""
    import oracle.layers as Lm
    om, sm = build_models(CIFAR("correct"))
    x, y = synth_classification(batch, (32, 32, 3), 10, 1239)
    e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(1e-3), batch, (32, 32, 3), (10,), device=0, precision=S.PREC_TF32)
    e.init_params()
    gl, gg = e.compute_grads(x, y)
    e.close()
    f64 = np.float64
    conv2d0, wgrad0, dgrad0 = Lm.conv2d, Lm.conv2d_wgrad, Lm.conv2d_dgrad
    on_tc = lambda cin: cin % 32 == 0
    t = lambda a_: _tf32_trunc(a_).astype(f64)
    setattr_(Lm, "conv2d", lambda x_, w_, *a, **k: conv2d0(t(x_), t(w_), *a, **k).astype(f32) if on_tc(x_.shape[-1]) else conv2d0(x_, w_, *a, **k))
    setattr_(Lm, "conv2d_wgrad", lambda x_, ws, g_, *a, **k: wgrad0(t(x_), ws, t(g_), *a, **k).astype(f32) if on_tc(x_.shape[-1]) else wgrad0(x_, ws, g_, *a, **k))
    setattr_(Lm, "conv2d_dgrad", lambda xs, w_, g_, *a, **k: dgrad0(xs, t(w_), t(g_), *a, **k).astype(f32) if on_tc(xs[-1]) else dgrad0(xs, w_, g_, *a, **k))
    _, mp, _ = O.init_params(om, O.Adam(1e-3), (batch, 32, 32, 3))
    ol, og, _ = O.LossModel(om, O.CategoricalCrossentropy).grad_stateful(x, y, mp)
    report = {"loss": (gl, float(ol))}
    for k in og:
        a_, b_ = gg[k].astype(f64).ravel(), np.asarray(og[k], f64).ravel()
        cos = float(a_ @ b_ / (np.linalg.norm(a_) * np.linalg.norm(b_) + 1e-300))
        rmax = float(np.max(np.abs(a_ - b_)) / (np.max(np.abs(b_)) + 1e-300))
        report[k] = (round(cos, 7), round(rmax, 6))
    return report


def test_config5_tf32_gradients_match_an_oracle_with_truncated_operands(monkeypatch):
    """BASELINE configs[4] at full size, TF32 mode, pinned through the contractions themselves: the oracle's Conv2D fprop / dgrad /
    wgrad truncate their operands to tf32 wherever the engine runs the tcgen05 kernels (Cin % 32 == 0: conv2, conv3; conv1 and the
    1024 -> 10 head stay fp32 on both sides) and accumulate in float64.  Loss 1e-4 (measured 4.6e-5); everything behind the last
    BatchNorm (its gamma / beta, the head) within 2e-3 of the largest entry (measured <= 5.5e-4); every tensor's direction cosine
    > 0.999 (measured >= 0.99944; the plain oracle allows 0.995).  In front of a BatchNorm no oracle pins single entries tighter: its
    backward subtracts nearly equal sums, which amplifies accumulation-order differences to per-cent level -- the same 1e-2 relative
    L2 on conv3's filter gradient separates two float32 implementations (module docstring)."""
    report = cfg5_tf32_emulation_report(256, monkeypatch.setattr)
    gl, ol = report.pop("loss")
    ok = abs(gl - ol) <= 1e-4 * abs(ol) and all(c > 0.999 for c, _ in report.values())
    ok = ok and all(report[k][1] < 2e-3 for k in ("10/gamma", "10/beta", "13/weights", "14/weights"))
    assert ok, "\n" + "\n".join(f"{k}: {v}" for k, v in report.items())


@pytest.mark.parametrize("precision", [S.PREC_TF32, S.PREC_3XTF32], ids=["tf32", "3xtf32"])
def test_config3_lstm_full_size_tensor_core_modes_against_the_oracle(precision):
    """BASELINE configs[2]: LSTM(256) >> Dense(1, Tanh), MSE, seq len 64, batch 1024 -- the tensor-core modes against the ORACLE
    (round 1 compared them with the repo's own fp32 engine only)"""
    om, sm = build_models([("lstm", 256), ("dense", 1, "tanh")])
    batch, t = 1024, 64
    rng = np.random.Generator(np.random.PCG64(1237))
    x = rng.uniform(0, 1, size=(batch * 2, t, 1)).astype(f32)
    y = rng.uniform(0, 1, size=(batch * 2, 1)).astype(f32)
    e = S.Engine(sm, S.MeanSquaredError, S.Adam(1e-3), batch, (t, 1), (1,), device=0, precision=precision)
    e.init_params()
    _, mp, op = O.init_params(om, O.Adam(1e-3), (batch, t, 1))
    for k, v in e.get_model_params().items():  # the Orthogonal recurrent init runs through the host QR: same values on both sides
        np.testing.assert_allclose(v, mp[k], rtol=0, atol=1e-6, err_msg=k)
    lm = O.LossModel(om, O.MeanSquaredError)
    gl, gg = e.compute_grads(x[:batch], y[:batch])
    ol, og, _ = lm.grad_stateful(x[:batch], y[:batch], mp)
    tol_l, tol_g = (4e-3, 0.3) if precision == S.PREC_TF32 else (5e-4, 2e-3)  # (fp32 run of this config: loss_rtol 5e-4)
    assert abs(gl - float(ol)) <= tol_l * abs(float(ol)), (gl, float(ol))
    gmax = max(float(np.max(np.abs(v))) for v in og.values())
    if precision == S.PREC_TF32:
        # 64 recurrent steps in plain TF32: kind::tf32 TRUNCATES its operands, so the error of h drifts instead of averaging out, and
        # the small gradient tensors (input kernels, biases: cancelling sums over 65 536 (row, step) terms) carry it in full --
        # measured 2e-2 .. 1.5e-1 of the model's largest gradient entry on single tensors.  What TF32 does pin: the direction of the
        # whole gradient (cosine over all tensors) and a coarse per-tensor bound; the 3xTF32 run below is the precise one.
        a = np.concatenate([gg[k].astype(np.float64).ravel() for k in og])
        b = np.concatenate([np.asarray(og[k], np.float64).ravel() for k in og])
        cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))
        nr = float(np.linalg.norm(a) / np.linalg.norm(b))
        assert cos > 0.98 and 0.8 < nr < 1.05, (cos, nr)  # measured: cosine 0.9912, norm ratio 0.860 (truncation shrinks every product)
    bad = []
    for k in og:
        scale = gmax if precision == S.PREC_TF32 else float(np.max(np.abs(og[k])))
        if k.startswith(("1/", "2/")):  # the output layer Dense(1) >> Bias: cancelling sums over the batch, natural scale sqrt(loss)
            scale = max(scale, float(np.sqrt(ol)))
        err = float(np.max(np.abs(gg[k].astype(np.float64) - np.asarray(og[k], np.float64))))
        if not err < tol_g * scale:
            bad.append(("grad", k, err, scale))
    oalg = O.Adam(1e-3)
    for s in (1, 2):
        xb, yb = x[(s - 1) * batch:s * batch], y[(s - 1) * batch:s * batch]
        loss = e.train_step(xb, yb, s)
        mp, op, olv = O.train_step(lm, oalg, xb, yb, mp, op, s)
        # TF32, second step: Adam moved every weight by ~rate along the SIGN of its gradient; the many near-zero gradient entries whose
        # sign the tf32 noise flips move the 256-term output sum coherently -- measured 6.3 % on the loss (fp32 / 3xTF32: 5e-4)
        tol_s = 0.1 if (precision == S.PREC_TF32 and s > 1) else tol_l
        if not abs(loss - float(olv)) <= tol_s * abs(float(olv)):
            bad.append(("loss", s, loss, float(olv)))
    got = e.get_model_params()
    for k, v in mp.items():
        d = float(np.max(np.abs(got[k] - v)))
        if not d <= 3 * 2 * 1e-3:
            bad.append(("param", k, d))
    e.close()
    assert not bad, bad  # every violated bound at once


def _tf32_rn(a):
    """cvt.rna.tf32.f32: round to the nearest tf32 value, ties away from zero (sign-magnitude: add half an ulp, drop 13 bits)"""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def test_config3_lstm_tf32_matches_an_oracle_that_rounds_where_the_engine_rounds(monkeypatch):
    """BASELINE configs[2] at full size in TF32 mode, pinned TIGHTLY.  In TF32 mode the engine stores h_t, dZ_t and the packed
    recurrent weights rounded to the nearest tf32 value (csrc/lstm.cu tf32_rn), so the operands of its tcgen05 products are exact
    tf32 numbers and the products exact in fp32.  The oracle cell below rounds at the same three places and accumulates in float64:
    what is left is accumulation order, the approximate exp of the TF32-mode gate kernels (~1e-6) and the rare rounding decision
    either of them flips.  Loss 1e-5 (measured 4e-6), every gradient 1e-3 of the model's largest entry (measured 3.8e-4, a gate bias) --
    against the direction-only bounds (cosine 0.98) the plain oracle allows in the test above."""
    import oracle.layers as Lm
    om, sm = build_models([("lstm", 256), ("dense", 1, "tanh")])
    batch, t = 1024, 64
    rng = np.random.Generator(np.random.PCG64(1237))
    x = rng.uniform(0, 1, size=(batch, t, 1)).astype(f32)
    y = rng.uniform(0, 1, size=(batch, 1)).astype(f32)
    e = S.Engine(sm, S.MeanSquaredError, S.Adam(1e-3), batch, (t, 1), (1,), device=0, precision=S.PREC_TF32)
    e.init_params()
    gl, gg = e.compute_grads(x, y)
    e.close()
    f64 = np.float64

    def step(self, x_, c_prev, h_prev, p):
        zs, acts = {}, {}
        for gname in self.GATES:
            z = (x_ @ p[f"{gname}/0/kernel_weights"]) + (h_prev.astype(f64) @ _tf32_rn(p[f"{gname}/0/recurrent_weights"]).astype(f64)).astype(f32)
            if self.bias:
                z = z + p[f"{gname}/1/weights"]
            zs[gname] = z
            acts[gname] = self._act(gname).fwd(z)
        f, i, g, o = (acts[n] for n in self.GATES)
        c = c_prev * f + i * g
        ac = self.activation.fwd(c)
        h = _tf32_rn(o * ac)
        return h, c, (x_, c_prev, h_prev, zs, acts, c, ac)

    def step_bwd(self, dh, dc_next, cache, p):
        x_, c_prev, h_prev, zs, acts, c, ac = cache
        f, i, g, o = (acts[n] for n in self.GATES)
        do = dh * ac
        dc = self.activation.bwd(dh * o, c, ac) + dc_next
        dacts = {"forget": dc * c_prev, "input": dc * g, "gate": dc * i, "output": do}
        dc_prev = dc * f
        dh_prev = np.zeros(h_prev.shape, f64)
        grads = {}
        for gname in self.GATES:
            dz = _tf32_rn(self._act(gname).bwd(dacts[gname], zs[gname], acts[gname]))
            grads[f"{gname}/0/kernel_weights"] = (x_.astype(f64).T @ dz.astype(f64)).astype(f32)
            grads[f"{gname}/0/recurrent_weights"] = (h_prev.astype(f64).T @ dz.astype(f64)).astype(f32)
            if self.bias:
                grads[f"{gname}/1/weights"] = np.sum(dz, axis=0, dtype=f64).astype(f32)
            dh_prev = dh_prev + dz.astype(f64) @ _tf32_rn(p[f"{gname}/0/recurrent_weights"]).astype(f64).T
        self._last_dx = np.zeros_like(x_)
        return dh_prev.astype(f32), dc_prev, grads

    monkeypatch.setattr(Lm.LSTMCell, "step", step)
    monkeypatch.setattr(Lm.LSTMCell, "step_bwd", step_bwd)
    _, mp, _ = O.init_params(om, O.Adam(1e-3), (batch, t, 1))
    ol, og, _ = O.LossModel(om, O.MeanSquaredError).grad_stateful(x, y, mp)
    assert abs(gl - float(ol)) <= 1e-5 * abs(float(ol)), (gl, float(ol))
    gmax = max(float(np.max(np.abs(v))) for v in og.values())
    worst = {k: float(np.max(np.abs(gg[k].astype(f64) - np.asarray(og[k], f64)))) / gmax for k in og}
    assert max(worst.values()) < 1e-3, worst


def test_maximize_walks_up_the_gradient():
    """Optimizer.maximize (Optimizer.scala:186,326-338): `w + delta` -- the same trajectory as the oracle's `minimizing=False`"""
    om, sm = build_models([("dense", 8, "sigmoid"), ("dense", 4, "softmax")])
    x, y = synth_classification(16 * 4, (12,), 4, 2)
    e = S.Engine(sm, S.CategoricalCrossentropy, S.SGD(0.1), 16, (12,), (4,), device=0, minimizing=False)
    e.init_params()
    oalg = O.SGD(0.1)
    _, mp, op = O.init_params(om, oalg, (16, 12))
    lm = O.LossModel(om, O.CategoricalCrossentropy)
    losses = []
    for t in range(1, 5):
        xb, yb = x[(t - 1) * 16:t * 16], y[(t - 1) * 16:t * 16]
        loss = e.train_step(xb, yb, t)
        mp, op, ol = O.train_step(lm, oalg, xb, yb, mp, op, t, minimizing=False)
        assert abs(loss - float(ol)) <= 2e-5 * abs(float(ol)) + 1e-7
        losses.append(loss)
    got = e.get_model_params()
    for k, v in mp.items():
        np.testing.assert_allclose(got[k], v, rtol=1e-3, atol=2e-5, err_msg=k)
    e.close()
