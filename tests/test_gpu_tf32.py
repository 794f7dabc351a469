"""Tensor-core (tcgen05 kind::tf32) contractions against the oracle and against the engine's own fp32 path.

Tolerance: kind::tf32 keeps 10 explicit mantissa bits of each operand (the low 13 bits of the fp32 word are ignored),
products are exact and accumulation is fp32.  A length-K dot product of O(1) terms therefore carries a relative error of
about 2^-10 / sqrt(K) .. 2^-10 per operand; tensors are compared with |err| <= 4e-3 * max|ref| (conv outputs and
gradients), the loss with rtol 2e-3, and parameters after n Adam steps with atol = 3 * n * rate (a sign flip of a
near-zero gradient moves a weight by one full Adam step)."""
import numpy as np
import pytest

import oracle as O
import scanet3_b200 as S
from helpers import OALG, OLOSS, SALG, SLOSS, build_models, synth_classification

pytestmark = pytest.mark.gpu
f32 = np.float32


def rel_to_max(got, ref):
    return float(np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64))) / (np.max(np.abs(ref)) + 1e-30))


def engines(spec, loss, batch, in_shape, classes, alg=None):
    om, sm = build_models(spec)
    alg = alg or S.Adam(0.001)
    e32 = S.Engine(sm, SLOSS[loss], alg, batch, in_shape, (classes,), device=0, precision=S.PREC_FP32)
    etc = S.Engine(sm, SLOSS[loss], alg, batch, in_shape, (classes,), device=0, precision=S.PREC_TF32)
    e32.init_params()
    etc.init_params()
    return om, e32, etc


CONV_STACKS = [
    # (name, spec, batch, in_shape): the second conv has Cin = 32 and Cout = 64 -> the tcgen05 path
    ("lenet28", [("conv", 32, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"),
                 ("conv", 64, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"), ("flatten",),
                 ("dense", 10, "softmax")], 100, (28, 28, 1)),
    ("tiny", [("conv", 32, "relu", (3, 3), (1, 1), "Valid", False), ("conv", 64, "relu", (3, 3), (1, 1), "Valid", False),
              ("flatten",), ("dense", 10, "softmax")], 3, (7, 9, 1)),
    ("k2x2_no_relu", [("conv", 32, "sigmoid", (2, 2), (1, 1), "Valid", False),
                      ("conv", 64, "identity", (2, 2), (1, 1), "Valid", False), ("flatten",), ("dense", 10, "softmax")], 5,
     (6, 11, 2)),
    ("cout128", [("conv", 64, "relu", (3, 3), (1, 1), "Valid", False), ("conv", 128, "relu", (2, 2), (1, 1), "Valid", False),
                 ("pool", (2, 2), (2, 2), "Valid"), ("flatten",), ("dense", 10, "softmax")], 4, (13, 13, 3)),
    # smooth activations: the same geometries with no mask flips => the kernels themselves are pinned at 4e-3
    ("lenet28_sigmoid", [("conv", 32, "sigmoid", (3, 3), (1, 1), "Valid", False), ("conv", 64, "sigmoid", (3, 3), (1, 1), "Valid", False),
                         ("flatten",), ("dense", 10, "softmax")], 100, (27, 27, 1)),
    ("tiny_linear", [("conv", 32, "identity", (3, 3), (1, 1), "Valid", False), ("conv", 64, "identity", (3, 3), (1, 1), "Valid", False),
                     ("flatten",), ("dense", 10, "softmax")], 3, (7, 9, 1)),
    ("cout128_tanh", [("conv", 64, "tanh", (3, 3), (1, 1), "Valid", False), ("conv", 128, "tanh", (2, 2), (1, 1), "Valid", False),
                      ("flatten",), ("dense", 10, "softmax")], 4, (13, 13, 3)),
    # BASELINE config 5 geometry (Cin 64 -> Cout 128 -> Cout 256, 3x3): the filter no longer fits shared memory (per-tap kernel),
    # taps*Cin > 512 TMEM columns and Cout = 256 (wgrad in tap groups x channel halves), tiny 6x6 -> 4x4 grids
    ("cifar_tail_tanh", [("conv", 64, "tanh", (3, 3), (1, 1), "Valid", False), ("conv", 128, "tanh", (3, 3), (1, 1), "Valid", False),
                         ("conv", 256, "tanh", (3, 3), (1, 1), "Valid", False), ("flatten",), ("dense", 10, "softmax")], 6, (10, 10, 3)),
    # the same with a batch that is not a multiple of the images per tile (whole-image tiles: 3 images of 6x6, 8 of 4x4): the last
    # tile's TMA box runs past the batch (zero fill) and its surplus rows must not be stored
    ("cifar_tail_tanh_b7", [("conv", 64, "tanh", (3, 3), (1, 1), "Valid", False), ("conv", 128, "tanh", (3, 3), (1, 1), "Valid", False),
                            ("conv", 256, "tanh", (3, 3), (1, 1), "Valid", False), ("flatten",), ("dense", 10, "softmax")], 7, (10, 10, 3)),
    ("same_pad_tanh", [("conv", 32, "tanh", (3, 3), (1, 1), "Valid", False), ("conv", 64, "tanh", (3, 3), (1, 1), "Same", False),
                       ("flatten",), ("dense", 10, "softmax")], 4, (9, 9, 2)),
    ("explicit_pad_tanh", [("conv", 32, "tanh", (3, 3), (1, 1), "Valid", False), ("conv", 64, "tanh", (3, 3), (1, 1), ((1, 2), (0, 1)), False),
                           ("flatten",), ("dense", 10, "softmax")], 4, (9, 9, 2)),
    ("ragged_1x3_tanh", [("conv", 32, "tanh", (2, 2), (1, 1), "Valid", False), ("conv", 64, "tanh", (1, 3), (1, 1), "Valid", False),
                         ("flatten",), ("dense", 10, "softmax")], 7, (5, 37, 2)),
]


@pytest.mark.parametrize("wgrad", ["halo", "per_tap"])  # filter-gradient kernel: one X box per chunk (default) / one box per tap
@pytest.mark.parametrize("name,spec,batch,in_shape", CONV_STACKS, ids=[c[0] for c in CONV_STACKS])
def test_tf32_conv_forward_and_gradients(name, spec, batch, in_shape, wgrad, monkeypatch):
    monkeypatch.setenv("SB_TC_HALO_WGRAD", "1" if wgrad == "halo" else "0")
    om, e32, etc = engines(spec, "cce", batch, in_shape, 10)
    x, y = synth_classification(batch, in_shape, 10, 21)
    p32, ptc = e32.predict(x), etc.predict(x)
    assert rel_to_max(ptc, p32) < 4e-3
    l32, g32 = e32.compute_grads(x, y)
    ltc, gtc = etc.compute_grads(x, y)
    assert abs(ltc - l32) <= 2e-3 * abs(l32)
    _, mp, _ = O.init_params(om, O.Adam(0.001), (batch,) + tuple(in_shape))
    ol, og, _ = O.LossModel(om, O.CategoricalCrossentropy).grad_stateful(x, y, mp)
    assert abs(ltc - float(ol)) <= 2e-3 * abs(float(ol))
    relu_net = any(it[0] == "conv" and it[2] == "relu" for it in spec)
    for k in g32:
        if not relu_net:
            assert rel_to_max(gtc[k], g32[k]) < 4e-3, (k, "vs fp32 engine")
            assert rel_to_max(gtc[k], og[k]) < 4e-3, (k, "vs oracle")
        else:
            # ReLU's derivative is discontinuous: a pre-activation within the tf32 error of zero (a fraction f ~ 4e-4 of
            # them at random init) flips its mask and changes that term of the gradient sum by 100%, i.e. a relative
            # L2 error of ~sqrt(f) = 2% that no contraction precision short of fp32 removes.  Smooth-activation variants
            # of the same geometries (above) pin the kernels at 4e-3; here the direction and norm are checked.
            a, b = gtc[k].astype(np.float64).ravel(), og[k].astype(np.float64).ravel()
            cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))
            assert cos > 0.995, (k, cos)
            assert abs(np.linalg.norm(a) / np.linalg.norm(b) - 1) < 0.05, k
    assert etc.launch_count() > 0
    prof_names = []
    etc.profile_enable(True)
    etc.compute_grads(x, y)
    prof_names = [r["name"] for r in etc.profile()]
    etc.profile_enable(False)
    assert any("tcgen05" in n for n in prof_names), prof_names  # the tensor-core kernels really ran
    e32.close()
    etc.close()


def _tf32_trunc(a):
    """what kind::tf32 keeps of an fp32 operand: the low 13 mantissa bits are ignored (truncation, not rounding)"""
    return (np.ascontiguousarray(a, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def test_tf32_lenet_gradients_match_an_oracle_with_truncated_operands(monkeypatch):
    """BASELINE config 2 at full size, TF32 mode, pinned TIGHTLY: the oracle is run with the operands of exactly the contractions the
    engine puts on tcgen05 (conv2 fprop / dgrad / wgrad: Cin = 32; the head 30976 -> 10 forward and backward) truncated to tf32 and
    accumulated in float64.  What is left between the two is accumulation order (and the handful of ReLU masks it flips), so the
    loss must agree to 2e-6 and every gradient to 5e-4 of its largest entry -- measured 4e-6 (head bias), 2.8e-4 (head weights),
    6.7e-5 (conv2), 4.1e-5 (conv1) -- not the 4e-3 / cosine 0.995 bounds the plain-fp32 oracle allows."""
    import oracle.layers as Lm
    spec = CONV_STACKS[0][1]
    om, sm = build_models(spec)
    batch = 100
    x, y = synth_classification(batch, (28, 28, 1), 10, 21)
    e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), batch, (28, 28, 1), (10,), device=0, precision=S.PREC_TF32)
    e.init_params()
    gl, gg = e.compute_grads(x, y)
    e.close()
    f64 = np.float64
    conv2d0, wgrad0, dgrad0 = Lm.conv2d, Lm.conv2d_wgrad, Lm.conv2d_dgrad
    on_tc = lambda cin: cin % 32 == 0  # the engine's rule for the tcgen05 conv kernels (tc.cu conv_fits)
    monkeypatch.setattr(Lm, "conv2d", lambda x_, w_, *a, **k: conv2d0(_tf32_trunc(x_).astype(f64), _tf32_trunc(w_).astype(f64), *a, **k).astype(np.float32)
                        if on_tc(x_.shape[-1]) else conv2d0(x_, w_, *a, **k))
    monkeypatch.setattr(Lm, "conv2d_wgrad", lambda x_, ws, g_, *a, **k: wgrad0(_tf32_trunc(x_).astype(f64), ws, _tf32_trunc(g_).astype(f64), *a, **k).astype(np.float32)
                        if on_tc(x_.shape[-1]) else wgrad0(x_, ws, g_, *a, **k))
    monkeypatch.setattr(Lm, "conv2d_dgrad", lambda xs, w_, g_, *a, **k: dgrad0(xs, _tf32_trunc(w_).astype(f64), _tf32_trunc(g_).astype(f64), *a, **k).astype(np.float32)
                        if on_tc(xs[-1]) else dgrad0(xs, w_, g_, *a, **k))
    t = lambda a_: _tf32_trunc(a_).astype(f64)
    monkeypatch.setattr(Lm.DenseKernel, "forward", lambda self, x_, p, training: ((t(x_) @ t(p["weights"])).astype(np.float32), x_, {}))
    monkeypatch.setattr(Lm.DenseKernel, "backward", lambda self, g_, x_, p, need_dx: (
        (t(g_) @ t(p["weights"]).T).astype(np.float32) if need_dx else None, {"weights": (t(x_).T @ t(g_)).astype(np.float32)}))
    _, mp, _ = O.init_params(om, O.Adam(0.001), (batch, 28, 28, 1))
    ol, og, _ = O.LossModel(om, O.CategoricalCrossentropy).grad_stateful(x, y, mp)
    assert abs(gl - float(ol)) <= 2e-6 * abs(float(ol)), (gl, float(ol))
    worst = {k: rel_to_max(gg[k], og[k]) for k in og}
    assert max(worst.values()) < 5e-4, worst


def test_tf32_lenet_trajectory_config2():
    spec = CONV_STACKS[0][1]
    om, sm = build_models(spec)
    batch, steps, rate = 100, 6, 0.001
    x, y = synth_classification(batch * steps, (28, 28, 1), 10, 3)
    e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(rate), batch, (28, 28, 1), (10,), device=0, precision=S.PREC_TF32)
    e.init_params()
    oalg = O.Adam(rate)
    _, mp, op = O.init_params(om, oalg, (batch, 28, 28, 1))
    lm = O.LossModel(om, O.CategoricalCrossentropy)
    for t in range(1, steps + 1):
        xb, yb = x[(t - 1) * batch:t * batch], y[(t - 1) * batch:t * batch]
        gl = e.train_step(xb, yb, t)
        mp, op, ol = O.train_step(lm, oalg, xb, yb, mp, op, t)
        assert abs(gl - float(ol)) <= 2e-3 * abs(float(ol)), (t, gl, float(ol))
    got = e.get_model_params()
    for k, v in mp.items():
        err = np.abs(got[k] - v)
        assert float(np.max(err)) <= 3 * steps * rate, k
        assert float(np.mean(err)) <= 0.05 * steps * rate, k  # the bulk of the weights agrees far better
    e.close()


def test_tf32_resident_epoch_runs_in_a_cuda_graph_and_matches_eager():
    spec = CONV_STACKS[0][1]
    _, sm = build_models(spec)
    x, y = synth_classification(100 * 4, (28, 28, 1), 10, 9)
    res = []
    for use_graph in (True, False):
        e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), 100, (28, 28, 1), (10,), device=0, precision=S.PREC_TF32,
                     use_graph=use_graph)
        e.init_params()
        e.upload_dataset(x, y)
        it, loss = e.train_epoch(0)
        it2, loss2 = e.train_epoch(it)
        res.append((it, loss, it2, loss2, e.get_model_params()))
        e.close()
    assert res[0][:4] == res[1][:4]  # deterministic: fixed split-K / reduction orders
    for k in res[0][4]:
        assert np.array_equal(res[0][4][k], res[1][4][k]), k


def test_tf32_head_only_model_closing_loss_sees_the_updated_weights():
    # the classifier head is the FIRST kernel of the epoch's closing loss, launched right behind the optimizer: its W slab must not be
    # staged ahead of the programmatic-dependency wait there (common.cuh prewait_ok); eager launches and graph replays must agree
    _, sm = build_models([("dense", 10, "softmax")])
    # 8192 features: K >= 32 * 148 puts the head forward on the tcgen05 split-K kernel (tc.cu tc_head_fwd)
    x, y = synth_classification(100 * 5, (8192,), 10, 11)
    res = []
    for use_graph in (True, False):
        e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), 100, (8192,), (10,), device=0, precision=S.PREC_TF32,
                     use_graph=use_graph)
        e.init_params()
        e.upload_dataset(x, y)
        it, loss = e.train_epoch(0)
        it2, loss2 = e.train_epoch(it)
        step_loss = e.train_step(x[:100], y[:100], it2 + 1)
        after = e.eval_loss(x[:100], y[:100])  # again right behind an optimizer launch
        res.append((it, loss, it2, loss2, step_loss, after, e.get_model_params()))
        e.close()
    assert all(np.isfinite(v) for v in res[0][:6])
    assert res[0][:6] == res[1][:6]
    for k in res[0][6]:
        assert np.array_equal(res[0][6][k], res[1][6][k]), k


# ---------------------------------------------------------------------------------------------------------------
# Dense (MatMul fwd / dgrad / wgrad) and the LSTM contractions on the tcgen05 GEMM
# ---------------------------------------------------------------------------------------------------------------
DENSE_STACKS = [
    # (name, spec, batch, in_features): smooth activations pin the kernels; ragged sizes exercise the TMA-clipped edges
    ("wide_mlp_small", [("dense", 256, "sigmoid"), ("dense", 128, "tanh"), ("dense", 10, "softmax")], 256, 784),
    ("ragged", [("dense", 72, "tanh"), ("dense", 200, "sigmoid"), ("dense", 10, "softmax")], 520, 300),
    ("relu_chain", [("dense", 512, "relu"), ("dense", 512, "relu"), ("dense", 10, "softmax")], 384, 256),
    ("split_k_wgrad", [("dense", 64, "tanh"), ("dense", 10, "softmax")], 4096, 64),
]


@pytest.mark.parametrize("name,spec,batch,nin", DENSE_STACKS, ids=[c[0] for c in DENSE_STACKS])
def test_tf32_dense_forward_and_gradients(name, spec, batch, nin):
    om, e32, etc = engines(spec, "cce", batch, (nin,), 10)
    x, y = synth_classification(batch, (nin,), 10, 33)
    x = (x - 0.5).astype(f32)  # centred features: no catastrophic common-mode in the dot products
    p32, ptc = e32.predict(x), etc.predict(x)
    assert rel_to_max(ptc, p32) < 4e-3
    l32, g32 = e32.compute_grads(x, y)
    ltc, gtc = etc.compute_grads(x, y)
    assert abs(ltc - l32) <= 2e-3 * abs(l32)
    relu_net = any(it[0] == "dense" and it[2] == "relu" for it in spec)
    for k in g32:
        if relu_net:
            a, b = gtc[k].astype(np.float64).ravel(), g32[k].astype(np.float64).ravel()
            cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))
            assert cos > 0.995, (k, cos)
        else:
            assert rel_to_max(gtc[k], g32[k]) < 6e-3, k
    etc.profile_enable(True)
    etc.compute_grads(x, y)
    names = [r["name"] for r in etc.profile()]
    etc.profile_enable(False)
    assert any(n.startswith("dense_") and "tcgen05" in n for n in names), names
    e32.close()
    etc.close()


def test_tf32_wide_mlp_trajectory_config4_pattern():  # BASELINE config 4 pattern: Dense(ReLU) x n >> Dense(10, Softmax), Adam
    spec = [("dense", 512, "relu"), ("dense", 512, "relu"), ("dense", 10, "softmax")]
    om, sm = build_models(spec)
    batch, steps, rate = 512, 5, 0.001
    x, y = synth_classification(batch * steps, (784,), 10, 4)
    e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(rate), batch, (784,), (10,), device=0, precision=S.PREC_TF32)
    e.init_params()
    oalg = O.Adam(rate)
    _, mp, op = O.init_params(om, oalg, (batch, 784))
    lm = O.LossModel(om, O.CategoricalCrossentropy)
    for t in range(1, steps + 1):
        xb, yb = x[(t - 1) * batch:t * batch], y[(t - 1) * batch:t * batch]
        gl = e.train_step(xb, yb, t)
        mp, op, ol = O.train_step(lm, oalg, xb, yb, mp, op, t)
        assert abs(gl - float(ol)) <= 3e-3 * abs(float(ol)), (t, gl, float(ol))
    got = e.get_model_params()
    for k, v in mp.items():
        err = np.abs(got[k] - v)
        assert float(np.max(err)) <= 3 * steps * rate, k
        assert float(np.mean(err)) <= 0.05 * steps * rate, k
    e.close()


@pytest.mark.parametrize("f,fused", [(1, "0"), (3, "0"), (1, "1"), (3, "1")])
def test_tf32_lstm_matches_fp32_engine(f, fused, monkeypatch):
    # fused = "1": gate math inside the tcgen05 GEMM epilogues (forward and backward; an option, see csrc/lstm.cu)
    monkeypatch.setenv("SB_LSTM_FUSED_EPILOGUE", fused)
    spec = [("lstm", 64), ("dense", 1, "tanh")]
    om, sm = build_models(spec)
    batch, t = 256, 12
    rng = np.random.Generator(np.random.PCG64(8))
    x = rng.uniform(0, 1, size=(batch, t, f)).astype(f32)
    y = rng.uniform(0, 1, size=(batch, 1)).astype(f32)
    res = []
    for prec in (S.PREC_FP32, S.PREC_TF32, S.PREC_3XTF32):
        e = S.Engine(sm, S.MeanSquaredError, S.Adam(1e-3), batch, (t, f), (1,), device=0, precision=prec)
        e.init_params()
        res.append((e.predict(x),) + e.compute_grads(x, y))
        e.close()
    (p32, l32, g32), (ptc, ltc, gtc), (p3x, l3x, g3x) = res
    assert rel_to_max(ptc, p32) < 4e-3
    assert abs(ltc - l32) <= 4e-3 * abs(l32)
    for k in g32:
        assert rel_to_max(gtc[k], g32[k]) < 1e-2, k
    # 3xTF32 holds fp32 tolerance through the 12-step recurrence
    assert rel_to_max(p3x, p32) < 2e-5
    assert abs(l3x - l32) <= 1e-5 * abs(l32)
    for k in g32:
        assert rel_to_max(g3x[k], g32[k]) < 1e-4, k


# ---------------------------------------------------------------------------------------------------------------
# 3xTF32: split-operand tensor-core GEMM that holds fp32 tolerance (hi = tf32(x), lo = x - hi; A_hi.B_lo + A_lo.B_hi + A_hi.B_hi)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,spec,batch,nin", DENSE_STACKS[:2] + DENSE_STACKS[3:], ids=[c[0] for c in DENSE_STACKS[:2] + DENSE_STACKS[3:]])
def test_3xtf32_dense_holds_fp32_tolerance(name, spec, batch, nin):
    om, sm = build_models(spec)
    e32 = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), batch, (nin,), (10,), device=0, precision=S.PREC_FP32)
    e3x = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), batch, (nin,), (10,), device=0, precision=S.PREC_3XTF32)
    e1x = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), batch, (nin,), (10,), device=0, precision=S.PREC_TF32)
    for e in (e32, e3x, e1x):
        e.init_params()
    x, y = synth_classification(batch, (nin,), 10, 33)
    x = (x - 0.5).astype(f32)
    p32, p3x, p1x = e32.predict(x), e3x.predict(x), e1x.predict(x)
    # the two fp32-class paths agree to a few 1e-6 of the largest value; plain TF32 is ~100x further away
    assert rel_to_max(p3x, p32) < 5e-6
    assert rel_to_max(p1x, p32) > 10 * rel_to_max(p3x, p32)
    l32, g32 = e32.compute_grads(x, y)
    l3x, g3x = e3x.compute_grads(x, y)
    assert abs(l3x - l32) <= 2e-6 * abs(l32)
    for k in g32:
        assert rel_to_max(g3x[k], g32[k]) < 2e-5, k
    e3x.profile_enable(True)
    e3x.compute_grads(x, y)
    names = [r["name"] for r in e3x.profile()]
    e3x.profile_enable(False)
    assert any(n.startswith("dense_") and "tcgen05" in n for n in names), names  # on the tensor cores, not the FFMA fallback
    for e in (e32, e3x, e1x):
        e.close()


# ---------------------------------------------------------------------------------------------------------------
# 3xTF32 on the convolution kernels (fprop / dgrad / wgrad on tcgen05 with split operands): north-star "a 3xTF32 option to hold
# fp32 tolerance" for Conv2D as implicit GEMM
# ---------------------------------------------------------------------------------------------------------------
X3_CONV_STACKS = [c for c in CONV_STACKS if c[0] in ("lenet28_sigmoid", "tiny_linear", "cout128_tanh", "cifar_tail_tanh", "same_pad_tanh",
                                                     "explicit_pad_tanh", "ragged_1x3_tanh", "k2x2_no_relu")]


@pytest.mark.parametrize("name,spec,batch,in_shape", X3_CONV_STACKS, ids=[c[0] for c in X3_CONV_STACKS])
def test_3xtf32_conv_holds_fp32_tolerance(name, spec, batch, in_shape):
    om, sm = build_models(spec)
    es = [S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), batch, in_shape, (10,), device=0, precision=p)
          for p in (S.PREC_FP32, S.PREC_3XTF32, S.PREC_TF32)]
    for e in es:
        e.init_params()
    e32, e3x, e1x = es
    x, y = synth_classification(batch, in_shape, 10, 21)
    p32, p3x, p1x = e32.predict(x), e3x.predict(x), e1x.predict(x)
    assert rel_to_max(p3x, p32) < 5e-6
    assert rel_to_max(p1x, p32) > 10 * rel_to_max(p3x, p32)  # plain TF32 is far outside this band: the split operands do the work
    l32, g32 = e32.compute_grads(x, y)
    l3x, g3x = e3x.compute_grads(x, y)
    assert abs(l3x - l32) <= 2e-6 * abs(l32)
    _, mp, _ = O.init_params(om, O.Adam(0.001), (batch,) + tuple(in_shape))
    _, og, _ = O.LossModel(om, O.CategoricalCrossentropy).grad_stateful(x, y, mp)
    for k in g32:
        assert rel_to_max(g3x[k], g32[k]) < 2e-5, (k, "vs the fp32 engine")
        assert rel_to_max(g3x[k], og[k]) < 5e-5, (k, "vs the oracle")
    e3x.profile_enable(True)
    e3x.compute_grads(x, y)
    names = [r["name"] for r in e3x.profile()]
    e3x.profile_enable(False)
    assert any(n.startswith("conv2d_fwd") and "tcgen05" in n for n in names), names
    assert any(n.startswith("conv2d_wgrad") and "tcgen05" in n for n in names), names
    assert any(n.startswith("conv2d_dgrad") and "tcgen05" in n for n in names), names
    for e in es:
        e.close()


def test_3xtf32_lenet_config2_full_size_trajectory():
    """BASELINE config 2 at full size with every contraction on the tensor cores in 3xTF32 against the oracle: loss at the fp32
    parity run's tolerance (rtol 5e-5, tests/test_gpu_parity.py::test_lenet_config2_full_size); parameters after 3 Adam steps by
    max |err| <= 2 * rate and at most 8 % of a tensor's elements outside (rtol 2e-3, atol 3e-5; measured 4.1 % on `3/weights`).  The gradients themselves agree
    with fp32 to 2e-5 of their maximum (test_3xtf32_conv_holds_fp32_tolerance); Adam divides every element by its own magnitude, so
    the ~2 % of conv2's filter taps whose gradient is that small move by a visibly different fraction of a step (measured: 1.7 % of
    `3/weights` beyond atol 1e-4, max 6.5e-4)."""
    from test_gpu_parity import LENET
    om, sm = build_models(LENET)
    batch, steps, rate = 100, 3, 0.001
    x, y = synth_classification(batch * steps, (28, 28, 1), 10, 3)
    e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(rate), batch, (28, 28, 1), (10,), device=0, precision=S.PREC_3XTF32)
    e.init_params()
    oalg = O.Adam(rate)
    _, mp, op = O.init_params(om, oalg, (batch, 28, 28, 1))
    lm = O.LossModel(om, O.CategoricalCrossentropy)
    for t in range(1, steps + 1):
        xb, yb = x[(t - 1) * batch:t * batch], y[(t - 1) * batch:t * batch]
        gl = e.train_step(xb, yb, t)
        mp, op, ol = O.train_step(lm, oalg, xb, yb, mp, op, t)
        assert abs(gl - float(ol)) <= 5e-5 * abs(float(ol)) + 1e-6, (t, gl, float(ol))
    got = e.get_model_params()
    for k, v in mp.items():
        err = np.abs(got[k] - v)
        assert float(err.max()) <= 2 * rate, (k, float(err.max()))
        assert float((err > 3e-5 + 2e-3 * np.abs(v)).mean()) < 0.08, k
    e.close()


@pytest.mark.parametrize("batch", [100, 37, 128])
def test_tensor_core_head_backward_matches_the_ffma_kernel(batch, monkeypatch):
    """csrc/tc.cu head_bwd_tc_kernel (TF32 mode, the config-2 head: 36864 -> 10): dW = X^T . G and dX = G . W^T on tcgen05, X tile by
    TMA as an MN-major operand (batch rows = the K dimension, rows past the batch are TMA zero fill), G / G^T / W written by the
    threads in the K-major swizzled layout, dX through swizzled staging + TMA stores.  Against SB_TC_HEAD_BWD=0 (the FFMA kernel): same forward, so the same loss; every
    gradient within tf32 rounding of the FFMA one (the head's inputs are post-ReLU pool outputs, its gradient feeds conv2/conv1)."""
    spec = [("conv", 32, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"),
            ("conv", 64, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"), ("flatten",), ("dense", 10, "softmax")]
    _, sm = build_models(spec)
    x, y = synth_classification(batch, (28, 28, 1), 10, 5)
    res = []
    for on in ("0", "1"):
        monkeypatch.setenv("SB_TC_HEAD_BWD", on)
        e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(1e-3), batch, (28, 28, 1), (10,), device=0, precision=S.PREC_TF32)
        e.init_params()
        res.append(e.compute_grads(x, y))
        e.close()
    (l0, g0), (l1, g1) = res
    assert l0 == l1
    for k in g0:
        assert rel_to_max(g1[k], g0[k]) < 3e-3, (k, rel_to_max(g1[k], g0[k]))


def test_persistent_lstm_time_loop_matches_the_per_step_path(monkeypatch):
    """csrc/lstm_persist.cu (opt-in, SB_LSTM_PERSIST=1): one launch walks all T forward steps with the recurrent weights resident in
    shared memory and a row-block step barrier.  Same TF32 products, same gate expressions as the per-step GEMM + gate kernels:
    saved state, output and gradients must agree to 1e-4 of the largest value (K is accumulated in the same order; the two paths
    differ only in where the x.Wk + b term is added)."""
    spec = [("lstm", 256), ("dense", 1, "tanh")]
    _, sm = build_models(spec)
    batch, t, f = 256, 16, 1
    rng = np.random.Generator(np.random.PCG64(12))
    x = rng.uniform(0, 1, size=(batch, t, f)).astype(f32)
    y = rng.uniform(0, 1, size=(batch, 1)).astype(f32)
    res = []
    for on in ("0", "1", "2"):  # per-step kernels | one launch, L2 step barrier | one launch, 8-CTA clusters + multicast TMA
        monkeypatch.setenv("SB_LSTM_PERSIST", on)
        e = S.Engine(sm, S.MeanSquaredError, S.Adam(1e-3), batch, (t, f), (1,), device=0, precision=S.PREC_TF32)
        e.init_params()
        res.append((e.predict(x),) + e.compute_grads(x, y))
        e.close()
    p0, l0, g0 = res[0]
    for p1, l1, g1 in res[1:]:
        assert rel_to_max(p1, p0) < 1e-4
        assert abs(l1 - l0) <= 1e-4 * abs(l0)
        for k in g0:
            assert rel_to_max(g1[k], g0[k]) < 1e-3, k
