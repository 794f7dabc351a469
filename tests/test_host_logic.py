"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the plan (param paths, shapes, arena
layout) agrees with the oracle's restatement of the reference, host-side averaging logic, and the world_size-2
gloo path.  No compute calls: there is no GPU here and the product has no CPU fallback."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle as O
import scanet3_b200 as S
from helpers import build_models
from scanet3_b200 import _native as N
from scanet3_b200.distributed import averaging_weights, ordered_weighted_sum, spark_chain_shape

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_the_header_declares():
    header = open(os.path.join(ROOT, "include", "scanet_b200.h")).read()
    declared = set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 38
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/scanet_b200.h but not exported"
    assert declared == set(N.PROTOTYPES.keys()), declared ^ set(N.PROTOTYPES.keys())
    assert N.lib().sb_abi_version() == 4


def test_no_cpu_fallback_without_a_device():
    if N.lib().sb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(S.NativeError) as ei:
        S.Engine(S.Dense(2), S.MeanSquaredError, S.SGD(), 2, (3,), (2,), device=0)
    assert ei.value.code == N.SB_ERR_CUDA and "no CPU fallback" in str(ei.value)
    e = S.Engine(S.Dense(2), S.MeanSquaredError, S.SGD(), 2, (3,), (2,), device=-1)
    for call in (e.init_params, lambda: e.train_epoch(0), lambda: e.predict(np.zeros((2, 3), np.float32))):
        with pytest.raises(S.NativeError) as ei:
            call()
        assert ei.value.code == N.SB_ERR_CUDA


def test_the_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "scanet3_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, re.M), f
                assert "oracle/" not in src or f.endswith(".cu"), f


SPECS = {
    "cfg1_mlp": ([("dense", 50, "sigmoid"), ("dense", 10, "softmax")], (784,), 1000),
    "cfg2_lenet": ([("conv", 32, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"),
                    ("conv", 64, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"), ("flatten",),
                    ("dense", 10, "softmax")], (28, 28, 1), 100),
    "cfg3_lstm": ([("lstm", 256), ("dense", 1, "tanh")], (64, 1), 1024),
    "cfg4_wide": ([("dense", 4096, "relu")] * 4 + [("dense", 10, "softmax")], (784,), 8192),
    "cfg5_cifar_bn": ([("conv", 64, "relu", (3, 3), (1, 1), "Valid", False), ("bn", 0.99, "reference"), ("pool", (2, 2), (2, 2), "Valid"),
                       ("conv", 128, "relu", (3, 3), (1, 1), "Valid", False), ("bn", 0.99, "reference"), ("pool", (2, 2), (2, 2), "Valid"),
                       ("conv", 256, "relu", (3, 3), (1, 1), "Valid", False), ("bn", 0.99, "reference"), ("pool", (2, 2), (2, 2), "Valid"),
                       ("flatten",), ("dense", 10, "softmax")], (32, 32, 3), 256),
}
N_PARAMS = {"cfg1_mlp": 39760, "cfg2_lenet": 328490, "cfg3_lstm": 264449, "cfg4_wide": 53600266}


@pytest.mark.parametrize("name", list(SPECS))
def test_plan_matches_the_reference_param_layout(name):
    spec, in_shape, batch = SPECS[name]
    om, sm = build_models(spec)
    out_dim = 1 if name == "cfg3_lstm" else 10
    e = S.Engine(sm, S.MeanSquaredError, S.Adam(), batch, in_shape, (out_dim,), device=-1)
    defs = O.model_params(om, (batch,) + in_shape)
    got = {p: i for p, i in e.params.items() if i.role != N.ROLE_OPT}
    assert list(got.keys()) == [k for k in defs if defs[k].trainable] + [k for k in defs if not defs[k].trainable] or \
        set(got.keys()) == set(defs.keys())
    for k, d in defs.items():
        assert tuple(got[k].shape) == tuple(d.shape), k
        assert (got[k].role == N.ROLE_WEIGHT) == d.trainable, k
        assert got[k].aggregation == (N.AGG_AVG if d.aggregation == "avg" else N.AGG_NONE), k
    # optimizer state `<weightPath>/<stateName>` for every trainable param (Optimizer.scala:118-121)
    opt = O.opt_param_defs(defs, O.Adam())
    assert set(p for p, i in e.params.items() if i.role == N.ROLE_OPT) == set(opt.keys())
    assert e.output_shape == (batch, out_dim)
    if name in N_PARAMS:
        assert sum(i.numel for i in got.values() if i.role == N.ROLE_WEIGHT) == N_PARAMS[name]
    # arena: 128-byte aligned tensors, weights first, then state, then one region per optimizer slot
    offs = [i.offset for i in e.params.values()]
    assert all(o % 32 == 0 for o in offs) and len(set(offs)) == len(offs)
    e.close()


@pytest.mark.parametrize("alg,slots", [(S.SGD(), ["velocity"]), (S.Adam(), ["momentum", "velocity"]), (S.AdaGrad(), ["grad_acc"]),
                                       (S.RMSProp(), ["avg_grad"]), (S.AdaDelta(), ["avg_grad", "avg_delta"]),
                                       (S.Adamax(), ["momentum_1", "momentum_2"]), (S.Nadam(), ["momentum", "velocity"]),
                                       (S.AMSGrad(), ["momentum", "velocity", "corrected_velocity"])])
def test_optimizer_state_names(alg, slots):
    e = S.Engine(S.Dense(2), S.MeanSquaredError, alg, 2, (3,), (2,), device=-1)
    assert e.opt_param_paths() == [f"{w}/{s}" for s in slots for w in ("0/weights", "1/weights")]
    e.close()


def test_invalid_models_raise_like_the_reference():
    with pytest.raises(S.IllegalArgument):  # Conv2D on a rank-2 input (layer/Conv2D.scala:84-86)
        S.Engine(S.Conv2D(4), S.MeanSquaredError, S.SGD(), 2, (3,), (2,), device=-1)
    with pytest.raises(S.IllegalArgument):  # output does not match the labels
        S.Engine(S.Dense(3), S.MeanSquaredError, S.SGD(), 2, (3,), (2,), device=-1)
    with pytest.raises(S.Unsupported):
        S.Conv2D(4, format="NCHW")
    assert S.maximize(S.Dense(1))._c.minimizing is False  # Optimizer.maximize (Optimizer.scala:326-338)
    with pytest.raises(ValueError, match="only \\[Same, Valid\\] paddings are supported"):
        S.Pool2D(padding=((1, 0), (1, 0)))
    with pytest.raises(TypeError):  # builder incomplete: run() must not be callable (Optimizer.scala:256-310)
        S.Dataset(np.zeros((2, 1)), np.zeros((2, 1))).train(S.Dense(1)).loss(S.MeanSquaredError).run()
    assert repr(S.Dense(4, S.Sigmoid) >> S.Dense(1, S.Sigmoid)) == "Dense(4) >> Bias(4) >> Sigmoid >> Dense(1) >> Bias(1) >> Sigmoid"
    assert repr(S.LinearRegression()) == "Dense(1) >> Bias(1)"  # RegressionSpec.scala:63-65
    assert repr(S.Conv2D(1, (2, 2))) == "Conv2D(1,(2,2),(1,1),Valid,NHWC)"  # Conv2DLayerSpec.scala:34
    assert repr(S.Pool2D((2, 2))) == "Pool2D((2,2),(1,1),Valid,NHWC,Max)"  # Pool2DLayerSpec.scala:28


def test_spark_chain_weights_agree_between_library_host_mirror_and_oracle():
    for k in range(1, 9):
        w_lib = S.spark_chain_weights(k)
        w_py, n_groups = spark_chain_shape(k)
        w_or = O.spark_chain_weights(k)
        assert np.allclose(w_lib, w_py) and np.allclose(w_py, w_or), k
        assert abs(sum(w_py) - 1.0) < 1e-12
    assert spark_chain_shape(2) == ([0.5, 0.5], 1)
    assert spark_chain_shape(4) == ([0.125, 0.125, 0.25, 0.5], 1)
    assert spark_chain_shape(8)[1] == 2 and spark_chain_shape(8)[0] == [1 / 16, 1 / 16, 1 / 16, 1 / 16, 1 / 8, 1 / 8, 1 / 4, 1 / 4]


def _fake_results(k, seed):
    rng = np.random.default_rng(seed)
    from collections import OrderedDict
    return [O.StepResult(3 + r, OrderedDict(w=rng.normal(size=37).astype(np.float32)),
                         OrderedDict(m=rng.normal(size=37).astype(np.float32)), float(rng.uniform())) for r in range(k)]


@pytest.mark.parametrize("k", [2, 4, 8])
def test_weighted_ordered_sum_reproduces_the_pairwise_chain_bit_for_bit(k):
    """(l+r)/2 chains == power-of-two weighted sums in the same order: what the fused P2P kernel computes."""
    res = _fake_results(k, k)
    chain = O.spark_tree_reduce(res)
    w, n_groups = spark_chain_shape(k)
    for key, src in (("w", "model_params"), ("m", "opt_params")):
        tot = None
        for g in range(n_groups):
            acc = None
            for q in range(g, k, n_groups):
                t = getattr(res[q], src)[key] * np.float32(w[q])
                acc = t if acc is None else acc + t
            tot = acc if tot is None else tot + acc
        assert np.array_equal(tot, getattr(chain, src)[key])
    assert ordered_weighted_sum([r.loss for r in res], w, n_groups) == pytest.approx(chain.loss, rel=1e-6)
    assert chain.iterations == sum(r.iterations for r in res)


def test_gloo_world_size_2_weighted_allreduce_and_scalar_combine(tmp_path):
    """N>1 host path on CPU: two ranks, gloo, 127.0.0.1 rendezvous."""
    script = tmp_path / "worker.py"
    script.write_text(f'''
import os, sys
sys.path.insert(0, {ROOT!r})
import numpy as np, torch, torch.distributed as dist
from scanet3_b200 import _native as N
from scanet3_b200.distributed import init_process_group_from_env, averaging_weights, allreduce_average_, combine_scalars
rank, world = init_process_group_from_env()
assert dist.get_backend() == "gloo" and world == 2
w, ng = averaging_weights(world, N.AVG_SPARK_CHAIN)
rng = np.random.default_rng(0)
arenas = [rng.normal(size=1000).astype(np.float32) for _ in range(world)]
t = torch.from_numpy(arenas[rank].copy())
allreduce_average_(t, w, rank)
expect = (arenas[0] + arenas[1]) / np.float32(2)
assert np.array_equal(t.numpy(), expect), "k=2 must equal the reference's (l+r)/2 exactly"
loss, iters = combine_scalars(1.0 + rank, 10 + rank, w, ng)
assert loss == 1.5 and iters == 21
dist.barrier()
print("RANK_OK", rank)
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2", CUDA_VISIBLE_DEVICES="")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK_OK {r}" in o, o


# ---------------------------------------------------------------------------------------------------------------
# parameter wire format (optimizers/KryoSerializers.scala:27-63): host-only entry points, no GPU needed
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(), (1,), (5,), (3, 4), (2, 3, 1, 5), (0, 7)])
def test_wire_records_match_the_kryo_layout(shape):
    from oracle import wire as W
    rng = np.random.default_rng(3)
    t = rng.standard_normal(shape).astype(np.float32)
    rec = S.wire_encode(t)
    assert rec == W.tensor_write(t)                      # byte for byte what TensorSerializer.write emits
    assert rec[:8] == bytes([0, 0, 0, 1, 0, 0, 0, len(shape)])  # big-endian type code DT_FLOAT, rank
    back, used = S.wire_decode(rec + b"trailing")
    assert used == len(rec) and back.shape == t.shape and np.array_equal(back, t)
    want, used2 = W.tensor_read(rec)
    assert used2 == len(rec) and np.array_equal(want, t)


def test_wire_decode_rejects_malformed_records():
    rec = S.wire_encode(np.arange(6, dtype=np.float32).reshape(2, 3))
    with pytest.raises(N.IllegalArgument):
        S.wire_decode(rec[:-1])                          # truncated payload
    with pytest.raises(N.IllegalArgument):
        S.wire_decode(rec[:10])                          # truncated header
    bad = bytearray(rec); bad[3] = 2                     # DT_DOUBLE
    with pytest.raises(N.Unsupported):
        S.wire_decode(bytes(bad))
    bad = bytearray(rec); bad[27] = 20                   # payload size disagrees with the shape
    with pytest.raises(N.IllegalArgument):
        S.wire_decode(bytes(bad))


# ---------------------------------------------------------------------------------------------------------------
# Model.describe (models/Model.scala:62-81), printed by Optimizer.run (optimizers/Optimizer.scala:65)
# ---------------------------------------------------------------------------------------------------------------
CNN_DESCRIPTION = """\
+----------------------------------+--------------+--------------+------------+------------+
|name                              |weights       |weights params|state params|output      |
+----------------------------------+--------------+--------------+------------+------------+
|Input                             |              |              |            |(_,24,24,1) |
|Conv2D(32,(3,3),(1,1),Valid,NHWC) |(3, 3, 1, 32) |288           |            |(_,22,22,32)|
|ReLU(0.0)                         |              |              |            |(_,22,22,32)|
|Pool2D((2,2),(1,1),Valid,NHWC,Max)|              |              |            |(_,21,21,32)|
|Conv2D(64,(3,3),(1,1),Valid,NHWC) |(3, 3, 32, 64)|18432         |            |(_,19,19,64)|
|ReLU(0.0)                         |              |              |            |(_,19,19,64)|
|Pool2D((2,2),(1,1),Valid,NHWC,Max)|              |              |            |(_,18,18,64)|
|Flatten                           |              |              |            |(_,20736)   |
|Dense(10)                         |(20736, 10)   |207360        |            |(_,10)      |
|Bias(10)                          |(10)          |10            |            |(_,10)      |
|Softmax                           |              |              |            |(_,10)      |
+----------------------------------+--------------+--------------+------------+------------+
Total weight params: 226090 (883.2 KB), state params: 0 (0 B)"""


def test_model_description_matches_the_reference_golden_table():
    # models/CNNSpec.scala:39-63 ("be described"), character for character
    model = (S.Conv2D(32, activation=S.ReLU()) >> S.Pool2D() >> S.Conv2D(64, activation=S.ReLU()) >> S.Pool2D()
             >> S.Flatten >> S.Dense(10, S.Softmax))
    assert model.describe((1, 24, 24, 1)) == CNN_DESCRIPTION


def test_description_helpers_follow_the_reference_rules():
    from scanet3_b200 import dsl
    assert dsl._format_size(0) == "0 B" and dsl._format_size(1023) == "1023 B"  # utils/Bytes.scala:4-8
    assert dsl._format_size(1024) == "1.0 KB" and dsl._format_size(904360) == "883.2 KB" and dsl._format_size(5 << 20) == "5.0 MB"
    assert dsl._group_concat([]) == "" and dsl._group_concat(["a"]) == "a" and dsl._group_concat(["a", "b"]) == "a, b"
    assert dsl._group_concat(["a", "a"]) == "[a]x2" and dsl._group_concat(["a", "b"] * 3) == "[a, b]x3"  # LayerInfo.scala:10-37
    assert dsl._shape_str(()) == "()" and dsl._shape_str((10,)) == "(10)" and dsl._shape_str((3, 3, 1, 32)) == "(3, 3, 1, 32)"
    # utils/TabulatorSpec.scala:10-23
    assert dsl._tabulate([["name", "age"], ["Pavlo", "31"], ["Yurii Dawg", "28"]]) == (
        "+----------+---+\n|name      |age|\n+----------+---+\n|Pavlo     |31 |\n|Yurii Dawg|28 |\n+----------+---+")


def _all_test_models():
    import test_gpu_lstm as TL
    import test_gpu_tf32 as TT
    for name, (spec, in_shape, batch) in SPECS.items():
        yield name, spec, (batch,) + tuple(in_shape), (1 if name == "cfg3_lstm" else 10)
    for name, spec, batch, in_shape in TT.CONV_STACKS:
        yield name, spec, (batch,) + tuple(in_shape), 10
    for name, spec, batch, nin in TT.DENSE_STACKS:
        yield name, spec, (batch, nin), 10
    for name, (spec, t, f, batch) in list(TL.SEQ_MODELS.items()) + list(TL.RNN_MODELS.items()):
        yield name, spec, (batch, t, f), 1


def test_described_shapes_and_counts_agree_with_the_engine_plan_for_every_test_model():
    """the Python shape rules behind `describe` against the C++ planner (plan-only engines): per-model weight / state element
    counts and the output shape, for the BASELINE configs and every model the GPU tests train"""
    n = 0
    for name, spec, full_in, out_dim in _all_test_models():
        _, sm = build_models(spec)
        infos = sm.info(full_in)
        e = S.Engine(sm, S.MeanSquaredError, S.SGD(), full_in[0], full_in[1:], (out_dim,), device=-1)
        plan = [i for i in e.params.values() if i.role != N.ROLE_OPT]
        assert sum(i.weightsTotal for i in infos) == sum(i.numel for i in plan if i.role == N.ROLE_WEIGHT), name
        assert sum(i.stateTotal for i in infos) == sum(i.numel for i in plan if i.role != N.ROLE_WEIGHT), name
        assert tuple(infos[-1].output) == tuple(e.output_shape) == tuple(sm.outputShape(full_in)), name
        assert sm.describe(full_in).count("\n") == len(infos) + 5, name
        e.close()
        n += 1
    assert n >= 25
