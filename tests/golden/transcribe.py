"""Writes tests/golden/reference_vectors.json: the reference's own known-answer vectors for the hot path, transcribed by hand
from the literals in its test specs (cited per case as file:line under /root/reference/src/test/scala/scanet/).  The
reference is Scala over TF-Java and cannot be executed here, so these literals -- TF-CPU outputs the reference's authors
recorded -- are what pins the oracle.  Re-run `python tests/golden/transcribe.py` after editing; nothing here reads
/root/reference at run time."""
import json
import os

IMG = [[2, 1, 2, 0, 1], [1, 3, 2, 2, 3], [1, 1, 3, 3, 0], [2, 2, 0, 1, 1], [0, 0, 3, 1, 2]]
FIL = [[2, 3], [0, 1]]
X4 = [[0, 0, 1], [0, 1, 1], [1, 0, 1], [1, 1, 1]]
W34 = [[1, 0.5, 1, 0.1], [0.1, 1, 1, 1], [1, 0, 0.2, 0.3]]

cases = []


def case(kind, ref, **kw):
    cases.append(dict(kind=kind, ref=ref, **kw))


# math/nn/KernelsSpec.scala -- Conv2D forward, all paddings / strides
for ref, padding, strides, exp in [
    ("math/nn/KernelsSpec.scala:45-58", "same", [1, 1], [[10, 10, 6, 6, 2], [12, 15, 13, 13, 6], [7, 11, 16, 7, 0], [10, 7, 4, 7, 2], [0, 9, 9, 8, 4]]),
    ("math/nn/KernelsSpec.scala:60-73", "same", [2, 2], [[10, 6, 2], [7, 16, 0], [0, 9, 4]]),
    ("math/nn/KernelsSpec.scala:77-89", "valid", [1, 1], [[10, 10, 6, 6], [12, 15, 13, 13], [7, 11, 16, 7], [10, 7, 4, 7]]),
    ("math/nn/KernelsSpec.scala:91-103", "valid", [2, 2], [[10, 6], [7, 16]]),
]:
    case("conv2d_fwd", ref, x=IMG, w=FIL, padding=padding, strides=strides, expected=exp)
case("conv2d_wgrad", "math/nn/KernelsSpec.scala:146-157", x=IMG, w_shape=[2, 2], expected=[[26, 25], [25, 27]])
case("conv2d_dgrad", "math/nn/KernelsSpec.scala:174-188", x_shape=[5, 5], w=FIL,
     expected=[[2, 5, 5, 5, 3], [2, 6, 6, 6, 4], [2, 6, 6, 6, 4], [2, 6, 6, 6, 4], [0, 1, 1, 1, 1]])
# math/nn/KernelsSpec.scala -- MaxPool / AvgPool forward and gradients
for ref, padding, strides, reduce, exp in [
    ("math/nn/KernelsSpec.scala:200-210", "same", [1, 1], "max", [[3, 3, 2, 3, 3], [3, 3, 3, 3, 3], [2, 3, 3, 3, 1], [2, 3, 3, 2, 2], [0, 3, 3, 2, 2]]),
    ("math/nn/KernelsSpec.scala:212-222", "same", [1, 1], "avg", [[1.75, 2.0, 1.5, 1.5, 2.0], [1.5, 2.25, 2.5, 2.0, 1.5], [1.5, 1.5, 1.75, 1.25, 0.5], [1.0, 1.25, 1.25, 1.25, 1.5], [0.0, 1.5, 2.0, 1.5, 2.0]]),
    ("math/nn/KernelsSpec.scala:224-233", "same", [2, 2], "max", [[3, 2, 3], [2, 3, 1], [0, 3, 2]]),
    ("math/nn/KernelsSpec.scala:237-246", "valid", [1, 1], "max", [[3, 3, 2, 3], [3, 3, 3, 3], [2, 3, 3, 3], [2, 3, 3, 2]]),
    ("math/nn/KernelsSpec.scala:248-256", "valid", [2, 2], "max", [[3, 2], [2, 3]]),
]:
    case("pool2d_fwd", ref, x=IMG, window=[2, 2], padding=padding, strides=strides, reduce=reduce, expected=exp)
case("pool2d_grad", "math/nn/KernelsSpec.scala:277-288", x=IMG, window=[2, 2], padding="same", strides=[1, 1], reduce="max",
     expected=[[0, 0, 1, 0, 0], [0, 4, 0, 0, 4], [0, 0, 3, 1, 0], [2, 0, 0, 0, 1], [1, 0, 4, 0, 4]])
case("pool2d_grad", "math/nn/KernelsSpec.scala:294-305", x=IMG, window=[2, 2], padding="same", strides=[1, 1], reduce="avg",
     expected=[[0.25, 0.5, 0.5, 0.5, 0.75], [0.5, 1, 1, 1, 1.5], [0.5, 1, 1, 1, 1.5], [0.5, 1, 1, 1, 1.5], [0.75, 1.5, 1.5, 1.5, 2.25]])
# models/layer/DenseLayerSpec.scala
case("dense_sigmoid_fwd", "models/layer/DenseLayerSpec.scala:17-41", x=X4, w=W34, b=[0, 0, 0, 0], round=6,
     expected=[[0.731059, 0.500000, 0.549834, 0.574442], [0.750260, 0.731059, 0.768525, 0.785835],
               [0.880797, 0.622459, 0.768525, 0.598688], [0.890903, 0.817574, 0.900250, 0.802184]])
case("dense_sigmoid_bce_grad", "models/layer/DenseLayerSpec.scala:55-78", x=X4,
     y=[[0.7310586, 0.5, 0.549834, 0.5744425], [0.7502601, 0.7310586, 0.76852477, 0.7858349],
        [0.8807971, 0.6224593, 0.76852477, 0.59868765], [0.89090323, 0.81757444, 0.9002496, 0.80218387]],
     expected_w=[[-0.048231270, -0.027502108, -0.041798398, -0.025054470], [-0.040072710, -0.034289565, -0.041798398, -0.036751173],
                 [-0.078313690, -0.041943270, -0.061695820, -0.047571808]],
     expected_b=[-0.078313690, -0.041943270, -0.061695820, -0.047571808])
# models/RegressionSpec.scala
case("linear_regression", "models/RegressionSpec.scala:13-61", x=[[1, 2], [2, 4]], y=[[6], [12]], w=[[2], [3]], b=[1],
     loss=17.0, result=[[9], [17]], grad_w_at_zero=[[-30], [-60]], grad_b_at_zero=[-18])
case("logistic_regression", "models/RegressionSpec.scala:67-105", x=[[0.34, 0.78], [0.6, 0.86]], y=[[0.402], [0.47800002]],
     w=[[0.2], [0.3]], b=[0.1], loss=0.7422824, result=[[0.599168], [0.617276]], grad_w=[[0.075301], [0.136784]],
     grad_b=[0.168222], round=6)
# models/layer/RNNLayerSpec.scala
case("simple_rnn_fwd", "models/layer/RNNLayerSpec.scala:17-34", x=[1, 2, 3], kernel=[[-0.5302748, -1.2049599]],
     recurrent=[[-0.77547526, -0.6313778], [-0.6313778, 0.7754754]], bias=[0, 0], expected=[[0.222901, -6.019066]], round=6)
case("lstm_fwd", "models/layer/RNNLayerSpec.scala:44-79", x=[1, 2, 3], round=6, expected=[[0.382158, 0.029766]],
     gates={"forget": [[[0.17034012, -0.39808878]], [[0.4032409, -0.37492862], [-0.0409357, 0.22285524]], [1, 1]],
            "input": [[[0.5514972, -0.6295431]], [[-0.50807524, -0.07556351], [0.02593641, 0.44693834]], [0, 0]],
            "gate": [[[0.39550006, 0.03499722]], [[-0.11996187, 0.5238996], [0.34965965, -0.09114401]], [0, 0]],
            "output": [[[-0.02363521, 0.1830315]], [[-0.37248448, 0.07327155], [-0.70414686, -0.3490578]], [0, 0]]})
# models/layer/BatchNormSpec.scala
case("batchnorm_generic_train", "models/layer/BatchNormSpec.scala:30-49", x=[[0, 0, 10], [0, 8, 4], [5, 0, 13], [20, 3, 8]],
     w=W34, momentum=0.5, round=3,
     expected=[[-0.595, -1.168, -1.042, -1.231], [-1.181, 0.422, -0.232, 1.313], [0.307, -0.671, -0.375, -0.656], [1.469, 1.417, 1.649, 0.574]],
     moving_mean=[[7.638, 2.938, 5.375, 3.0]], moving_variance=[[39.828, 13.148, 35.764, 3.47]])
# models/InitializerSpec.scala (TF Philox, seeded)
case("random_uniform_seeded", "models/InitializerSpec.scala:23-30", min=-1, max=1, seed=1, shape=[3], round=2, expected=[0.63, -0.84, 0.78])

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
with open(out, "w") as f:
    json.dump({"source": "pashashiz/scanet3 src/test/scala/scanet (transcribed literals)", "cases": cases}, f, indent=1)
print(len(cases), "cases ->", out)
