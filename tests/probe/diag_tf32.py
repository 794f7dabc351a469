"""Developer diagnostic: per-tensor error of the TF32 engine vs the FP32 engine for a few conv geometries."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scanet3_b200 as S
from helpers import build_models, synth_classification


def rel(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b)) / (np.max(np.abs(b)) + 1e-30))


CASES = [
    ("3x3 linear", [("conv", 32, "identity", (3, 3), (1, 1), "Valid", False), ("conv", 64, "identity", (3, 3), (1, 1), "Valid", False),
                    ("flatten",), ("dense", 10, "softmax")], 3, (7, 9, 1)),
    ("3x3 sigmoid", [("conv", 32, "sigmoid", (3, 3), (1, 1), "Valid", False), ("conv", 64, "sigmoid", (3, 3), (1, 1), "Valid", False),
                     ("flatten",), ("dense", 10, "softmax")], 3, (7, 9, 1)),
    ("3x3 relu", [("conv", 32, "relu", (3, 3), (1, 1), "Valid", False), ("conv", 64, "relu", (3, 3), (1, 1), "Valid", False),
                  ("flatten",), ("dense", 10, "softmax")], 3, (7, 9, 1)),
    ("3x3 relu big", [("conv", 32, "relu", (3, 3), (1, 1), "Valid", False), ("conv", 64, "relu", (3, 3), (1, 1), "Valid", False),
                      ("flatten",), ("dense", 10, "softmax")], 16, (28, 28, 1)),
    ("3x3 linear big", [("conv", 32, "identity", (3, 3), (1, 1), "Valid", False), ("conv", 64, "identity", (3, 3), (1, 1), "Valid", False),
                        ("flatten",), ("dense", 10, "softmax")], 16, (28, 28, 1)),
    ("1x3 linear", [("conv", 32, "identity", (3, 3), (1, 1), "Valid", False), ("conv", 64, "identity", (1, 3), (1, 1), "Valid", False),
                    ("flatten",), ("dense", 10, "softmax")], 3, (7, 9, 1)),
    ("3x1 linear", [("conv", 32, "identity", (3, 3), (1, 1), "Valid", False), ("conv", 64, "identity", (3, 1), (1, 1), "Valid", False),
                    ("flatten",), ("dense", 10, "softmax")], 3, (7, 9, 1)),
]
for name, spec, batch, in_shape in CASES:
    _, sm = build_models(spec)
    x, y = synth_classification(batch, in_shape, 10, 21)
    out = []
    for prec in (S.PREC_FP32, S.PREC_TF32):
        e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), batch, in_shape, (10,), device=0, precision=prec)
        e.init_params()
        pred = e.predict(x)
        loss, g = e.compute_grads(x, y)
        out.append((pred, loss, g))
        e.close()
    print(f"{name:16s} pred {rel(out[1][0], out[0][0]):.2e} loss {abs(out[1][1] - out[0][1]) / abs(out[0][1]):.2e} " +
          " ".join(f"{k}:{rel(out[1][2][k], out[0][2][k]):.2e}" for k in out[0][2]))
