import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import scanet3_b200 as S
from helpers import build_models, synth_classification
spec=[("conv", 32, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"),
      ("conv", 64, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"), ("flatten",), ("dense", 10, "softmax")]
_, sm = build_models(spec)
x, y = synth_classification(100*8, (28,28,1), 10, 9)
e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), 100, (28,28,1), (10,), device=0, precision=S.PREC_TF32, use_graph=True)
e.init_params(); e.upload_dataset(x, y)
print(e.train_epoch(0))
print(e.train_epoch(8))
e.close()
