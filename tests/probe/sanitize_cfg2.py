import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import scanet3_b200 as S
from helpers import build_models, synth_classification
spec=[("conv", 32, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"),
      ("conv", 64, "relu", (3, 3), (1, 1), "Valid", False), ("pool", (2, 2), (1, 1), "Valid"), ("flatten",), ("dense", 10, "softmax")]
_, sm = build_models(spec)
B=20
x, y = synth_classification(B*10+3, (28,28,1), 10, 9)
e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.001), B, (28,28,1), (10,), device=0, precision=S.PREC_TF32, use_graph=True)
e.init_params(); e.upload_dataset(x, y)
print(e.train_epoch(0))
print(e.train_batches(x, y, 10)[:3])
print(e.estimate(S.native.EST_ACCURACY), e.estimate(S.native.EST_SQUARED_ERROR, x, y))
e.save_checkpoint("/tmp/ck.sb3", 20, 2); print(e.load_checkpoint("/tmp/ck.sb3"))
e.close()
print("done")
