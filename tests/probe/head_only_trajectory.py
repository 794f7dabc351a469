"""debug: head-only model (Dense(10,Softmax) as the first layer), per-step loss vs the oracle in every precision / switch"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
from helpers import *
import oracle as O
import scanet3_b200 as S

K = int(os.environ.get("DBG_K", "8192"))
rate = float(os.environ.get("DBG_RATE", "0.001"))
om, sm = build_models([("dense", 10, "softmax")])
x, y = synth_classification(500, (K,), 10, 11)
oalg = O.Adam(rate)
_, mp, op = O.init_params(om, oalg, (100, K))
lm = O.LossModel(om, O.CategoricalCrossentropy)
ol = []
for t in range(1, 6):
    mp, op, l = O.train_step(lm, oalg, x[(t-1)*100:t*100], y[(t-1)*100:t*100], mp, op, t)
    ol.append(float(l))
print("oracle", ol)
for prec, name in ((S.PREC_FP32, "fp32"), (S.PREC_TF32, "tf32")):
    for g in (True, False):
        e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(rate), 100, (K,), (10,), device=0, precision=prec, use_graph=g)
        e.init_params()
        gl = [e.train_step(x[(t-1)*100:t*100], y[(t-1)*100:t*100], t) for t in range(1, 6)]
        after = e.eval_loss(x[:100], y[:100])
        w = e.get_model_params()
        print(name, "graph" if g else "eager", gl, after, {k: float(np.abs(v).max()) for k, v in w.items()})
        e.close()
