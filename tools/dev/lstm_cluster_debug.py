"""Developer check of csrc/lstm_persist.cu: per-step path vs the one-launch forms, on the whole h sequence (returnSequence)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scanet3_b200 as S
from helpers import build_models

def run(on, batch, t, f=1):
    os.environ["SB_LSTM_PERSIST"] = on
    _, sm = build_models([("lstm", 256, True)])
    rng = np.random.Generator(np.random.PCG64(12))
    x = rng.uniform(0, 1, size=(batch, t, f)).astype(np.float32)
    e = S.Engine(sm, S.MeanSquaredError, S.Adam(1e-3), batch, (t, f), (t, 256), device=0, precision=S.PREC_TF32)
    e.init_params()
    p = e.predict(x)
    e.close()
    return p.reshape(batch, t, 256)

batch, t = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 3
p0 = run("0", batch, t)
p2 = run("2", batch, t)
d = np.abs(p2 - p0)
for s in range(t):
    ds = d[:, s, :]
    print(f"step {s}: max {ds.max():.3e}; bad rows {np.nonzero(ds.max(axis=1) > 1e-5)[0].tolist()[:70]}")
    print(f"         bad units by chunk of 32: {[int((ds[:, c*32:(c+1)*32].max() > 1e-5)) for c in range(8)]}  per-unit-in-chunk: {np.nonzero(ds.max(axis=0) > 1e-5)[0].tolist()[:40]}")
