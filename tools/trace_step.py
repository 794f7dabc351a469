"""In-situ timeline of the train step inside the replayed CUDA graph (SB_TRACE facility of csrc/common.cuh): every kernel stamps
%globaltimer right after its programmatic-dependency wait; the start-to-start intervals are what each kernel costs the step with
the launch overlap included.  Needs the trace build of the library: `make trace`.
usage: python tools/trace_step.py [cfg2] [steps] > profiles/r2_trace_cfg2.txt"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SB_LIB", os.path.join(ROOT, "scanet3_b200", "libscanet_b200_trace.so"))  # built by `make trace`
import bench  # noqa: E402
import scanet3_b200 as S  # noqa: E402
from scanet3_b200 import _native as N  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    bench.select_config(cfg)
    x, y = bench.synth(bench.BATCH * (steps + 16), 7)
    loss = S.MeanSquaredError if bench.CFG["loss"] == "mse" else S.CategoricalCrossentropy
    e = S.Engine(bench.build_model(S, cfg), loss, S.Adam(bench.CFG["rate"]), bench.BATCH, bench.IN_SHAPE, (bench.CLASSES,), device=0,
                 precision=S.PREC_TF32, use_graph=True)
    e.init_params()
    e.upload_dataset(x, y)
    e.warm()
    e.train_steps_async(0, 0, 16)
    e.sync()
    lib = N.lib()
    lib.sb_internal_trace_begin.restype = C.c_int
    lib.sb_internal_trace_read.restype = C.c_int
    lib.sb_internal_trace_read.argtypes = [C.c_void_p, C.c_int]
    N.check(lib.sb_internal_trace_begin())
    e.train_steps_async(16, 16, steps)
    e.sync()
    buf = np.zeros(1 << 16, np.uint64)
    N.check(lib.sb_internal_trace_read(buf.ctypes.data, buf.size))
    n = int(buf[0])
    ev = buf[2:2 + 2 * n].reshape(-1, 2)
    order = np.argsort(ev[:, 0], kind="stable")
    ev = ev[order]
    t = ev[:, 0].astype(np.int64)
    tag = [(int(v) >> 16, int(v) & 0xFFFF) for v in ev[:, 1]]
    per_step = n // steps
    print(f"# {cfg}: {n} kernel starts over {steps} graph-replayed steps = {per_step} per step; total {(t[-1] - t[0]) / 1000 / steps:.2f} us/step")
    # aggregate by position within the step (kernels of one step are identified by their (grid, block) signature sequence)
    start = per_step * 8  # skip the first 8 steps
    rows = {}
    seq = []
    for i in range(start, n - 1):
        k = tag[i]
        d = (t[i + 1] - t[i]) / 1000.0
        if k not in rows:
            rows[k] = []
            seq.append(k)
        rows[k].append(d)
    print("# (grid, block)  starts/step  median start-to-next-start us   (sum of medians x starts/step = step)")
    tot = 0.0
    for k in seq:
        v = np.array(rows[k])
        per = len(v) / max(1, (steps - 8))
        tot += float(np.median(v)) * per
        print(f"{str(k):>16s}  {per:5.2f}  {np.median(v):7.2f}  (p10 {np.percentile(v, 10):6.2f}  p90 {np.percentile(v, 90):6.2f})")
    print(f"# sum {tot:.2f} us")
    e.close()


if __name__ == "__main__":
    main()
