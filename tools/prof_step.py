"""Tiny driver for ncu captures: a few eager (no CUDA graph) train steps of a BASELINE config, so that `ncu -k regex:<kernel>`
sees plain launches.  usage: python tools/prof_step.py [cfg2] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import scanet3_b200 as S  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    bench.select_config(cfg)
    x, y = bench.synth(bench.BATCH * 4, 7)
    loss = S.MeanSquaredError if bench.CFG["loss"] == "mse" else S.CategoricalCrossentropy
    e = S.Engine(bench.build_model(S, cfg), loss, S.Adam(bench.CFG["rate"]), bench.BATCH, bench.IN_SHAPE, (bench.CLASSES,), device=0,
                 precision=S.PREC_TF32, use_graph=False)
    e.init_params()
    e.upload_dataset(x, y)
    e.train_steps_async(0, 0, min(steps, 4))
    e.sync()
    print("ok", e.read_loss())
    e.close()


if __name__ == "__main__":
    main()
