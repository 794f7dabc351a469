"""Two PROCESSES sharing GPU 0 (gloo rendezvous; the GPU time-slices the two contexts) drive the device-synchronised averaging
kernel of csrc/group.cu over CUDA-IPC mappings -- the only way to exercise `sb_group_create_ipc` / `sb_group_average_device`
on a single-GPU box.  Modes:
  ok       both ranks train 2 epochs and average; parameters, loss and the DEVICE-side global step counter must match the oracle's
           2-partition run (Optimizer.scala:59-104,196-230 restated), and a second RankAverager on the same engines must work
           (stale barrier flags of the first group are zeroed at creation)
  timeout  rank 1 never enters the averaging: rank 0's kernel gives up after SB_P2P_TIMEOUT_MS, the next sync returns
           SB_ERR_TIMEOUT, and the CUDA context is still usable afterwards
Launched by tests/test_gpu_averaging.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch.distributed as dist  # noqa: E402

import oracle as O  # noqa: E402
import scanet3_b200 as S  # noqa: E402
from helpers import build_models, synth_classification  # noqa: E402
from scanet3_b200 import _native as N  # noqa: E402
from scanet3_b200.distributed import RankAverager  # noqa: E402


def main():
    mode = sys.argv[1]
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    spec = [("dense", 16, "sigmoid"), ("dense", 4, "softmax")]
    om, sm = build_models(spec)
    batch, rows = 8, world * (8 * 3 + 5)
    x, y = synth_classification(rows, (12,), 4, 17)
    a, b = rank * rows // world, (rank + 1) * rows // world
    e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.01), batch, (12,), (4,), device=0)
    e.init_params()
    e.upload_dataset(x[a:b], y[a:b])
    ok = True
    if mode == "timeout":
        avg = RankAverager(e, N.AVG_MEAN, N.COLL_P2P)
        it, loss = e.train_epoch(0)
        if rank == 0:
            try:
                avg.average(loss, it)   # rank 1 never arrives
                ok = False
                print("IPC_CHECK timeout: the averaging returned although the peer never arrived", flush=True)
            except S.NativeError as ex:
                ok = ex.code == N.SB_ERR_TIMEOUT
                print(f"IPC_CHECK timeout: got [{ex.code}] {ex.message}", flush=True)
            # the context survived: a plain epoch still runs and syncs
            it2, loss2 = e.train_epoch(it)
            ok = ok and it2 == it and np.isfinite(loss2)
        dist.barrier()
    else:
        epochs, git, hist = 2, 0, []
        for round_ in range(2):  # the second round builds a NEW averager on the same engines (epochs restart at 1)
            avg = RankAverager(e, N.AVG_SPARK_CHAIN, N.COLL_P2P)
            for ep in range(epochs):
                first = round_ == 0 and ep == 0
                it, loss = e.train_epoch(0 if first else N.ITER_FROM_DEVICE)
                loss, total = avg.average(loss, it)
                git += total
                hist.append((len(hist) + 1, git, loss))
            avg.close()
        got = e.get_model_params()
        if rank == 0:
            mp, _, ohist = O.run_training(om, O.CategoricalCrossentropy, O.Adam(0.01), x, y, batch, 2 * epochs, world, "spark_chain")
            ok = [(h[0], h[1]) for h in hist] == [(h[0], h[1]) for h in ohist]
            for (_, _, gl), (_, _, ol) in zip(hist, ohist):
                ok = ok and abs(gl - ol) <= 5e-5 * abs(ol) + 1e-7
            for k2, v in mp.items():
                ok = ok and np.allclose(got[k2], v, rtol=1e-3, atol=2e-5)
            print(f"IPC_CHECK ok-mode {'OK' if ok else 'MISMATCH'} hist={hist} oracle={ohist}", flush=True)
        flat = np.concatenate([v.ravel() for v in got.values()])
        gathered = [None] * world
        dist.all_gather_object(gathered, flat.tobytes())
        ok = ok and all(g == gathered[0] for g in gathered)
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    e.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("IPC_CHECK RESULT", "OK" if all(flags) else "FAIL", flush=True)
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
