"""One process per GPU (torchrun): k ranks train their shards and average at every epoch boundary with the
device-synchronised fused P2P kernel; rank 0 compares parameters, loss and the global iteration counter with the oracle's
k-partition run (Optimizer.scala:59-104,196-230 restated).  Launched by tests/test_gpu_multi.py or by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py [spark_chain|mean|nccl]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle as O  # noqa: E402
import scanet3_b200 as S  # noqa: E402
from helpers import build_models, synth_classification  # noqa: E402
from scanet3_b200 import _native as N  # noqa: E402
from scanet3_b200.distributed import RankAverager, init_process_group_from_env  # noqa: E402


def run_check(mode, rank, world, local, log=None):
    """3 epochs of a small MLP over `world` shards, averaged at every epoch boundary through RankAverager (the very path
    bench.py and the host runtime use); rank 0 compares with the oracle's k-partition run.  -> dict on every rank."""
    spec = [("dense", 16, "sigmoid"), ("dense", 4, "softmax")]
    om, sm = build_models(spec)
    batch, epochs, rows = 8, 3, world * (8 * 3 + 5)
    x, y = synth_classification(rows, (12,), 4, 17)
    a, b = rank * rows // world, (rank + 1) * rows // world
    e = S.Engine(sm, S.CategoricalCrossentropy, S.Adam(0.01), batch, (12,), (4,), device=local)
    e.init_params()
    e.upload_dataset(x[a:b], y[a:b])
    coll = N.COLL_NCCL if mode == "nccl" else N.COLL_P2P
    avg_mode = N.AVG_MEAN if mode in ("mean", "nccl") else N.AVG_SPARK_CHAIN
    avg = RankAverager(e, avg_mode, coll)
    git, hist = 0, []
    for ep in range(1, epochs + 1):
        # device-synchronised P2P averaging: from the second epoch on the DEVICE's own step counter continues -- the averaging
        # kernel added the peers' iteration counts to it (Adam's bias corrections would expose a wrong count in the parameters)
        it, loss = e.train_epoch(N.ITER_FROM_DEVICE if (coll == N.COLL_P2P and ep > 1) else git)
        loss, total = avg.average(loss, it)
        git += total
        hist.append((ep, git, loss))
    got = e.get_model_params()
    ok, max_rel = True, 0.0
    if rank == 0:
        mp, _, ohist = O.run_training(om, O.CategoricalCrossentropy, O.Adam(0.01), x, y, batch, epochs, world,
                                      "mean" if avg_mode == N.AVG_MEAN else "spark_chain")
        ok = [(h[0], h[1]) for h in hist] == [(h[0], h[1]) for h in ohist]
        for (_, _, gl), (_, _, ol) in zip(hist, ohist):
            ok = ok and abs(gl - ol) <= 5e-5 * abs(ol) + 1e-7
        for k2, v in mp.items():
            ok = ok and np.allclose(got[k2], v, rtol=1e-3, atol=2e-5)
            max_rel = max(max_rel, float(np.max(np.abs(got[k2] - v) / (np.abs(v) + 1e-3))))
        print(f"MULTI_GPU_CHECK world={world} mode={mode} {'OK' if ok else 'MISMATCH'} max_rel_err={max_rel:.2e} hist={hist} "
              f"oracle={ohist}", flush=True, file=log or sys.stdout)
    # every rank must end up with identical parameters
    t = torch.from_numpy(np.concatenate([v.ravel() for v in got.values()])).cuda()
    ref = t.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(t, ref))
    flag = torch.tensor([int(ok and same)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    avg.close()
    e.close()
    return {"k": world, "mode": mode, "ok": int(flag.item()) == 1, "max_rel_err": max_rel}


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "spark_chain"
    rank, world = init_process_group_from_env()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    res = run_check(mode, rank, world, local)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
