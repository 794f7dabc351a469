"""One-process-per-GPU plumbing (torchrun): the epoch-end model averaging across ranks.

Reference: `treeReduce(aggregateParams)` (optimizers/Optimizer.scala:86,196-230): every model and optimizer-state
tensor becomes a fixed convex combination of the k workers' tensors, the loss combines the same way and the
iteration counts add up.  torch.distributed is used for rendezvous, barriers and the NCCL baseline only; the
default data path is the engine's own fused P2P kernel over NVLink peer memory (csrc/group.cu)."""
import ctypes as C
import math
import os
from typing import List, Optional, Sequence, Tuple

from . import _native as N


def spark_chain_shape(k: int) -> Tuple[List[float], int]:
    """(weights, n_groups) of Spark 3.3.1 `RDD.treeReduce(f, depth=2)` for k partitions under ascending completion
    order: members of group j are ranks q with q % n_groups == j, chained in ascending order with
    pairwise (l+r)/2; the group results are chained the same way.  Pure host logic (no GPU, no library)."""
    def chain(n):
        return [1.0] if n == 1 else [0.5 ** (n - 1)] + [0.5 ** (n - i) for i in range(1, n)]
    w = [1.0] * k
    groups = [[i] for i in range(k)]
    parts, n_groups = k, 1
    scale = max(int(math.ceil(parts ** 0.5)), 2)
    while parts > scale + math.ceil(parts / scale):
        parts //= scale
        buckets = [[] for _ in range(parts)]
        for i, g in enumerate(groups):
            buckets[i % parts].append(g)
        merged = []
        for lst in buckets:
            cw = chain(len(lst))
            m = []
            for c, g in zip(cw, lst):
                for idx in g:
                    w[idx] *= c
                m.extend(g)
            merged.append(m)
        groups, n_groups = merged, parts
    for c, g in zip(chain(len(groups)), groups):
        for idx in g:
            w[idx] *= c
    return w, n_groups


def averaging_weights(k: int, avg_mode: int, weights: Optional[Sequence[float]] = None) -> Tuple[List[float], int]:
    if avg_mode == N.AVG_MEAN:
        return [1.0 / k] * k, 1
    if avg_mode == N.AVG_SPARK_CHAIN:
        return spark_chain_shape(k)
    if weights is None or len(weights) != k:
        raise ValueError("AVG_WEIGHTS needs k weights")
    return list(weights), 1


def ordered_weighted_sum(values: Sequence[float], weights: Sequence[float], n_groups: int) -> float:
    """fp32 sum in the reduce's own order (group-wise chains, then across groups)."""
    import numpy as np
    tot = None
    for g in range(n_groups):
        acc = None
        for q in range(g, len(values), n_groups):
            t = np.float32(values[q]) * np.float32(weights[q])
            acc = t if acc is None else np.float32(acc + t)
        tot = acc if tot is None else np.float32(tot + acc)
    return float(tot)


def init_process_group_from_env():
    """RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* as set by torchrun; NCCL when CUDA is present, else gloo."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            try:
                kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
            except Exception:
                kw = {}
        dist.init_process_group(backend=backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def allreduce_average_(tensor, weights: Sequence[float], rank: int):
    """NCCL/gloo baseline: pre-multiply by this rank's weight, then all-reduce(sum) in place."""
    import torch.distributed as dist
    tensor.mul_(float(weights[rank]))
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def combine_scalars(loss: float, iterations: int, weights: Sequence[float], n_groups: int) -> Tuple[float, int]:
    """loss with the averaging weights, iterations as a plain sum (Optimizer.scala:221-223)."""
    import torch.distributed as dist
    world = dist.get_world_size()
    gathered = [None] * world
    dist.all_gather_object(gathered, (float(loss), int(iterations)))
    return ordered_weighted_sum([g[0] for g in gathered], weights, n_groups), sum(g[1] for g in gathered)


class RankAverager:
    """Averages this rank's engine arena with its peers' at an epoch boundary."""

    def __init__(self, engine, avg_mode: int = N.AVG_SPARK_CHAIN, collective: int = N.COLL_P2P,
                 weights: Optional[Sequence[float]] = None, device_sync: bool = True):
        import torch.distributed as dist
        self.engine, self.collective = engine, collective
        # device_sync: the cross-rank barriers and the loss combination run INSIDE the averaging kernel (flags over NVLink
        # peer memory), so an epoch boundary costs one kernel and no host collective.  Needs every rank on its own GPU.
        self.device_sync = device_sync
        self._epoch = 0
        self._iters_cache = None
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.weights, self.n_groups = averaging_weights(self.world, avg_mode, weights)
        self._g = None
        self._arena = None
        if self.world > 1 and collective == N.COLL_P2P:
            handles = [None] * self.world
            dist.all_gather_object(handles, engine.arena_ipc_handle())
            blob = b"".join(handles)
            w = (C.c_float * self.world)(*self.weights)
            g = C.c_void_p()
            mode = avg_mode
            # The barrier flags live in the arena tail and outlive a group: an earlier RankAverager on the same engine left
            # flags >= this one's epochs.  No peer may still be inside an old averaging kernel when the tails are zeroed
            # (barrier 1), and every rank must have zeroed its tail before anyone raises a flag again (barrier 2).
            engine.sync()
            dist.barrier()
            N.check(N.lib().sb_group_create_ipc(engine.handle, self.rank, self.world, blob, int(mode), w, C.byref(g)))
            dist.barrier()
            self._g = g
        elif self.world > 1:
            if self.n_groups != 1:
                raise ValueError("the NCCL baseline sums in NCCL's own order; use COLL_P2P for the Spark k=8 shape")
            self._arena = engine.arena_tensor()

    def average(self, loss: float, iterations: int) -> Tuple[float, int]:
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return loss, iterations
        if self._g is not None and self.device_sync:
            return self._average_device(iterations)
        self.engine.sync()  # this rank's epoch is done
        if self._g is not None:
            dist.barrier()  # every rank's epoch is done: peers' arenas are final
            N.check(N.lib().sb_group_average_local(self._g))
            self.engine.sync()
            dist.barrier()  # every slice is stored everywhere
        else:
            allreduce_average_(self._arena, self.weights, self.rank)
            torch.cuda.synchronize()
        self.engine.reset_unaggregated()
        return combine_scalars(loss, iterations, self.weights, self.n_groups)

    def average_async(self, iterations: int) -> None:
        """Enqueue the device-synchronised averaging: no host collective, no sync.  The kernel combines the losses and adds the
        peers' iteration counts to the engine's DEVICE step counter, so `train_steps_async(ITER_FROM_DEVICE, ...)` /
        `train_epoch(ITER_FROM_DEVICE)` continue at the reference's global `iter` (Optimizer.scala:88,223) whatever the host
        believes.  `result()` reads the combined loss and the summed count back (one sync)."""
        if self.world == 1 or self._g is None:
            return
        self._epoch += 1
        N.check(N.lib().sb_group_average_device(self._g, self._epoch, int(iterations)))

    def result(self) -> Tuple[float, int]:
        """(combined loss, summed iteration count) of the latest `average_async`; raises NativeError(SB_ERR_TIMEOUT) if a
        peer never reached the barrier."""
        loss, total = C.c_float(), C.c_int64()
        N.check(N.lib().sb_group_read_result(self._g, C.byref(loss), C.byref(total)))
        return float(loss.value), int(total.value)

    def _average_device(self, iterations: int) -> Tuple[float, int]:
        self.average_async(iterations)
        return self.result()

    def close(self):
        if self._g is not None:
            self.engine.sync()
            N.lib().sb_group_destroy(self._g)
            self._g = None
