#!/usr/bin/env python
"""Benchmark of the hot path: train samples/sec of the MNIST-shape LeNet CNN (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one minibatch train step (batch 100 per GPU) of the per-worker loop; an epoch is the rank's shard
(60 000 / N rows => 600 / N steps) and ends with the synchronous model averaging across the N ranks.  One JSON line
is printed by rank 0.  See DESIGN.md section 6 for what each key means and how it is measured."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train samples/sec (MNIST-shape CNN)"
# BASELINE.json configs; cfg2 (configs[1]) is the workload the metric is quoted on and the default.  The others are run with
# --config for profiles/ only.  flop = algorithmic FLOP per sample of one train step.
CONFIGS = {
    "cfg1": dict(batch=1000, in_shape=(784,), out=10, rows=30000, flop=159_800, loss="cce", rate=0.01,
                 workload="cfg1 MLP Dense(50,Sigmoid)>>Dense(10,Softmax), CCE, Adam(0.01), batch 1000/GPU, synthetic 30000x784"),
    "cfg2": dict(batch=100, in_shape=(28, 28, 1), out=10, rows=60000, flop=61_140_480, loss="cce", rate=0.001,
                 workload="cfg2 LeNet Conv2D(32,ReLU)>>Pool2D>>Conv2D(64,ReLU)>>Pool2D>>Flatten>>Dense(10,Softmax), CCE, "
                          "Adam(0.001), batch 100/GPU, synthetic 60000x28x28x1 sharded over the ranks"),
    "cfg3": dict(batch=1024, in_shape=(64, 1), out=1, rows=8192, flop=101_058_048, loss="mse", rate=0.001,
                 workload="cfg3 LSTM(256)>>Dense(1,Tanh), MSE, Adam(0.001), seq len 64, batch 1024/GPU, synthetic 8192x64x1"),
    "cfg4": dict(batch=8192, in_shape=(784,), out=10, rows=32768, flop=315_080_704, loss="cce", rate=0.001,
                 workload="cfg4 wide MLP 4xDense(4096,ReLU)>>Dense(10,Softmax), CCE, Adam(0.001), batch 8192/GPU, synthetic 32768x784"),
    "cfg5": dict(batch=256, in_shape=(32, 32, 3), out=10, rows=5120, flop=109_353_984, loss="cce", rate=0.001,
                 workload="cfg5 CIFAR-shape CNN Conv2D(64,ReLU)>>BatchNorm>>Pool2D(s2)>>Conv2D(128,ReLU)>>BatchNorm>>Pool2D(s2)>>"
                          "Conv2D(256,ReLU)>>BatchNorm>>Pool2D(s2)>>Flatten>>Dense(10,Softmax), CCE, Adam(0.001), batch 256/GPU, "
                          "synthetic 5120x32x32x3"),
}
CFG = CONFIGS["cfg2"]
BATCH, IN_SHAPE, CLASSES, ROWS_TOTAL = CFG["batch"], CFG["in_shape"], CFG["out"], CFG["rows"]
WORKLOAD, FLOP_PER_SAMPLE = CFG["workload"], CFG["flop"]


def select_config(name):
    global CFG, BATCH, IN_SHAPE, CLASSES, ROWS_TOTAL, WORKLOAD, FLOP_PER_SAMPLE
    CFG = CONFIGS[name]
    BATCH, IN_SHAPE, CLASSES, ROWS_TOTAL = CFG["batch"], CFG["in_shape"], CFG["out"], CFG["rows"]
    WORKLOAD, FLOP_PER_SAMPLE = CFG["workload"], CFG["flop"]


def synth(rows, seed):
    """features U[1/256,1] (MNIST pixels are (ubyte+1)/256), one-hot labels class = row % 10 (regression: targets U[0,1],
    the min-max scaled sunspot prep); PCG64."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.uniform(1.0 / 256, 1.0, size=(rows,) + IN_SHAPE).astype(np.float32)
    if CFG["loss"] == "mse":
        return x, rng.uniform(0.0, 1.0, size=(rows, CLASSES)).astype(np.float32)
    y = np.zeros((rows, CLASSES), np.float32)
    y[np.arange(rows), np.arange(rows) % CLASSES] = 1.0
    return x, y


def build_model(S, name, seed=1):
    """the BASELINE model in the product DSL (S = scanet3_b200) or, with the snake_case keywords, in the oracle's"""
    kw = (lambda i: {"kernelInitializer": i}) if hasattr(S, "Engine") else (lambda i: {"kernel_initializer": i})
    g = S.GlorotUniform(seed)
    if name == "cfg1":
        return S.Dense(50, S.Sigmoid, **kw(g)) >> S.Dense(10, S.Softmax, **kw(g))
    if name == "cfg2":
        return (S.Conv2D(32, activation=S.ReLU(), **kw(g)) >> S.Pool2D() >> S.Conv2D(64, activation=S.ReLU(), **kw(g))
                >> S.Pool2D() >> S.Flatten >> S.Dense(10, S.Softmax, **kw(g)))
    if name == "cfg3":
        rk = {"recurrentInitializer": S.Orthogonal(seed=seed)} if hasattr(S, "Engine") else {"recurrent_initializer": S.Orthogonal(seed=seed)}
        return S.LSTM(256, **kw(g), **rk) >> S.Dense(1, S.Tanh, **kw(g))
    if name == "cfg4":
        m = S.Dense(4096, S.ReLU(), **kw(g))
        for _ in range(3):
            m = m >> S.Dense(4096, S.ReLU(), **kw(g))
        return m >> S.Dense(10, S.Softmax, **kw(g))
    if name == "cfg5":
        s2 = dict(strides=(2, 2))
        return (S.Conv2D(64, activation=S.ReLU(), **kw(g)) >> S.BatchNorm() >> S.Pool2D(**s2)
                >> S.Conv2D(128, activation=S.ReLU(), **kw(g)) >> S.BatchNorm() >> S.Pool2D(**s2)
                >> S.Conv2D(256, activation=S.ReLU(), **kw(g)) >> S.BatchNorm() >> S.Pool2D(**s2)
                >> S.Flatten >> S.Dense(10, S.Softmax, **kw(g)))
    raise ValueError(name)


def lenet(S, seed=1):
    return build_model(S, "cfg2", seed)


def lenet_oracle(O, seed=1):
    return build_model(O, "cfg2", seed)


def cpu_port_samples_per_sec(max_steps, budget_s):
    """The reference's CPU path cannot run here; its stand-in is the oracle's torch-CPU port of the SAME
    train step (oracle/torch_port.py: oneDNN / MKL kernels, all host threads -- the kernel class TF-CPU uses), pinned against the
    NumPy oracle by tests/test_oracle_goldens.py.  Returns (samples/s, steps, seconds, description)."""
    import oracle as O
    om = lenet_oracle(O)
    alg = O.Adam(0.001)
    _, mp, op = O.init_params(om, alg, (BATCH,) + IN_SHAPE)
    x, y = synth(BATCH * 4, 1235)
    try:
        import torch
        from oracle.torch_port import TorchLeNetPort
        port = TorchLeNetPort(mp, rate=0.001, threads=os.cpu_count())
        step = lambda j, t: port.train_step(x[j * BATCH:(j + 1) * BATCH], y[j * BATCH:(j + 1) * BATCH], t)
        what = f"torch {torch.__version__} CPU port (oneDNN/MKL, {torch.get_num_threads()} threads)"
    except Exception:  # torch unavailable: the NumPy/BLAS oracle itself
        lm = O.LossModel(om, O.CategoricalCrossentropy)
        state = {"mp": mp, "op": op}

        def step(j, t):
            state["mp"], state["op"], _ = O.train_step(lm, alg, x[j * BATCH:(j + 1) * BATCH], y[j * BATCH:(j + 1) * BATCH],
                                                      state["mp"], state["op"], t)
        what = "NumPy/BLAS oracle port"
    for w in range(3):
        step(w % 4, w + 1)  # warm-up (thread pools, allocations, oneDNN primitive cache)
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (time.perf_counter() - t0) < budget_s:
        step(n % 4, n + 4)
        n += 1
    dt = time.perf_counter() - t0
    return n * BATCH / dt, n, dt, what


class ClockSampler:
    """SM clock / throttle reasons of one GPU, sampled IN-PROCESS through NVML by a thread that is started well before the first
    timed window (no process is spawned anywhere near a timed region).  `mark(on)` brackets the timed windows: `sm_mhz` is the
    median over the samples taken while a window was open (falls back to all samples of the run)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, period_s=0.002):
        self.gpu, self.period = gpu_index, period_s
        self.samples, self.in_window, self.stop_flag, self.thread, self.err = [], False, False, None, None
        self.nvml = self.h = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # CUDA_VISIBLE_DEVICES may renumber devices: match by PCI bus id of the CUDA device
            try:
                import torch
                bus = torch.cuda.get_device_properties(self.gpu).pci_bus_id
                dom = torch.cuda.get_device_properties(self.gpu).pci_domain_id
                dev = torch.cuda.get_device_properties(self.gpu).pci_device_id
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08X}:{bus:02X}:{dev:02X}.0".encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as ex:  # no NVML: say so instead of inventing numbers
            self.err = f"{type(ex).__name__}: {ex}"

    def _run(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((self.in_window, mhz, rs))
            except Exception as ex:
                self.err = f"{type(ex).__name__}: {ex}"
                return
            time.sleep(self.period)

    def mark(self, on):
        self.in_window = bool(on)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"NVML unavailable ({self.err})"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=2)
        win = [s for s in self.samples if s[0]] or self.samples
        reasons = set()
        for _, _, rs in win:
            for bit, name in self.REASONS.items():
                if rs & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median([s[1] for s in win])) if win else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(win), "samples_total": len(self.samples),
                "how": "in-process NVML thread, 2 ms period, samples taken while a timed window was open"}


def make_config(world, collective="p2p", precision="tf32"):
    """the `config` object of BOTH arms (ours and --impl reference): identical keys and values, so the driver's same_config
    check compares like with like; arm-specific facts (precision, CUDA graphs) live in `dtype` / `impl_detail`"""
    rows = ROWS_TOTAL // world
    return {"workload": WORKLOAD, "global_batch": BATCH * world, "rows_per_rank": rows, "steps_per_epoch": rows // BATCH,
            "averaging": "none (1 worker)" if world == 1 else f"model averaging every {rows // BATCH} steps (epoch end)",
            "l2": f"inputs larger than L2: the shard ({rows * int(np.prod(IN_SHAPE)) * 4 / 1e6:.0f} MB per rank) is streamed batch by "
                  "batch, every step reads rows no earlier step touched; the step's own activations are L2-resident by nature",
            "flop_per_sample": FLOP_PER_SAMPLE}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  The reference (Scala + TF-Java + Spark)
    cannot run here, so this times the oracle port with all host threads on bounded samples:
    windows of `steps` train steps, repeated until >= 2 s have been timed; value = median window."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    windows, total_t, total_n, what = [], 0.0, 0, ""
    deadline = time.perf_counter() + 150.0
    cpu_port_samples_per_sec(max_steps=max(1, args.warmup), budget_s=30.0)  # warm-up steps (thread pools, oneDNN primitives)
    while (total_t < 2.0 or len(windows) < 3) and time.perf_counter() < deadline:
        sps, n, dt, what = cpu_port_samples_per_sec(max_steps=max(1, args.steps), budget_s=60.0)
        windows.append(sps)
        total_t += dt
        total_n += n
    sps = float(np.median(windows))
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * BATCH / sps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.gpus),
        "impl_detail": {"note": "the reference (Scala + TF-Java 0.5.0 + Spark) cannot run in this image (no JVM); this is the oracle's "
                                "torch-CPU port of the same train step on the host cores, ONE worker (rank 0) whatever --gpus says",
                        "windows": len(windows), "window_spread": (max(windows) - min(windows)) / sps if windows else None},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{len(windows)} windows of {args.steps} train steps of batch {BATCH} ({total_n} steps, "
                                   f"{total_t:.1f} s timed), median window; {what}"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch

    import scanet3_b200 as S
    from scanet3_b200 import _native as N
    from scanet3_b200.distributed import RankAverager, init_process_group_from_env

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = local_rank = 0
    dist = None
    if world > 1:
        import torch.distributed as dist
        rank, world = init_process_group_from_env()
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun for N>1")
    torch.cuda.set_device(local_rank)
    precision = {"fp32": N.PREC_FP32, "tf32": N.PREC_TF32, "3xtf32": N.PREC_3XTF32}[args.precision]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # seconds before the first timed window (setup, warm-up and the parity check come first)

    # ---- cross-GPU averaging parity, OUTSIDE any timed region: the 3-epoch trajectory of tests/multi_gpu_check.py through the
    #      same RankAverager the timed run uses, against the oracle's k-partition run (the oracle is only the checker here) ----
    parity = None
    if world > 1 and not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from multi_gpu_check import run_check
        parity = [run_check(mode, rank, world, local_rank, log=sys.stderr) for mode in ("spark_chain", "mean")]  # stdout: the JSON line only

    rows = ROWS_TOTAL // world
    steps_per_epoch = rows // BATCH
    K = args.steps
    x, y = synth(rows, 1234 + 2 + 1000 * rank)  # seed = 1234 + config index (+ a per-rank stream for its shard)
    loss = S.MeanSquaredError if CFG["loss"] == "mse" else S.CategoricalCrossentropy
    eng = S.Engine(build_model(S, args.config), loss, S.Adam(CFG["rate"]), BATCH, IN_SHAPE, (CLASSES,), device=local_rank,
                   precision=precision, use_graph=True)
    eng.init_params()
    eng.upload_dataset(x, y)
    eng.warm()  # every CUDA graph the timed regions replay is captured, instantiated and uploaded here, not inside them
    coll = N.COLL_NCCL if args.collective == "nccl" else N.COLL_P2P
    avg = RankAverager(eng, N.AVG_SPARK_CHAIN if coll == N.COLL_P2P or world <= 4 else N.AVG_MEAN, coll) if world > 1 else None
    stream = torch.cuda.ExternalStream(eng.stream(), device=local_rank)

    state = {"iter": 0, "batch": 0, "averagings": 0, "dev_iter": False}

    def epoch_average():
        """epoch boundary: the fused P2P kernel synchronises the ranks on the device (flag barriers over NVLink), combines the
        losses and adds up the iteration counts there, so nothing here blocks the host; the NCCL arm keeps its host barriers."""
        state["averagings"] += 1
        if coll == N.COLL_P2P:
            state["dev_iter"] = True   # from now on the device's own step counter is the global one
            avg.average_async(steps_per_epoch)
            return world * steps_per_epoch  # host mirror of the device-side sum: the shards are equal by construction
        return avg.average(0.0, steps_per_epoch)[1]

    def at_epoch_end():
        state["batch"] = 0
        if avg is not None:
            total = epoch_average()
            state["iter"] += total - steps_per_epoch  # the step counter is global (Optimizer.scala:88)

    def run_steps(n):
        """n resident steps; epoch boundaries trigger the model averaging (weak scaling: K steps per rank)."""
        done = 0
        while done < n:
            m = min(n - done, steps_per_epoch - state["batch"])
            eng.train_steps_async(N.ITER_FROM_DEVICE if state["dev_iter"] else state["iter"], state["batch"], m)
            state["iter"] += m
            state["batch"] += m
            done += m
            if state["batch"] == steps_per_epoch:
                at_epoch_end()

    xh = torch.from_numpy(x[:steps_per_epoch * BATCH]).pin_memory()
    yh = torch.from_numpy(y[:steps_per_epoch * BATCH]).pin_memory()
    xb, yb = BATCH * int(np.prod(IN_SHAPE)) * 4, BATCH * CLASSES * 4
    loss_h = torch.zeros(max(K, steps_per_epoch) + 64, dtype=torch.float32).pin_memory()

    def run_steps_e2e(n):
        # the public host-fed call: consecutive pinned batches, H2D on the copy stream + one loss D2H per step, up to the epoch end
        j = 0
        while j < n:
            b = state["batch"]
            m = min(n - j, steps_per_epoch - b)
            # host-fed steps carry their global iteration explicitly; after a device-side averaging the host's mirror of the
            # counter (state["iter"], advanced by the summed counts) is the same number
            eng.train_batches_async_ptr(xh.data_ptr() + b * xb, yh.data_ptr() + b * yb, m, state["iter"],
                                        (0 if os.environ.get("SB_BENCH_NO_LOSS") else loss_h.data_ptr() + 4 * j))
            state["dev_iter"] = False
            state["iter"] += m
            state["batch"] += m
            j += m
            if state["batch"] == steps_per_epoch:
                at_epoch_end()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # Where a timed window starts inside the epoch.  N > 1: the window straddles the epoch boundary (K/2 steps, the fused
    # averaging, K/2 steps), so every window contains the exchange step the metric is defined with; N = 1: no exchange exists,
    # windows simply follow one another through the shard.
    straddle = world > 1 and K < steps_per_epoch
    start_batch = (steps_per_epoch - K // 2) % steps_per_epoch if straddle else None

    def position(runner):
        if start_batch is None:
            return
        gap = (start_batch - state["batch"]) % steps_per_epoch
        if gap:
            runner(gap)

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_windows(runner, sync_engine, reps_cap):
        """`K` steps per window, device-timed on the engine's stream between barrier + synchronize pairs, max over ranks per
        window; returns the per-window milliseconds and the number of averagings that fell inside the windows"""
        out, avgs = [], 0
        budget_t0 = time.perf_counter()
        for r in range(reps_cap):
            position(runner)
            barrier()
            a0 = state["averagings"]
            if rank == 0:
                sampler.mark(True)
            ev0.record(stream)
            runner(K)
            if sync_engine:
                eng.sync()  # compute stream AND the copy stream carrying the loss read-backs: the D2H reads are inside the window
            ev1.record(stream)
            barrier()
            if rank == 0:
                sampler.mark(False)
            out.append(max_over_ranks(ev0.elapsed_time(ev1)))
            avgs += state["averagings"] - a0
            # every rank takes the same decision (the clock is rank 0's)
            stop = 1.0 if (r + 1 >= args.min_repeats and time.perf_counter() - budget_t0 > args.window_budget_s) else 0.0
            if dist is not None:
                t = torch.tensor([stop], device=f"cuda:{local_rank}")
                dist.broadcast(t, 0)
                stop = float(t.item())
            if stop:
                break
        return out, avgs

    W = max(args.warmup, 3)
    # ---- device-timed region: inputs resident in HBM ----
    run_steps(W)
    if world > 1:  # one full epoch incl. its averaging before any clock starts (module load of the averaging kernel, peer mappings)
        run_steps(steps_per_epoch)
    barrier()
    launches0 = eng.launch_count()
    wins, avgs_in = timed_windows(run_steps, False, args.max_repeats)
    launches = (eng.launch_count() - launches0)
    ms = float(np.median(wins))
    value = world * K * BATCH / (ms / 1000.0)

    # ---- end to end through the public C-ABI call with HOST buffers (pinned), H2D + loss D2H every step ----
    run_steps_e2e(W)
    if world > 1:
        run_steps_e2e(steps_per_epoch)
    barrier()
    t0 = time.perf_counter()
    e2e_wins, e2e_avgs = timed_windows(run_steps_e2e, True, args.max_repeats)
    e2e_wall = time.perf_counter() - t0
    e2e_ms = float(np.median(e2e_wins))
    e2e_value = world * K * BATCH / max(e2e_ms / 1000.0, 1e-9)
    eng.sync()
    last_loss = float(loss_h[0].item())

    # ---- after the timed runs: finish the epoch, average, and compare the ranks' arenas (they must be bit-identical) ----
    arena_equal = None
    if avg is not None:
        run_steps(steps_per_epoch - state["batch"])
        eng.sync()
        a = eng.arena_tensor()
        cks = torch.stack([a.double().sum(), (a.double() * a.double()).sum()])
        got = [torch.zeros_like(cks) for _ in range(world)]
        dist.all_gather(got, cks)
        arena_equal = all(bool(torch.equal(g, got[0])) for g in got)

    # ---- the exchange step alone: fused P2P averaging kernel (flag barriers + reduce-scatter + all-gather over NVLink) ----
    averaging = None
    if avg is not None and coll == N.COLL_P2P:
        reps = 10
        avg.average_async(steps_per_epoch)
        barrier()
        ev0.record(stream)
        for _ in range(reps):
            avg.average_async(steps_per_epoch)
        ev1.record(stream)
        barrier()
        a_ms = max_over_ranks(ev0.elapsed_time(ev1) / reps)
        arena_bytes = eng.arena()[1] * 4
        bus = 2.0 * (world - 1) / world * arena_bytes / (a_ms / 1000.0) / 1e9  # per GPU, each direction carries (k-1)/k of it
        averaging = {"ms": a_ms, "arena_bytes": arena_bytes, "bus_GBps_per_gpu": bus, "per_direction_GBps": bus / 2.0,
                     "peak_per_direction_GBps": 770.0, "frac": bus / 2.0 / 770.0, "nccl_allreduce_busbw_ref_GBps": 725.0,
                     "frac_of_nccl_busbw": bus / 725.0,
                     "note": "one kernel per rank incl. two cross-GPU flag barriers, the loss combine and the iteration sum; bus = "
                             "2(k-1)/k x arena / time (NCCL's bus-bandwidth convention); references: measured NVLink peer copy 770 GB/s "
                             "per direction and 8-rank NCCL all-reduce bus bandwidth 725 GB/s at 1 GiB (B200_PROFILING.md)"}

    # ---- per-kernel-class device times over the same steps (eager launches bracketed by CUDA events on the engine
    #      stream) -> the dominant kernel and its roofline ----
    roofline = None
    prof_rows = []
    psteps = max(1, min(K, 50, steps_per_epoch))
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
        which = "of measured" if peaks else "of fallback"
        # TF32 denominator: cuBLAS TF32 8192^3 measured on this pool's B200 with MEASURED_PEAKS.json's protocol
        # (tools/measure_tf32_peak.py -> profiles/r1_tf32_peak.json); the step draws little power (clocks stay at max), so the
        # burst figure applies.  Fallback: half of the measured bf16 figure.
        tf32_peak, tf32_note = bf16 / 2.0, f"TF32 dense peak taken as 1/2 x bf16 cuBLAS sustained ({bf16:.0f} TF/s) {which}"
        try:
            tp = json.load(open(os.path.join(ROOT, "profiles", "r1_tf32_peak.json")))
            tf32_peak = float(tp["tf32_tflops"])
            tf32_note = f"cuBLAS TF32 8192^3 burst measured on this pool's B200 ({tf32_peak:.0f} TF/s; sustained {tp['tf32_tflops_sustained']:.0f})"
        except Exception:
            pass
        eng.profile_enable(True)
        eng.profile_reset()
        eng.train_steps_async(state["iter"], 0, psteps)
        eng.sync()
        prof_rows = sorted(eng.profile(), key=lambda r: -r["total_ms"])
        eng.profile_enable(False)
        tot = sum(r["total_ms"] for r in prof_rows) or 1.0
        top = prof_rows[0]
        n_launch = max(top["launches"], 1)                     # launches of the class over the profiled steps
        per_launch_s = top["total_ms"] / 1000.0 / n_launch     # average launch duration (CUDA events, engine stream)
        traffic = None
        for tf in ("r2_traffic.json", "r1_traffic.json"):
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", tf))).get(top["name"], {}).get("dram_bytes_per_launch")
                if traffic:
                    break
            except Exception:
                pass
        tensor_bound = top["flops"] > 0 and "tcgen05" in top["name"]
        if tensor_bound:
            ach = top["flops"] / n_launch / per_launch_s / 1e12
            peak = tf32_peak
            roofline = {"bound": "tensor", "kernel": top["name"], "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                        "frac": ach / peak, "traffic": traffic, "share_of_step": top["total_ms"] / tot, "peak_note": tf32_note}
        else:
            ach = top["bytes"] / n_launch / per_launch_s / 1e9
            roofline = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": hbm, "unit": "GB/s",
                        "frac": ach / hbm, "traffic": traffic, "share_of_step": top["total_ms"] / tot,
                        "peak_note": f"HBM copy bandwidth {which}"}
        # every tensor-core class of the step against the same TF32 figure (the north-star's own yardstick for the contractions)
        roofline["tensor_classes"] = [
            {"kernel": r["name"], "avg_launch_us": r["total_ms"] * 1e3 / max(r["launches"], 1),
             "achieved_TFLOPs": r["flops"] / max(r["launches"], 1) / (r["total_ms"] / 1e3 / max(r["launches"], 1)) / 1e12,
             "frac": r["flops"] / max(r["launches"], 1) / (r["total_ms"] / 1e3 / max(r["launches"], 1)) / 1e12 / tf32_peak}
            for r in prof_rows if r["flops"] > 0 and "tcgen05" in r["name"] and r["total_ms"] > 0]
        roofline["avg_launch_us"] = per_launch_s * 1e6
        roofline["launches_per_step"] = n_launch / max(psteps, 1)
        roofline["algorithmic_per_launch"] = (top["flops"] if tensor_bound else top["bytes"]) / n_launch
        roofline["traffic_note"] = ("dram bytes per launch from profiles/ (one ncu --set full capture, cold L2)"
                                    if traffic else "no ncu capture of this kernel class")
        # whole step against the same HBM figure: the algorithmic bytes of every kernel class of one step (what the engine
        # attributes to each launch: logical inputs read once + outputs written once) over the graph-replayed step time
        try:
            step_bytes = sum(float(r["bytes"]) for r in prof_rows) / max(psteps, 1)
            step_s = ms / K / 1000.0
            roofline["whole_step"] = {"algorithmic_bytes": step_bytes, "achieved": step_bytes / step_s / 1e9, "unit": "GB/s",
                                      "frac": step_bytes / step_s / 1e9 / hbm,
                                      "note": "sum of the kernel classes' algorithmic bytes per step / ms_per_step of the timed region"}
        except Exception:
            pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config == "cfg2":
        sps, n, dt, what = cpu_port_samples_per_sec(max_steps=2000, budget_s=15.0)
        cpu = {"value": sps, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{n} train steps of batch {BATCH} ({dt:.1f} s) of the same workload, {what}"}

    if rank == 0:
        clocks = sampler.stop()
        spread = lambda w: {"windows": len(w), "median_ms": float(np.median(w)), "min_ms": float(min(w)), "max_ms": float(max(w)),
                            "p10_ms": float(np.percentile(w, 10)), "p90_ms": float(np.percentile(w, 90))}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if precision == N.PREC_FP32 else ("f32 storage, tf32 tensor-core contractions (fp32 accumulate)"
                                                            if precision == N.PREC_TF32 else "f32 (3xTF32 contractions)"),
            "data": "synthetic",
            "config": make_config(world, args.collective, args.precision),
            "impl_detail": {"precision": args.precision, "cuda_graph": True, "collective": args.collective if world > 1 else None,
                            "timing": f"every window = exactly {K} steps between barrier+synchronize pairs, CUDA events on the engine's "
                                      "stream, max over ranks per window; value = median window; all CUDA graphs captured before "
                                      "the first window (sb_warm)" + ("; every window straddles an epoch boundary: K/2 steps, the "
                                      "fused averaging, K/2 steps" if straddle else "")},
            "windows": spread(wins), "averagings_in_timed_region": avgs_in,
            "averagings_per_window": avgs_in / max(len(wins), 1),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": xb + yb, "d2h_bytes_per_step": 4,
                    "wall_s": e2e_wall, "last_loss": last_loss, "windows": spread(e2e_wins), "averagings_in_timed_region": e2e_avgs},
            "gpu_launches": int(round(launches / max(len(wins), 1))), "gpu_launches_all_windows": int(launches),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "averaging": averaging,
            "parity": parity, "arena_equal_across_ranks": arena_equal,
            "model_tflops": value * FLOP_PER_SAMPLE / 1e12,
            "kernels": [{"name": r["name"], "ms_per_step": r["total_ms"] / psteps,
                         "launches_per_step": r["launches"] / psteps} for r in prof_rows[:16]],
        }
        if value < 0.98 * e2e_value:
            line["warning"] = (f"device-timed value {value:.0f} < 0.98 x e2e {e2e_value:.0f}: a steady state cannot be slower resident "
                               "than host-fed -- one of the two windows is contaminated")
        if parity is not None and not all(p["ok"] for p in parity):
            line["warning_parity"] = "cross-GPU averaging parity check FAILED"
        if arena_equal is False:
            line["warning_arena"] = "ranks' arenas differ after the closing averaging"
        print(json.dumps(line), flush=True)
    if avg is not None:
        avg.close()
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=60)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32", "3xtf32"])
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the cross-GPU averaging parity check that runs before the timed region (N > 1)")
    ap.add_argument("--min-repeats", type=int, default=5, help="timed windows of --steps steps: at least this many ...")
    ap.add_argument("--max-repeats", type=int, default=200, help="... at most this many ...")
    ap.add_argument("--window-budget-s", type=float, default=1.5, help="... and no more once this much wall time went into them")
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE config (cfg2 = the metric's workload)")
    args = ap.parse_args()
    select_config(args.config)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
