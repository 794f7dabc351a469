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
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  The reference (Scala + TF-Java + Spark)
    cannot run here, so this times the oracle port with all host threads on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step_budget = 240.0 / max(1, args.steps + args.warmup)
    sps, n, dt, what = cpu_port_samples_per_sec(max_steps=max(1, args.steps), budget_s=min(120.0, per_step_budget * args.steps))
    line = {
        "impl": "reference", "metric": "train samples/sec (MNIST-shape CNN)", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(n, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "bounded sample of the same train step on the host cores"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{n} train steps of batch {BATCH} ({dt:.1f} s), {what}"},
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

    rows = ROWS_TOTAL // world
    steps_per_epoch = rows // BATCH
    x, y = synth(rows, 1234 + 2 + 1000 * rank)  # seed = 1234 + config index (+ a per-rank stream for its shard)
    loss = S.MeanSquaredError if CFG["loss"] == "mse" else S.CategoricalCrossentropy
    eng = S.Engine(build_model(S, args.config), loss, S.Adam(CFG["rate"]), BATCH, IN_SHAPE, (CLASSES,), device=local_rank,
                   precision=precision, use_graph=True)
    eng.init_params()
    eng.upload_dataset(x, y)
    coll = N.COLL_NCCL if args.collective == "nccl" else N.COLL_P2P
    avg = RankAverager(eng, N.AVG_SPARK_CHAIN if coll == N.COLL_P2P or world <= 4 else N.AVG_MEAN, coll) if world > 1 else None
    stream = torch.cuda.ExternalStream(eng.stream(), device=local_rank)

    state = {"iter": 0, "batch": 0}

    def epoch_average():
        """epoch boundary: the fused P2P kernel synchronises the ranks on the device (flag barriers over NVLink) and combines
        the losses there, so nothing here blocks the host; the NCCL arm keeps its host-side barriers."""
        if coll == N.COLL_P2P:
            return avg.average_async(steps_per_epoch)
        return avg.average(0.0, steps_per_epoch)[1]

    def run_steps(n):
        """n resident steps; epoch boundaries trigger the model averaging (weak scaling: K steps per rank)."""
        done = 0
        while done < n:
            m = min(n - done, steps_per_epoch - state["batch"])
            eng.train_steps_async(state["iter"], state["batch"], m)
            state["iter"] += m
            state["batch"] += m
            done += m
            if state["batch"] == steps_per_epoch:
                state["batch"] = 0
                if avg is not None:
                    total = epoch_average()
                    state["iter"] += total - steps_per_epoch  # the step counter is global (Optimizer.scala:88)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed region: inputs resident in HBM ----
    run_steps(max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    run_steps(args.steps)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps * BATCH / (ms / 1000.0)

    # ---- end to end through the public C-ABI call with HOST buffers (pinned), H2D + loss D2H every step ----
    xh = torch.from_numpy(x[:steps_per_epoch * BATCH]).pin_memory()
    yh = torch.from_numpy(y[:steps_per_epoch * BATCH]).pin_memory()
    e2e_warm = max(args.warmup, 19)  # at least two 8-step groups (one graph per staging buffer) + single steps get captured untimed
    loss_h = torch.zeros(args.steps + e2e_warm + 8, dtype=torch.float32).pin_memory()
    xb, yb = BATCH * int(np.prod(IN_SHAPE)) * 4, BATCH * CLASSES * 4

    def run_steps_e2e(n, slot0):
        # the public host-fed call: consecutive pinned batches, H2D on the copy stream + one loss D2H per step, up to the epoch end
        j = 0
        while j < n:
            b = state["batch"]
            m = min(n - j, steps_per_epoch - b)
            eng.train_batches_async_ptr(xh.data_ptr() + b * xb, yh.data_ptr() + b * yb, m, state["iter"],
                                        (0 if os.environ.get("SB_BENCH_NO_LOSS") else loss_h.data_ptr() + 4 * (slot0 + j)))
            state["iter"] += m
            state["batch"] += m
            j += m
            if state["batch"] == steps_per_epoch:
                state["batch"] = 0
                if avg is not None:
                    total = epoch_average()
                    state["iter"] += total - steps_per_epoch
    run_steps_e2e(e2e_warm, 0)
    barrier()
    t0 = time.perf_counter()
    ev0.record(stream)
    run_steps_e2e(args.steps, e2e_warm)
    eng.sync()  # compute stream AND the copy stream that carries the loss read-backs: the D2H reads are inside the timed region
    ev1.record(stream)
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = max(ev0.elapsed_time(ev1), 0.0)
    if dist is not None:
        t = torch.tensor([e2e_ms, e2e_wall * 1000.0], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_wall = float(t[0].item()), float(t[1].item()) / 1000.0
    e2e_value = world * args.steps * BATCH / max(e2e_ms / 1000.0, 1e-9)
    last_loss = float(loss_h[e2e_warm + args.steps - 1].item())

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
        a_ms = ev0.elapsed_time(ev1) / reps
        t = torch.tensor([a_ms], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        a_ms = float(t.item())
        arena_bytes = eng.arena()[1] * 4
        bus = 2.0 * (world - 1) / world * arena_bytes / (a_ms / 1000.0) / 1e9  # per GPU, each direction carries (k-1)/k of it
        averaging = {"ms": a_ms, "arena_bytes": arena_bytes, "bus_GBps_per_gpu": bus, "per_direction_GBps": bus / 2.0,
                     "peak_per_direction_GBps": 770.0, "frac": bus / 2.0 / 770.0, "nccl_allreduce_busbw_ref_GBps": 725.0,
                     "frac_of_nccl_busbw": bus / 725.0,
                     "note": "one kernel per rank incl. two cross-GPU flag barriers and the loss combine; bus = 2(k-1)/k x arena / time "
                             "(NCCL's bus-bandwidth convention); references: measured NVLink peer copy 770 GB/s per direction and "
                             "8-rank NCCL all-reduce bus bandwidth 725 GB/s at 1 GiB (B200_PROFILING.md)"}

    # ---- per-kernel-class device times over the same steps (eager launches bracketed by CUDA events on the engine
    #      stream) -> the dominant kernel and its roofline ----
    roofline = None
    prof_rows = []
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
        psteps = max(1, min(args.steps, 50, steps_per_epoch))
        eng.train_steps_async(state["iter"], 0, psteps)
        eng.sync()
        prof_rows = sorted(eng.profile(), key=lambda r: -r["total_ms"])
        eng.profile_enable(False)
        tot = sum(r["total_ms"] for r in prof_rows) or 1.0
        top = prof_rows[0]
        n_launch = max(top["launches"], 1)                     # launches of the class over the profiled steps
        per_launch_s = top["total_ms"] / 1000.0 / n_launch     # average launch duration (CUDA events, engine stream)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(top["name"], {}).get("dram_bytes_per_launch")
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
        roofline["avg_launch_us"] = per_launch_s * 1e6
        roofline["launches_per_step"] = n_launch / max(psteps, 1)
        roofline["algorithmic_per_launch"] = (top["flops"] if tensor_bound else top["bytes"]) / n_launch
        roofline["traffic_note"] = ("dram bytes per launch from profiles/r1_traffic.json (one ncu --set full capture, cold L2)"
                                    if traffic else "no ncu capture of this kernel class")
        # whole step against the same HBM figure: the algorithmic bytes of every kernel class of one step (what the engine
        # attributes to each launch: logical inputs read once + outputs written once) over the graph-replayed step time
        try:
            step_bytes = sum(float(r["bytes"]) for r in prof_rows) / max(psteps, 1)
            step_s = ms / args.steps / 1000.0
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
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if precision == N.PREC_FP32 else ("f32 storage, tf32 tensor-core contractions (fp32 accumulate)"
                                                            if precision == N.PREC_TF32 else "f32 (3xTF32 contractions)"),
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "precision": args.precision, "steps_per_epoch": steps_per_epoch,
                       "averaging": "none (1 worker)" if world == 1 else f"{args.collective} every {steps_per_epoch} steps",
                       "l2": f"shard {rows * int(np.prod(IN_SHAPE)) * 4 / 1e6:.0f} MB per rank is streamed batch by batch (cfg2: 188 MB/N > L2); "
                             "per-step activations are L2-resident by nature",
                       "cuda_graph": True, "flop_per_sample": FLOP_PER_SAMPLE},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": xb + yb, "d2h_bytes_per_step": 4,
                    "wall_s": e2e_wall, "last_loss": last_loss},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "averaging": averaging,
            "model_tflops": value * FLOP_PER_SAMPLE / 1e12,
            "kernels": [{"name": r["name"], "ms_per_step": r["total_ms"] / max(1, min(args.steps, 50, steps_per_epoch)),
                         "launches_per_step": r["launches"] / max(1, min(args.steps, 50, steps_per_epoch))} for r in prof_rows[:16]],
        }
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
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE config (cfg2 = the metric's workload)")
    args = ap.parse_args()
    select_config(args.config)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
