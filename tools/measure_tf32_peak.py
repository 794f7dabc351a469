#!/usr/bin/env python
"""cuBLAS TF32 GEMM throughput on this GPU, measured the way MEASURED_PEAKS.json measures bf16 (torch.matmul 8192^3, best of 10 =
burst; back to back for 4 s = sustained).  Writes profiles/r1_tf32_peak.json -- the denominator bench.py uses for tensor-bound
kernels when present (otherwise 1/2 of the measured bf16 figure)."""
import json
import os
import time

import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
for _ in range(3):
    (a @ b)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    (a @ b)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n ** 3 / (best / 1e3) / 1e12
t0 = time.time()
it = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(20):
        (a @ b)
    it += 20
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
sustained = 2 * n ** 3 * it / (e0.elapsed_time(e1) / 1e3) / 1e12
out = {"tf32_tflops": burst, "tf32_tflops_sustained": sustained, "gpu_name": torch.cuda.get_device_name(0),
       "how": "torch.matmul fp32 inputs with allow_tf32 (cuBLAS), 8192^3: best of 10 (burst) and back to back for 4 s (sustained)",
       "torch": torch.__version__}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r1_tf32_peak.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out))
