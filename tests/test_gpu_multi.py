"""Multi-GPU averaging parity (one process per GPU, torchrun).  Skipped on a box with a single GPU; the single-GPU
k-engine tests in test_gpu_averaging.py cover the same arithmetic through the in-process group."""
import os
import subprocess
import sys

import pytest

import scanet3_b200 as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["spark_chain", "mean", "nccl"])
def test_ranks_average_like_the_oracle(mode):
    n = S.native.lib().sb_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    k = 2 if n < 4 else (4 if n < 8 else 8)
    if mode == "nccl" and k > 4:
        k = 4  # the NCCL baseline arm cannot reproduce the grouped Spark chain; its mean is covered at k = 4
    port = 29600 + {"spark_chain": 0, "mean": 1, "nccl": 2}[mode]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={k}", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py"), mode],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    try:  # keep the log (profiles/ is committed; gpurun_out/ travels back from the GPU box)
        for d in ("profiles", "gpurun_out"):
            os.makedirs(os.path.join(ROOT, d), exist_ok=True)
            with open(os.path.join(ROOT, d, f"r2_multi_gpu_check_{mode}_k{k}.log"), "w") as f:
                f.write("\n".join(l for l in r.stdout.splitlines() if "MULTI_GPU_CHECK" in l) + "\n")
    except OSError:
        pass
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-3000:]
