"""CPU oracle for the scanet3 data-parallel train path.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy restatement of the reference's
algorithm (pashashiz/scanet3, `src/main/scala/scanet/...`) for the one hot path the
B200 engine accelerates: the per-worker minibatch train step
(`optimizers/Optimizer.scala:106-161`) plus the epoch-end pairwise model averaging
(`optimizers/Optimizer.scala:196-230`).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it, and only as the checker or the timed CPU baseline.
The product (`scanet3_b200/`, `libscanet_b200.so`) never imports or links it.

Why a restatement: the reference is Scala over TensorFlow-Java 0.5.0 (TF 2.10 C++
kernels, `build.sbt:13`) and Spark 3.3.1 (`build.sbt:16`); neither a JVM nor those
artefacts exist offline, so the reference cannot be executed here.

Pinning status (see tests/test_oracle_goldens.py):
  * PINNED by the reference's own golden vectors: Dense/Bias/Activate/Flatten forward,
    BCE/MSE losses + gradients, Conv2D fwd (SAME/VALID/stride 2/EXPLICIT) + wgrad + dgrad,
    Max/Avg pool fwd + grads (first-max tie rule), generic BatchNorm train/inference,
    SimpleRNN and LSTM forward, softmax/tanh/ReLU, TF-Philox seeded initialisers,
    decayingAvg/boost/sqrtZeroSafe, TensorIterator batching, and the convergence
    thresholds of all 8 optimizers on the reference's deterministic CSV fixtures.
  * PARITY UNPINNED (no reference test exists): fused (rank-4) BatchNorm values and
    gradients, LSTM gradients (checked here against float64 finite differences only),
    exact per-step optimizer numbers, and multi-partition (k>1) averaging order.
"""
from .philox import philox_uniform, philox_normal, philox_truncated_normal  # noqa: F401
from .initializers import *  # noqa: F401,F403
from .layers import *  # noqa: F401,F403
from .losses import *  # noqa: F401,F403
from .optimizers import *  # noqa: F401,F403
from .train import *  # noqa: F401,F403
from . import estimators  # noqa: F401,E402
