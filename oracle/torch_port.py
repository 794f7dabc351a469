"""CPU port of the cfg2 (LeNet) train step on torch CPU ops (oneDNN / MKL, all host threads) -- TEST INFRASTRUCTURE and the timed
CPU baseline of bench.py only; the product never imports it.

Why it exists: the reference's own CPU path (TensorFlow-Java 0.5.0 = TF 2.10 Eigen/oneDNN kernels under Spark local[*]) cannot run
here (no JVM).  The NumPy oracle (oracle/train.py) is written for exactness checks, not speed; this port
runs the SAME step -- Conv2D/ReLU/MaxPool(2x2, stride 1)/Conv2D/ReLU/MaxPool/Flatten/Dense/Bias/Softmax, CategoricalCrossentropy
with eps = 1e-8 inside the log, the reference's Adam (eps inside the sqrt, after bias correction) -- on the same class of CPU
kernels TF-CPU uses, so the reported CPU baseline is not handicapped by im2col in NumPy.  Pinned against the NumPy oracle by
tests/test_oracle_goldens.py::test_torch_port_matches_numpy_oracle.

Reference lines: models/layer/{Conv2D,Pool2D,Dense,Bias,Flatten}.scala, models/Activation.scala:63-83, models/Loss.scala:44-57,
optimizers/Adam.scala:29-41, optimizers/Optimizer.scala:169-194."""
import numpy as np
import torch
import torch.nn.functional as F


class TorchLeNetPort:
    def __init__(self, params, rate=0.001, beta1=0.9, beta2=0.999, eps=1e-7, threads=None):
        """params: the oracle's parameter dict ('0/weights' HWIO conv1, '3/weights' HWIO conv2, '7/weights' [30976,10], '8/weights' [10])"""
        if threads:
            torch.set_num_threads(int(threads))
        self.rate, self.b1, self.b2, self.eps = rate, beta1, beta2, eps
        self.p = {k: torch.tensor(np.asarray(v), dtype=torch.float32, requires_grad=True) for k, v in params.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.p.items()}

    def forward_loss(self, x, y):
        """x: [B,28,28,1] NHWC float32, y: [B,10] one-hot"""
        p = self.p
        h = torch.as_tensor(x).permute(0, 3, 1, 2)                                     # NCHW for oneDNN
        h = F.conv2d(h, p["0/weights"].permute(3, 2, 0, 1))                            # HWIO -> OIHW, VALID, stride 1
        h = F.max_pool2d(torch.clamp_min(h, 0.0), kernel_size=2, stride=1)             # ReLU(0) >> Pool2D((2,2),(1,1),Valid)
        h = F.conv2d(h, p["3/weights"].permute(3, 2, 0, 1))
        h = F.max_pool2d(torch.clamp_min(h, 0.0), kernel_size=2, stride=1)
        h = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)                              # Flatten in NHWC order
        z = h @ p["7/weights"] + p["8/weights"]
        e = torch.exp(z)                                                               # un-stabilised softmax (Activation.scala:63-69)
        prob = e / e.sum(dim=1, keepdim=True)
        yt = torch.as_tensor(y)
        return -(yt * torch.log(prob + 1e-8)).sum() / yt.shape[0]                      # Loss.scala:44-57

    def train_step(self, x, y, t):
        loss = self.forward_loss(x, y)
        grads = torch.autograd.grad(loss, list(self.p.values()))
        tf = float(np.float32(t))
        with torch.no_grad():
            for (k, w), g in zip(self.p.items(), grads):
                self.m[k] = self.b1 * self.m[k] + (1.0 - self.b1) * g                  # decayingAvg
                self.v[k] = self.b2 * self.v[k] + (1.0 - self.b2) * g * g
                mhat = self.m[k] / (1.0 - self.b1 ** tf)                               # boost
                vhat = self.v[k] / (1.0 - self.b2 ** tf)
                w -= (self.rate / torch.sqrt(vhat + self.eps)) * mhat                  # eps inside the sqrt (Adam.scala:38-40)
        return float(loss.detach())

    def params_numpy(self):
        return {k: v.detach().numpy().copy() for k, v in self.p.items()}
