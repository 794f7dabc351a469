"""Losses restated from `src/main/scala/scanet/models/Loss.scala:10-58` (test infrastructure only).
Gradients follow autodiff through the reference's op graph (`math/alg/AllKernels.scala` localGrads)."""
import numpy as np

__all__ = ["MeanSquaredError", "BinaryCrossentropy", "CategoricalCrossentropy", "IdentityLoss"]


class _Loss:
    name = "?"

    def __repr__(self):
        return self.name


class _MSE(_Loss):
    """(predicted - expected).sqr.mean over ALL elements (Loss.scala:18-24)."""
    name = "MeanSquaredError"

    def value(self, p, y):
        d = p - y
        return np.mean(d * d, dtype=p.dtype)

    def grad(self, p, y):
        t = p.dtype.type
        return (t(2) * (p - y)) / t(p.size)


class _BCE(_Loss):
    """mean( -(y*log(p+eps)) - (1-y)*log(1-(p-eps)) ), eps = 1e-8f (Loss.scala:26-42)."""
    name = "BinaryCrossentropy"

    def value(self, p, y):
        t = p.dtype.type
        eps, one = t(np.float32(1e-8)), t(1)
        left = y * np.log(p + eps)
        right = (one - y) * np.log(one - (p - eps))
        return np.mean(-left - right, dtype=p.dtype)

    def grad(self, p, y):
        t = p.dtype.type
        eps, one = t(np.float32(1e-8)), t(1)
        n = t(p.size)
        return (-(y / (p + eps)) + (one - y) / (one - (p - eps))) / n


class _CCE(_Loss):
    """-sum_all( y*log(p+eps) ) / B, eps = 1e-8f, B = y.shape[0] (Loss.scala:44-57)."""
    name = "CategoricalCrossentropy"

    def value(self, p, y):
        t = p.dtype.type
        eps = t(np.float32(1e-8))
        return -np.sum(y * np.log(p + eps), dtype=p.dtype) / t(y.shape[0])

    def grad(self, p, y):
        t = p.dtype.type
        eps = t(np.float32(1e-8))
        return -(y / (p + eps)) / t(y.shape[0])


class _Identity(_Loss):
    """loss = predicted (Loss.scala:11-16); used by the x^2 fixture (SGDSpec.scala:16-31)."""
    name = "Identity"

    def value(self, p, y):
        return p.reshape(()) if p.size == 1 else p

    def grad(self, p, y):
        return np.ones_like(p)


MeanSquaredError = _MSE()
BinaryCrossentropy = _BCE()
CategoricalCrossentropy = _CCE()
IdentityLoss = _Identity()
