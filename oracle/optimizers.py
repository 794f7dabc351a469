"""The 8 update rules restated from `src/main/scala/scanet/optimizers/{SGD,Adam,AdaGrad,RMSProp,AdaDelta,
Adamax,Nadam,AMSGrad}.scala` with helpers `decayingAvg` / `boost` / `sqrtZeroSafe`
(`math/alg/AllKernels.scala:491-492, 541-549`).  Test infrastructure only.

`build(g, state, t)` returns `(delta, next_state)`; the caller applies `w - delta`
(`optimizers/Optimizer.scala:186`).  `t` is the GLOBAL step `globalIter + localIter + 1`
(`Optimizer.scala:143`), fed as int32 and cast to float.  Operation order is kept as written in the
reference because fp32 rounding depends on it."""
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np

from . import initializers as I

__all__ = ["SGD", "Adam", "AdaGrad", "RMSProp", "AdaDelta", "Adamax", "Nadam", "AMSGrad", "decaying_avg", "boost",
           "sqrt_zero_safe"]


def _c(x, like):
    """float constant `x.const.cast[T]`: the Scala literal is a Float, so round to fp32 first."""
    return like.dtype.type(np.float32(x))


def decaying_avg(avg, nxt, decay):
    """(decay*avg) + ((1-decay)*next), `1-decay` evaluated in T (alg/AllKernels.scala:541-544)."""
    one = avg.dtype.type(1)
    return (decay * avg) + ((one - decay) * nxt)


def boost(x, rate, t):
    """x / (1 - rate^float(t)) (alg/AllKernels.scala:546-549).  TF computes powf; restated as the
    correctly-rounded value via float64 (glibc powf is correctly rounded in all but rare cases)."""
    ty = x.dtype.type
    p = ty(np.float64(rate) ** np.float64(np.float32(t)))
    return x / (ty(1) - p)


def sqrt_zero_safe(x, eps):
    return np.sqrt(x + eps)


class Algorithm:
    def state_defs(self):
        """-> OrderedDict name -> initializer (all `aggregation = Some(Avg)`)."""
        raise NotImplementedError

    def build(self, g, state, t):
        raise NotImplementedError


@dataclass(frozen=True)
class SGD(Algorithm):
    rate: float = 0.01
    momentum: float = 0.0
    nesterov: bool = False

    def state_defs(self):
        return OrderedDict(velocity=I.Zeros)

    def build(self, g, s, t):
        m, r = _c(self.momentum, g), _c(self.rate, g)
        v = (r * g) - (m * s["velocity"])  # note the sign (SGD.scala:21)
        delta = (r * g) - (m * v) if self.nesterov else v
        return delta, {"velocity": v}


@dataclass(frozen=True)
class Adam(Algorithm):
    rate: float = 0.001
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-7

    def state_defs(self):
        return OrderedDict(momentum=I.Zeros, velocity=I.Zeros)

    def build(self, g, s, t):
        b1, b2, eps, r = _c(self.beta1, g), _c(self.beta2, g), _c(self.epsilon, g), _c(self.rate, g)
        m = decaying_avg(s["momentum"], g, b1)
        v = decaying_avg(s["velocity"], g * g, b2)
        # eps INSIDE the sqrt, applied after bias correction (Adam.scala:36-38)
        delta = (r / np.sqrt(boost(v, b2, t) + eps)) * boost(m, b1, t)
        return delta, {"momentum": m, "velocity": v}


@dataclass(frozen=True)
class AdaGrad(Algorithm):
    rate: float = 0.001
    epsilon: float = 1e-7

    def state_defs(self):
        return OrderedDict(grad_acc=I.Zeros)

    def build(self, g, s, t):
        acc = s["grad_acc"] + g * g
        delta = (_c(self.rate, g) / sqrt_zero_safe(acc, _c(self.epsilon, g))) * g
        return delta, {"grad_acc": acc}


@dataclass(frozen=True)
class RMSProp(Algorithm):
    rate: float = 0.001
    rho: float = 0.9
    epsilon: float = 1e-7

    def state_defs(self):
        return OrderedDict(avg_grad=I.Zeros)

    def build(self, g, s, t):
        a = decaying_avg(s["avg_grad"], g * g, _c(self.rho, g))
        delta = (_c(self.rate, g) / sqrt_zero_safe(a, _c(self.epsilon, g))) * g
        return delta, {"avg_grad": a}


@dataclass(frozen=True)
class AdaDelta(Algorithm):
    rate: float = 1.0
    rho: float = 0.9
    epsilon: float = 1e-7

    def state_defs(self):
        return OrderedDict(avg_grad=I.Zeros, avg_delta=I.Zeros)

    def build(self, g, s, t):
        rho, eps = _c(self.rho, g), _c(self.epsilon, g)
        a = decaying_avg(s["avg_grad"], g * g, rho)
        delta = _c(self.rate, g) * ((sqrt_zero_safe(s["avg_delta"], eps) / sqrt_zero_safe(a, eps)) * g)
        d = decaying_avg(s["avg_delta"], delta * delta, rho)
        return delta, {"avg_grad": a, "avg_delta": d}


@dataclass(frozen=True)
class Adamax(Algorithm):
    rate: float = 0.001
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-7
    init_acc: float = 0.001

    def state_defs(self):
        return OrderedDict(momentum_1=I.Const(self.init_acc), momentum_2=I.Const(self.init_acc))

    def build(self, g, s, t):
        b1, b2, eps, r = _c(self.beta1, g), _c(self.beta2, g), _c(self.epsilon, g), _c(self.rate, g)
        one = g.dtype.type(1)
        m1 = decaying_avg(s["momentum_1"], g, b1)
        m2 = np.maximum(b2 * s["momentum_2"], (one - b2) * np.abs(g))
        delta = (r * boost(m1, b1, t)) / (m2 + eps)
        return delta, {"momentum_1": m1, "momentum_2": m2}


@dataclass(frozen=True)
class Nadam(Algorithm):
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-7
    init_acc: float = 0.001

    def state_defs(self):
        return OrderedDict(momentum=I.Const(self.init_acc), velocity=I.Const(self.init_acc))

    def build(self, g, s, t):
        b1, b2, eps = _c(self.beta1, g), _c(self.beta2, g), _c(self.epsilon, g)
        one = g.dtype.type(1)
        m = decaying_avg(s["momentum"], g, b1)
        v = decaying_avg(s["velocity"], g * g, b2)
        m_nesterov = (b1 * boost(m, b1, t)) + ((one - b1) * boost(g, b1, t))
        delta = m_nesterov / (np.sqrt(boost(v, b2, t)) + eps)  # no learning rate (Nadam.scala:19-23)
        return delta, {"momentum": m, "velocity": v}


@dataclass(frozen=True)
class AMSGrad(Algorithm):
    rate: float = 0.001
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-7

    def state_defs(self):
        return OrderedDict(momentum=I.Zeros, velocity=I.Zeros, corrected_velocity=I.Zeros)

    def build(self, g, s, t):
        b1, b2, eps, r = _c(self.beta1, g), _c(self.beta2, g), _c(self.epsilon, g), _c(self.rate, g)
        m = decaying_avg(s["momentum"], g, b1)
        v_cur = decaying_avg(s["velocity"], g * g, b2)
        v = np.maximum(s["corrected_velocity"], v_cur)
        delta = (r * m) / (np.sqrt(v) + eps)  # no bias correction (AMSGrad.scala:43)
        return delta, {"momentum": m, "velocity": v_cur, "corrected_velocity": v}
