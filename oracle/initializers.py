"""Initializers restated from `src/main/scala/scanet/models/Initializer.scala:7-220` (test infrastructure only)."""
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .philox import philox_normal, philox_truncated_normal, philox_uniform

__all__ = ["Zeros", "Ones", "Const", "RandomUniform", "RandomNormal", "RandomNormalTruncated", "GlorotUniform",
           "GlorotNormal", "HeUniform", "HeNormal", "LecunUniform", "LecunNormal", "Orthogonal", "fans"]

_TRUNC_COEF = np.float32(0.87962566103423978)


def fans(shape: Tuple[int, ...]) -> Tuple[int, int]:
    """`Initializer.fans` (Initializer.scala:16-25): rank>=3 takes d0 as 'out', d1 as 'in' (sic)."""
    dims = tuple(shape)
    if len(dims) == 0:
        return 1, 1
    if len(dims) == 1:
        return dims[0], dims[0]
    if len(dims) == 2:
        return dims[0], dims[1]
    out, inn = dims[0], dims[1]
    size = int(np.prod(dims[2:]))
    return size * inn, size * out


def _rand(shape, dist, seed, scale=None, shift=None):
    """`rand(shape, dist, seed, scale, shift) = (sample * scale) + shift` (rand/AllKernels.scala:68-81)."""
    if seed is None:
        raise ValueError("the oracle only supports seeded random initializers (SURVEY App. B-Q6)")
    n = int(np.prod(shape)) if len(shape) else 1
    sample = {"uniform": philox_uniform, "normal": philox_normal, "truncated": philox_truncated_normal}[dist](seed, n)
    sample = sample.reshape(shape).astype(np.float32)
    if scale is not None:
        sample = sample * np.float32(scale)
    if shift is not None:
        sample = sample + np.float32(shift)
    return sample.astype(np.float32)


class Initializer:
    def build(self, shape) -> np.ndarray:
        raise NotImplementedError


@dataclass(frozen=True)
class _Zeros(Initializer):
    def build(self, shape):
        return np.zeros(shape, np.float32)


@dataclass(frozen=True)
class _Ones(Initializer):
    def build(self, shape):
        return np.ones(shape, np.float32)


Zeros = _Zeros()
Ones = _Ones()


@dataclass(frozen=True)
class Const(Initializer):
    value: float

    def build(self, shape):
        return np.full(shape, np.float32(self.value), np.float32)


@dataclass(frozen=True)
class RandomUniform(Initializer):
    min: float = -0.05
    max: float = 0.05
    seed: Optional[int] = None

    def build(self, shape):
        lo, hi = np.float32(self.min), np.float32(self.max)
        return _rand(shape, "uniform", self.seed, scale=hi - lo, shift=lo)


@dataclass(frozen=True)
class RandomNormal(Initializer):
    mean: float = 0.0
    std: float = 0.05
    seed: Optional[int] = None

    def build(self, shape):
        return _rand(shape, "normal", self.seed, scale=self.std, shift=self.mean)


@dataclass(frozen=True)
class RandomNormalTruncated(Initializer):
    mean: float = 0.0
    std: float = 0.05
    seed: Optional[int] = None

    def build(self, shape):
        return _rand(shape, "truncated", self.seed, scale=self.std, shift=self.mean)


def _f32_sqrt_of_ratio(num: float, den: int) -> np.float32:
    # `math.sqrt(6f / (fanIn + fanOut)).toFloat`: fp32 division, double sqrt, cast back (Initializer.scala:86)
    return np.float32(math.sqrt(float(np.float32(num) / np.float32(den))))


@dataclass(frozen=True)
class GlorotUniform(Initializer):
    seed: Optional[int] = None

    def build(self, shape):
        fi, fo = fans(shape)
        limit = _f32_sqrt_of_ratio(6.0, fi + fo)
        return _rand(shape, "uniform", self.seed, scale=np.float32(2) * limit, shift=-limit)


@dataclass(frozen=True)
class GlorotNormal(Initializer):
    seed: Optional[int] = None

    def build(self, shape):
        fi, fo = fans(shape)
        std = _f32_sqrt_of_ratio(2.0, fi + fo) / _TRUNC_COEF
        return _rand(shape, "truncated", self.seed, scale=std)


@dataclass(frozen=True)
class HeUniform(Initializer):
    seed: Optional[int] = None

    def build(self, shape):
        fi, _ = fans(shape)
        limit = _f32_sqrt_of_ratio(6.0, fi)
        return _rand(shape, "uniform", self.seed, scale=np.float32(2) * limit, shift=-limit)


@dataclass(frozen=True)
class HeNormal(Initializer):
    seed: Optional[int] = None

    def build(self, shape):
        fi, _ = fans(shape)
        return _rand(shape, "truncated", self.seed, scale=_f32_sqrt_of_ratio(2.0, fi) / _TRUNC_COEF)


@dataclass(frozen=True)
class LecunUniform(Initializer):
    seed: Optional[int] = None

    def build(self, shape):
        fi, _ = fans(shape)
        limit = _f32_sqrt_of_ratio(3.0, fi)
        return _rand(shape, "uniform", self.seed, scale=np.float32(2) * limit, shift=-limit)


@dataclass(frozen=True)
class LecunNormal(Initializer):
    seed: Optional[int] = None

    def build(self, shape):
        fi, _ = fans(shape)
        return _rand(shape, "truncated", self.seed, scale=_f32_sqrt_of_ratio(1.0, fi) / _TRUNC_COEF)


@dataclass(frozen=True)
class Orthogonal(Initializer):
    """Initializer.scala:202-219: QR of a normal matrix, Q made unique with sign(diag R)."""
    gain: Optional[float] = None
    seed: Optional[int] = None

    def build(self, shape):
        assert len(shape) >= 2
        rows = int(np.prod(shape[:-1]))
        cols = int(shape[-1])
        flat = (max(rows, cols), min(rows, cols))
        init = _rand(flat, "normal", self.seed)
        q, r = np.linalg.qr(init.astype(np.float32))
        qu = q * np.sign(np.diag(r))
        qt = qu.T if rows < cols else qu
        if self.gain is not None:
            qt = qt * np.float32(self.gain)
        return np.ascontiguousarray(qt.reshape(shape), dtype=np.float32)
