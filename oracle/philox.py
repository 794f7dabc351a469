"""TensorFlow's seeded random ops restated (test infrastructure only).

The reference emits TF `RandomUniform` / `RandomStandardNormal` / `TruncatedNormal` ops and sets only
the `seed` attr (`seed2` stays 0) -- `src/main/scala/scanet/math/rand/AllKernels.scala:12-43` -- and
evaluates every initializer in a fresh session (`core/Eval.scala:14-17`), so the generator always
starts at counter 0.  The arithmetic lives in TensorFlow 2.10 (`tensorflow-core-platform:0.5.0`,
`build.sbt:13`), absent from /root/reference; this file restates its published algorithm:
Philox-4x32-10 (`tensorflow/core/lib/random/philox_random.h`), `Uint32ToFloat`, `BoxMullerFloat`,
`TruncatedNormalDistribution` (`random_distributions.h`) and the CPU fill order
(`kernels/random_op_cpu.h`).  Pinned by `InitializerSpec.scala:23-66` (tests/test_oracle_goldens.py).
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(counters: np.ndarray, key0: int, key1: int) -> np.ndarray:
    """counters: uint64 [n,4] holding 32-bit words; returns uint32 [n,4] after 10 rounds."""
    c = counters.astype(np.uint64) & _MASK
    c0, c1, c2, c3 = c[:, 0].copy(), c[:, 1].copy(), c[:, 2].copy(), c[:, 3].copy()
    k0, k1 = key0 & 0xFFFFFFFF, key1 & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def _counter_block(start: int, n: int, seed2: int) -> np.ndarray:
    """128-bit counters start..start+n-1 (low 64 bits count; seed2 sits in words 2,3)."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    lo = idx & _MASK
    hi = idx >> np.uint64(32)
    base2 = np.uint64(seed2 & 0xFFFFFFFF)
    base3 = np.uint64((seed2 >> 32) & 0xFFFFFFFF)
    # Skip() adds to the 128-bit counter little-endian: words 0,1 carry into 2,3 (never happens here)
    return np.stack([lo, hi, np.full(n, base2, np.uint64), np.full(n, base3, np.uint64)], axis=1)


def philox_words(seed: int, n_words: int, seed2: int = 0, start_counter: int = 0) -> np.ndarray:
    n_blocks = (n_words + 3) // 4
    out = philox4x32(_counter_block(start_counter, n_blocks, seed2), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return out.reshape(-1)[:n_words]


def uint32_to_float(x: np.ndarray) -> np.ndarray:
    bits = (np.uint32(127) << np.uint32(23)) | (x.astype(np.uint32) & np.uint32(0x7FFFFF))
    return bits.view(np.float32) - np.float32(1.0)


def philox_uniform(seed: int, n: int) -> np.ndarray:
    """TF RandomUniform(float32) in [0,1): element 4i+j = word j of philox(counter i)."""
    return uint32_to_float(philox_words(seed, n))


def _box_muller(x0: np.ndarray, x1: np.ndarray):
    u1 = np.maximum(uint32_to_float(x0), np.float32(1.0e-7))
    # `const float v1 = 2.0f * M_PI * Uint32ToFloat(x1)`: the product is evaluated in double, rounded once
    v1 = (np.float64(2.0 * np.pi) * uint32_to_float(x1).astype(np.float64)).astype(np.float32)
    u2 = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    return (np.sin(v1) * u2).astype(np.float32), (np.cos(v1) * u2).astype(np.float32)


def philox_normal(seed: int, n: int) -> np.ndarray:
    """TF RandomStandardNormal(float32): Box-Muller on word pairs (0,1),(2,3) of each block."""
    w = philox_words(seed, 4 * ((n + 3) // 4)).reshape(-1, 4)
    f0, f1 = _box_muller(w[:, 0], w[:, 1])
    f2, f3 = _box_muller(w[:, 2], w[:, 3])
    return np.stack([f0, f1, f2, f3], axis=1).reshape(-1)[:n]


def philox_truncated_normal(seed: int, n: int) -> np.ndarray:
    """TF TruncatedNormal(float32): variable samples per output; each group of 4 outputs restarts the
    generator at counter group*256 and draws single words until 4 values with |x| < 2 are found."""
    out = np.empty(4 * ((n + 3) // 4), np.float32)
    for g in range(len(out) // 4):
        got, blk = 0, 0
        stream = np.empty(0, np.uint32)
        pos = 0
        while got < 4:
            if pos + 2 > len(stream):
                more = philox_words(seed, 64, start_counter=g * 256 + blk)
                blk += 16
                stream = np.concatenate([stream[pos:], more])
                pos = 0
            f0, f1 = _box_muller(stream[pos:pos + 1], stream[pos + 1:pos + 2])
            pos += 2
            for f in (f0[0], f1[0]):
                if got < 4 and abs(f) < np.float32(2.0):
                    out[4 * g + got] = f
                    got += 1
    return out[:n]
