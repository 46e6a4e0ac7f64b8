"""ORACLE / TEST INFRASTRUCTURE ONLY.  Platform-exact pseudo-random tensors (integer hash -> float32),
so that golden outputs can be committed without their inputs."""
import numpy as np


def det_uniform(shape, seed, lo=-1.0, hi=1.0):
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        x = (idx * np.uint64(6364136223846793005) + np.uint64(seed) * np.uint64(1442695040888963407)
             + np.uint64(1013904223))
        x ^= x >> np.uint64(29)
        x = x * np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(32)
    u = (x >> np.uint64(40)).astype(np.float64) / float(1 << 24)          # 24-bit mantissa, exact
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def unpack(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(np.asarray(bits))[:n].astype(bool).reshape(shape)
