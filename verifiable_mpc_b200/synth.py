"""Deterministic synthetic inputs: counter-based SplitMix64 stream, vectorised with numpy.

    word(seed, i, j) = mix(seed + GOLDEN*(4i + j + 1));  v = 256-bit LE of the four words
    scalar(seed, i)  = (v mod 2^253), minus l when >= l          (Ed25519 group order l)
The CUDA twin is csrc/kernels.cuh:synth_scalar_ed (KSynthScalars / KFixedBase); oracle/prng.py restates it
independently for the tests.
"""
import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
ED_L = 2**252 + 27742317777372353535851937790883648493
_L_WORDS = np.array([(ED_L >> (64 * j)) & (2**64 - 1) for j in range(4)], dtype=np.uint64)


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def scalars_ed25519(seed, n, start=0, out=None):
    """n x 32 uint8 (little-endian scalars below l) for indices start .. start+n-1."""
    with np.errstate(over="ignore"):
        i = np.arange(start, start + n, dtype=np.uint64)
        w = np.empty((n, 4), dtype=np.uint64)
        for j in range(4):
            w[:, j] = _mix(np.uint64(seed) + GOLDEN * (np.uint64(4) * i + np.uint64(j + 1)))
        w[:, 3] &= np.uint64((1 << 61) - 1)  # keep 253 bits
        # d = w - l with borrow propagation
        d = np.empty_like(w)
        borrow = np.zeros(n, dtype=np.uint64)
        for j in range(4):
            t = w[:, j] - _L_WORDS[j]
            b1 = (w[:, j] < _L_WORDS[j]).astype(np.uint64)
            t2 = t - borrow
            b2 = (t < borrow).astype(np.uint64)
            d[:, j] = t2
            borrow = b1 | b2
        ge = borrow == 0  # no borrow: w >= l
        w[ge] = d[ge]
    raw = w.view(np.uint8).reshape(n, 32)
    if out is not None:
        out[:] = raw
        return out
    return raw
