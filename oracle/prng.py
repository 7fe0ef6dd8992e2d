"""Counter-based synthetic-input generator (TEST INFRASTRUCTURE ONLY; the product has its own numpy/CUDA
implementation of the same spec in verifiable_mpc_b200/synth.py and csrc/ -- this file is the independent
restatement the tests compare them with).

Spec (SURVEY.md 8d "Synthetic inputs"):
    word(seed, i, j) = splitmix64_mix(seed + GOLDEN * (4*i + j + 1))          j = 0..3, 64-bit wrap-around
    v(seed, i)       = word0 | word1<<64 | word2<<128 | word3<<192            (little-endian 256-bit)
    scalar(seed, i)  = (v mod 2^253) - (l if (v mod 2^253) >= l else 0)       for Ed25519 (l = group order)
    scalar_bn(seed,i)= v - (n if v >= n else 0)                               for BN256 (n has its top bit set)
Scalars use seed S, known discrete logs of the synthetic bases use seed S+1:  g_i = scalar(S+1, i) * B.
"""

MASK64 = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15


def splitmix64_mix(z):
    z &= MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def word(seed, i, j):
    return splitmix64_mix(seed + GOLDEN * (4 * i + j + 1))


def v256(seed, i):
    return sum(word(seed, i, j) << (64 * j) for j in range(4))


ED_L = 2**252 + 27742317777372353535851937790883648493
BN_N = 65000549695646603732796438742359905742570406053903786389881062969044166799969


def scalar(seed, i):
    v = v256(seed, i) & ((1 << 253) - 1)
    return v - ED_L if v >= ED_L else v


def scalar_bn(seed, i):
    v = v256(seed, i)
    return v - BN_N if v >= BN_N else v
