"""Generates tests/golden/ed25519_msm_fold.json with the pure-Python oracle (oracle/ed25519.py), i.e. with the
reference's own algorithm (per-term double-and-add + tree product, pivot.py:143 / compressed_pivot.py:64).
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ed25519 as E  # noqa: E402
from oracle import prng  # noqa: E402

N = 32
dl = [prng.scalar(0x601D, i) for i in range(N)]
pts = [E.scalar_mul(E.B, r) for r in dl]
pts[5] = pts[4]            # repeated base
pts[9] = E.IDENTITY        # identity as a base
pts[11] = E.affine_neg(pts[10])


def enc(sc):
    return [[hex(abs(s)), s < 0] for s in sc]


cases = []
for name, sc in [
    ("uniform32", [prng.scalar(0x601E, i) for i in range(32)]),
    ("uniform5", [prng.scalar(0x601F, i) for i in range(5)]),
    ("edge", [0, 1, E.L - 1, 2, -1, -2, E.L, E.L + 1, 2**252, 2**253 - 1, 2**16, 2**16 - 1, 2**15, 2**15 + 1,
              prng.scalar(1, 1) ** 2, -(2**300)]),
    ("all_equal_digits", [int("1" * 63, 16) % E.L] * 12),
    ("single", [prng.scalar(0x6020, 0)]),
    ("empty", []),
]:
    cases.append({"name": name, "scalars": enc(sc), "expect": [hex(v) for v in E.msm_naive(sc, pts)]})

c = prng.scalar(0x6021, 0)
fold = {"n": 16, "c": hex(c), "expect": [[hex(x), hex(y)] for x, y in E.fold(pts[:16], c)]}
out = {"generator": "tests/golden/make_golden.py (oracle/ed25519.py)", "dlogs": [hex(r) for r in dl],
       "points": [[hex(x), hex(y)] for x, y in pts], "msm": cases, "fold": fold}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "ed25519_msm_fold.json"), "w"), indent=1)
print("wrote", len(cases), "msm cases")
