"""Generates tests/golden/pynocchio_big_<k>.json: the UNMODIFIED reference ``compute_proof``
(/root/reference/verifiable_mpc/trinocchio/pynocchio.py:228-273 on oracle/mpyc_shim) on a synthetic evaluation key
with 2^k mid wires and 2^k quotient coefficients, so the BN256 G1/G2 MSM kernels are pinned at sizes that leave the
single-block paths (the demo QAP of pynocchio_proof.json has 6 mid wires).

The key is synthetic with KNOWN discrete logs: every entry is ``e * generator`` with ``e`` drawn from
``random.Random(seed)`` in the order of tests/golden/seeded_inputs.py::pynocchio_draw_inputs (the front-end cannot build QAPs of this size, SURVEY
F10; compute_proof only reads the key as a dict of points).  Nothing but seed, proof and a few spot checks is stored.

Run from the repo root in the build container (1 core; ~0.5 min at k = 8, ~1.5 min at k = 10, ~25 min at k = 14):
    python tests/golden/make_pynocchio_big_golden.py 10
"""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "mpyc_shim"))
sys.path.insert(0, "/root/reference")

from mpyc.fingroups import EllipticCurve  # noqa: E402
import verifiable_mpc.trinocchio.pynocchio as pynocchio  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from seeded_inputs import PYN_FIRST_MID as FIRST_MID  # noqa: E402
from seeded_inputs import PYN_MID_TEMPLATES as MID_TEMPLATES  # noqa: E402
from seeded_inputs import PYN_SEEDS as SEEDS  # noqa: E402
from seeded_inputs import pynocchio_draw_inputs as draw_inputs  # noqa: E402


class QapStub:
    def __init__(self, mid):
        self.indices_mid = mid


class PolyStub:
    def __init__(self, coeffs):
        self.coeffs = coeffs

    def __len__(self):
        return len(self.coeffs)


class DeltasStub:
    def __init__(self, d):
        self.v, self.w, self.y = d["v"], d["w"], d["y"]


def enc(p):
    a = p.affine()
    if a is None:
        return None
    x, y = a
    return [[hex(v) for v in x], [hex(v) for v in y]] if isinstance(x, tuple) else [hex(x), hex(y)]


def main(k):
    t0 = time.time()
    g1c = EllipticCurve("BN256", "jacobian")
    g2c = EllipticCurve("BN256_twist", "jacobian")
    for cls in (g1c, g2c):
        cls.is_additive, cls.is_multiplicative = True, False
    seed = SEEDS[k]
    mid, key_exps, c, h, deltas = draw_inputs(k, seed, g1c.order)
    evalkey = {name: e * (g2c.generator if name.endswith("g2") else g1c.generator) for name, e in key_exps.items()}
    print(f"k={k}: key {time.time() - t0:.0f} s", flush=True)
    t1 = time.time()
    proof = pynocchio.compute_proof(QapStub(mid), c, PolyStub(h), evalkey, DeltasStub(deltas))
    prove_s = time.time() - t1
    proof_nozk = pynocchio.compute_proof(QapStub(mid), c, PolyStub(h), evalkey, None)
    spots = [MID_TEMPLATES[0].format(i=mid[0]), MID_TEMPLATES[1].format(i=mid[-1]), f"s^{(1 << k) - 1}*g1", "r_w*t*g2"]
    out = {"generator": "tests/golden/make_pynocchio_big_golden.py: unmodified reference compute_proof on oracle/mpyc_shim",
           "log2m": k, "seed": seed, "first_mid": FIRST_MID,
           "proof": {name: enc(v) for name, v in proof.items()},
           "proof_nozk": {name: enc(v) for name, v in proof_nozk.items()},
           "evalkey_spots": {name: enc(evalkey[name]) for name in spots},
           "reference_cpu_seconds": {"compute_proof": round(prove_s, 2),
                                     "note": "1 core, pure-Python ints on the MPyC look-alike (no gmpy2), this container"}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"pynocchio_big_{k}.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, f"total {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main(int(sys.argv[1]))
