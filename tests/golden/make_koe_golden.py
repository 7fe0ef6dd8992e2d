"""Generates tests/golden/koe_proof.json by running the UNMODIFIED reference KoE pivot
(/root/reference/verifiable_mpc/ac20/knowledge_of_exponent.py: trusted_setup, opening_linear_form_prover and the
pairing-based opening_linear_form_verifier) on top of oracle/mpyc_shim with seeded randomness, in the multiplicative
'projective' setting of verifiable_mpc/ac20/test/test_koe.py:16-40.
Run from the repo root in the build container:   python tests/golden/make_koe_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "mpyc_shim"))
sys.path.insert(0, "/root/reference")

from mpyc.finfields import GF  # noqa: E402
from mpyc.fingroups import EllipticCurve  # noqa: E402
import verifiable_mpc.ac20.knowledge_of_exponent as koe  # noqa: E402
import verifiable_mpc.ac20.pivot as ref_pivot  # noqa: E402


def enc(p):
    a = p.affine()
    if a is None:
        return None
    x, y = a
    return [[hex(v) for v in x], [hex(v) for v in y]] if isinstance(x, tuple) else [hex(x), hex(y)]


def build(n=4, seed=77):
    group1 = EllipticCurve("BN256", "projective")
    group2 = EllipticCurve("BN256_twist", "projective")
    for grp in (group1, group2):
        grp.is_additive, grp.is_multiplicative = False, True
    order = group1.order
    gf = GF(modulus=order)
    rng = random.Random(seed)
    koe.prng = rng
    pp = koe.trusted_setup(group1.generator, group2.generator, n, order)
    x = [gf(rng.randrange(order)) for _ in range(n)]
    gamma = gf(rng.randrange(order))
    L = ref_pivot.LinearForm([gf(rng.randrange(order)) for _ in range(n)])
    proof, u = koe.opening_linear_form_prover(L, x, gamma, pp)
    checks = koe.opening_linear_form_verifier(L, pp, proof, u)
    assert all(checks.values()), checks
    return group1, group2, gf, pp, x, gamma, L, proof, u


if __name__ == "__main__":
    g1, g2, gf, pp, x, gamma, L, proof, u = build()
    out = {"generator": "tests/golden/make_koe_golden.py: unmodified reference on oracle/mpyc_shim",
           "pp_lhs": [enc(p) for p in pp["pp_lhs"]], "pp_rhs": [enc(p) for p in pp["pp_rhs"]],
           "x": [hex(v.value) for v in x], "gamma": hex(gamma.value), "L": [hex(v.value) for v in L.coeffs],
           "u": hex(u.value), "proof": {k: enc(v) for k, v in proof.items()}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "koe_proof.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)
