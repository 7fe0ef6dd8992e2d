"""Generates tests/golden/ac20_compressed_pivot.json by running the UNMODIFIED reference
(/root/reference/verifiable_mpc/ac20/{pivot,compressed_pivot}.py) on top of oracle/mpyc_shim with seeded randomness.

The reference cannot travel to the GPU box, so its outputs are committed as fixtures: for each case the inputs
(generator exponents, witness x, gamma, linear form L, PRNG seed) and the complete proof + public commitment P.
Run from the repo root in the build container:   python tests/golden/make_ac20_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "mpyc_shim"))
sys.path.insert(0, "/root/reference")

from mpyc.finfields import GF  # noqa: E402  (the shim)
from mpyc.fingroups import EllipticCurve  # noqa: E402
import verifiable_mpc.ac20.compressed_pivot as ref_cp  # noqa: E402
import verifiable_mpc.ac20.pivot as ref_pivot  # noqa: E402


def enc_pt(p):
    x, y = p.affine()
    return [hex(x), hex(y)]


def run_case(n, seed):
    group = EllipticCurve("Ed25519", "projective")
    group.is_additive, group.is_multiplicative = False, True
    gf = GF(group.order)
    setup = random.Random(seed)
    exps = [setup.randrange(1, group.order) for _ in range(n)]
    k_exp = setup.randrange(1, group.order)
    h = group.generator
    g = [h ** e for e in exps]
    k = h ** k_exp
    x = [gf(setup.randrange(gf.order)) for _ in range(n)]
    if n >= 7:  # a few small / negative-looking witnesses like real circuits have
        x[1], x[2], x[3] = gf(0), gf(1), gf(-1)
    gamma = gf(setup.randrange(gf.order))
    L = ref_pivot.LinearForm([gf(setup.randrange(gf.order)) for _ in range(n)])
    y = L(x)
    P = ref_pivot.vector_commitment(x, gamma, g, h)
    generators = {"g": g, "h": h, "k": k}
    rng = random.Random(seed + 1)
    ref_cp.prng = rng
    ref_pivot.prng = rng
    proof = ref_cp.protocol_5_prover(generators, P, L, y, x, gamma, gf)
    assert ref_cp.protocol_5_verifier(generators, P, L, y, proof, gf) is True
    # basic pivot (protocol 2) on the same statement
    rng2 = random.Random(seed + 2)
    ref_pivot.prng = rng2
    z, phi, c = ref_pivot.prove_linear_form_eval(g, h, P, L, y, x, int(gamma), gf)
    assert ref_pivot.verify_linear_form_proof(g, h, P, L, y, z, phi, c) is True
    rounds = sum(1 for key in proof if key.startswith("A") and key != "A")
    return {
        "n": n, "seed": seed, "exponents": [hex(e) for e in exps], "k_exponent": hex(k_exp),
        "x": [hex(v.value) for v in x], "gamma": hex(gamma.value), "L": [hex(c_.value) for c_ in L.coeffs],
        "y": hex(y.value), "P": enc_pt(P),
        "proof": {"t": hex(proof["t"].value), "A": enc_pt(proof["A"]),
                  "A_i": [enc_pt(proof[f"A{i}"]) for i in range(rounds)],
                  "B_i": [enc_pt(proof[f"B{i}"]) for i in range(rounds)],
                  "z_prime": [hex(v.value) for v in proof["z_prime"]]},
        "pivot": {"z": [hex(v.value) for v in z], "phi": hex(phi), "c": hex(c)},
    }


out = {"generator": "tests/golden/make_ac20_golden.py: unmodified reference on oracle/mpyc_shim",
       "cases": [run_case(n, seed) for n, seed in ((3, 11), (7, 12), (31, 13))]}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ac20_compressed_pivot.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path, [c["n"] for c in out["cases"]])
