"""Generates tests/golden/ac20_big_<k>.json: compressed-pivot proofs of the UNMODIFIED reference
(/root/reference/verifiable_mpc/ac20/{pivot,compressed_pivot}.py on oracle/mpyc_shim) at the sizes BASELINE.json
configures (N = 2^k generators, k = 10, 12, 16), so that the GPU twins are pinned bit for bit above the 2^13 fold-kernel
switch and the 256-entry host hand-off, not only at N <= 32.

Inputs are NOT stored (3 * 2^16 scalars would be 12 MB of hex): they are the draws of ``random.Random(seed)`` in the
order of tests/golden/seeded_inputs.py::ac20_draw_inputs, which tests/ac20_cases.py replays.  Stored: commitment P, y, the complete proof, a few
generator spot checks and the sha256 of the canonical proof text.

Run from the repo root in the build container (single core, pure Python: ~12 s at k = 10, ~50 s at k = 12,
~17 min at k = 16):   python tests/golden/make_ac20_big_golden.py 16
"""
import hashlib
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "mpyc_shim"))
sys.path.insert(0, "/root/reference")

from mpyc.finfields import GF  # noqa: E402  (the shim)
from mpyc.fingroups import EllipticCurve  # noqa: E402
import verifiable_mpc.ac20.compressed_pivot as ref_cp  # noqa: E402
import verifiable_mpc.ac20.pivot as ref_pivot  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from seeded_inputs import AC20_N_BOOLEAN as N_BOOLEAN  # noqa: E402
from seeded_inputs import AC20_SEEDS as SEEDS  # noqa: E402
from seeded_inputs import ac20_draw_inputs as draw_inputs  # noqa: E402
from seeded_inputs import canonical_proof_text  # noqa: E402


def enc_pt(p):
    x, y = p.affine()
    return [hex(x), hex(y)]


def main(k):
    t0 = time.time()
    group = EllipticCurve("Ed25519", "projective")
    group.is_additive, group.is_multiplicative = False, True
    gf = GF(group.order)
    seed = SEEDS[k]
    exps, k_exp, xi, gi, Li = draw_inputs(k, seed, group.order)
    n = len(exps)
    h = group.generator
    g = [h ** e for e in exps]
    kk = h ** k_exp
    x = [gf(v) for v in xi]
    gamma = gf(gi)
    L = ref_pivot.LinearForm([gf(v) for v in Li])
    y = L(x)
    P = ref_pivot.vector_commitment(x, gamma, g, h)
    print(f"k={k}: inputs + commitment {time.time() - t0:.0f} s", flush=True)
    generators = {"g": g, "h": h, "k": kk}
    rng = random.Random(seed + 1)
    ref_cp.prng = rng
    ref_pivot.prng = rng
    t1 = time.time()
    proof = ref_cp.protocol_5_prover(generators, P, L, y, x, gamma, gf)
    prove_s = time.time() - t1
    print(f"k={k}: reference protocol_5_prover {prove_s:.0f} s", flush=True)
    t1 = time.time()
    assert ref_cp.protocol_5_verifier(generators, P, L, y, proof, gf) is True
    verify_s = time.time() - t1
    rounds = sum(1 for key in proof if key.startswith("A") and key != "A")
    enc = {"t": hex(proof["t"].value), "A": enc_pt(proof["A"]),
           "A_i": [enc_pt(proof[f"A{i}"]) for i in range(rounds)],
           "B_i": [enc_pt(proof[f"B{i}"]) for i in range(rounds)],
           "z_prime": [hex(v.value) for v in proof["z_prime"]]}
    spots = [0, 1, n // 2, n - 1]
    out = {"generator": "tests/golden/make_ac20_big_golden.py: unmodified reference on oracle/mpyc_shim",
           "log2N": k, "n": n, "seed": seed, "n_boolean": N_BOOLEAN,
           "y": hex(y.value), "P": enc_pt(P), "proof": enc,
           "proof_sha256": hashlib.sha256(canonical_proof_text(enc).encode()).hexdigest(),
           "generator_spots": {str(i): enc_pt(g[i]) for i in spots}, "k_point": enc_pt(kk),
           "reference_cpu_seconds": {"prove": round(prove_s, 2), "verify": round(verify_s, 2),
                                     "note": "1 core, pure-Python ints on the MPyC look-alike (no gmpy2), this container"}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"ac20_big_{k}.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, f"total {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main(int(sys.argv[1]))
