"""Generates tests/golden/ac20_demo_n128.json.gz: BASELINE config 1.  The reference's own driver
(/root/reference/demos/demo_zkp_ac20.py --elliptic: circuit builder -> circuit_sat_cb.circuit_sat_prover ->
compressed pivot, N = 128 generators, 6 folding rounds) runs UNMODIFIED on oracle/mpyc_shim; the statement it hands to
``compressed_pivot.protocol_5_prover`` (circuit_sat_cb.py:264) and the commitment it makes with
``pivot.vector_commitment`` (circuit_sat_cb.py:103) are captured with their results, so the GPU box (no reference tree)
can replay exactly the calls the reference's driver makes.  The pivot's prng is re-seeded at the capture point so the
prover's draws are replayable.   Run from the repo root:   python tests/golden/make_ac20_demo_golden.py
"""
import gzip
import importlib.util
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "mpyc_shim"))
sys.path.insert(0, REF)
sys.argv = ["demo_zkp_ac20.py"]

spec = importlib.util.spec_from_file_location("demo_zkp_ac20_golden", os.path.join(REF, "demos", "demo_zkp_ac20.py"))
demo = importlib.util.module_from_spec(spec)
spec.loader.exec_module(demo)
import verifiable_mpc.ac20.circuit_builder as rcb  # noqa: E402
import verifiable_mpc.ac20.circuit_sat_cb as rcs  # noqa: E402
import verifiable_mpc.ac20.circuit_sat_r1cs as rr1cs  # noqa: E402
import verifiable_mpc.ac20.compressed_pivot as rcp  # noqa: E402
import verifiable_mpc.ac20.pivot as rpivot  # noqa: E402

PIVOT_SEED = 4242


class _Quiet:
    def pprint(self, *a, **k):
        pass


def enc_pt(p):
    x, y = p.affine()
    return [hex(x), hex(y)]


def fe(v):
    """Exact operand as the driver passed it: ["f", hex residue] for a field element, ["i", hex of the (possibly
    negative, possibly thousands of bits long, unreduced) Python int] otherwise -- the Fiat-Shamir pre-image contains
    the decimal repr of these very objects (pivot.py:131-136), so their types are part of the statement."""
    return ["f", hex(int(v.value))] if hasattr(v, "value") else ["i", hex(int(v))]


cap = {"commitments": []}
orig_vc, orig_p5 = rpivot.vector_commitment, rcp.protocol_5_prover


def vc(x, gamma, g, h):
    out = orig_vc(x, gamma, g, h)
    if len(x) > 64 and not cap["commitments"]:  # the z commitment of circuit_sat_cb.py:103 (first large call)
        cap["commitments"].append({"x": [fe(v) for v in x], "gamma": fe(gamma), "n_g": len(g), "out": enc_pt(out)})
    return out


def p5(generators, P, L, y, x, gamma, gf):
    rcp.prng = random.Random(PIVOT_SEED)
    proof = orig_p5(generators, P, L, y, x, gamma, gf)
    rounds = sum(1 for key in proof if key.startswith("A") and key != "A")
    cap["pivot"] = {
        "g": [enc_pt(p) for p in generators["g"]], "h": enc_pt(generators["h"]), "k": enc_pt(generators["k"]),
        "P": enc_pt(P), "L": [fe(c) for c in L.coeffs], "L_constant": fe(L.constant),
        "L_type": type(L).__name__, "y": fe(y), "x": [fe(v) for v in x], "gamma": fe(gamma),
        "prng_seed": PIVOT_SEED,
        "proof": {"t": fe(proof["t"]), "A": enc_pt(proof["A"]),
                  "A_i": [enc_pt(proof[f"A{i}"]) for i in range(rounds)],
                  "B_i": [enc_pt(proof[f"B{i}"]) for i in range(rounds)],
                  "z_prime": [fe(v) for v in proof["z_prime"]]}}
    return proof


demo.GROUP = "Elliptic"
demo.pp = _Quiet()
for mod in (rpivot, rcp, rcs, rr1cs, rcb):
    if hasattr(mod, "prng"):
        mod.prng = random.Random(42)
rpivot.vector_commitment = vc
rcp.protocol_5_prover = p5
checks = demo.main(demo.cs.PivotChoice.compressed, 3)
assert all(checks.values()), checks
assert len(cap["pivot"]["g"]) == 127 and len(cap["pivot"]["proof"]["A_i"]) == 6
out = {"generator": "tests/golden/make_ac20_demo_golden.py: demos/demo_zkp_ac20.py --elliptic, unmodified, on oracle/mpyc_shim",
       "verification": {k: bool(v) for k, v in checks.items()}, **cap}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ac20_demo_n128.json.gz")
with gzip.GzipFile(path, "wb", mtime=0) as f:  # the unreduced integer coefficients of L are ~2900 digits each
    f.write(json.dumps(out, indent=1).encode())
print("wrote", path, "N =", len(cap["pivot"]["g"]) + 1, "commitments", len(cap["commitments"]))
