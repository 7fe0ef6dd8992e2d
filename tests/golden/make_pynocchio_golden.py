"""Generates tests/golden/pynocchio_proof.json by running the UNMODIFIED reference Pinocchio flow
(/root/reference/verifiable_mpc/trinocchio/pynocchio.py: keygen, compute_proof, pairing-based verify; QAP tools and
ac20/pairing.py likewise unmodified) on top of oracle/mpyc_shim with seeded randomness, for the demo program of
demos/demo_zkp_pynocchio.py:46-49.  Stores the evaluation-key entries compute_proof reads, the witness, the quotient
polynomial, the zero-knowledge deltas and the resulting proof (all points in canonical affine coordinates).
Run from the repo root in the build container:   python tests/golden/make_pynocchio_golden.py
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
import verifiable_mpc.tools.code_to_qap as c2q  # noqa: E402
import verifiable_mpc.tools.qap_creator as qc  # noqa: E402
import verifiable_mpc.trinocchio.pynocchio as pynocchio  # noqa: E402

CODE = """
def qeval(x):
    y = x**3 + x**2 + x
    return y + x + 5
"""


def enc(p):
    a = p.affine()
    if a is None:
        return None
    x, y = a
    return [[hex(v) for v in x], [hex(v) for v in y]] if isinstance(x, tuple) else [hex(x), hex(y)]


def build(seed=2024):
    bn_curve = EllipticCurve("BN256", "jacobian")
    bn_twist = EllipticCurve("BN256_twist", "jacobian")
    g1, g2 = bn_curve.generator, bn_twist.generator
    modulus = bn_curve.order
    gf = GF(modulus=modulus)
    gf.is_signed = False
    pynocchio.prng = random.Random(seed)
    qap = c2q.QAP(CODE, gf)
    td = pynocchio.Trapdoor(modulus)
    gen = pynocchio.Generators(td, g1, g2)
    evalkey = pynocchio.generate_evalkey(td, qap, gen)
    verikey = pynocchio.generate_verikey(td, qap, gen)
    c = qap.calculate_witness([gf(3)])
    p = pynocchio.compute_p_poly(qap, c)
    h, r = p / qap.t
    assert r == qc.Poly([0] * qap.d)
    deltas = pynocchio.SampleDeltas(modulus)
    h = h + pynocchio.compute_h_zk_terms(qap, c, deltas)
    proof = pynocchio.compute_proof(qap, c, h, evalkey, deltas)
    proof_nozk = pynocchio.compute_proof(qap, c, (p / qap.t)[0], evalkey, None)
    checks = pynocchio.verify(qap, verikey, proof, c[: qap.out_ix + 1])
    assert all(checks.values()), checks
    return qap, evalkey, verikey, c, h, (p / qap.t)[0], deltas, proof, proof_nozk, td


if __name__ == "__main__":
    qap, evalkey, verikey, c, h, h_nozk, deltas, proof, proof_nozk, td = build()

    def poly(p):
        return [hex(int(v)) for v in p.coeffs]

    out = {
        "generator": "tests/golden/make_pynocchio_golden.py: unmodified reference on oracle/mpyc_shim",
        "indices_mid": list(qap.indices_mid), "m": qap.m, "d": qap.d,
        "c": [hex(int(v)) for v in c], "h": [hex(int(v)) for v in h.coeffs], "h_nozk": [hex(int(v)) for v in h_nozk.coeffs],
        "deltas": {"v": hex(deltas.v), "w": hex(deltas.w), "y": hex(deltas.y)},
        "evalkey": {k: enc(v) for k, v in evalkey.items()},
        # inputs of generate_evalkey (:101-167): the trapdoor and the QAP polynomials it evaluates at s
        "trapdoor": {k: hex(getattr(td, k)) for k in ("r_v", "r_w", "r_y", "s", "alpha_v", "alpha_w", "alpha_y", "beta", "gamma")},
        "qap_polys": {"v": {str(i): poly(qap.v[i]) for i in qap.indices_mid},
                      "w": {str(i): poly(qap.w[i]) for i in qap.indices_mid},
                      "y": {str(i): poly(qap.y[i]) for i in qap.indices_mid}, "t": poly(qap.t)},
        "proof": {k: enc(v) for k, v in proof.items()},
        "proof_nozk": {k: enc(v) for k, v in proof_nozk.items()},
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pynocchio_proof.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, "mid", out["indices_mid"], "len(h)", len(out["h"]))
