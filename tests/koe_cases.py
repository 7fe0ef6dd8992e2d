"""Shared helper: rebuild the KoE golden case with the PRODUCT's group / field types and run the prover twin."""
import json
import os

from pynocchio_cases import dec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "koe_proof.json")


def check_koe(g1_group, g2_group):
    from verifiable_mpc_b200.ac20 import knowledge_of_exponent as twin
    from verifiable_mpc_b200.ac20 import pivot
    from verifiable_mpc_b200.finfields import GF

    gold = json.load(open(GOLDEN))
    for grp in (g1_group, g2_group):
        grp.is_additive, grp.is_multiplicative = False, True
    try:
        gf = GF(g1_group.order)
        pp = {"pp_lhs": [g1_group._make(dec(p)) for p in gold["pp_lhs"]],
              "pp_rhs": [g2_group._make(dec(p)) for p in gold["pp_rhs"]]}
        x = [gf(int(v, 16)) for v in gold["x"]]
        gamma = gf(int(gold["gamma"], 16))
        L = pivot.LinearForm([gf(int(v, 16)) for v in gold["L"]])
        proof, u = twin.opening_linear_form_prover(L, x, gamma, pp)
        assert u.value == int(gold["u"], 16)
        assert sorted(proof) == sorted(gold["proof"])
        for key, val in gold["proof"].items():
            assert proof[key].affine() == dec(val), key
        P, pi = twin.restriction_argument_prover(range(len(x)), x, gamma, pp)
        assert P == proof["P"] and pi == proof["pi"]
        # the same with the parameters resident on the device, and a non-contiguous index set
        ppd = twin.PreparedPP(pp)
        try:
            proof_d, u_d = twin.opening_linear_form_prover(L, x, gamma, ppd)
            assert u_d == u and all(proof_d[k] == proof[k] for k in proof)
            assert twin.linear_form_R(L, ppd, u) == twin.linear_form_R(L, pp, u)
            S = [0, 2]
            assert twin.restriction_argument_prover(S, x, gamma, ppd) == twin.restriction_argument_prover(S, x, gamma, pp)
        finally:
            ppd.free()
        # trusted setup twin: seeded like tests/golden/make_koe_golden.py (seed 77, the setup's three draws come first),
        # it reproduces the reference's public parameters (two fixed-base batches instead of 4n sequential powers)
        import random

        twin.prng = random.Random(77)
        pp2 = twin.trusted_setup(g1_group.generator, g2_group.generator, len(x), g1_group.order)
        assert [p.affine() for p in pp2["pp_lhs"]] == [dec(p) for p in gold["pp_lhs"]]
        assert [p.affine() for p in pp2["pp_rhs"]] == [dec(p) for p in gold["pp_rhs"]]
        twin.prng = random.Random(77)  # a base that is not the standard generator takes the per-element path
        pp3 = twin.trusted_setup(g1_group.generator ** 2, g2_group.generator, 1, g1_group.order)
        assert pp3["pp_lhs"][0] == pp2["pp_lhs"][0] ** 2 and pp3["pp_rhs"][:2] == pp2["pp_rhs"][:2]
        # group-type operators in multiplicative notation
        assert (pp["pp_lhs"][0] ** 3) * pp["pp_lhs"][0] == pp["pp_lhs"][0] ** 4
        return proof, u, L, pp
    finally:
        for grp in (g1_group, g2_group):
            grp.is_additive, grp.is_multiplicative = True, False
