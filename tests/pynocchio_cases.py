"""Shared helpers: rebuild the Pinocchio golden case with the PRODUCT's group types and run the compute_proof twin."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pynocchio_proof.json")


def dec(p):
    if p is None:
        return None
    if isinstance(p[0], list):
        return (tuple(int(v, 16) for v in p[0]), tuple(int(v, 16) for v in p[1]))
    return (int(p[0], 16), int(p[1], 16))


class QapStub:
    def __init__(self, indices_mid):
        self.indices_mid = indices_mid


class PolyStub:
    def __init__(self, coeffs):
        self.coeffs = coeffs

    def __len__(self):
        return len(self.coeffs)


class DeltasStub:
    def __init__(self, d):
        self.v, self.w, self.y = (int(d[k], 16) for k in ("v", "w", "y"))


def load():
    return json.load(open(GOLDEN))


def check_compute_proof(g1_group, g2_group):
    from verifiable_mpc_b200.trinocchio import pynocchio as twin

    gold = load()
    evalkey = {}
    for key, val in gold["evalkey"].items():
        group = g2_group if key.endswith("g2") else g1_group
        evalkey[key] = group._make(dec(val))
    qap = QapStub(gold["indices_mid"])
    c = [int(v, 16) for v in gold["c"]]
    h = PolyStub([int(v, 16) for v in gold["h"]])
    h_nozk = PolyStub([int(v, 16) for v in gold["h_nozk"]])
    deltas = DeltasStub(gold["deltas"])

    proof = twin.compute_proof(qap, c, h, evalkey, deltas)
    assert sorted(proof) == sorted(gold["proof"])
    for key, val in gold["proof"].items():
        assert proof[key].affine() == dec(val), key
    proof2 = twin.compute_proof(qap, c, h_nozk, evalkey, None)
    for key, val in gold["proof_nozk"].items():
        assert proof2[key].affine() == dec(val), key
    # bases resident on the device
    prepared = twin.PreparedEvalKey(qap, evalkey, h_len=len(h))
    proof3 = twin.compute_proof(qap, c, h, prepared, deltas)
    for key, val in gold["proof"].items():
        assert proof3[key].affine() == dec(val), key
    return proof


class _EvalPoly:
    """QAP polynomial stand-in: coefficients mod n with the reference's ``eval`` (qap_creator.py:47-48,108-109)."""

    def __init__(self, coeffs, order):
        self.coeffs, self.order = coeffs, order

    def eval(self, x):
        return sum(c * pow(x, i, self.order) for i, c in enumerate(self.coeffs)) % self.order


def check_generate_evalkey(g1_group, g2_group):
    """generate_evalkey twin (two fixed-base batches) against the evaluation key the unmodified reference produced
    for the same trapdoor and QAP (tests/golden/pynocchio_proof.json): same keys, same order, same points."""
    from verifiable_mpc_b200.trinocchio import pynocchio as twin

    gold = load()
    order = g1_group.order

    class Td:
        pass

    td = Td()
    for k, v in gold["trapdoor"].items():
        setattr(td, k, int(v, 16))

    class Qap:
        indices_mid = gold["indices_mid"]
        d = gold["d"]
        v = {int(i): _EvalPoly([int(c, 16) for c in cs], order) for i, cs in gold["qap_polys"]["v"].items()}
        w = {int(i): _EvalPoly([int(c, 16) for c in cs], order) for i, cs in gold["qap_polys"]["w"].items()}
        y = {int(i): _EvalPoly([int(c, 16) for c in cs], order) for i, cs in gold["qap_polys"]["y"].items()}
        t = _EvalPoly([int(c, 16) for c in gold["qap_polys"]["t"]], order)

    class Gen:
        g1, g2 = g1_group.generator, g2_group.generator

    key = twin.generate_evalkey(td, Qap, Gen)
    assert list(key) == list(gold["evalkey"])
    for name, val in gold["evalkey"].items():
        assert key[name].affine() == dec(val), name
        assert isinstance(key[name], g2_group if name.endswith("g2") else g1_group)
    return key
