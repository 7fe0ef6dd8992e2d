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


def check_big_compute_proof(k, g1_group, g2_group, via_dict=False):
    """2^k mid wires and quotient coefficients (k = 8 .. 14; BASELINE config 4 is 2^14): synthetic key with known
    discrete logs replayed from the seed (tests/golden/seeded_inputs.py), bases generated on the device, proof compared
    with what the UNMODIFIED reference compute_proof returned (tests/golden/pynocchio_big_<k>.json)."""
    from golden.seeded_inputs import PYN_MID_TEMPLATES, pynocchio_draw_inputs
    from verifiable_mpc_b200.trinocchio import pynocchio as twin

    gold = json.load(open(os.path.join(os.path.dirname(GOLDEN), f"pynocchio_big_{k}.json")))
    order = g1_group.order
    mid, key_exps, c, h, deltas = pynocchio_draw_inputs(k, gold["seed"], order)
    assert mid[0] == gold["first_mid"]
    ctx = g1_group._ctx()
    groups, bases = {}, {}
    assert [t[1] for t in twin._MID_SUMS] == list(PYN_MID_TEMPLATES)
    for name, template, delta_terms in twin._MID_SUMS:
        group = g2_group if name.endswith("g2") else g1_group
        exps = [key_exps[template.format(i=i)] for i in mid] + [key_exps[key] for _, key in delta_terms]
        groups[name] = group
        bases[name] = ctx.fixed_base(scalars=exps, curve=group.curve_id)
    groups["h*g1"] = g1_group
    bases["h*g1"] = ctx.fixed_base(scalars=[key_exps[f"s^{i}*g1"] for i in range(len(h))], curve=g1_group.curve_id)
    # the device-generated key is the reference's key (spot checks stored with the fixture)
    spots = gold["evalkey_spots"]
    assert bases["r_v*v_mid*g1"].tolist(0, 1)[0] == dec(spots[PYN_MID_TEMPLATES[0].format(i=mid[0])])
    assert bases["r_w*w_mid*g2"].tolist(len(mid) - 1, 1)[0] == dec(spots[PYN_MID_TEMPLATES[1].format(i=mid[-1])])
    assert bases["r_w*w_mid*g2"].tolist(len(mid), 1)[0] == dec(spots["r_w*t*g2"])
    assert bases["h*g1"].tolist(len(h) - 1, 1)[0] == dec(spots[f"s^{len(h) - 1}*g1"])
    prepared = twin.PreparedEvalKey.from_device_bases(mid, len(h), groups, bases)
    qap, hp, dl = QapStub(mid), PolyStub(h), DeltasStub({a: hex(v) for a, v in deltas.items()})
    try:
        proof = twin.compute_proof(qap, c, hp, prepared, dl)
        assert list(proof) == list(gold["proof"])
        for key, val in gold["proof"].items():
            assert proof[key].affine() == dec(val), key
        proof = twin.compute_proof(qap, c, hp, prepared, None)
        for key, val in gold["proof_nozk"].items():
            assert proof[key].affine() == dec(val), key
        # key tables (PreparedEvalKey.precompute): the same proofs without a single doubling in the eight sums
        if hasattr(prepared.bases["h*g1"], "precompute"):
            prepared.precompute()
            proof = twin.compute_proof(qap, c, hp, prepared, dl)
            for key, val in gold["proof"].items():
                assert proof[key].affine() == dec(val), ("tables", key)
            proof = twin.compute_proof(qap, c, hp, prepared, None)
            for key, val in gold["proof_nozk"].items():
                assert proof[key].affine() == dec(val), ("tables", key)
        if via_dict:  # the reference's calling convention: a dict of group elements, uploaded per call
            evalkey = {}
            for name, template, delta_terms in twin._MID_SUMS:
                pts = bases[name].tolist()
                for i, p in zip(mid, pts):
                    evalkey[template.format(i=i)] = groups[name]._make(p)
                for (_, key), p in zip(delta_terms, pts[len(mid):]):
                    evalkey[key] = groups[name]._make(p)
            for i, p in enumerate(bases["h*g1"].tolist()):
                evalkey[f"s^{i}*g1"] = g1_group._make(p)
            proof = twin.compute_proof(qap, c, hp, evalkey, dl)
            for key, val in gold["proof"].items():
                assert proof[key].affine() == dec(val), key
    finally:
        for dev in bases.values():
            dev.free()
