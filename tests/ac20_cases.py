"""Shared helpers: rebuild the inputs of a golden AC20 case with the PRODUCT's types and run the GPU twins."""
import json
import os
import random

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ac20_compressed_pivot.json")


def load_cases():
    return json.load(open(GOLDEN))["cases"]


def build_inputs(case, group, gf):
    from verifiable_mpc_b200.ac20 import generators as gens
    from verifiable_mpc_b200.ac20 import pivot

    exps = [int(e, 16) for e in case["exponents"]]
    generators = gens.create_generators(len(exps), group, with_k=False, exponents=exps)
    generators["k"] = group.generator ** int(case["k_exponent"], 16)
    x = [gf(int(v, 16)) for v in case["x"]]
    gamma = gf(int(case["gamma"], 16))
    L = pivot.LinearForm([gf(int(v, 16)) for v in case["L"]])
    y = L(x)
    assert y.value == int(case["y"], 16)
    return generators, x, gamma, L, y


def pt(p):
    return (int(p[0], 16), int(p[1], 16))


def check_case(case, group, gf):
    """Runs commitment, compressed-pivot prover/verifier and basic pivot; asserts equality with the reference."""
    from verifiable_mpc_b200.ac20 import compressed_pivot as cp
    from verifiable_mpc_b200.ac20 import pivot

    generators, x, gamma, L, y = build_inputs(case, group, gf)
    g, h = generators["g"], generators["h"]
    P = pivot.vector_commitment(x, gamma, g, h)
    assert P.affine() == pt(case["P"])

    rng = random.Random(case["seed"] + 1)
    cp.prng = rng
    pivot.prng = rng
    proof = cp.protocol_5_prover(generators, P, L, y, x, gamma, gf)
    ref = case["proof"]
    assert proof["t"].value == int(ref["t"], 16)
    assert proof["A"].affine() == pt(ref["A"])
    rounds = len(ref["A_i"])
    assert sorted(proof) == sorted(["t", "A", "z_prime"] + [f"A{i}" for i in range(rounds)] + [f"B{i}" for i in range(rounds)])
    for i in range(rounds):
        assert proof[f"A{i}"].affine() == pt(ref["A_i"][i]), i
        assert proof[f"B{i}"].affine() == pt(ref["B_i"][i]), i
    assert [v.value for v in proof["z_prime"]] == [int(v, 16) for v in ref["z_prime"]]
    # the generators handed in are still intact (the prover folds a private copy)
    assert len(generators["g"]) == case["n"]
    assert cp.protocol_5_verifier(generators, P, L, y, proof, gf) is True

    # the reference's proof (rebuilt from the fixture) verifies with the twin verifier; a tampered one does not
    ref_proof = {"t": gf(int(ref["t"], 16)), "A": group._make(pt(ref["A"])),
                 "z_prime": [gf(int(v, 16)) for v in ref["z_prime"]]}
    for i in range(rounds):
        ref_proof[f"A{i}"] = group._make(pt(ref["A_i"][i]))
        ref_proof[f"B{i}"] = group._make(pt(ref["B_i"][i]))
    assert cp.protocol_5_verifier(generators, P, L, y, ref_proof, gf) is True
    bad = dict(ref_proof)
    bad["z_prime"] = [ref_proof["z_prime"][0] + 1] + ref_proof["z_prime"][1:]
    assert cp.protocol_5_verifier(generators, P, L, y, bad, gf) is False
    bad = dict(ref_proof)
    bad["A0"] = ref_proof["B0"]
    assert cp.protocol_5_verifier(generators, P, L, y, bad, gf) is False

    # basic pivot (protocol 2)
    pivot.prng = random.Random(case["seed"] + 2)
    z, phi, c = pivot.prove_linear_form_eval(g, h, P, L, y, x, int(gamma), gf)
    assert [v.value for v in z] == [int(v, 16) for v in case["pivot"]["z"]]
    assert phi == int(case["pivot"]["phi"], 16) and c == int(case["pivot"]["c"], 16)
    assert pivot.verify_linear_form_proof(g, h, P, L, y, z, phi, c) is True
    assert pivot.verify_linear_form_proof(g, h, P, L, y, z, phi + 1, c) is False


def check_binary_transcript(group, gf, n=31, seed=3):
    """Opt-in binary Fiat-Shamir transcript (pivot.TRANSCRIPT = "binary"): prover and verifier of this package agree
    on it for every representation of the same statement, a proof is bound to the mode it was made in, and the
    reference mode is left untouched."""
    from verifiable_mpc_b200.ac20 import compressed_pivot as cp
    from verifiable_mpc_b200.ac20 import generators as gens
    from verifiable_mpc_b200.ac20 import pivot

    rng = random.Random(seed)
    gens.prng = rng
    generators = gens.create_generators(n, group)
    x = [gf(rng.randrange(gf.order)) for _ in range(n)]
    gamma = gf(rng.randrange(gf.order))
    L = pivot.LinearForm([gf(rng.randrange(gf.order)) for _ in range(n)])
    y = L(x)
    P = pivot.vector_commitment(x, gamma, generators["g"], generators["h"])
    host_generators = dict(generators, g=list(generators["g"]))  # plain Python list of group elements
    old = (pivot.TRANSCRIPT, cp.DEVICE_SCALAR_MIN, cp.DEVICE_SCALAR_MIN_PROVER, cp.FAST_INT_PATH)
    try:
        cp.prng = random.Random(seed + 1)
        ref_proof = cp.protocol_5_prover(generators, P, L, y, x, gamma, gf)
        assert cp.protocol_5_verifier(generators, P, L, y, ref_proof, gf) is True
        pivot.TRANSCRIPT = "binary"
        proofs = []
        for dev_min, fast in ((2, True), (1 << 30, True), (1 << 30, False)):  # device scalars / host ints / objects
            cp.DEVICE_SCALAR_MIN = cp.DEVICE_SCALAR_MIN_PROVER = dev_min
            cp.FAST_INT_PATH = fast
            cp.prng = random.Random(seed + 1)
            proofs.append(cp.protocol_5_prover(generators, P, L, y, x, gamma, gf))
            for gset in (generators, host_generators):
                assert cp.protocol_5_verifier(gset, P, L, y, proofs[-1], gf) is True
        for other in proofs[1:]:
            assert sorted(other) == sorted(proofs[0]) and all(other[k] == proofs[0][k] for k in other)
        assert any(proofs[0][k] != ref_proof[k] for k in ref_proof if k not in ("t", "A"))  # different challenges
        assert cp.protocol_5_verifier(generators, P, L, y, ref_proof, gf) is False   # made for the other transcript
        assert cp.protocol_5_verifier(generators, P, L, y + 1, proofs[0], gf) is False
        bad = dict(proofs[0])
        bad["A0"], bad["B0"] = bad["B0"], bad["A0"]
        assert cp.protocol_5_verifier(generators, P, L, y, bad, gf) is False
        # basic pivot in the same mode
        pivot.prng = random.Random(seed + 2)
        z, phi, c = pivot.prove_linear_form_eval(generators["g"], generators["h"], P, L, y, x, int(gamma), gf)
        assert pivot.verify_linear_form_proof(generators["g"], generators["h"], P, L, y, z, phi, c) is True
        assert pivot.verify_linear_form_proof(host_generators["g"], generators["h"], P, L, y, z, phi, c) is True
        pivot.TRANSCRIPT = "reference"
        assert pivot.verify_linear_form_proof(generators["g"], generators["h"], P, L, y, z, phi, c) is False
        assert cp.protocol_5_verifier(generators, P, L, y, proofs[0], gf) is False
        assert cp.protocol_5_verifier(generators, P, L, y, ref_proof, gf) is True
    finally:
        pivot.TRANSCRIPT, cp.DEVICE_SCALAR_MIN, cp.DEVICE_SCALAR_MIN_PROVER, cp.FAST_INT_PATH = old


# ------------------------------------------------------------------------------------------------ configured sizes
def _golden(name):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    if name.endswith(".gz"):
        import gzip

        return json.loads(gzip.open(path).read())
    return json.load(open(path))


def _ref_proof(group, gf, ref):
    rounds = len(ref["A_i"])
    proof = {"t": gf(int(ref["t"], 16)), "A": group._make(pt(ref["A"])),
             "z_prime": [gf(int(v, 16)) for v in ref["z_prime"]]}
    for i in range(rounds):
        proof[f"A{i}"] = group._make(pt(ref["A_i"][i]))
        proof[f"B{i}"] = group._make(pt(ref["B_i"][i]))
    return proof


def _assert_proof_equals(proof, ref):
    rounds = len(ref["A_i"])
    assert sorted(proof) == sorted(["t", "A", "z_prime"] + [f"A{i}" for i in range(rounds)] + [f"B{i}" for i in range(rounds)])
    assert proof["t"].value == int(ref["t"], 16)
    assert proof["A"].affine() == pt(ref["A"])
    for i in range(rounds):
        assert proof[f"A{i}"].affine() == pt(ref["A_i"][i]), f"A{i}"
        assert proof[f"B{i}"].affine() == pt(ref["B_i"][i]), f"B{i}"
    assert [v.value for v in proof["z_prime"]] == [int(v, 16) for v in ref["z_prime"]]


def _enc_proof(proof):
    rounds = sum(1 for key in proof if key.startswith("A") and key != "A")

    def enc(p):
        x, y = p.affine()
        return [hex(x), hex(y)]

    return {"t": hex(proof["t"].value), "A": enc(proof["A"]), "A_i": [enc(proof[f"A{i}"]) for i in range(rounds)],
            "B_i": [enc(proof[f"B{i}"]) for i in range(rounds)], "z_prime": [hex(v.value) for v in proof["z_prime"]]}


def check_big_case(k, group, gf, precomputed=False):
    """N = 2^k (k = 10, 12, 16): the inputs are replayed from the seed (tests/golden/seeded_inputs.py), the commitment
    and the whole proof must equal what the UNMODIFIED reference produced (tests/golden/ac20_big_<k>.json), and the
    reference's proof must verify here."""
    import hashlib

    from golden.seeded_inputs import ac20_draw_inputs, canonical_proof_text
    from verifiable_mpc_b200.ac20 import compressed_pivot as cp
    from verifiable_mpc_b200.ac20 import generators as gens
    from verifiable_mpc_b200.ac20 import pivot

    gold = _golden(f"ac20_big_{k}.json")
    exps, k_exp, xi, gi, Li = ac20_draw_inputs(k, gold["seed"], group.order)
    generators = gens.create_generators(len(exps), group, with_k=False, exponents=exps)
    generators["k"] = group.generator ** k_exp
    assert generators["k"].affine() == pt(gold["k_point"])
    for i, p in gold["generator_spots"].items():
        assert generators["g"][int(i)].affine() == pt(p), f"generator {i}"
    if precomputed and hasattr(generators["g"], "precompute"):
        generators["g"].precompute()
    x = [gf(v) for v in xi]
    gamma = gf(gi)
    L = pivot.LinearForm([gf(v) for v in Li])
    y = L(x)
    assert y.value == int(gold["y"], 16)
    P = pivot.vector_commitment(x, gamma, generators["g"], generators["h"])
    assert P.affine() == pt(gold["P"])
    rng = random.Random(gold["seed"] + 1)
    cp.prng = rng
    pivot.prng = rng
    proof = cp.protocol_5_prover(generators, P, L, y, x, gamma, gf)
    _assert_proof_equals(proof, gold["proof"])
    assert hashlib.sha256(canonical_proof_text(_enc_proof(proof)).encode()).hexdigest() == gold["proof_sha256"]
    assert len(generators["g"]) == gold["n"]
    assert cp.protocol_5_verifier(generators, P, L, y, _ref_proof(group, gf, gold["proof"]), gf) is True
    bad = _ref_proof(group, gf, gold["proof"])
    last = len(gold["proof"]["A_i"]) - 1
    bad[f"B{last}"] = bad[f"A{last}"]
    assert cp.protocol_5_verifier(generators, P, L, y, bad, gf) is False


def check_demo_case(group, gf):
    """BASELINE config 1 (demos/demo_zkp_ac20.py --elliptic, N = 128): the calls the reference's driver makes --
    pivot.vector_commitment at circuit_sat_cb.py:103 and compressed_pivot.protocol_5_prover at :264 -- replayed on the
    captured statement with the operand TYPES the driver uses (unreduced Python ints as coefficients of L, ints and
    field elements mixed in the witness); results must equal the unmodified reference's
    (tests/golden/ac20_demo_n128.json.gz)."""
    from verifiable_mpc_b200.ac20 import compressed_pivot as cp
    from verifiable_mpc_b200.ac20 import pivot
    from verifiable_mpc_b200.fingroups import DevicePointList

    def val(e):
        return gf(int(e[1], 16)) if e[0] == "f" else int(e[1], 16)

    gold = _golden("ac20_demo_n128.json.gz")
    pv = gold["pivot"]
    g_host = [group._make(pt(p)) for p in pv["g"]]
    h, kk = group._make(pt(pv["h"])), group._make(pt(pv["k"]))
    assert h == group.generator
    com = gold["commitments"][0]
    ref_proof_enc = dict(pv["proof"], t=pv["proof"]["t"][1], z_prime=[v[1] for v in pv["proof"]["z_prime"]])
    for g in (g_host, DevicePointList.from_points(group, g_host)):  # list of elements, as the reference passes; resident
        z_commitment = pivot.vector_commitment([val(v) for v in com["x"]], val(com["gamma"]), g, h)
        assert z_commitment.affine() == pt(com["out"])
        generators = {"g": g, "h": h, "k": kk}
        P = group._make(pt(pv["P"]))
        # an AffineForm, as circuit_sat_cb.py:264 passes it (protocol 5 makes it linear itself, compressed_pivot.py:97)
        assert pv["L_type"] == "AffineForm"
        L = pivot.AffineForm([val(v) for v in pv["L"]], val(pv["L_constant"]))
        x = [val(v) for v in pv["x"]]
        y, gamma = val(pv["y"]), val(pv["gamma"])
        assert L(x) == y
        cp.prng = random.Random(pv["prng_seed"])
        proof = cp.protocol_5_prover(generators, P, L, y, x, gamma, gf)
        _assert_proof_equals(proof, ref_proof_enc)
        assert cp.protocol_5_verifier(generators, P, L, y, _ref_proof(group, gf, ref_proof_enc), gf) is True
        assert cp.protocol_5_verifier(generators, P, L, y + 1, _ref_proof(group, gf, ref_proof_enc), gf) is False
