"""CPU: the sparse AC20 form builder (verifiable_mpc_b200/ac20/sparse_forms.py) against the UNMODIFIED
reference's dense one (verifiable_mpc/ac20/circuit_builder.py:417-545 on oracle/mpyc_shim): same Lagrange vectors, same
forms coefficient by coefficient (exact unreduced integers: their text enters the Fiat-Shamir pre-images), same
multiplication triples; and the whole circuit-sat prover of the reference with the builder bound in produces the same
proof.  Needs the reference tree (skipped on the GPU box)."""
import os
import random
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "verifiable_mpc")), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    for p in (os.path.join(ROOT, "oracle", "mpyc_shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import verifiable_mpc.ac20.circuit_builder as cb
    import verifiable_mpc.ac20.recombine as recombine
    from mpyc.finfields import GF

    return cb, recombine, GF(2**252 + 27742317777372353535851937790883648493)


def demo_circuit(cb, n=3):
    """The circuit of demos/demo_zkp_ac20.py:54-67 (add / mul / scalar-mul gates, the != and >= gadgets)."""
    circuit = cb.Circuit()
    b = cb.CircuitVar(1, circuit, "b")
    c = cb.CircuitVar(2, circuit, "c")
    d = c + c + c * c + c * c * 1 + 1 + b
    e = d * d + c ** n + 10
    f = d * c + e
    f.label_output("f")
    g = f != 100
    g.label_output("g")
    h = g >= 10
    h.label_output("h")
    return circuit


def chain_circuit(cb, m, gf, seed=1):
    """m multiplication gates with shared sub-expressions, constants and scalar multiples: v <- (v + a) * (3 * v + b) + 1
    (field-element inputs: with plain ints the wire values would double in length at every gate)."""
    rnd = random.Random(seed)
    circuit = cb.Circuit()
    a = cb.CircuitVar(gf(rnd.randrange(1, 50)), circuit, "a")
    b = cb.CircuitVar(gf(rnd.randrange(1, 50)), circuit, "b")
    v = a * b
    for i in range(m - 1):
        s = v + a
        v = s * (3 * v + b) + 1 if i % 3 else s * b + v
    v.label_output("out")
    (v + a).label_output("out2")
    return circuit


def test_lagrange_vector_matches_recombination_vectors(ref):
    cb, recombine, gf = ref
    from verifiable_mpc_b200.ac20 import sparse_forms as sp

    rnd = random.Random(5)
    for count in (1, 2, 3, 7, 22, 45):
        for _ in range(3):
            c = rnd.randrange(gf.modulus)
            want = recombine._recombination_vectors(gf, tuple(range(count)), (c,))[0]
            assert sp.lagrange_vector(gf.modulus, count, c) == want, count
            assert sp.lagrange(gf, range(count), c) == cb.lagrange(gf, range(count), c)
    with pytest.raises(ZeroDivisionError):
        sp.lagrange_vector(gf.modulus, 5, 3)


@pytest.mark.parametrize("which", ["demo", "chain40", "chain160"])
def test_forms_equal_the_reference(ref, which):
    cb, recombine, gf = ref
    from verifiable_mpc_b200.ac20 import sparse_forms as sp

    circuit = demo_circuit(cb) if which == "demo" else chain_circuit(cb, int(which[5:]), gf)
    forms = sp.SparseCircuitForms(circuit)
    c = random.Random(9).randrange(gf.modulus)
    for wire in (0, 1):
        want = cb.calculate_fg_form(circuit, wire, c, gf)
        got = sp.calculate_fg_form(circuit, wire, c, gf, forms)
        assert got.coeffs == want.coeffs and got.constant == want.constant
        assert [type(v) for v in got.coeffs] == [type(v) for v in want.coeffs]
        assert repr(got) == repr(want)  # what enters the Fiat-Shamir pre-image
    assert sp.calculate_h_form(circuit, c, gf).coeffs == cb.calculate_h_form(circuit, c, gf).coeffs
    want = cb.calculate_circuit_forms(circuit)
    got = sp.calculate_circuit_forms(circuit, forms)
    assert [f.coeffs for f in got] == [f.coeffs for f in want] and [f.constant for f in got] == [f.constant for f in want]
    x = circuit.initial_inputs()
    assert forms.multiplication_triples(x) == circuit.multiplication_triples(x)
    # every wire of every multiplication gate, dense form by dense form
    for g in circuit.mul_gates()[:: max(1, circuit.mul_ct // 25)]:
        for wire in (0, 1):
            want = cb.construct_affine_form(g, circuit, wire)
            got = forms.dense(forms.wire_form(g, wire), ac20=False)
            assert got.coeffs == want.coeffs and got.constant == want.constant


def test_reference_prover_with_sparse_builder_gives_the_same_proof(ref, monkeypatch):
    """circuit_builder.calculate_fg_form / calculate_h_form / calculate_circuit_forms / lagrange and
    Circuit.multiplication_triples rebound to the sparse twins inside the unmodified circuit_sat_cb driver (QR group of
    the shim: no elliptic-curve work needed to compare transcripts) -- same proof dictionary, verification passes."""
    cb, recombine, gf0 = ref
    import verifiable_mpc.ac20.circuit_sat_cb as cs
    import verifiable_mpc.ac20.circuit_sat_r1cs as r1cs
    import verifiable_mpc.ac20.compressed_pivot as rcp
    import verifiable_mpc.ac20.pivot as rpivot
    from mpyc.finfields import GF
    from mpyc.fingroups import QuadraticResidues
    from verifiable_mpc_b200.ac20 import sparse_forms as sp

    group = QuadraticResidues(l=64)
    group.is_additive, group.is_multiplicative = False, True
    gf = GF(modulus=group.order)

    def run():
        for mod in (rpivot, rcp, cs, r1cs, cb):
            if hasattr(mod, "prng"):
                monkeypatch.setattr(mod, "prng", random.Random(77))
        circuit = demo_circuit(cb)
        x = circuit.initial_inputs()
        check, padding, g_length = cs.check_input_length_power_of_2(x, circuit)
        [cb.CircuitVar(0, circuit, "unused_" + str(i)) for i in range(padding)]
        x = circuit.initial_inputs()
        generators = cs.create_generators(g_length, cs.PivotChoice.compressed, group)
        proof = cs.circuit_sat_prover(generators, circuit, x, gf, cs.PivotChoice.compressed)
        checks = cs.circuit_sat_verifier(proof, generators, circuit, gf, cs.PivotChoice.compressed)
        return proof, checks

    proof_a, checks_a = run()
    assert all(checks_a.values())
    cache = {}

    def forms_of(circuit):
        if id(circuit) not in cache:
            cache[id(circuit)] = sp.SparseCircuitForms(circuit)
        return cache[id(circuit)]

    # the forms must be of the class the driver's own forms have (they are added to each other)
    monkeypatch.setattr(sp, "AffineForm", rpivot.AffineForm)
    monkeypatch.setattr(sp, "LinearForm", rpivot.LinearForm)
    monkeypatch.setattr(cb, "calculate_fg_form", lambda circuit, wire, challenge, gf: sp.calculate_fg_form(circuit, wire, challenge, gf, forms_of(circuit)))
    monkeypatch.setattr(cb, "calculate_h_form", sp.calculate_h_form)
    monkeypatch.setattr(cb, "calculate_circuit_forms", lambda circuit: sp.calculate_circuit_forms(circuit, forms_of(circuit)))
    monkeypatch.setattr(cb, "lagrange", sp.lagrange)
    monkeypatch.setattr(cb.Circuit, "multiplication_triples", lambda self, inputs: forms_of(self).multiplication_triples(inputs))
    proof_b, checks_b = run()
    assert all(checks_b.values())
    assert repr(proof_a) == repr(proof_b)


def test_builder_scales_linearly(ref):
    """m = 2^11 multiplication gates: both wire forms, the h form and the triples in well under a second of host time
    (the reference's dense builder needs ~m^2 list slots per wire: 0.09 s at m = 512, minutes and gigabytes at 2^15)."""
    cb, recombine, gf = ref
    from verifiable_mpc_b200.ac20 import sparse_forms as sp

    m = 1 << 11
    circuit = chain_circuit(cb, m, gf, seed=3)
    assert circuit.mul_ct == m
    t0 = time.perf_counter()
    forms = sp.SparseCircuitForms(circuit)
    c = random.Random(1).randrange(gf.modulus)
    f = sp.calculate_fg_form(circuit, 0, c, gf, forms)
    g = sp.calculate_fg_form(circuit, 1, c, gf, forms)
    h = sp.calculate_h_form(circuit, c, gf)
    x = circuit.initial_inputs()
    a, b, prod = forms.multiplication_triples(x)
    dt = time.perf_counter() - t0
    assert len(f.coeffs) == len(g.coeffs) == len(h.coeffs) == circuit.input_ct + 3 + 2 * m
    assert dt < 5.0, dt
