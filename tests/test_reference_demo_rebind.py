"""The drop-in claim, end to end on the reference's own driver: ``demos/demo_zkp_ac20.py --elliptic`` (BASELINE
config 1: circuit builder -> ``circuit_sat_cb.circuit_sat_prover`` / ``_verifier``, compressed pivot, N = 128) is run
UNMODIFIED on the MPyC look-alike, then again with the group type and the pivot functions rebound to this package as
INTEGRATION.md section 2 prescribes (device = tests/fake_engine.py here; the GPU box has no reference tree).  With the
same seeds both runs must produce the same proof, entry by entry, and both verifications must pass."""
import importlib.util
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "verifiable_mpc")), reason="reference tree not mounted")


class _Quiet:
    def pprint(self, *a, **k):
        pass


def _norm(v):
    if hasattr(v, "affine"):
        return ("pt", tuple(v.affine()))
    if hasattr(v, "normalize") and hasattr(v, "x"):
        p = v.normalize()
        return ("pt", (int(p.x), int(p.y)))
    if isinstance(v, dict):
        return {k: _norm(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_norm(x) for x in v]
    if hasattr(v, "value"):
        return ("fe", int(v.value))
    return repr(v)


@pytest.mark.parametrize("device_scalar_min", [256, 2])
def test_reference_demo_with_rebound_pivot_gives_the_same_proof(device_scalar_min, capsys, monkeypatch):
    for p in (os.path.join(ROOT, "oracle", "mpyc_shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    monkeypatch.setattr(sys, "argv", ["demo_zkp_ac20.py"])
    spec = importlib.util.spec_from_file_location("demo_zkp_ac20_under_test", os.path.join(REF, "demos", "demo_zkp_ac20.py"))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    import verifiable_mpc.ac20.circuit_builder as rcb
    import verifiable_mpc.ac20.circuit_sat_cb as rcs
    import verifiable_mpc.ac20.circuit_sat_r1cs as rr1cs
    import verifiable_mpc.ac20.compressed_pivot as rcp
    import verifiable_mpc.ac20.pivot as rpivot
    from fake_engine import FakeContext
    from verifiable_mpc_b200 import fingroups
    import verifiable_mpc_b200.ac20.compressed_pivot as gcp
    import verifiable_mpc_b200.ac20.pivot as gpivot

    demo.GROUP = "Elliptic"
    demo.pp = _Quiet()
    captured = {}
    orig_prover = rcs.circuit_sat_prover

    def capture(*a, **k):
        captured["proof"] = orig_prover(*a, **k)
        return captured["proof"]

    monkeypatch.setattr(rcs, "circuit_sat_prover", capture)

    def seed_all(s):
        for mod in (rpivot, rcp, rcs, rr1cs, rcb):
            if hasattr(mod, "prng"):
                monkeypatch.setattr(mod, "prng", random.Random(s))

    # run A: the reference as it is
    seed_all(42)
    checks_a = demo.main(demo.cs.PivotChoice.compressed, 3)
    proof_a = _norm(captured["proof"])
    assert all(checks_a.values())

    # run B: INTEGRATION.md section 2 -- this package's group type and pivot functions bound into the reference
    monkeypatch.setattr(fingroups.Ed25519Point, "context", FakeContext())
    gpivot._single_cache.clear()
    monkeypatch.setattr(demo, "EllipticCurve", fingroups.EllipticCurve)
    monkeypatch.setattr(rpivot, "vector_commitment", gpivot.vector_commitment)
    monkeypatch.setattr(rpivot, "list_mul", gpivot.list_mul)
    for name in ("protocol_4_prover", "protocol_4_verifier", "protocol_5_prover", "protocol_5_verifier"):
        monkeypatch.setattr(rcp, name, getattr(gcp, name))
    monkeypatch.setattr(gcp, "DEVICE_SCALAR_MIN", device_scalar_min)
    monkeypatch.setattr(gcp, "DEVICE_SCALAR_MIN_PROVER", device_scalar_min)
    calls = FakeContext.calls
    seed_all(42)
    monkeypatch.setattr(gcp, "prng", rcp.prng)
    monkeypatch.setattr(gpivot, "prng", rpivot.prng)
    checks_b = demo.main(demo.cs.PivotChoice.compressed, 3)
    proof_b = _norm(captured["proof"])
    gpivot._single_cache.clear()
    capsys.readouterr()
    assert FakeContext.calls > calls, "the rebound functions did not reach the engine"
    assert all(checks_b.values()) and checks_a == checks_b
    assert sorted(proof_a) == sorted(proof_b)
    for key in proof_a:
        assert proof_a[key] == proof_b[key], key
