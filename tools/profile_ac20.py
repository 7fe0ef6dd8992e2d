#!/usr/bin/env python
"""cProfile of the compressed-pivot prover / verifier twins at N = 2^logn (development tool): where the HOST time goes."""
import cProfile
import io
import os
import pstats
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    from verifiable_mpc_b200 import fingroups
    from verifiable_mpc_b200.ac20 import compressed_pivot as cp
    from verifiable_mpc_b200.ac20 import generators as gens
    from verifiable_mpc_b200.ac20 import pivot
    from verifiable_mpc_b200.finfields import GF

    group = fingroups.EllipticCurve("Ed25519", "projective")
    group.is_additive, group.is_multiplicative = False, True
    gf = GF(group.order)
    n = (1 << logn) - 1
    rng = random.Random(logn)
    gens.prng = rng
    generators = gens.create_generators(n, group)
    x = [gf(rng.randrange(gf.order)) for _ in range(n)]
    gamma = gf(rng.randrange(gf.order))
    L = pivot.LinearForm([gf(rng.randrange(gf.order)) for _ in range(n)])
    y = L(x)
    P = pivot.vector_commitment(x, gamma, generators["g"], generators["h"])
    cp.prng = random.Random(1)
    cp.protocol_5_prover(generators, P, L, y, x, gamma, gf)  # warm-up
    for what in ("prove", "verify"):
        cp.prng = random.Random(2)
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        if what == "prove":
            proof = cp.protocol_5_prover(generators, P, L, y, x, gamma, gf)
        else:
            ok = cp.protocol_5_verifier(generators, P, L, y, proof, gf)
        pr.disable()
        dt = time.perf_counter() - t0
        out = io.StringIO()
        pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(22)
        print(f"==== {what} N=2^{logn}: {dt * 1e3:.1f} ms (under cProfile)")
        print("\n".join(out.getvalue().splitlines()[4:40]))
    assert ok
    # timeline marks of the device-resident prover (compressed_pivot.TRACE), without the profiler
    for rep in range(2):
        cp.prng = random.Random(3)
        cp.TRACE = []
        t0 = time.perf_counter()
        cp.protocol_5_prover(generators, P, L, y, x, gamma, gf)
        t1 = time.perf_counter()
        marks, cp.TRACE = cp.TRACE, None
    agg, prev = {}, t0
    for label, t in marks:
        agg[label] = agg.get(label, 0.0) + (t - prev)
        prev = t
    agg["(after last mark)"] = t1 - prev
    print(f"==== prove N=2^{logn}: {1e3 * (t1 - t0):.1f} ms; time up to each mark, summed over rounds (ms):")
    for label, v in agg.items():
        print(f"   {label:28s} {1e3 * v:8.2f}")
    print(f"   (before p5:start: draws, checks, g_hat clone) {1e3 * (marks[0][1] - t0):.2f}")


if __name__ == "__main__":
    main()
