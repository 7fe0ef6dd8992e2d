#!/usr/bin/env python
"""AC20 compressed-pivot prove / verify latency on the device twins (BASELINE.json config 3: N = 2^16 generators).

Synthetic statement below the circuit front-end: generators g_i = r_i*B (device fixed-base kernel),
witness x, blinding gamma and linear form L uniform, P = commitment; then protocol_5_prover / protocol_5_verifier.
Prints one JSON line per size with a host/device breakdown.
"""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


_ACC = {"hash": 0.0, "commit": 0.0, "fold": 0.0, "lincomb": 0.0}
_instrumented = False


def instrument(group):
    """Accumulate the time spent hashing (transcript text / bytes + SHA-256) and inside the device-call wrappers."""
    global _instrumented
    from verifiable_mpc_b200.ac20 import compressed_pivot as cp
    from verifiable_mpc_b200.ac20 import pivot

    if _instrumented:
        return
    _instrumented = True

    def timed(name, fn):
        def wrapper(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                _ACC[name] += time.perf_counter() - t0
        return wrapper

    pivot.fiat_shamir_hash = timed("hash", pivot.fiat_shamir_hash)
    pivot.fiat_shamir_prefix = timed("hash", pivot.fiat_shamir_prefix)  # transcript text (device + host) + SHA-256
    pivot.binary_prefix = timed("hash", pivot.binary_prefix)
    pivot.vector_commitment = timed("commit", pivot.vector_commitment)
    cp._fold_generators = timed("fold", cp._fold_generators)
    group.lincomb = classmethod(timed("lincomb", group.lincomb.__func__))


def measure(group, gf, logn, repeat=2, transcript="reference", precompute=False):
    """Best of `repeat` prove / verify runs of the compressed pivot at N = 2^logn generators (synthetic statement
    below the circuit front-end); returns the record tools/bench_ac20.py prints and bench.py embeds."""
    from verifiable_mpc_b200.ac20 import compressed_pivot as cp
    from verifiable_mpc_b200.ac20 import generators as gens
    from verifiable_mpc_b200.ac20 import pivot

    instrument(group)
    acc = _ACC
    ctx = group._ctx()
    old_transcript, pivot.TRANSCRIPT = pivot.TRANSCRIPT, transcript
    try:
        N = 1 << logn
        n = N - 1
        rng = random.Random(logn)
        gens.prng = rng
        t0 = time.perf_counter()
        generators = gens.create_generators(n, group)
        ctx.sync()
        t_gen = time.perf_counter() - t0
        t_pre = None
        if precompute:
            t0 = time.perf_counter()
            generators["g"].precompute()
            ctx.sync()
            t_pre = time.perf_counter() - t0
        x = [gf(rng.randrange(gf.order)) for _ in range(n)]
        gamma = gf(rng.randrange(gf.order))
        L = pivot.LinearForm([gf(rng.randrange(gf.order)) for _ in range(n)])
        y = L(x)
        t0 = time.perf_counter()
        P = pivot.vector_commitment(x, gamma, generators["g"], generators["h"])
        t_commit = time.perf_counter() - t0
        best = None
        for rep in range(repeat):
            cp.prng = random.Random(1000 + rep)
            for k in acc:
                acc[k] = 0.0
            t0 = time.perf_counter()
            proof = cp.protocol_5_prover(generators, P, L, y, x, gamma, gf)
            t_prove = time.perf_counter() - t0
            prove_parts = dict(acc)
            for k in acc:
                acc[k] = 0.0
            t0 = time.perf_counter()
            ok = cp.protocol_5_verifier(generators, P, L, y, proof, gf)
            t_verify = time.perf_counter() - t0
            rec = {"N": N, "transcript": transcript, "rounds": logn - 1, "prove_s": t_prove, "verify_s": t_verify,
                   "verified": bool(ok), "create_generators_s": t_gen, "z_commitment_s": t_commit,
                   "precomputed_generator_table": bool(precompute), "precompute_s": t_pre,
                   "prove_breakdown_s": {k: round(v, 4) for k, v in prove_parts.items()},
                   "prove_host_other_s": round(t_prove - sum(prove_parts.values()), 4),
                   "verify_breakdown_s": {k: round(v, 4) for k, v in acc.items()}}
            if best is None or rec["prove_s"] < best["prove_s"]:
                best = rec
        generators["g"].dev.free()
        return best
    finally:
        pivot.TRANSCRIPT = old_transcript


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, nargs="+", default=[10, 13, 16])
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--scalar-min", type=int, default=-1, help="experiment: compressed_pivot.DEVICE_SCALAR_MIN")
    ap.add_argument("--transcript", default="reference", choices=["reference", "binary"],
                    help="binary: the opt-in canonical-bytes Fiat-Shamir transcript (not verifiable by the reference)")
    ap.add_argument("--precompute", action="store_true", help="fixed generators with a table (DevicePointList.precompute)")
    ap.add_argument("--opt", type=int, nargs=2, action="append", default=[], metavar=("KEY", "VALUE"),
                    help="experiment: vmsm_ctx_set_option(KEY, VALUE) before measuring (repeatable)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    from verifiable_mpc_b200 import fingroups
    from verifiable_mpc_b200.ac20 import compressed_pivot as cp
    from verifiable_mpc_b200.finfields import GF

    if args.scalar_min >= 0:
        cp.DEVICE_SCALAR_MIN = args.scalar_min
    group = fingroups.EllipticCurve("Ed25519", "projective")
    group.is_additive, group.is_multiplicative = False, True
    gf = GF(group.order)
    ctx = group._ctx()
    for key, value in args.opt:
        ctx.set_option(key, value)
    out = open(args.out, "a") if args.out else None
    for logn in args.log2n:
        best = measure(group, gf, logn, args.repeat, args.transcript, args.precompute)
        best["options"] = args.opt
        print(json.dumps(best), flush=True)
        if out:
            out.write(json.dumps(best) + "\n")
    ctx.close()


if __name__ == "__main__":
    main()
