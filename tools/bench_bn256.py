#!/usr/bin/env python
"""BN256 prover-key multi-exponentiations (BASELINE.json config 4): device-timed G1 / G2 MSMs and the wall time of the
compute_proof twin for a synthetic 2^14-constraint QAP (|mid| = len(h) = 2^14, bases r_i*G generated on the device).
Prints JSON lines."""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_proof_case(ctx, m, seed=7):
    """A synthetic compute_proof input below the QAP front end: |mid| = len(h) = m, every base vector of the evaluation
    key generated on the device (r_i * G) instead of uploaded.  Returns (qap, witness, h class, deltas, prepared key)."""
    from verifiable_mpc_b200 import fingroups
    from verifiable_mpc_b200.engine import BN_N
    from verifiable_mpc_b200.trinocchio import pynocchio as twin

    g1 = fingroups.EllipticCurve("BN256", "jacobian")
    g2 = fingroups.EllipticCurve("BN256_twist", "jacobian")
    fingroups.BN256Point.context = ctx
    rng = random.Random(seed)

    class Q:
        indices_mid = range(3, 3 + m)

    class H:
        coeffs = [rng.randrange(BN_N) for _ in range(m)]

        def __len__(self):
            return len(self.coeffs)

    class D:
        v, w, y = (rng.randrange(BN_N) for _ in range(3))

    c = [rng.randrange(BN_N) for _ in range(m + 3)]

    class Prepared(twin.PreparedEvalKey):
        def __init__(self):
            self.indices_mid, self.h_len, self.groups, self.bases = list(Q.indices_mid), m, {}, {}
            for k, (name, _, deltas) in enumerate(twin._MID_SUMS):
                group = g2 if name.endswith("g2") else g1
                self.groups[name] = group
                self.bases[name] = ctx.fixed_base(seed=100 + k, n=m + len(deltas), curve=group.curve_id)
            self.groups["h*g1"] = g1
            self.bases["h*g1"] = ctx.fixed_base(seed=99, n=m, curve=1)

    return Q, c, H, D, Prepared()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, nargs="+", default=[10, 12, 14])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--windows", type=int, nargs="+", default=[0])
    ap.add_argument("--radix", type=int, default=-1, help="experiment: VMSM_OPT_REDUCE_RADIX (log2 of the tree radix)")
    ap.add_argument("--quad-acc", type=int, default=-1, help="experiment: VMSM_OPT_BN_QUAD_ACC")
    ap.add_argument("--curves", type=int, nargs="+", default=[1, 2], help="1 = G1, 2 = G2")
    ap.add_argument("--no-proof", action="store_true", help="skip the compute_proof part")
    ap.add_argument("--table-windows", type=int, nargs="+", default=[0], help="window bits of the key tables (0 = default)")
    ap.add_argument("--profile", action="store_true", help="cProfile of one compute_proof call (host side)")
    ap.add_argument("--bn-sets", type=int, nargs="+", default=[0], help="VMSM_OPT_BN_PRE_SETS values for the table MSMs")
    ap.add_argument("--bn-seg-len", type=int, nargs="+", default=[0], help="VMSM_OPT_BN_SEG_LEN values for the table MSMs")
    ap.add_argument("--seg-mode", type=int, nargs="+", default=[1], help="VMSM_OPT_SEG_MODE values for the table MSMs")
    ap.add_argument("--quad-fix", type=int, default=-1, help="experiment: VMSM_OPT_BN_QUAD_FIX")
    ap.add_argument("--plain-seg-mode", type=int, nargs="+", default=[1], help="VMSM_OPT_SEG_MODE values for the plain-path MSMs")
    ap.add_argument("--plain-seg-len", type=int, nargs="+", default=[0], help="VMSM_OPT_BN_SEG_LEN values for the plain-path MSMs")
    ap.add_argument("--table-log2n", type=int, nargs="+", default=[], help="sizes of the table MSM sweep (default: max log2n)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from verifiable_mpc_b200 import Context, _lib, fingroups
    from verifiable_mpc_b200.engine import BN_N
    from verifiable_mpc_b200.trinocchio import pynocchio as twin

    ctx = Context(0)
    ctx.set_option(_lib.OPT_PHASE_TIMING, 1)
    if args.quad_acc >= 0:
        ctx.set_option(_lib.OPT_BN_QUAD_ACC, args.quad_acc)
    if args.quad_fix >= 0:
        ctx.set_option(_lib.OPT_BN_QUAD_FIX, args.quad_fix)
    if args.radix >= 0:
        ctx.set_option(_lib.OPT_REDUCE_RADIX, args.radix)
    out = open(args.out, "a") if args.out else None

    def emit(rec):
        print(json.dumps(rec), flush=True)
        if out:
            out.write(json.dumps(rec) + "\n")

    for logn in args.log2n:
        n = 1 << logn
        for curve, name in [(cv, "G%d" % cv) for cv in args.curves]:
            sets = [(ctx.fixed_base(seed=0x5EEE + 16 * k, n=n, curve=curve), ctx.synth_scalars(0x5EED + 16 * k, n, curve=curve))
                    for k in range(3)]
            for c, psm, psl in [(c_, a, b) for c_ in args.windows for a in args.plain_seg_mode for b in args.plain_seg_len]:
                ctx.set_option(_lib.OPT_WINDOW_BITS, c)
                ctx.set_option(_lib.OPT_SEG_MODE, psm)
                ctx.set_option(_lib.OPT_BN_SEG_LEN, psl)
                for w in range(3):
                    ctx.msm_dev(*sets[w % 3], slot=0)
                ctx.sync()
                ctx.phase_times()
                ctx.timer_start()
                for s in range(args.steps):
                    ctx.msm_dev(*sets[s % 3], slot=s % 32)
                ms = ctx.timer_stop() / args.steps
                ph, calls = ctx.phase_times()
                emit({"bench": "bn256_msm", "group": name, "log2n": logn, "window": c, "seg_mode": psm, "seg_len": psl,
                      "ms": ms, "Mpts_s": n / ms / 1e3,
                      "phase_ms": {k: round(v / max(calls, 1), 4) for k, v in ph.items()}})
            ctx.set_option(_lib.OPT_WINDOW_BITS, 0)
            ctx.set_option(_lib.OPT_SEG_MODE, 1)
            ctx.set_option(_lib.OPT_BN_SEG_LEN, 0)
            for p, s in sets:
                p.free()
                s.free()

    if not args.no_proof:
        m = 1 << max(args.log2n)
        Q, c, H, D, prepared = synthetic_proof_case(ctx, m)
        for tables in (False, True):
            if tables:
                t0 = time.perf_counter()
                prepared.precompute()
                ctx.sync()
                t_tab = time.perf_counter() - t0
            ref = twin.compute_proof(Q, c, H(), prepared, D)
            best = 1e9
            for _ in range(5):
                t0 = time.perf_counter()
                proof = twin.compute_proof(Q, c, H(), prepared, D)
                best = min(best, time.perf_counter() - t0)
            if tables:
                assert all(proof[k] == first[k] for k in first), "tables changed the proof"
            first = proof
            emit({"bench": "pynocchio_compute_proof", "mid": m, "len_h": m, "prove_s": best, "key_tables": tables,
                  "table_build_s": t_tab if tables else None,
                  "note": "8 MSMs (6 G1 + 1 G2 over mid wires, 1 G1 over h), bases resident, scalars packed on the host per call"})
        if args.profile:
            twin.TRACE = []
            twin.compute_proof(Q, c, H(), prepared, D)
            t0 = twin.TRACE[0][1]
            print("compute_proof timeline (ms):", [(lab, round(1e3 * (t - t0), 3)) for lab, t in twin.TRACE])
            twin.TRACE = None
            import cProfile
            import io
            import pstats

            pr = cProfile.Profile()
            pr.enable()
            twin.compute_proof(Q, c, H(), prepared, D)
            pr.disable()
            buf = io.StringIO()
            pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(18)
            print(buf.getvalue())
    # device-timed MSMs over the tables
    for tlog in (args.table_log2n or [max(args.log2n)]):
        tm = 1 << tlog
        for tc, curve, name in [(tc_, cv, nm) for tc_ in args.table_windows for cv, nm in ((1, "G1"), (2, "G2")) if cv in args.curves]:
            sets = [(ctx.fixed_base(seed=0x5EEE + 16 * k, n=tm, curve=curve).precompute(tc), ctx.synth_scalars(0x5EED + 16 * k, tm, curve=curve))
                    for k in range(3)]
            for mode, bsets, slen in [(a, b, c_) for a in args.seg_mode for b in args.bn_sets for c_ in args.bn_seg_len]:
                ctx.set_option(_lib.OPT_SEG_MODE, mode)
                ctx.set_option(_lib.OPT_BN_PRE_SETS, bsets)
                ctx.set_option(_lib.OPT_BN_SEG_LEN, slen)
                for w in range(3):
                    ctx.msm_dev(*sets[w % 3], slot=0)
                ctx.sync()
                ctx.phase_times()
                ctx.timer_start()
                for s in range(args.steps):
                    ctx.msm_dev(*sets[s % 3], slot=s % 32)
                ms = ctx.timer_stop() / args.steps
                ph, calls = ctx.phase_times()
                emit({"bench": "bn256_msm_key_tables", "group": name, "table_window": tc, "seg_mode": mode, "sets": bsets,
                      "seg_len": slen, "log2n": tlog, "ms": ms, "Mpts_s": tm / ms / 1e3,
                      "phase_ms": {k: round(v / max(calls, 1), 4) for k, v in ph.items()}})
            ctx.set_option(_lib.OPT_SEG_MODE, 1)
            ctx.set_option(_lib.OPT_BN_PRE_SETS, 0)
            ctx.set_option(_lib.OPT_BN_SEG_LEN, 0)
            for p, sc in sets:
                p.free()
                sc.free()
    ctx.close()


if __name__ == "__main__":
    main()
