#!/usr/bin/env python
"""Device-timed generator fold (compressed_pivot.py:64): g'_j = c*g_j + g_{half+j} in place, incl. normalisation."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verifiable_mpc_b200 import Context  # noqa: E402

import time  # noqa: E402

from verifiable_mpc_b200 import _lib  # noqa: E402

LP_FOLD = 253 * (4 * 72 + 4 * 44) + 85 * 504 + 504  # doublings + NAF additions + final addition (limb products)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2half", type=int, nargs="+", default=[4, 8, 10, 12, 14, 15, 16, 18])
    ap.add_argument("--out", default="")
    ap.add_argument("--quad-max", type=int, default=-1, help="experiment: VMSM_OPT_FOLD_QUAD_MAX")
    args = ap.parse_args()
    ctx = Context(0)
    if args.quad_max >= 0:
        ctx.set_option(_lib.OPT_FOLD_QUAD_MAX, args.quad_max)
    peak = ctx.imad_peak()
    out = open(args.out, "a") if args.out else None
    c = 0x0EADBEEFCAFEBABE123456789ABCDEF0123456789ABCDEF0123456789ABCDEF % (2**252 + 27742317777372353535851937790883648493)
    for lg in args.log2half:
        half = 1 << lg
        best = None
        for rep in range(3):
            pts = ctx.fixed_base(seed=0x5EEE + rep, n=2 * half)
            ctx.sync()
            ctx.timer_start()
            pts.fold(c)
            ms = ctx.timer_stop()
            best = ms if best is None else min(best, ms)
            pts.free()
        rec = {"bench": "fold", "half": half, "ms": best, "Melem_s": half / best / 1e3,
               "frac_of_imad_peak": half * LP_FOLD / (best * 1e-3) / (peak * 1e12)}
        print(json.dumps(rec), flush=True)
        if out:
            out.write(json.dumps(rec) + "\n")
    # the field-vector halvings that accompany every generator fold (z' = z_L + c z_R, L' = c L_L + L_R), the cross-term
    # dot product and the coefficient text: HBM-side figures.  Algorithmic bytes per output element:
    # fold 64 B read + 32 B written, dot 64 B read, text 32 B read + ~78 B written.
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0
    for lg in args.log2half:
        half = 1 << lg
        sc = ctx.synth_scalars(0x5EED, 2 * half)
        other = ctx.synth_scalars(0x5EEF, 2 * half)
        res = {}
        for name, fn, nbytes in (
                ("scalars_fold", lambda: sc.fold(half, c, _lib.FOLD_WITNESS), 96 * half),
                ("scalars_dot", lambda: ctx.scalars_dot(sc, 0, other, half, half), 64 * half)):
            fn()
            ctx.sync()
            best = None
            for rep in range(5):
                ctx.sync()
                if name == "scalars_dot":  # synchronous call (returns the value): wall clock around it
                    t0 = time.perf_counter()
                    fn()
                    ms = (time.perf_counter() - t0) * 1e3
                else:
                    ctx.timer_start()
                    fn()
                    ms = ctx.timer_stop()
                best = ms if best is None else min(best, ms)
            res[name] = {"ms": best, "GB_s": nbytes / (best * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (best * 1e-3) / 1e9 / hbm}
        sc.text_bytes(0, half)  # first call of a new maximum size allocates the (pinned) text buffers
        t0 = time.perf_counter()
        txt = sc.text_bytes(0, half)
        ms = (time.perf_counter() - t0) * 1e3
        res["scalars_text"] = {"ms_incl_d2h": ms, "bytes_out": len(txt)}
        rec = {"bench": "scalar_vectors", "half": half, "hbm_peak_GB_s": hbm, **res}
        print(json.dumps(rec), flush=True)
        if out:
            out.write(json.dumps(rec) + "\n")
        sc.free()
        other.free()
    ctx.close()


if __name__ == "__main__":
    main()
