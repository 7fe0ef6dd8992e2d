#!/usr/bin/env python
"""Device-timed generator fold (compressed_pivot.py:64): g'_j = c*g_j + g_{half+j} in place, incl. normalisation."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verifiable_mpc_b200 import Context  # noqa: E402

LP_FOLD = 253 * (4 * 72 + 4 * 44) + 85 * 504 + 504  # doublings + NAF additions + final addition (limb products)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2half", type=int, nargs="+", default=[4, 8, 10, 12, 14, 15, 16, 18])
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    ctx = Context(0)
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
    ctx.close()


if __name__ == "__main__":
    main()
