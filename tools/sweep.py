#!/usr/bin/env python
"""Device-timed MSM sweep over n and tuning options (development tool; writes JSON lines).
    python tools/sweep.py --logn 10 12 14 16 18 20 22 --windows 0 --out gpurun_out/sweep.jsonl
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verifiable_mpc_b200 import Context, _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, nargs="+", default=[10, 12, 14, 16, 18, 20])
    ap.add_argument("--windows", type=int, nargs="+", default=[0])
    ap.add_argument("--sort", type=int, nargs="+", default=[1])
    ap.add_argument("--radix", type=int, nargs="+", default=[3])
    ap.add_argument("--cap", type=int, nargs="+", default=[0])
    ap.add_argument("--async-tail", type=int, nargs="+", default=[1])
    ap.add_argument("--sort-blocks", type=int, nargs="+", default=[-1])
    ap.add_argument("--dual-head", type=int, default=1, help="VMSM_OPT_DUAL_HEAD")
    ap.add_argument("--async-sort", type=int, default=1, help="0: counting sort on the main stream (serial phase times)")
    ap.add_argument("--quad-threshold", type=int, default=-1, help="experiment: VMSM_OPT_QUAD_THRESHOLD")
    ap.add_argument("--precompute", type=int, nargs="+", default=[-1],
                    help="-1: plain path; 0 / 8..16: tables of 2^(c*w)*P_i with this window (0 = by size)")
    ap.add_argument("--pre-sets", type=int, nargs="+", default=[0], help="VMSM_OPT_PRE_SETS values to sweep")
    ap.add_argument("--seg-len", type=int, nargs="+", default=[0], help="VMSM_OPT_SEG_LEN values to sweep (0 = whole waves)")
    ap.add_argument("--seg-mode", type=int, nargs="+", default=[1], help="VMSM_OPT_SEG_MODE values to sweep")
    ap.add_argument("--block-sort", type=int, nargs="+", default=[0], help="VMSM_OPT_BLOCK_SORT values to sweep")
    ap.add_argument("--block-sort-min", type=int, default=-1, help="VMSM_OPT_BLOCK_SORT_MIN")
    ap.add_argument("--acc-carveout", type=int, default=-2, help="VMSM_OPT_ACC_CARVEOUT (-1 = driver default)")
    ap.add_argument("--shard-self", type=int, default=0,
                    help="1: every MSM is issued as the only shard of a 1-rank mailbox (push + gather kernels, device normalisation)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    ctx = Context(0)
    ctx.set_option(_lib.OPT_PHASE_TIMING, 1)
    if args.quad_threshold >= 0:
        ctx.set_option(_lib.OPT_QUAD_THRESHOLD, args.quad_threshold)
    ctx.set_option(_lib.OPT_ASYNC_SORT, args.async_sort)
    ctx.set_option(_lib.OPT_DUAL_HEAD, args.dual_head)
    if args.block_sort_min >= 0:
        ctx.set_option(_lib.OPT_BLOCK_SORT_MIN, args.block_sort_min)
    if args.acc_carveout >= -1:
        ctx.set_option(_lib.OPT_ACC_CARVEOUT, args.acc_carveout)
    if args.shard_self:
        ctx.mailbox_create(1)
    shard_seq = [0]

    def issue(pts, sc, slot):
        if args.shard_self:
            shard_seq[0] += 1
            ctx.set_option(_lib.OPT_SHARD_SEQ, shard_seq[0])
        ctx.msm_dev(pts, sc, slot=slot)

    peak = ctx.imad_peak()
    out = open(args.out, "a") if args.out else None
    for logn in args.logn:
        n = 1 << logn
        sets = [(ctx.fixed_base(seed=0x5EEE + 16 * k, n=n), ctx.synth_scalars(0x5EED + 16 * k, n)) for k in range(3)]
        for pre, c in [(p_, c_) for p_ in args.precompute for c_ in args.windows]:
            if pre >= 0:
                for pts, _ in sets:
                    pts.precompute(pre)
            ctx.set_option(_lib.OPT_PRE_MIN_TERMS, 256 if pre >= 0 else 1 << 30)
            for sort in args.sort:
                for radix, cap, at, sb, ps, sl, bs in [(r, cp, a, b, q, z, y) for r in args.radix for cp in args.cap
                                                       for a in args.async_tail for b in args.sort_blocks
                                                       for q in (args.pre_sets if pre >= 0 else [0]) for z in args.seg_len
                                                       for y in args.block_sort]:
                  for sm in args.seg_mode:
                    ctx.set_option(_lib.OPT_SEG_MODE, sm)
                    ctx.set_option(_lib.OPT_BLOCK_SORT, bs)
                    ctx.set_option(_lib.OPT_PRE_SETS, ps)
                    ctx.set_option(_lib.OPT_SEG_LEN, sl)
                    if sb >= 0:
                        ctx.set_option(_lib.OPT_SORT_BLOCKS, sb)
                    if cap:
                        ctx.set_option(_lib.OPT_CAP_FACTOR, cap)
                    ctx.set_option(_lib.OPT_ASYNC_TAIL, at)
                    ctx.set_option(_lib.OPT_WINDOW_BITS, c)
                    ctx.set_option(_lib.OPT_SORT_BUCKETS, sort)
                    ctx.set_option(_lib.OPT_REDUCE_RADIX, radix)
                    for w in range(3):
                        issue(*sets[w % 3], slot=w)
                    ctx.sync()
                    ctx.phase_times()
                    ctx.timer_start()
                    h0 = time.perf_counter()
                    for s in range(args.steps):
                        if args.shard_self and s and s % 32 == 0:
                            ctx.sync()  # a mailbox entry may not be pushed again before the owner has gathered it
                        issue(*sets[s % 3], slot=s % 32)
                    host_ms = 1e3 * (time.perf_counter() - h0) / args.steps  # host time to ISSUE one MSM (no sync)
                    ms = ctx.timer_stop() / args.steps
                    ph, calls = ctx.phase_times()
                    rec = {"log2n": logn, "shard_self": args.shard_self, "seg_mode": sm, "block_sort": bs, "acc_carveout": args.acc_carveout, "precompute": pre, "pre_sets": ps, "seg_len": sl, "window": c, "sort": sort, "radix": radix, "cap": cap, "async_tail": at, "sort_blocks": sb, "ms": ms, "host_issue_ms": round(host_ms, 4), "Mpts_s": n / ms / 1e3,
                           "imad_peak_tlps": peak, "phase_ms": {k: round(v / calls, 5) for k, v in ph.items()}}
                    print(json.dumps(rec), flush=True)
                    if out:
                        out.write(json.dumps(rec) + "\n")
                        out.flush()
        for p, s in sets:
            p.free()
            s.free()
    ctx.close()


if __name__ == "__main__":
    main()
