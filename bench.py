#!/usr/bin/env python
"""bench.py -- Ed25519 MSM throughput (points/s) on 1..8 B200s, the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log2n 20] [--impl reference] [--sweep]

N = 1 runs in-process; N > 1 is launched by the driver under torchrun (one rank per GPU; torch.distributed is used
only for the barrier / max-over-ranks / 128-byte partial exchange -- the data path is libvmsm.so).  A "step" is one
MSM over one batch of synthetic seeded scalars with the bases resident in HBM.  Multi-GPU: index-range split of ONE
(N * 2^log2n)-term MSM, every rank computes the partial sum of its slice, rank 0 adds the N partials ("weak" scaling:
per-GPU work fixed).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ed25519_msm_points_per_s"
UNIT = "points/s"
SEED_SCALARS, SEED_BASES = 0x5EED, 0x5EEE
NSETS = 3  # distinct (bases, scalars) sets rotated between steps so no step finds its inputs in L2
E2E_DEPTH = 4  # MSMs in flight in the end-to-end loop

# SURVEY.md 8(d)/App. E work model: limb products per point at the LP-minimising window c*(n)
M, S = 72, 44
MADD, ADD, DBL = 7 * M, 9 * M, 4 * M + 4 * S


def lp_msm(n, c):
    W = -(-254 // c)
    return n * W * MADD + W * 2 * (1 << (c - 1)) * ADD + (W - 1) * c * DBL + W * ADD


def lp_per_point(n):
    return min(lp_msm(n, c) for c in range(2, 25)) / n


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=subprocess.PIPE, text=True,
                                         stderr=subprocess.DEVNULL)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.perf_counter()

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        mhz = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": float(rows[0][2]) if rows and rows[0][2].replace(".", "").isdigit() else None,
                "samples": len(rows), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    seed_s, seed_r, start, cnt = args
    from oracle import ed25519 as E
    from oracle import prng

    pts = [E.scalar_mul(E.B, 1 + ((start + i) % 64)) for i in range(cnt)]  # cheap bases (small multiples of B)
    sc = [prng.scalar(seed_s, start + i) for i in range(cnt)]
    t0 = time.perf_counter()
    E.msm_naive(sc, pts)  # the reference's algorithm: per-term double-and-add + tree product (pivot.py:143)
    return time.perf_counter() - t0


def cpu_reference_rate(points_per_core, cores):
    """Points/s of the pure-Python restatement of the reference's MSM on `cores` processes (a sharded commitment)."""
    import multiprocessing as mp

    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(SEED_SCALARS, SEED_BASES, k * points_per_core, points_per_core) for k in range(cores)])
        dt = time.perf_counter() - t0
    return cores * points_per_core / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # exactly K steps, each a bounded sample of the 2^log2n-term MSM (the cost is linear in the number of terms):
    # ~1.4 s per step at 1024 terms per core, shrunk for large K so that the whole run stays within ~2 minutes
    steps = max(1, args.steps)
    per_core = max(32, min(args.cpu_points, args.cpu_points * 90 // steps))
    for _ in range(min(args.warmup, 1)):
        cpu_reference_rate(max(8, per_core // 8), cores)
    t_tot, pts_tot = 0.0, 0
    for _ in range(steps):
        rate, dt = cpu_reference_rate(per_core, cores)
        t_tot += dt
        pts_tot += per_core * cores
    value = pts_tot / t_tot
    n = 1 << args.log2n
    sample = (f"{steps} steps x {cores} processes x {per_core} terms of the 2^{args.log2n}-term MSM; pure-Python restatement "
              f"of pivot.vector_commitment (MPyC/gmpy2 are not installable in this image), cost is linear in n")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t_tot / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "python-int", "data": "synthetic",
            "config": {"workload": f"ed25519_msm_2^{args.log2n}", "n_per_gpu": n, "bounded_sample": True},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def dlog_sum(seed_s, seed_r, n, start=0):
    from verifiable_mpc_b200 import synth

    L = synth.ED_L
    s = synth.scalars_ed25519(seed_s, n, start)
    r = synth.scalars_ed25519(seed_r, n, start)
    tot = 0
    for i in range(n):
        tot += int.from_bytes(s[i].tobytes(), "little") * int.from_bytes(r[i].tobytes(), "little")
    return tot % L


def run_gpu(args, rank, world, dist):
    import ctypes

    import numpy as np

    from verifiable_mpc_b200 import Context, _lib, shard, synth

    local = int(os.environ.get("LOCAL_RANK", rank))
    ctx = Context(local)
    n = 1 << args.log2n
    if args.window:
        ctx.set_option(_lib.OPT_WINDOW_BITS, args.window)
    if args.sort_blocks >= 0:
        ctx.set_option(_lib.OPT_SORT_BLOCKS, args.sort_blocks)
    ctx.set_option(_lib.OPT_PHASE_TIMING, 1)

    # inputs: NSETS distinct (bases, scalars) sets per rank, all resident in HBM before the timed region.
    # set k of rank r covers global indices [r*n, (r+1)*n) of stream (seed + 16*k).
    base_off = rank * n
    bases, scal, host_scal = [], [], []
    for k in range(NSETS):
        rs = synth.scalars_ed25519(SEED_BASES + 16 * k, n, start=base_off)  # known dlogs r_i
        bases.append(ctx.fixed_base(scalars=rs))                              # g_i = r_i * B on the device
        hs = synth.scalars_ed25519(SEED_SCALARS + 16 * k, n, start=base_off)
        scal.append(ctx.upload_scalars(hs))
        pin = ctx.pinned(n * 32)
        pin.array[:] = hs.reshape(-1)
        host_scal.append(pin)
        del rs, hs

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    # multi-GPU: index-range split; partials go GPU -> GPU through the owner's mailbox (peer stores over NVLink,
    # mapped with CUDA IPC); torch.distributed only carries the 64-byte IPC handle, the barrier and the timing max
    seq_counter = [0]
    if dist is not None:
        shard.setup_mailbox(ctx, dist, rank, world)

    def issue(kind, k, slot):
        """One step on this rank: its shard of the (world * n)-term MSM (or the whole MSM when world == 1)."""
        if dist is not None:
            seq_counter[0] += 1
            ctx.set_option(_lib.OPT_SHARD_SEQ, seq_counter[0])
        if kind == "dev":
            ctx.msm_dev(bases[k], scal[k], slot=slot)
        else:
            ctx.msm_async(bases[k], host_scal[k].ptr, 0, n, slot=slot)

    def recycle_slots(s):
        """Result slots / mailbox entries are reused every 48 steps.  A single GPU needs nothing (an unread result is
        simply overwritten); with a mailbox a rank may not push step s into an entry the owner has not yet gathered
        for step s - 48, so once per 48 steps the owner drains its pipeline and every rank waits for it."""
        if dist is not None and s and s % 48 == 0:
            ctx.sync()
            dist.barrier()

    def combine(slot):
        """Owner: the sum over all ranks (gathered on the device); other ranks: their own partial."""
        return ctx.result(slot)

    # warm-up
    for w in range(args.warmup):
        issue("dev", w % NSETS, 0)
        combine(0)
    barrier()
    ctx.phase_times()
    l0 = ctx.launch_count()

    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    tm0 = time.perf_counter()
    ctx.timer_start()
    for s in range(args.steps):
        recycle_slots(s)
        issue("dev", s % NSETS, s % 48)
    ms = ctx.timer_stop()
    barrier()
    tm1 = time.perf_counter()
    clocks = sampler.stop(tm0, tm1) if sampler else None
    launches = ctx.launch_count() - l0
    phases, calls = ctx.phase_times()
    if dist is not None:
        ms = shard.max_over_ranks(dist, ms)
        launches = shard.sum_over_ranks(dist, launches)

    # per-kernel figures for the roofline: the same workload once more with the three streams joined (no overlap of
    # the next MSM's counting sort / the previous MSM's tail with the accumulate kernel), CUDA events around each phase
    ctx.set_option(_lib.OPT_ASYNC_SORT, 0)
    ctx.set_option(_lib.OPT_ASYNC_TAIL, 0)
    for w in range(2):
        issue("dev", w % NSETS, 0)
    ctx.sync()
    ctx.phase_times()
    serial_steps = max(1, min(args.steps, 10))
    ctx.timer_start()
    for s in range(serial_steps):
        issue("dev", s % NSETS, s % 48)
    serial_ms = ctx.timer_stop() / serial_steps
    serial_phases, serial_calls = ctx.phase_times()
    ctx.set_option(_lib.OPT_ASYNC_SORT, 1)
    ctx.set_option(_lib.OPT_ASYNC_TAIL, 1)
    barrier()

    # correctness of what was just timed: set (steps-1) % NSETS, against the known-dlog identity
    last = (args.steps - 1) % NSETS
    got = combine((args.steps - 1) % 48)
    checked = None
    if args.check:
        e = dlog_sum(SEED_SCALARS + 16 * last, SEED_BASES + 16 * last, n, start=base_off)
        if dist is not None:
            import torch

            parts = [None] * world
            dist.all_gather_object(parts, e)
            e = sum(parts) % synth.ED_L
        if rank == 0:
            # e*B through a different device path (fixed-base comb) than the Pippenger pipeline being checked
            exp = ctx.fixed_base(scalars=[e]).tolist()[0]
            checked = bool(got == exp)

    # end-to-end through the public host API: every step copies that step's scalars from pinned host memory to the
    # device (H2D on the library's copy stream) and reads the resulting group element back on the host (the final
    # kernel writes it into mapped pinned memory).  Steps are pipelined E2E_DEPTH deep: the host fetches result s-3 after
    # issuing step s, so the copy and the counting sort of step s overlap the accumulate kernels of earlier steps.
    def e2e_loop(steps):
        last = None
        lag = E2E_DEPTH - 1
        for s in range(steps):
            recycle_slots(s)
            issue("async", s % NSETS, s % 48)
            if s >= lag:
                last = combine((s - lag) % 48)
        for s in range(max(steps - lag, 0), steps):
            last = combine(s % 48)
        return last

    e2e_loop(min(args.warmup, 3))
    barrier()
    e0 = time.perf_counter()
    e2e_last = e2e_loop(args.steps)
    barrier()
    e2e_s = time.perf_counter() - e0
    if dist is not None:
        e2e_s = shard.max_over_ranks(dist, e2e_s)
    e2e_ok = None
    if rank == 0 and checked is not None and (args.steps - 1) % NSETS == last:
        e2e_ok = bool(e2e_last == got)

    if rank != 0:
        return
    total_pts = world * n * args.steps
    value = total_pts / (ms * 1e-3)
    peak_tlps = ctx.imad_peak()
    acc_ms = serial_phases["accumulate"] / max(serial_calls, 1)
    # the accumulate kernel performs exactly one 7M mixed addition per non-zero digit: n * W of them per launch
    c_auto = _choose_window(n) if not args.window else args.window
    W_c = -(-254 // c_auto)
    acc_lp = n * W_c * MADD
    achieved = acc_lp / (acc_ms * 1e-3) / 1e12 if acc_ms > 0 else None
    msm_frac = value / world * lp_per_point(n) / (peak_tlps * 1e12)
    # HBM-side view of the counting sort (SURVEY 8d asks for GB/s next to the integer roofline): per scalar it reads the
    # 32-byte scalar twice (histogram, scatter) and writes W 4-byte index entries, plus three passes over the counters
    sp = {k: v / max(serial_calls, 1) for k, v in serial_phases.items()}
    sort_ms = sp["digits"] + sp["scan"] + sp["scatter"] + sp["order"]
    sort_bytes = n * (32 + 32 + 4 * W_c) + 3 * 4 * W_c * (1 << (c_auto - 1))
    hbm_peak, hbm_src = _hbm_peak_gbs()
    sort_gbs = sort_bytes / (sort_ms * 1e-3) / 1e9 if sort_ms > 0 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"ed25519_msm_2^{args.log2n}", "n_per_gpu": n, "n_total": world * n,
                   "window_bits": c_auto, "windows": W_c, "bases": "g_i = r_i*B, device generated, niels form resident",
                   "l2": f"inputs rotate over {NSETS} distinct (bases, scalars) sets ({NSETS * n * 128 >> 20} MiB) > 126 MB L2",
                   "multi_gpu": ("index-range split of one N*n-term MSM; partials pushed into rank 0's HBM mailbox over "
                                 "NVLink peer stores (CUDA IPC), summed by a gather kernel; no NCCL") if world > 1 else "single GPU",
                   "result_checked_vs_known_dlog": checked, "e2e_result_matches": e2e_ok},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": total_pts / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * 32 * world,
                "d2h_bytes_per_step": 64 * world, "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "Context.msm_async(points, pinned_scalars, slot) + Context.result(slot), pipelined %d deep" % E2E_DEPTH},
        "roofline": {"bound": "imad", "kernel": "vmsm_kernel<KAccumulate>", "achieved": achieved, "peak": peak_tlps,
                     "unit": "T limb-products/s", "frac": (achieved / peak_tlps) if achieved else None,
                     "traffic": _ncu_traffic_bytes() if args.log2n == 20 else None, "traffic_unit": "DRAM bytes per launch (ncu --set full, profiles/)",
                     "algorithmic_gather_bytes_per_launch": n * W_c * 96,
                     "peak_source": "measured live: vmsm_microbench_imad (independent carry-chained IMAD.WIDE.U32 multiply-accumulate chains, all SMs; every IMAD.WIDE form is half-rate on B200)",
                     "kernel_ms": acc_ms, "algorithmic_lp_per_launch": acc_lp,
                     "whole_msm_frac": msm_frac, "whole_msm_lp_per_point": lp_per_point(n),
                     "measured_in": f"{serial_steps} extra steps of the same workload with the library's streams joined "
                                    "(serial), CUDA events around every phase; the headline loop overlaps phases of "
                                    "consecutive MSMs, so its per-phase times are not additive",
                     "serial_ms_per_step": serial_ms, "kernel_share_of_serial_step": acc_ms / serial_ms,
                     "phase_ms": {k: v / max(serial_calls, 1) for k, v in serial_phases.items()},
                     "overlapped_phase_ms": {k: v / max(calls, 1) for k, v in phases.items()},
                     "hbm_phases": {"counting_sort": {
                         "algorithmic_bytes": sort_bytes, "ms": sort_ms, "GB_s": sort_gbs, "hbm_peak_GB_s": hbm_peak,
                         "frac": (sort_gbs / hbm_peak) if sort_gbs else None, "peak_source": hbm_src,
                         "note": "bound by L2 atomics (16 fetch-and-adds per scalar), not by HBM bandwidth; in the "
                                 "headline loop it runs as a thin slice under the accumulate kernel"}}},
    }
    if args.cpu_baseline:
        cores = os.cpu_count() or 1
        rate, dt = cpu_reference_rate(args.cpu_points, cores)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{cores} processes x {args.cpu_points} terms ({dt:.1f} s) of the same MSM with the "
                                          "pure-Python restatement of pivot.vector_commitment (oracle/ed25519.py)"}
    print(json.dumps(line), flush=True)


def _hbm_peak_gbs():
    """Measured HBM copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the recipe's figure."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "of fallback (B200_PROFILING.md: 6.65 TB/s; MEASURED_PEAKS.json absent)"


def _ncu_traffic_bytes():
    """DRAM bytes per KAccumulate launch from the committed `ncu --set full` capture (profiles/), or None."""
    import csv
    import glob

    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*", "ncu_full_KAccumulate.csv")), reverse=True):
        try:
            rows = list(csv.reader(open(path)))
            hdr = rows[0]
            unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot = []
            for r in rows[1:]:
                if not r:
                    continue
                b = 0.0
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    col = [i for i, h in enumerate(hdr) if h.startswith(key)][0]
                    b += float(r[col]) * unit[hdr[col].split("[")[1].rstrip("]")]
                tot.append(b)
            if tot:
                return sum(tot) / len(tot)
        except Exception:
            continue
    return None


def _choose_window(n):
    """Mirror of csrc/pipeline.cuh:choose_window for 253-bit scalars (reporting only)."""
    best, best_c = None, 8
    for c in range(3, 18):
        r = 252 % c
        if r == 0 or c - r > 4 or c < 11 or (n >= 49152 and c < 15):
            continue
        W = -(-254 // c)
        cost = n * W * 504.0 + W * (1 << (c - 1)) * 2.0 * 648.0 * 1.3 + (W - 1) * c * 464.0
        if best is None or cost < best:
            best, best_c = cost, c
    return best_c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--sort-blocks", type=int, default=-1, help="experiment: VMSM_OPT_SORT_BLOCKS (-1 = library default)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-check", dest="check", action="store_false")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-points", type=int, default=1024, help="terms per host core in the CPU baseline sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("gloo", rank=rank, world_size=world)
        dist = dist_mod
    try:
        run_gpu(args, rank, world, dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
