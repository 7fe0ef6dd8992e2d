#!/usr/bin/env python
"""bench.py -- Ed25519 MSM throughput (points/s) on 1..8 B200s, the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log2n 20] [--impl reference] [--sweep]

N = 1 runs in-process; N > 1 is launched by the driver under torchrun (one rank per GPU; torch.distributed is used
only for the barrier / max-over-ranks -- the data path is libvmsm.so, partial results travel GPU -> GPU through a peer
mailbox).  A "step" is one batch of MSMS_PER_STEP MSMs of 2^log2n terms each over synthetic seeded scalars with the
bases resident in HBM (a batch, so that the driver's K = 20 steps time ~0.7 s instead of 30 ms).  Multi-GPU: index-range
split of (N * 2^log2n)-term MSMs, every rank computes the partial sum of its slice, rank 0 adds the N partials
("weak" scaling: per-GPU work fixed).  Prints ONE JSON line on rank 0.  Beside the headline the same line carries the
rest of BASELINE.json's metric: `sweep` (2^16, 2^18 points/s and roofline fraction), `fixed_generators` (the same sizes
over precomputed generator tables), `ac20` (compressed-pivot prove / verify latency at N = 2^16 with a CPU figure),
`bn256` (the BN256 G1 / G2 prover-key MSMs and the compute_proof twin for a 2^14-constraint QAP) and `strong` (one fixed
2^20-term MSM split over the N GPUs).
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
MSMS_PER_STEP = 24  # one step = this many back-to-back MSMs (each over the next of the NSETS input sets)
STRONG_LOG2N = 20  # the fixed-size MSM of the strong-scaling block

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


def time_msms(ctx, pairs, count, warm=3, pre_issue=None):
    """ms per MSM of `count` device-resident MSMs issued back to back over the rotating (points, scalars) pairs."""
    for w in range(warm):
        ctx.msm_dev(*pairs[w % len(pairs)], slot=0)
    ctx.sync()
    ctx.timer_start()
    for s in range(count):
        if pre_issue:
            pre_issue(s)
        ctx.msm_dev(*pairs[s % len(pairs)], slot=s % 48)
    return ctx.timer_stop() / count


def synth_pairs(ctx, n, nsets, seed_off=0):
    """`nsets` device-generated (bases g_i = r_i*B, scalars) pairs of n terms; seeds as oracle/prng.py specifies."""
    return [(ctx.fixed_base(seed=SEED_BASES + seed_off + 16 * k, n=n), ctx.synth_scalars(SEED_SCALARS + seed_off + 16 * k, n))
            for k in range(nsets)]


def sets_beyond_l2(n, bytes_per_term=128):
    """Enough distinct input sets that a rotation never finds its inputs in the 126 MB L2."""
    return max(NSETS, -(-160 * (1 << 20) // (n * bytes_per_term)))


def sweep_block(ctx, peak_tlps, log2ns=(16, 18)):
    """Points/s at the other sizes of the metric's range, plain path and over precomputed generator tables; the result
    of every configuration is checked against the known-dlog identity (plain) / the plain result (tables)."""
    from verifiable_mpc_b200 import _lib

    out, fixed = [], []
    for logn in log2ns:
        n = 1 << logn
        nsets = sets_beyond_l2(n)
        pairs = synth_pairs(ctx, n, nsets, seed_off=0x1000 * logn)
        count = max(48, min(512, (1 << 27) // n))
        ms = time_msms(ctx, pairs, count)
        last = (count - 1) % len(pairs)
        got = ctx.result((count - 1) % 48)
        e = dlog_sum(SEED_SCALARS + 0x1000 * logn + 16 * last, SEED_BASES + 0x1000 * logn + 16 * last, n)
        ok = bool(got == ctx.fixed_base(scalars=[e]).tolist()[0])
        pps = n / (ms * 1e-3)
        out.append({"log2n": logn, "points_per_s": pps, "ms_per_msm": ms, "msms_timed": count, "input_sets": nsets,
                    "whole_msm_lp_per_point": lp_per_point(n), "whole_msm_frac": pps * lp_per_point(n) / (peak_tlps * 1e12),
                    "result_checked_vs_known_dlog": ok})
        # the same MSMs over tables of 2^(16 w) * g_i (fixed generators): table footprint 16 x 96 B per point
        psets = pairs[:max(NSETS, -(-160 * (1 << 20) // (n * 16 * 96)))]
        t0 = time.perf_counter()
        for pts, _ in psets:
            pts.precompute()
        ctx.sync()
        t_pre = (time.perf_counter() - t0) / len(psets)
        ms_p = time_msms(ctx, psets, count)
        lastp = (count - 1) % len(psets)
        gotp = ctx.result((count - 1) % 48)
        ctx.set_option(_lib.OPT_PRE_MIN_TERMS, 1 << 30)  # the plain path on the same inputs, for the comparison
        ctx.msm_dev(*psets[lastp], slot=50)
        okp = bool(gotp == ctx.result(50))
        ctx.set_option(_lib.OPT_PRE_MIN_TERMS, 256)
        ppsp = n / (ms_p * 1e-3)
        fixed.append({"log2n": logn, "points_per_s": ppsp, "ms_per_msm": ms_p, "table_build_ms": 1e3 * t_pre,
                      "table_bytes_per_point": 16 * 96, "input_sets": len(psets),
                      "lp_per_point_executed": (16 * n * MADD) / n,
                      "frac_of_peak_on_executed_work": ppsp * 16 * MADD / (peak_tlps * 1e12),
                      "frac_on_survey_work_model": ppsp * lp_per_point(n) / (peak_tlps * 1e12),
                      "equals_plain_path_result": okp})
        for pts, sc in pairs:
            pts.free()
            sc.free()
    return out, fixed


def ac20_block(ctx):
    """AC20 compressed-pivot prove / verify latency at N = 2^16 through the reference-facing Python API
    (protocol_5_prover / protocol_5_verifier twins; host lists of field elements in, proof dict out), wall clock."""
    from oracle import ed25519 as E
    from tools import bench_ac20
    from verifiable_mpc_b200 import fingroups
    from verifiable_mpc_b200.finfields import GF

    group = fingroups.EllipticCurve("Ed25519", "projective")
    group.is_additive, group.is_multiplicative = False, True
    old = fingroups.Ed25519Point.context
    fingroups.Ed25519Point.context = ctx
    try:
        gf = GF(group.order)
        bench_ac20.measure(group, gf, 12, 1)  # warm-up (pools, pinned buffers, lazy tables)
        rec = bench_ac20.measure(group, gf, 16, 3)
        rec_pre = bench_ac20.measure(group, gf, 16, 3, precompute=True)
    finally:
        fingroups.Ed25519Point.context = old
    # CPU figure beside it: the reference's prover is (MSM terms + fold elements) independent double-and-add scalar
    # multiplications (pivot.py:143, compressed_pivot.py:41-42,64,110) -- operation count x measured cost per operation
    import random
    rnd = random.Random(3)
    pt = E.scalar_mul(E.B, rnd.randrange(E.L))
    t0 = time.perf_counter()
    reps = 64
    for _ in range(reps):
        E.scalar_mul(pt, rnd.randrange(E.L))
    t_mul = (time.perf_counter() - t0) / reps
    N, rounds = 1 << 16, 15
    msm_terms = N + sum(2 * ((N >> (i + 1)) + 1) for i in range(rounds))
    fold_elems = N - 2
    ops = msm_terms + fold_elems + 3 * rounds + 3
    cores = os.cpu_count() or 1
    recorded = None
    try:
        with open(os.path.join(ROOT, "tests", "golden", "ac20_big_16.json")) as f:
            recorded = json.load(f)["reference_cpu_seconds"]
    except Exception:
        pass
    return {"N": N, "api": "compressed_pivot.protocol_5_prover / protocol_5_verifier (reference signatures)",
            "prove_ms": 1e3 * rec["prove_s"], "verify_ms": 1e3 * rec["verify_s"], "verified": rec["verified"],
            "prove_breakdown_ms": {k: 1e3 * v for k, v in rec["prove_breakdown_s"].items()},
            "prove_host_other_ms": 1e3 * rec["prove_host_other_s"],
            "verify_breakdown_ms": {k: 1e3 * v for k, v in rec["verify_breakdown_s"].items()},
            "transcript": rec["transcript"],
            "with_generator_table": {"prove_ms": 1e3 * rec_pre["prove_s"], "verify_ms": 1e3 * rec_pre["verify_s"],
                                     "z_commitment_ms": 1e3 * rec_pre["z_commitment_s"],
                                     "table_build_ms": 1e3 * (rec_pre["precompute_s"] or 0.0)},
            "z_commitment_ms": 1e3 * rec["z_commitment_s"],
            "cpu_baseline": {"kind": "port", "prove_s_1_core": ops * t_mul, "prove_s_all_cores_ideal": ops * t_mul / cores,
                             "cores": cores, "scalar_mult_ms_1_core": 1e3 * t_mul, "scalar_mults_per_proof": ops,
                             "recorded_full_run_of_the_unmodified_reference": recorded,
                             "sample": f"{reps} double-and-add scalar multiplications of oracle/ed25519.py on 1 core x the "
                                       "reference prover's operation count (N-term announcement, 2 x (half+1) terms and "
                                       "half fold elements per round); pure-Python ints, no gmpy2 (not installable "
                                       "here); the recorded run is the whole unmodified protocol_5_prover on the MPyC "
                                       "look-alike in the build container (tests/golden/make_ac20_big_golden.py 16)"}}


# BN256 work model: Montgomery multiplication 136 limb products; Jacobian mixed addition 7M+4S, addition
# 11M+5S, doubling 2M+5S; G2 (Karatsuba Fp2) x 3
BN_M = 136
BN_MADD, BN_ADD, BN_DBL = 11 * BN_M, 16 * BN_M, 7 * BN_M


def lp_msm_bn(n, c, g2=False):
    W = -(-257 // c)
    lp = n * W * BN_MADD + W * 2 * (1 << (c - 1)) * BN_ADD + (W - 1) * c * BN_DBL + W * BN_ADD
    return 3 * lp if g2 else lp


def bn256_block(ctx, peak_tlps, logn=14):
    """BASELINE.json config 4: the prover-key multi-exponentiations of a 2^14-constraint QAP.  Device-timed G1 / G2 MSMs
    over the key tables (and the plain path beside them), and the wall time of the compute_proof twin (reference
    signature: witness and quotient coefficients as host lists in, proof dict out) on a synthetic prepared key."""
    from tools import bench_bn256
    from verifiable_mpc_b200 import _lib, fingroups
    from verifiable_mpc_b200.trinocchio import pynocchio as twin

    n = 1 << logn
    out = {"log2n": logn, "msm": []}
    for curve, name in ((1, "G1"), (2, "G2")):
        pairs = [(ctx.fixed_base(seed=SEED_BASES + 0x7000 + 16 * k, n=n, curve=curve),
                  ctx.synth_scalars(SEED_SCALARS + 0x7000 + 16 * k, n, curve=curve)) for k in range(NSETS)]
        count = 96
        ms_plain = time_msms(ctx, pairs, count)
        plain_last = ctx.result((count - 1) % 48, curve=curve)
        for pts, _ in pairs:
            pts.precompute()
        ctx.sync()
        ms = time_msms(ctx, pairs, count)
        ok = bool(ctx.result((count - 1) % 48, curve=curve) == plain_last)
        lp_min = min(lp_msm_bn(n, c, curve == 2) for c in range(4, 17))
        out["msm"].append({"group": name, "ms_per_msm_key_tables": ms, "ms_per_msm_plain": ms_plain, "msms_timed": count,
                           "points_per_s": n / (ms * 1e-3), "equals_plain_path_result": ok,
                           "frac_on_survey_work_model": lp_min / (ms * 1e-3) / (peak_tlps * 1e12),
                           "frac_on_work_of_the_executed_window": lp_msm_bn(n, 13, curve == 2) / (ms * 1e-3) / (peak_tlps * 1e12),
                           "lp_survey_model": lp_min, "lp_executed_window_c13": lp_msm_bn(n, 13, curve == 2)})
        for pts, sc in pairs:
            pts.free()
            sc.free()
    old = fingroups.BN256Point.context
    try:
        Q, c, H, D, prepared = bench_bn256.synthetic_proof_case(ctx, n)
        prepared.precompute()
        ctx.sync()
        twin.compute_proof(Q, c, H(), prepared, D)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            twin.compute_proof(Q, c, H(), prepared, D)
            best = min(best, time.perf_counter() - t0)
        out["compute_proof_ms"] = 1e3 * best
        out["compute_proof"] = ("trinocchio.pynocchio.compute_proof twin, |mid| = len(h) = 2^%d, prepared key with tables, witness "
                                "and h as host lists of ints (packed and uploaded inside the timed call), 8 MSMs (6 G1 + 1 G2 "
                                "over the mid wires, 1 G1 over h) with the zero-knowledge delta terms; best of 5" % logn)
        for dev in prepared.bases.values():
            dev.free()
    finally:
        fingroups.BN256Point.context = old
    return out


def run_gpu(args, rank, world, dist):
    from tools import dist_util
    from verifiable_mpc_b200 import Context, _lib, shard, synth

    local = int(os.environ.get("LOCAL_RANK", rank))
    ctx = Context(local)
    n = 1 << args.log2n
    R = args.msms_per_step
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

    def issue(kind, k, slot, pts=None, sc=None):
        """One MSM on this rank: its shard of the (world * n)-term MSM (or the whole MSM when world == 1)."""
        if dist is not None:
            seq_counter[0] += 1
            ctx.set_option(_lib.OPT_SHARD_SEQ, seq_counter[0])
        if kind == "dev":
            ctx.msm_dev(pts[k] if pts else bases[k], sc[k] if sc else scal[k], slot=slot)
        else:
            ctx.msm_async(bases[k], host_scal[k].ptr, 0, n, slot=slot)

    def recycle_slots(i):
        """Result slots / mailbox entries are reused every 48 MSMs.  A single GPU needs nothing (an unread result is
        simply overwritten); with a mailbox a rank may not push MSM i into an entry the owner has not yet gathered
        for MSM i - 48.  Rolling flow control, no pipeline drain: every 24 MSMs the owner waits for its results of MSMs
        i-32 .. i-25 (issued long ago; one per tail stream, so everything up to MSM i-25 has been gathered), then all
        ranks meet at a host barrier while their GPUs keep working on the >= 24 MSMs still queued; the entries of
        MSMs i-48 .. i-25 are then free for MSMs i .. i+23."""
        if dist is not None and i >= 48 and i % 24 == 0:
            if rank == 0:
                for j in range(i - 32, i - 24):
                    ctx.result(j % 48)
            dist.barrier()

    def combine(slot):
        """Owner: the sum over all ranks (gathered on the device); other ranks: their own partial."""
        return ctx.result(slot)

    # warm-up: W steps of the same shape as the timed ones
    for w in range(args.warmup * R):
        recycle_slots(w)
        issue("dev", w % NSETS, w % 48)
    combine((args.warmup * R - 1) % 48)
    barrier()
    ctx.phase_times()
    l0 = ctx.launch_count()

    total_msms = args.steps * R
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    tm0 = time.perf_counter()
    ctx.timer_start()
    for i in range(total_msms):
        recycle_slots(i)
        issue("dev", i % NSETS, i % 48)
    ms = ctx.timer_stop()
    barrier()
    tm1 = time.perf_counter()
    clocks = sampler.stop(tm0, tm1) if sampler else None
    launches = ctx.launch_count() - l0
    phases, calls = ctx.phase_times()
    if dist is not None:
        ms = dist_util.max_over_ranks(dist, ms)
        launches = dist_util.sum_over_ranks(dist, launches)

    # per-kernel figures for the roofline: the same workload once more with the three streams joined (no overlap of
    # the next MSM's counting sort / the previous MSM's tail with the accumulate kernel), CUDA events around each phase
    ctx.set_option(_lib.OPT_ASYNC_SORT, 0)
    ctx.set_option(_lib.OPT_ASYNC_TAIL, 0)
    ctx.set_option(_lib.OPT_DUAL_HEAD, 0)
    for w in range(2):
        issue("dev", w % NSETS, 0)
    ctx.sync()
    ctx.phase_times()
    serial_msms = 12
    ctx.timer_start()
    for i in range(serial_msms):
        issue("dev", i % NSETS, i % 48)
    serial_ms = ctx.timer_stop() / serial_msms
    serial_phases, serial_calls = ctx.phase_times()
    ctx.set_option(_lib.OPT_ASYNC_SORT, 1)
    ctx.set_option(_lib.OPT_ASYNC_TAIL, 1)
    ctx.set_option(_lib.OPT_DUAL_HEAD, 1)
    barrier()

    # correctness of what was just timed: the last MSM's input set, against the known-dlog identity
    last = (total_msms - 1) % NSETS
    got = combine((total_msms - 1) % 48)
    checked = None
    if args.check:
        e = dlog_sum(SEED_SCALARS + 16 * last, SEED_BASES + 16 * last, n, start=base_off)
        if dist is not None:
            parts = [None] * world
            dist.all_gather_object(parts, e)
            e = sum(parts) % synth.ED_L
        if rank == 0:
            # e*B through a different device path (fixed-base comb) than the Pippenger pipeline being checked
            exp = ctx.fixed_base(scalars=[e]).tolist()[0]
            checked = bool(got == exp)

    # end-to-end through the public host API: every MSM copies its scalars from pinned host memory to the device (H2D
    # on the library's copy stream) and its resulting group element is read back on the host (the final kernel writes
    # it into mapped pinned memory).  MSMs are pipelined E2E_DEPTH deep: the host fetches result i-3 after issuing MSM
    # i, so the copy and the counting sort of MSM i overlap the accumulate kernels of earlier ones.
    def e2e_loop(count):
        last_pt = None
        lag = E2E_DEPTH - 1
        for i in range(count):
            recycle_slots(i)
            issue("async", i % NSETS, i % 48)
            if i >= lag:
                last_pt = combine((i - lag) % 48)
        for i in range(max(count - lag, 0), count):
            last_pt = combine(i % 48)
        return last_pt

    e2e_loop(min(args.warmup, 3) * R)
    barrier()
    e0 = time.perf_counter()
    e2e_last = e2e_loop(total_msms)
    barrier()
    e2e_s = time.perf_counter() - e0
    if dist is not None:
        e2e_s = dist_util.max_over_ranks(dist, e2e_s)
    e2e_ok = None
    if rank == 0 and checked is not None:
        e2e_ok = bool(e2e_last == got)

    peak_tlps = ctx.imad_peak() if rank == 0 else None

    # ---- strong scaling: ONE fixed 2^20-term MSM split by index range over the `world` GPUs, plain and over
    # precomputed generator tables (every rank holds the table of its own slice)
    strong = strong_block(ctx, dist, rank, world, issue_seq=seq_counter, bases_full=bases if args.log2n == STRONG_LOG2N else None,
                          scal_full=scal if args.log2n == STRONG_LOG2N else None, recycle=recycle_slots, barrier=barrier)

    sweep = fixed = ac20 = bn256 = None
    if world == 1 and args.extras:
        sweep, fixed = sweep_block(ctx, peak_tlps)
        ac20 = ac20_block(ctx)
        bn256 = bn256_block(ctx, peak_tlps)

    if rank != 0:
        return
    total_pts = world * n * total_msms
    value = total_pts / (ms * 1e-3)
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
                   "msms_per_step": R, "ms_per_msm": ms / total_msms,
                   "step": f"a batch of {R} back-to-back MSMs of 2^{args.log2n} terms per GPU (timed region = steps x {R} MSMs)",
                   "window_bits": c_auto, "windows": W_c, "bases": "g_i = r_i*B, device generated, niels form resident, no precomputed tables",
                   "l2": f"inputs rotate over {NSETS} distinct (bases, scalars) sets ({NSETS * n * 128 >> 20} MiB) > 126 MB L2",
                   "multi_gpu": ("index-range split of N*n-term MSMs; partials pushed into rank 0's HBM mailbox over "
                                 "NVLink peer stores (CUDA IPC), summed by a gather kernel; no NCCL") if world > 1 else "single GPU",
                   "result_checked_vs_known_dlog": checked, "e2e_result_matches": e2e_ok},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": total_pts / e2e_s, "unit": UNIT, "h2d_bytes_per_step": R * n * 32 * world,
                "d2h_bytes_per_step": R * 64 * world, "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "Context.msm_async(points, pinned_scalars, slot) + Context.result(slot), pipelined %d deep" % E2E_DEPTH},
        "roofline": {"bound": "imad", "kernel": "vmsm_kernel<KAccumulateT<false>> (one thread per bucket, plain bases)", "achieved": achieved, "peak": peak_tlps,
                     "unit": "T limb-products/s", "frac": (achieved / peak_tlps) if achieved else None,
                     "traffic": _ncu_traffic_bytes() if args.log2n == 20 else None, "traffic_unit": "DRAM bytes per launch (ncu --set full, profiles/)",
                     "algorithmic_gather_bytes_per_launch": n * W_c * 96,
                     "peak_source": "measured live: vmsm_microbench_imad (independent carry-chained IMAD.WIDE.U32 multiply-accumulate chains, all SMs; every IMAD.WIDE form is half-rate on B200)",
                     "imad_wide_tlps": peak_tlps, "imad_nominal_tlps_at_sm_mhz": (148 * 32 * clocks["sm_mhz"] * 1e6 / 1e12) if clocks and clocks.get("sm_mhz") else None,
                     "kernel_ms": acc_ms, "algorithmic_lp_per_launch": acc_lp,
                     "whole_msm_frac": msm_frac, "whole_msm_lp_per_point": lp_per_point(n),
                     "measured_in": f"{serial_msms} extra MSMs of the same workload with the library's streams joined "
                                    "(serial), CUDA events around every phase; the headline loop overlaps phases of "
                                    "consecutive MSMs, so its per-phase times are not additive",
                     "serial_ms_per_msm": serial_ms, "kernel_share_of_serial_step": acc_ms / serial_ms,
                     "phase_ms": {k: v / max(serial_calls, 1) for k, v in serial_phases.items()},
                     "overlapped_phase_ms": {k: v / max(calls, 1) for k, v in phases.items()},
                     "hbm_phases": {"counting_sort": {
                         "algorithmic_bytes": sort_bytes, "ms": sort_ms, "GB_s": sort_gbs, "hbm_peak_GB_s": hbm_peak,
                         "frac": (sort_gbs / hbm_peak) if sort_gbs else None, "peak_source": hbm_src,
                         "note": "bound by L2 atomics (16 fetch-and-adds per scalar), not by HBM bandwidth; in the "
                                 "headline loop it runs as a thin slice under the accumulate kernel"}}},
        "strong": strong,
    }
    if sweep is not None:
        line["sweep"] = sweep
        line["fixed_generators"] = fixed
        line["ac20"] = ac20
        line["bn256"] = bn256
    if args.cpu_baseline:
        cores = os.cpu_count() or 1
        rate, dt = cpu_reference_rate(args.cpu_points, cores)
        rate1, dt1 = cpu_reference_rate(max(64, args.cpu_points // 4), 1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "one_core_points_per_s": rate1,
                                "sample": f"{cores} processes x {args.cpu_points} terms ({dt:.1f} s) of the same MSM with the "
                                          "pure-Python restatement of pivot.vector_commitment (oracle/ed25519.py); "
                                          f"1 core: {max(64, args.cpu_points // 4)} terms ({dt1:.1f} s); plain Python ints -- "
                                          "gmpy2 / MPyC are not installable in this image and /root/reference does not "
                                          "exist on the GPU box"}
    print(json.dumps(line), flush=True)


def strong_block(ctx, dist, rank, world, issue_seq, bases_full, scal_full, recycle, barrier):
    """ms per ONE fixed 2^STRONG_LOG2N-term MSM split by index range over `world` GPUs (pipelined issue, device
    resident, max over ranks), plain path and over precomputed tables, and the same MSM on one GPU for the ratio."""
    from tools import dist_util
    from verifiable_mpc_b200 import _lib, synth

    n_total = 1 << STRONG_LOG2N
    count = 480  # 0.1 s (8 GPUs) .. 0.66 s (1 GPU) per timed loop: the fill and drain of the pipeline (~1 ms) stay below 1 %
    own_full = bases_full is None

    def single_gpu_times():
        nonlocal bases_full, scal_full
        if own_full:
            bases_full, scal_full = zip(*synth_pairs(ctx, n_total, NSETS))
        pairs = list(zip(bases_full, scal_full))
        t_plain = time_msms(ctx, pairs, count)
        ref = ctx.result((count - 1) % 48)
        for pts in bases_full:
            pts.precompute()
        ctx.sync()
        t_pre = time_msms(ctx, pairs, count)
        same = bool(ctx.result((count - 1) % 48) == ref)
        return t_plain, t_pre, same

    if world == 1:
        t_plain, t_pre, same = single_gpu_times()
        return {"n_total": n_total, "n_gpus": 1, "ms_per_msm": t_plain, "ms_per_msm_generator_tables": t_pre,
                "x_vs_1_gpu": 1.0, "x_vs_1_gpu_generator_tables": 1.0, "tables_equal_plain_result": same,
                "msms_timed": count}

    # this rank's slice of the fixed MSM: indices [rank*m, (rank+1)*m) of the same seeded streams as the headline sets
    m = n_total // world
    sl_bases, sl_scal = [], []
    for k in range(NSETS):
        sl_bases.append(ctx.fixed_base(scalars=synth.scalars_ed25519(SEED_BASES + 16 * k, m, start=rank * m)))
        sl_scal.append(ctx.upload_scalars(synth.scalars_ed25519(SEED_SCALARS + 16 * k, m, start=rank * m)))

    def shard_issue(i):
        issue_seq[0] += 1
        ctx.set_option(_lib.OPT_SHARD_SEQ, issue_seq[0])
        ctx.msm_dev(sl_bases[i % NSETS], sl_scal[i % NSETS], slot=i % 48)

    def timed_sharded():
        for i in range(6):
            shard_issue(i)
        barrier()
        ctx.timer_start()
        for i in range(count):
            recycle(i)
            shard_issue(i)
        t = ctx.timer_stop() / count
        barrier()
        return dist_util.max_over_ranks(dist, t), ctx.result((count - 1) % 48)

    t_plain_n, res_plain = timed_sharded()
    for pts in sl_bases:
        pts.precompute()
    ctx.sync()
    t_pre_n, res_pre = timed_sharded()
    # the same fixed MSM on ONE GPU (rank 0, which holds indices [0, 2^20) of the same streams), others idle
    one = None
    if rank == 0:
        one = single_gpu_times()
        ok = None
        if not own_full or True:
            ctx.msm_dev(bases_full[(count - 1) % NSETS], scal_full[(count - 1) % NSETS], slot=51)
            ok = bool(ctx.result(51) == res_plain == res_pre)
    barrier()
    if rank != 0:
        return None
    t1_plain, t1_pre, same = one
    return {"n_total": n_total, "n_gpus": world, "terms_per_gpu": m, "ms_per_msm": t_plain_n,
            "ms_per_msm_generator_tables": t_pre_n, "one_gpu_ms_per_msm": t1_plain,
            "one_gpu_ms_per_msm_generator_tables": t1_pre, "x_vs_1_gpu": t1_plain / t_plain_n,
            "x_vs_1_gpu_generator_tables": t1_pre / t_pre_n, "x_generator_tables_vs_1_gpu_plain": t1_plain / t_pre_n,
            "sharded_result_equals_one_gpu_result": ok, "msms_timed": count,
            "split": "index range, one slice (and its table) per GPU; 128-byte partials pushed into rank 0's mailbox "
                     "by each final kernel and summed by its gather kernel (no NCCL, no host hop)"}


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

    # newest first: the round-2 summaries (tools/ncu_summarize.py: one row per launch, dram_rd_MB / dram_wr_MB columns)
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*", "ncu_ed_accumulate_plain_2p20.csv")), reverse=True):
        try:
            rows = [r for r in csv.DictReader(open(path)) if "KAccumulate" in (r.get("kernel") or "")]
            tot = [1e6 * (float(r["dram_rd_MB"]) + float(r["dram_wr_MB"])) for r in rows if r.get("dram_rd_MB")]
            if tot:
                return sum(tot) / len(tot)
        except Exception:
            continue
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
    ap.add_argument("--msms-per-step", type=int, default=MSMS_PER_STEP)
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip the sweep / AC20 blocks (N = 1)")
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
