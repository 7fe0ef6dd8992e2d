#!/usr/bin/env python
"""Share-local commitments of the MPC prover, one process and one GPU context per party (BASELINE config 5 shape:
demos/demo_zkp_mpc_ac20.py -M3 --elliptic).  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 3 --master-addr 127.0.0.1 tools/demo_mpc_parties.py --log2n 16

Party i holds Shamir shares (degree t) of the exponent vector, computes its factor of the Pedersen commitment with ONE
MSM on its own device (mpc_ac20.local_commitment_share), the factors are exchanged (here: torch.distributed gloo
all_gather of 64-byte encodings, standing in for MPyC's transfer), and every party multiplies them.  Party 0 checks the
result against the known discrete logs.  `--fake` runs the same flow on the CPU oracle (tests).
"""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=12)
    ap.add_argument("--threshold", type=int, default=1)
    ap.add_argument("--fake", action="store_true", help="CPU oracle instead of the GPU (tests only)")
    args = ap.parse_args()
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from verifiable_mpc_b200 import fingroups
    from verifiable_mpc_b200.ac20 import generators as gens
    from verifiable_mpc_b200.ac20 import mpc_ac20

    if args.fake:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from fake_engine import FakeContext
        fingroups.Ed25519Point.context = FakeContext()
    else:
        from verifiable_mpc_b200 import Context
        from verifiable_mpc_b200._lib import load
        ndev = __import__("ctypes").c_int32()
        load().vmsm_device_count(__import__("ctypes").byref(ndev))
        fingroups.Ed25519Point.context = Context(int(os.environ.get("LOCAL_RANK", rank)) % max(ndev.value, 1))
    group = fingroups.EllipticCurve("Ed25519", "projective")
    group.is_additive, group.is_multiplicative = False, True
    order = group.order
    n = (1 << args.log2n) - 1

    # public generators with known discrete logs (every party derives the same ones), secret exponents x, gamma
    pub = random.Random(2020)
    dlogs = [pub.randrange(1, order) for _ in range(n)]
    generators = gens.create_generators(n, group, with_k=False, exponents=dlogs)
    g, h = generators["g"], generators["h"]
    dealer = random.Random(7)  # stands in for the parties' joint randomness: every process replays the same dealing
    x = [dealer.randrange(order) for _ in range(n)]
    gamma = dealer.randrange(order)
    t = args.threshold
    my_shares = []
    for v in x + [gamma]:
        coeffs = [v] + [dealer.randrange(order) for _ in range(t)]
        my_shares.append(sum(c * pow(rank + 1, k, order) for k, c in enumerate(coeffs)) % order)
    lam = mpc_ac20.recombine_at_zero(order, list(range(1, world + 1)))

    dist.barrier()
    t0 = time.perf_counter()
    mpc_ac20.local_commitment_share(my_shares[:-1], my_shares[-1], g, h, lam[rank])
    t_cold = time.perf_counter() - t0  # first MSM of the process: workspace allocation, module load
    dist.barrier()
    t0 = time.perf_counter()
    part = mpc_ac20.local_commitment_share(my_shares[:-1], my_shares[-1], g, h, lam[rank])
    t_local = time.perf_counter() - t0
    parts = [None] * world
    dist.all_gather_object(parts, part.affine())
    commitment = mpc_ac20.combine_commitment_shares([group._make(p) for p in parts])
    t_all = time.perf_counter() - t0
    ok = None
    if rank == 0:
        e = (sum(a * b for a, b in zip(x, dlogs)) + gamma) % order  # h = B, g_j = dlog_j * B
        ok = commitment == group.generator ** e
        print(json.dumps({"demo": "mpc_share_local_commitment", "parties": world, "threshold": t, "n": n,
                          "local_msm_s": t_local, "local_msm_first_call_s": t_cold, "commit_incl_exchange_s": t_all, "matches_known_dlog": bool(ok),
                          "device": "cpu-oracle" if args.fake else "cuda"}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
