"""world_size-2 gloo test (CPU) of the multi-GPU host logic: slice ownership, mailbox-handle broadcast, reductions, and
the algebra of the split -- the sum of the per-rank partial MSMs equals the MSM of the whole vector."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    sys.path.insert(0, os.path.join(%r, "tests"))
    import torch.distributed as dist
    from oracle import ed25519 as E, prng
    from verifiable_mpc_b200 import shard
    from tools import dist_util
    from fake_engine import FakeContext

    class Ctx(FakeContext):
        opened = None
        def mailbox_create(self, world):
            return bytes(range(64))
        def mailbox_open_ipc(self, handle, rank, world):
            Ctx.opened = (handle, rank, world)

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = Ctx()
    handle = shard.setup_mailbox(ctx, dist, rank, world)
    assert handle == bytes(range(64))
    assert (Ctx.opened is None) == (rank == 0)
    if rank:
        assert Ctx.opened == (bytes(range(64)), rank, world)

    n_total = 37                                    # odd on purpose: uneven slices
    slices = [shard.rank_slice(n_total, world, r) for r in range(world)]
    assert slices[0][0] == 0 and sum(c for _, c in slices) == n_total
    assert all(slices[r][0] + slices[r][1] == slices[r + 1][0] for r in range(world - 1))
    start, cnt = slices[rank]
    dl = [prng.scalar(0x5EEE, i) for i in range(start, start + cnt)]
    sc = [prng.scalar(0x5EED, i) for i in range(start, start + cnt)]
    pts = ctx.fixed_base(scalars=dl)
    partial = ctx.msm(pts, sc)                      # this rank's slice (oracle arithmetic standing in for the GPU)
    gathered = [None] * world
    dist.all_gather_object(gathered, partial)
    assert dist_util.max_over_ranks(dist, float(rank)) == world - 1
    assert dist_util.sum_over_ranks(dist, rank + 1) == world * (world + 1) // 2
    if rank == 0:
        total = ctx.lincomb(gathered, [1] * world)
        full_dl = [prng.scalar(0x5EEE, i) for i in range(n_total)]
        full_sc = [prng.scalar(0x5EED, i) for i in range(n_total)]
        assert total == E.msm_known_dlog(full_sc, full_dl)
        print("SHARD_OK")
    dist.destroy_process_group()
""") % (ROOT, ROOT)


def test_two_rank_split_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29600 + os.getpid() % 300
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (out, err) in zip(procs, outs):
        assert p.returncode == 0, err[-2000:]
    assert "SHARD_OK" in outs[0][0]


def test_rank_slice_properties():
    from verifiable_mpc_b200.shard import rank_slice

    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 4, 8):
            sl = [rank_slice(n, world, r) for r in range(world)]
            assert sl[0][0] == 0 and sum(c for _, c in sl) == n
            assert max(c for _, c in sl) - min(c for _, c in sl) <= 1


def test_three_party_share_local_commitment_over_gloo():
    """tools/demo_mpc_parties.py with three processes (the CPU oracle standing in for each party's GPU): Shamir shares,
    per-party local MSM, exchange of the factors, product == commitment with the known discrete logs."""
    port = 29900 + os.getpid() % 90
    procs = []
    for rank in range(3):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="3", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PYTHONPATH=ROOT)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "demo_mpc_parties.py"), "--log2n", "4",
                                       "--fake"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (out, err) in zip(procs, outs):
        assert p.returncode == 0, err[-2000:]
    assert '"matches_known_dlog": true' in outs[0][0]
