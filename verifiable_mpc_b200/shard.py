"""Host-side helpers for the index-range split of one MSM over the GPUs of a box.

One process per GPU.  The process group (any ``torch.distributed``-like object with ``broadcast_object_list``,
``all_reduce`` and ``barrier``; gloo is enough) carries only the 64-byte CUDA-IPC handle of the owner's mailbox, the
barrier and scalar reductions of timings -- never point or scalar data.
"""


def rank_slice(n_total, world, rank):
    """Contiguous slice [start, start+count) of an n_total-term MSM owned by `rank` (earlier ranks get the remainder)."""
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def setup_mailbox(ctx, dist, rank, world):
    """Rank 0 creates the mailbox in its HBM and broadcasts the IPC handle; the other ranks map it."""
    obj = [ctx.mailbox_create(world) if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    if rank != 0:
        ctx.mailbox_open_ipc(obj[0], rank, world)
    dist.barrier()
    return obj[0]


def max_over_ranks(dist, value):
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, value):
    import torch

    t = torch.tensor([int(value)], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
