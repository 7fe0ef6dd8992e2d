"""Host-side helpers for the index-range split of one MSM over the GPUs of a box.

One process per GPU.  The process group (any ``torch.distributed``-like object with ``broadcast_object_list``,
``barrier``; gloo is enough) carries only the 64-byte CUDA-IPC handle of the owner's mailbox and the barrier -- never
point or scalar data.  (The timing reductions of bench.py live in tools/dist_util.py: this package imports no torch.)
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
