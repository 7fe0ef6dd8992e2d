"""torch.distributed glue for bench.py / the gloo tests (timing reductions only; never point or scalar data).
Kept outside the product package: verifiable_mpc_b200 imports no PyTorch."""


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
