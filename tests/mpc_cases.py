"""Shared body of the MPC share-local commitment tests (CPU with the fake engine, GPU with real contexts)."""
import random


def shamir_shares(values, m, t, order, rng):
    """rows[i][j] = share of values[j] held by party i (evaluation point i + 1), degree-t polynomials."""
    rows = [[] for _ in range(m)]
    for v in values:
        coeffs = [v % order] + [rng.randrange(order) for _ in range(t)]
        for i in range(m):
            x = i + 1
            rows[i].append(sum(c * pow(x, k, order) for k, c in enumerate(coeffs)) % order)
    return rows


def check_share_local_commitments(group, gf, n=33, m=3, t=1, seed=5):
    from verifiable_mpc_b200.ac20 import generators as gens
    from verifiable_mpc_b200.ac20 import mpc_ac20, pivot

    rng = random.Random(seed)
    order = group.order
    gens.prng = rng
    generators = gens.create_generators(n, group)
    g, h = generators["g"], generators["h"]
    x = [rng.randrange(order) for _ in range(n)]
    x[:4] = [0, 1, order - 1, 2]
    gamma = rng.randrange(order)
    rows = shamir_shares(x + [gamma], m, t, order, rng)
    lam = mpc_ac20.recombine_at_zero(order, list(range(1, m + 1)))
    assert sum(l * rows[i][0] for i, l in enumerate(lam)) % order == x[0]
    want = pivot.vector_commitment([gf(v) for v in x], gf(gamma), g, h)

    parts = [mpc_ac20.local_commitment_share(rows[i][:-1], rows[i][-1], g, h, lam[i]) for i in range(m)]
    assert mpc_ac20.combine_commitment_shares(parts) == want
    assert all(p != want for p in parts)  # no single party's factor is the commitment
    # field-element shares, the one-call simulation, and a subset of t + 1 parties with its own recombination vector
    got = mpc_ac20.vector_commitment_from_shares([[gf(v) for v in r[:-1]] for r in rows], [gf(r[-1]) for r in rows], g, h)
    assert got == want
    sub = list(range(0, m, 2))[:t + 1]
    assert len(sub) == t + 1
    got = mpc_ac20.vector_commitment_from_shares([rows[i][:-1] for i in sub], [rows[i][-1] for i in sub], g, h,
                                                 xs=[i + 1 for i in sub])
    assert got == want
    return want
