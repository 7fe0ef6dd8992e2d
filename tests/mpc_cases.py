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

    from oracle import ed25519 as E

    rng = random.Random(seed)
    order = group.order
    gens.prng = rng
    dlogs = [rng.randrange(1, order) for _ in range(n)]  # known discrete logs: the oracle computes every expected value
    generators = gens.create_generators(n, group, exponents=dlogs)
    g, h = generators["g"], generators["h"]
    x = [rng.randrange(order) for _ in range(n)]
    x[:4] = [0, 1, order - 1, 2]
    gamma = rng.randrange(order)
    rows = shamir_shares(x + [gamma], m, t, order, rng)
    lam = mpc_ac20.recombine_at_zero(order, list(range(1, m + 1)))
    assert sum(l * rows[i][0] for i, l in enumerate(lam)) % order == x[0]
    want = pivot.vector_commitment([gf(v) for v in x], gf(gamma), g, h)
    # expected values from the ORACLE (not from this package): h = B, g_j = dlogs[j] * B
    assert want.affine() == E.msm_known_dlog(x + [gamma], dlogs + [1])

    parts = [mpc_ac20.local_commitment_share(rows[i][:-1], rows[i][-1], g, h, lam[i]) for i in range(m)]
    for i, part in enumerate(parts):  # the reference's party-local factor: prod_j base_j ** (lambda_i * share_i(x_j))
        assert part.affine() == E.msm_known_dlog([lam[i] * v % order for v in rows[i]], dlogs + [1]), f"party {i}"
    if n <= 40:  # and by the reference's own algorithm (per-term double-and-add, tree product) on the host points
        host_g = [E.scalar_mul(E.B, d) for d in dlogs] + [E.B]
        assert parts[0].affine() == E.msm_naive([lam[0] * v % order for v in rows[0]], host_g)
    assert mpc_ac20.combine_commitment_shares(parts) == want
    acc = E.IDENTITY
    for part in parts:
        acc = E.affine_add(acc, part.affine())
    assert want.affine() == acc
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


def check_list_mul(group, n=150, seed=3):
    """pivot.list_mul (pivot.py:26-28) against the oracle: product of n elements with known discrete logs, and the
    reference's tree product itself on a short list with repeated / inverse / identity elements."""
    from oracle import ed25519 as E
    from verifiable_mpc_b200.ac20 import pivot

    rng = random.Random(seed)
    order = group.order
    dlogs = [rng.randrange(order) for _ in range(n)]
    dlogs[3], dlogs[4], dlogs[5] = dlogs[2], (order - dlogs[2]) % order, 0
    elems = [group.generator ** d for d in dlogs]
    for m in (1, 2, 62, 63, 64, 65, 127, n):
        assert pivot.list_mul(elems[:m]).affine() == E.scalar_mul(E.B, sum(dlogs[:m]) % order), m
    short = [e.affine() for e in elems[:9]]
    assert pivot.list_mul(elems[:9]).affine() == E.tree_reduce(E.affine_add, short, E.IDENTITY)
    assert pivot.list_mul(elems[2:5]).affine() == E.scalar_mul(E.B, dlogs[2])  # P * P * P^-1
