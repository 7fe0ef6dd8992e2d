"""Seeded input generators shared by the fixture makers (tests/golden/make_*_big_golden.py, which run the unmodified
reference in the build container) and by the tests that replay the same inputs on the GPU box (where the reference
tree does not exist).  Nothing here imports the reference, the oracle or the product: plain ``random.Random`` draws in
a fixed order."""
import random

# ------------------------------------------------------------------------------------------------ AC20 (Ed25519)
AC20_SEEDS = {10: 1010, 12: 1212, 16: 1616}
AC20_N_BOOLEAN = 64  # leading witnesses drawn from {0, 1}: real circuits are full of bits (long-bucket path)


def ac20_draw_inputs(k, seed, order):
    """N = 2^k generators: exponents of g (g_i = h ** e_i), exponent of k, witness x, gamma, linear form L."""
    n = (1 << k) - 1
    rng = random.Random(seed)
    exps = [rng.randrange(1, order) for _ in range(n)]
    k_exp = rng.randrange(1, order)
    x = [rng.randrange(order) for _ in range(n)]
    x[:AC20_N_BOOLEAN] = [rng.randrange(2) for _ in range(AC20_N_BOOLEAN)]
    x[AC20_N_BOOLEAN] = order - 1
    gamma = rng.randrange(order)
    L = [rng.randrange(order) for _ in range(n)]
    return exps, k_exp, x, gamma, L


def canonical_proof_text(proof):
    """Proof (hex-encoded fixture form) -> one text line; its sha256 is stored beside the proof."""
    parts = [proof["t"], *proof["A"]]
    for a, b in zip(proof["A_i"], proof["B_i"]):
        parts += [*a, *b]
    parts += proof["z_prime"]
    return ",".join(parts)


# ------------------------------------------------------------------------------------------------ Pinocchio (BN256)
PYN_SEEDS = {8: 808, 10: 1010, 12: 1212, 14: 1414}
PYN_MID_TEMPLATES = ("r_v*v{i}*g1", "r_w*w{i}*g2", "r_y*y{i}*g1", "r_v*alpha_v*v{i}*g1", "r_w*alpha_w*w{i}*g1",
                     "r_y*alpha_y*y{i}*g1", "r_v*beta*v+r_w*beta*w+r_y*beta*y{i}_g1")
PYN_ZK_KEYS = ("r_v*t*g1", "r_w*t*g2", "r_y*t*g1", "r_v*alpha_v*t*g1", "r_w*alpha_w*t*g1", "r_y*alpha_y*t*g1",
               "r_v*beta*t*g1", "r_w*beta*t*g1", "r_y*beta*t*g1")
PYN_FIRST_MID = 3  # indices_mid = range(3, 3 + 2^k), like the demo QAP (one, inputs, outputs come first)


def pynocchio_draw_inputs(k, seed, order):
    """Synthetic evaluation key with known discrete logs (entry = e * generator), witness c (with a run of booleans),
    quotient coefficients h, zero-knowledge deltas.  Draw order: key exponents template by template, the s-powers,
    the ZK entries, then c, h, deltas."""
    m = 1 << k
    rng = random.Random(seed)
    mid = list(range(PYN_FIRST_MID, PYN_FIRST_MID + m))
    key_exps = {}
    for t in PYN_MID_TEMPLATES:
        for i in mid:
            key_exps[t.format(i=i)] = rng.randrange(1, order)
    for i in range(m):
        key_exps[f"s^{i}*g1"] = rng.randrange(1, order)
    for name in PYN_ZK_KEYS:
        key_exps[name] = rng.randrange(1, order)
    c = [rng.randrange(order) for _ in range(PYN_FIRST_MID + m)]
    for j in range(PYN_FIRST_MID, PYN_FIRST_MID + 32):
        c[j] = rng.randrange(2)
    c[PYN_FIRST_MID + 32] = order - 1
    h = [rng.randrange(order) for _ in range(m)]
    deltas = {a: rng.randrange(order) for a in ("v", "w", "y")}
    return mid, key_exps, c, h, deltas
