"""Pure-Python restatement of the reference's Ed25519 group arithmetic (TEST INFRASTRUCTURE ONLY).

What it restates (reference file:line, relative to /root/reference):
  * ``g[i] ** x_i`` per-element scalar multiplication + product of the list
    (verifiable_mpc/ac20/pivot.py:139-145 ``vector_commitment``, pivot.py:26-28 ``list_mul``)
  * ``(g_hat_l[i] ** c) * g_hat_r[i]`` generator fold (verifiable_mpc/ac20/compressed_pivot.py:64, :178)
  * the group itself is MPyC's ``EllipticCurve('Ed25519', 'projective')`` (demos/demo_zkp_ac20.py:46),
    which is third-party code absent from this image: twisted Edwards curve
    -x^2 + y^2 = 1 + d x^2 y^2 over GF(2^255-19) in projective (X:Y:Z) coordinates with the EFD
    add-2008-bbjlp / dbl-2008-bbjlp formulas, ``repeat`` = right-to-left binary double-and-add,
    ``mpctools.reduce`` = binary-tree reduction.  Parity with real MPyC internals is UNPINNED; parity is
    defined on canonical affine (x, y).

Everything is plain Python ints; points are tuples.  Two independent formula sets are provided
(affine and projective) so tests can cross-check them.
"""

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493  # group order (prime)
D = (-121665 * pow(121666, -1, P)) % P
D2 = (2 * D) % P
BX = 15112221349535400772501151409588531511454012693041857206046113283949847762202
BY = 46316835694926478169428394003475163141307993866256225615783033603165251855960
B = (BX, BY)
IDENTITY = (0, 1)


# ------------------------------------------------------------------ affine (formula set 1)
def on_curve(pt):
    x, y = pt
    return (-x * x + y * y - 1 - D * x * x * y * y) % P == 0


def affine_add(p1, p2):
    """Unified affine twisted-Edwards addition (a = -1); complete because d is a non-square."""
    x1, y1 = p1
    x2, y2 = p2
    t = D * x1 * x2 * y1 * y2 % P
    x3 = (x1 * y2 + y1 * x2) * pow(1 + t, -1, P) % P
    y3 = (y1 * y2 + x1 * x2) * pow(1 - t, -1, P) % P
    return (x3, y3)


def affine_neg(p1):
    return ((-p1[0]) % P, p1[1])


# ------------------------------------------------------------------ projective (formula set 2)
def to_projective(pt):
    return (pt[0], pt[1], 1)


def normalize(pp):
    """(X:Y:Z) -> canonical affine (x, y), the form every parity comparison uses."""
    X, Y, Z = pp
    zi = pow(Z, -1, P)
    return (X * zi % P, Y * zi % P)


PROJ_IDENTITY = (0, 1, 1)


def proj_add(p1, p2):
    """EFD add-2008-bbjlp with a = -1 (10M + 1S + 1D)."""
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    A = Z1 * Z2 % P
    Bq = A * A % P
    C = X1 * X2 % P
    Dd = Y1 * Y2 % P
    E = D * C * Dd % P
    F = (Bq - E) % P
    G = (Bq + E) % P
    X3 = A * F * ((X1 + Y1) * (X2 + Y2) - C - Dd) % P
    Y3 = A * G * (Dd + C) % P  # D - a*C with a = -1
    Z3 = F * G % P
    return (X3, Y3, Z3)


def proj_dbl(p1):
    """EFD dbl-2008-bbjlp with a = -1 (3M + 4S)."""
    X1, Y1, Z1 = p1
    Bq = (X1 + Y1) ** 2 % P
    C = X1 * X1 % P
    Dd = Y1 * Y1 % P
    E = (-C) % P
    F = (E + Dd) % P
    H = Z1 * Z1 % P
    J = (F - 2 * H) % P
    X3 = (Bq - C - Dd) * J % P
    Y3 = F * (E - Dd) % P
    Z3 = F * J % P
    return (X3, Y3, Z3)


def proj_neg(p1):
    return ((-p1[0]) % P, p1[1], p1[2])


def proj_eq(p1, p2):
    return (p1[0] * p2[2] - p2[0] * p1[2]) % P == 0 and (p1[1] * p2[2] - p2[1] * p1[2]) % P == 0


def proj_repeat(p1, n):
    """Right-to-left binary double-and-add, the shape of MPyC's generic ``repeat`` (UNVERIFIED recollection,
    SURVEY.md App. B.3): negative n -> inverse first; n == 0 -> identity."""
    if n < 0:
        p1 = proj_neg(p1)
        n = -n
    if n == 0:
        return PROJ_IDENTITY
    d = p1
    c = PROJ_IDENTITY
    for i in range(n.bit_length() - 1):
        if (n >> i) & 1:
            c = proj_add(c, d)
        d = proj_dbl(d)
    return proj_add(c, d)


def affine_repeat(p1, n):
    """Left-to-right double-and-add on affine formulas: an independent second path for cross-checks."""
    if n < 0:
        p1 = affine_neg(p1)
        n = -n
    acc = IDENTITY
    for i in reversed(range(n.bit_length())):
        acc = affine_add(acc, acc)
        if (n >> i) & 1:
            acc = affine_add(acc, p1)
    return acc


def tree_reduce(op, xs, initial):
    """Binary-tree reduction in the style of mpyc.mpctools.reduce (pivot.py:28 call site)."""
    xs = [initial] + list(xs)
    while len(xs) > 1:
        odd = len(xs) % 2
        xs[odd:] = [op(xs[i], xs[i + 1]) for i in range(odd, len(xs), 2)]
    return xs[0]


# ------------------------------------------------------------------ the hot-path functions
def scalar_mul(pt, n):
    """Affine in, canonical affine out; via the projective path (the reference's coordinates)."""
    return normalize(proj_repeat(to_projective(pt), n))


def msm_naive(scalars, points):
    """Restates pivot.py:143 ``list_mul([g[i] ** _int(x_i) ...])``: n independent double-and-add scalar
    multiplications then a tree product.  Scalars are arbitrary Python ints (negative / unreduced allowed,
    as the reference passes them).  Returns canonical affine (x, y)."""
    assert len(points) >= len(scalars), "Not enough generators."
    terms = [proj_repeat(to_projective(points[i]), int(s)) for i, s in enumerate(scalars)]
    return normalize(tree_reduce(proj_add, terms, PROJ_IDENTITY))


def vector_commitment(x, gamma, g, h):
    """pivot.py:139-145: ``(h ** gamma) * list_mul([g[i] ** x_i])`` on canonical affine points."""
    assert len(g) >= len(x), "Not enough generators."
    terms = [proj_repeat(to_projective(g[i]), int(s)) for i, s in enumerate(x)]
    prod = tree_reduce(proj_add, terms, PROJ_IDENTITY)
    c = proj_add(proj_repeat(to_projective(h), int(gamma)), prod)
    return normalize(c)


def fold(points, c):
    """compressed_pivot.py:64: ``g_prime[i] = (g_hat_l[i] ** c) * g_hat_r[i]``, canonical affine out."""
    half = len(points) // 2
    out = []
    for i in range(half):
        t = proj_add(proj_repeat(to_projective(points[i]), int(c)), to_projective(points[half + i]))
        out.append(normalize(t))
    return out


def msm_known_dlog(scalars, dlogs):
    """Size-independent check for bases g_i = r_i * B:  MSM(s, g) = (sum s_i r_i mod l) * B  (SURVEY.md 8c(3))."""
    e = sum(int(s) * int(r) for s, r in zip(scalars, dlogs)) % L
    return scalar_mul(B, e)


# ------------------------------------------------------------------ encodings shared with the C ABI
def point_to_bytes(pt):
    """64-byte canonical affine little-endian x || y (include/vmsm.h wire format)."""
    return int(pt[0]).to_bytes(32, "little") + int(pt[1]).to_bytes(32, "little")


def point_from_bytes(b):
    return (int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little"))


def scalar_to_bytes(s):
    return (int(s) % L).to_bytes(32, "little")
