"""Pure-Python restatement of the BN256 group arithmetic used by the Pinocchio prover (TEST INFRASTRUCTURE ONLY).

Reference call sites: trinocchio/pynocchio.py:228-273 ``compute_proof`` (8 multi-exponentiations in ADDITIVE notation:
``int(c[i]) * evalkey[key]`` then ``apply_to_list(point_add, ...)`` :82-91), groups ``EllipticCurve('BN256','jacobian')``
and ``EllipticCurve('BN256_twist','jacobian')`` (demos/demo_zkp_pynocchio.py:27-29).  Parameters from
verifiable_mpc/ac20/pairing.py:51-56: u = 1868033^3, p = 36u^4+36u^3+24u^2+6u+1, Fp2 = Fp[i]/(i^2+1), xi = i+3,
E: y^2 = x^3 + 3, E': y^2 = x^3 + 3/xi.  This is the 256-bit Barreto-Naehrig curve of Go's x/crypto/bn256, not
Ethereum's alt_bn128.  MPyC's own formulas are unavailable here (parity unpinned at that boundary); results are
compared on canonical affine coordinates, with the group laws pinned by curve-membership, order and
known-discrete-log checks (tests/test_oracle_kat.py).

Points: None = identity, else affine (x, y) with x, y ints (G1) or pairs (re, im) (G2).
"""

U = 1868033 ** 3
P = 36 * U ** 4 + 36 * U ** 3 + 24 * U ** 2 + 6 * U + 1
N = 36 * U ** 4 + 36 * U ** 3 + 18 * U ** 2 + 6 * U + 1
assert P == 65000549695646603732796438742359905742825358107623003571877145026864184071783
assert N == 65000549695646603732796438742359905742570406053903786389881062969044166799969

G1 = (1, P - 2)
G2 = ((64746500191241794695844075326670126197795977525365406531717464316923369116492,
       21167961636542580255011770066570541300993051739349375019639421053990175267184),
      (17778617556404439934652658462602675281523610326338642107814333856843981424549,
       20666913350058776956210519119118544732556678129809273996262322366050359951122))


# ------------------------------------------------------------------ Fp2 = Fp[i]/(i^2+1), elements (re, im)
def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * d % P, (-a[1]) * d % P)


def f2_scale(a, k):
    return (a[0] * k % P, a[1] * k % P)


B1 = 3
B2 = f2_mul((3, 0), f2_inv((3, 1)))  # 3 / (i + 3)


class Field:
    """Uniform view of Fp and Fp2 so the curve code below is written once."""

    def __init__(self, ext):
        self.ext = ext
        self.zero = (0, 0) if ext else 0
        self.one = (1, 0) if ext else 1

    def add(self, a, b):
        return f2_add(a, b) if self.ext else (a + b) % P

    def sub(self, a, b):
        return f2_sub(a, b) if self.ext else (a - b) % P

    def mul(self, a, b):
        return f2_mul(a, b) if self.ext else a * b % P

    def inv(self, a):
        return f2_inv(a) if self.ext else pow(a, -1, P)

    def small(self, a, k):
        return f2_scale(a, k) if self.ext else a * k % P


FP, FP2 = Field(False), Field(True)


def on_curve(F, pt):
    if pt is None:
        return True
    x, y = pt
    b = B2 if F.ext else B1
    return F.sub(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), (b if F.ext else b))) == F.zero


# ------------------------------------------------------------------ affine group law (formula set 1)
def affine_add(F, p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if F.add(y1, y2) == F.zero:
            return None
        lam = F.mul(F.small(F.mul(x1, x1), 3), F.inv(F.small(y1, 2)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    return (x3, F.sub(F.mul(lam, F.sub(x1, x3)), y1))


def affine_neg(F, p1):
    return None if p1 is None else (p1[0], F.sub(F.zero, p1[1]))


# ------------------------------------------------------------------ Jacobian (formula set 2), identity = Z == 0
def to_jac(F, pt):
    return (F.one, F.one, F.zero) if pt is None else (pt[0], pt[1], F.one)


def normalize(F, j):
    X, Y, Z = j
    if Z == F.zero:
        return None
    zi = F.inv(Z)
    zi2 = F.mul(zi, zi)
    return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))


def jac_dbl(F, p1):
    X, Y, Z = p1
    if Z == F.zero or Y == F.zero:
        return (F.one, F.one, F.zero)
    A = F.mul(X, X)
    Bq = F.mul(Y, Y)
    C = F.mul(Bq, Bq)
    t = F.add(X, Bq)
    D = F.small(F.sub(F.sub(F.mul(t, t), A), C), 2)
    Ee = F.small(A, 3)
    Ff = F.mul(Ee, Ee)
    X3 = F.sub(Ff, F.small(D, 2))
    Y3 = F.sub(F.mul(Ee, F.sub(D, X3)), F.small(C, 8))
    Z3 = F.small(F.mul(Y, Z), 2)
    return (X3, Y3, Z3)


def jac_add(F, p1, p2):
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    if Z1 == F.zero:
        return p2
    if Z2 == F.zero:
        return p1
    Z1Z1 = F.mul(Z1, Z1)
    Z2Z2 = F.mul(Z2, Z2)
    U1 = F.mul(X1, Z2Z2)
    U2 = F.mul(X2, Z1Z1)
    S1 = F.mul(Y1, F.mul(Z2, Z2Z2))
    S2 = F.mul(Y2, F.mul(Z1, Z1Z1))
    if U1 == U2:
        if S1 == S2:
            return jac_dbl(F, p1)
        return (F.one, F.one, F.zero)
    H = F.sub(U2, U1)
    R = F.sub(S2, S1)
    HH = F.mul(H, H)
    HHH = F.mul(H, HH)
    V = F.mul(U1, HH)
    X3 = F.sub(F.sub(F.mul(R, R), HHH), F.small(V, 2))
    Y3 = F.sub(F.mul(R, F.sub(V, X3)), F.mul(S1, HHH))
    Z3 = F.mul(F.mul(Z1, Z2), H)
    return (X3, Y3, Z3)


def jac_repeat(F, p1, n):
    """Right-to-left double-and-add (shape of MPyC's generic repeat; negative n -> inverse first)."""
    if n < 0:
        p1 = (p1[0], F.sub(F.zero, p1[1]), p1[2])
        n = -n
    acc = (F.one, F.one, F.zero)
    d = p1
    while n:
        if n & 1:
            acc = jac_add(F, acc, d)
        d = jac_dbl(F, d)
        n >>= 1
    return acc


def scalar_mul(F, pt, n):
    return normalize(F, jac_repeat(F, to_jac(F, pt), int(n)))


def apply_to_list(op, inputs):
    """Binary-tree application, restating trinocchio/pynocchio.py:82-91."""
    n = len(inputs)
    if n == 1:
        return inputs[0]
    return op(apply_to_list(op, inputs[: n // 2]), apply_to_list(op, inputs[n // 2:]))


def msm_naive(F, scalars, points):
    """pynocchio.py:229-246: terms = [int(c[i]) * key_i ...]; apply_to_list(point_add, terms).  Empty -> identity."""
    if not scalars:
        return None
    terms = [jac_repeat(F, to_jac(F, points[i]), int(s)) for i, s in enumerate(scalars)]
    return normalize(F, apply_to_list(lambda a, b: jac_add(F, a, b), terms))


def msm_known_dlog(F, scalars, dlogs):
    e = sum(int(s) * int(r) for s, r in zip(scalars, dlogs)) % N
    return scalar_mul(F, G2 if F.ext else G1, e)


# ------------------------------------------------------------------ wire formats of include/vmsm.h
def fp_to_bytes(v):
    return int(v).to_bytes(32, "little")


def point_to_bytes(F, pt):
    """G1: 64 B x||y; G2: 128 B x.re||x.im||y.re||y.im; identity = all zero bytes ((0,0) is not on either curve)."""
    if pt is None:
        return bytes(128 if F.ext else 64)
    if F.ext:
        return b"".join(fp_to_bytes(v) for v in (pt[0][0], pt[0][1], pt[1][0], pt[1][1]))
    return fp_to_bytes(pt[0]) + fp_to_bytes(pt[1])


def point_from_bytes(F, b):
    vals = [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]
    if not any(vals):
        return None
    return ((vals[0], vals[1]), (vals[2], vals[3])) if F.ext else (vals[0], vals[1])
