"""GPU-backed twin of the PROVER side of ``verifiable_mpc/ac20/knowledge_of_exponent.py`` (KoE pivot over BN256).

``restriction_argument_prover`` (:75-91): two multi-exponentiations with the same exponents, one over the G1 powers
``pp_lhs`` and one over the G2 powers ``pp_rhs``;  ``opening_linear_form_prover`` (:101-131): a 2n-term G1
multi-exponentiation with the coefficients of c(X) = (gamma + x_1 X + ...)(L_n + L_{n-1} X + ...).  Each becomes ONE
device MSM (libvmsm.so on the BN256 curves).  The polynomial product and the pairing-based verifier stay host-side /
the reference's; ``linear_form_R`` offers the verifier's n-term G2 product (:146) as a device MSM as well.
Same names, arguments and return values as the reference functions.
"""
from random import SystemRandom

from ..engine import BN_N, pack_scalars
from . import pivot

prng = SystemRandom()


def _msm(points, scalars):
    group = type(points[0])
    ctx = group._ctx()
    dev = ctx.upload_points([p.affine() for p in points], curve=group.curve_id)
    try:
        return group._make(ctx.msm(dev, pack_scalars(scalars, BN_N)))
    finally:
        dev.free()


class PreparedPP(dict):
    """The public parameters of ``trusted_setup`` with ``pp_lhs`` / ``pp_rhs`` ALSO resident on the device, so the
    prover functions below skip the upload of up to 2n bases per call (they only depend on the setup).  It is the
    reference's ``pp`` dict (same keys, same lists) and can be passed wherever ``pp`` is expected."""

    def __init__(self, pp):
        super().__init__(pp)
        self.dev = {}
        for key in ("pp_lhs", "pp_rhs"):
            group = type(pp[key][0])
            self.dev[key] = group._ctx().upload_points([p.affine() for p in pp[key]], curve=group.curve_id)

    def free(self):
        for dev in self.dev.values():
            dev.free()
        self.dev = {}


def _msm_pp(pp, key, first, count, scalars):
    """sum_j scalars[j] * pp[key][first + j], from the resident copy when ``pp`` is a PreparedPP."""
    dev = getattr(pp, "dev", {}).get(key)
    if dev is None:
        return _msm(list(pp[key][first:first + count]), scalars)
    group = type(pp[key][0])
    return group._make(group._ctx().msm(dev, pack_scalars(scalars, BN_N), off=first, n=count))


def list_mul(x):
    return _msm(list(x), [1] * len(x))


def vector_commitment(x, gamma, g, h):
    """``h**gamma * prod g[i]**x[i]`` over a BN256 group (reference :29-38)."""
    assert len(g) >= len(x), "Not enough generators."
    return _msm(list(g[: len(x)]) + [h], [int(v) for v in x] + [int(gamma)])


def trusted_setup(_g1, _g2, n, order, progress_bar=False):
    """Public parameters ``pp_lhs[i] = g1**(z**(i+1))``, ``pp_rhs[i] = g2**(z**(i+1))`` with ``g1 = _g1**g_exp``,
    ``g2 = _g2**(g_exp*alpha)`` (reference :50-72: 4n sequential scalar multiplications).  Same three draws from
    ``prng``; the 2n + 2n powers are two fixed-base batches on the device when ``_g1`` / ``_g2`` are the groups'
    standard generators (how every caller invokes it: circuit_sat_r1cs.py:84-91), otherwise one device call each."""
    g_exp = prng.randrange(1, order)
    alpha = prng.randrange(order)
    z = prng.randrange(order)
    group1, group2 = type(_g1), type(_g2)
    e1, e2, zp = [], [], 1
    for _ in range(2 * n):
        zp = zp * z % order
        e1.append(g_exp * zp % order)
        e2.append(g_exp * alpha % order * zp % order)
    pp = {}
    for key, group, base, exps in (("pp_lhs", group1, _g1, e1), ("pp_rhs", group2, _g2, e2)):
        if base == group.generator:
            dev = group._ctx().fixed_base(scalars=exps, curve=group.curve_id)
            try:
                pp[key] = [group._make(pt) for pt in dev.tolist()]
            finally:
                dev.free()
        else:
            pp[key] = [group.lincomb([base], [e]) for e in exps]
    return pp


def restriction_argument_prover(S, x, gamma, pp):
    """Restriction argument [Gro10], prover: (P, pi) over the S-indices of x (reference :75-91)."""
    S = list(S)
    scalars = [int(gamma)] + [int(x[i]) for i in S]
    if S == list(range(len(S))):  # the usual case (all of x): a contiguous range of the parameters
        return _msm_pp(pp, "pp_lhs", 0, len(S) + 1, scalars), _msm_pp(pp, "pp_rhs", 0, len(S) + 1, scalars)
    P = _msm([pp["pp_lhs"][0]] + [pp["pp_lhs"][i + 1] for i in S], scalars)
    pi = _msm([pp["pp_rhs"][0]] + [pp["pp_rhs"][i + 1] for i in S], scalars)
    return P, pi


def _poly_mul_mod(a, b, q):
    out = [0] * (len(a) + len(b) - 1)
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                out[i + j] = (out[i + j] + ai * bj) % q
    return out


def opening_linear_form_prover(L, x, gamma, pp, P=None, pi=None):
    """ZK argument of knowledge for the opening of a linear form (reference :101-131)."""
    proof = {}
    n = len(x)
    S = range(n)
    assert 2 * n - 1 <= len(pp["pp_lhs"]), \
        "Requirement does not hold: 2*len(x)-1 <= number of generators in first group."
    if P is None:
        P, pi = restriction_argument_prover(S, x, gamma, pp)
    proof["P"] = P
    proof["pi"] = pi
    u = L(x)
    L_linear, u_linear = pivot.affine_to_linear(L, u, n)
    q = BN_N
    lhs = [int(gamma) % q] + [int(x_i) % q for x_i in x]
    rhs = [int(L_linear.coeffs[n - (j + 1)]) % q for j in range(n)]
    c_bar = _poly_mul_mod(lhs, rhs, q)
    assert int(u_linear) % q == c_bar[n], "L(x) not equal to n-th coefficient of c_poly"
    c_bar[n] = 0
    assert len(pp["pp_lhs"]) == 2 * n
    proof["Q"] = _msm_pp(pp, "pp_lhs", 0, 2 * n, [-c for c in c_bar])
    return proof, u


def linear_form_R(L, pp, u):
    """The verifier's ``R = prod pp_rhs[j] ** L_linear.coeffs[n-(j+1)]`` (reference :146) as one G2 MSM."""
    n = len(L.coeffs)
    L_linear, _ = pivot.affine_to_linear(L, u, n)
    return _msm_pp(pp, "pp_rhs", 0, n, [int(L_linear.coeffs[n - (j + 1)]) for j in range(n)])


def prove_nullity_koe(pp, lin_forms, x, gamma, gf, P, pi):
    """Nullity protocol on top of the linear-form opening (reference :152-162)."""
    rho = pivot.fiat_shamir_hash([P, lin_forms], gf.order)
    L = sum((linform_i) * (rho ** i) for i, linform_i in enumerate(lin_forms))
    L = pivot.LinearForm([gf(c) if isinstance(c, int) else c for c in L.coeffs])
    proof, u = opening_linear_form_prover(L, x, gamma, pp, P, pi)
    return proof, L, u
