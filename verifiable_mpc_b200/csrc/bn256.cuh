// BN256 G1 (over Fp) and G2 (the sextic twist over Fp2) group arithmetic, written once over a field policy
// (FpBN / Fp2BN, fbn256.cuh).  Short Weierstrass y^2 = x^3 + b with a = 0, Jacobian coordinates (X : Y : Z),
// x = X/Z^2, y = Y/Z^3, identity Z = 0.  The formulas (EFD madd-2007-bl, add-2007-bl, dbl-2009-l) are incomplete,
// so every exceptional case (identity operands, P = Q, P = -Q) is branched on explicitly -- repeated bases and
// cancelling terms are legal MSM inputs.
//
// Replaces the per-term `int(c[i]) * evalkey[key]` double-and-add and the `apply_to_list(point_add, ...)` tree of
// trinocchio/pynocchio.py:229-246 (BN256 'jacobian' groups, demos/demo_zkp_pynocchio.py:27-29); results are
// compared on canonical affine coordinates.
#pragma once
#include "fbn256.cuh"

namespace vmsm {

template <class F>
struct wjac {
    typename F::T X, Y, Z;
};
// stored base: affine, Montgomery form; the identity is encoded as (0, 0), which is on neither curve (b != 0)
template <class F>
struct waff {
    typename F::T x, y;
};

template <class F>
VMSM_HD wjac<F> wj_identity() {
    wjac<F> r = {F::one(), F::one(), F::zero()};
    return r;
}
template <class F>
VMSM_HD bool wa_is_identity(const waff<F> &a) {
    return F::is_zero(a.x) && F::is_zero(a.y);
}
template <class F>
VMSM_HD wjac<F> wa_to_jac(const waff<F> &a) {
    if (wa_is_identity(a)) return wj_identity<F>();
    wjac<F> r = {a.x, a.y, F::one()};
    return r;
}

// 2P, dbl-2009-l (a = 0): 2M + 5S.  Z = 0 stays Z = 0.
template <class F>
VMSM_HD wjac<F> wj_dbl(const wjac<F> &p) {
    typedef typename F::T T;
    T A = F::sqr(p.X), B = F::sqr(p.Y), C = F::sqr(B);
    T t = F::sqr(F::add(p.X, B));
    T D = F::dbl(F::sub(F::sub(t, A), C));
    T E = F::add(F::dbl(A), A);
    T Fq = F::sqr(E);
    wjac<F> r;
    r.X = F::sub(Fq, F::dbl(D));
    T C8 = F::dbl(F::dbl(F::dbl(C)));
    r.Y = F::sub(F::mul(E, F::sub(D, r.X)), C8);
    r.Z = F::dbl(F::mul(p.Y, p.Z));
    return r;
}

// P + (neg ? -Q : Q) with Q affine, madd-2007-bl: 7M + 4S
template <class F>
VMSM_HD wjac<F> wj_madd(const wjac<F> &p, const waff<F> &q, bool neg) {
    typedef typename F::T T;
    if (wa_is_identity(q)) return p;
    T qy = neg ? F::neg(q.y) : q.y;
    if (F::is_zero(p.Z)) {
        wjac<F> r = {q.x, qy, F::one()};
        return r;
    }
    T Z1Z1 = F::sqr(p.Z);
    T U2 = F::mul(q.x, Z1Z1);
    T S2 = F::mul(F::mul(qy, p.Z), Z1Z1);
    T H = F::sub(U2, p.X);
    T rr = F::sub(S2, p.Y);
    if (F::is_zero(H)) {
        if (F::is_zero(rr)) return wj_dbl(p);
        return wj_identity<F>();
    }
    T HH = F::sqr(H);
    T I = F::dbl(F::dbl(HH));
    T J = F::mul(H, I);
    T r2 = F::dbl(rr);
    T V = F::mul(p.X, I);
    wjac<F> r;
    r.X = F::sub(F::sub(F::sqr(r2), J), F::dbl(V));
    r.Y = F::sub(F::mul(r2, F::sub(V, r.X)), F::dbl(F::mul(p.Y, J)));
    r.Z = F::sub(F::sub(F::sqr(F::add(p.Z, H)), Z1Z1), HH);
    return r;
}

// P + Q, both Jacobian, add-2007-bl: 11M + 5S
template <class F>
VMSM_HD wjac<F> wj_add(const wjac<F> &p, const wjac<F> &q) {
    typedef typename F::T T;
    if (F::is_zero(p.Z)) return q;
    if (F::is_zero(q.Z)) return p;
    T Z1Z1 = F::sqr(p.Z), Z2Z2 = F::sqr(q.Z);
    T U1 = F::mul(p.X, Z2Z2), U2 = F::mul(q.X, Z1Z1);
    T S1 = F::mul(F::mul(p.Y, q.Z), Z2Z2), S2 = F::mul(F::mul(q.Y, p.Z), Z1Z1);
    T H = F::sub(U2, U1);
    T rr = F::sub(S2, S1);
    if (F::is_zero(H)) {
        if (F::is_zero(rr)) return wj_dbl(p);
        return wj_identity<F>();
    }
    T I = F::sqr(F::dbl(H));
    T J = F::mul(H, I);
    T r2 = F::dbl(rr);
    T V = F::mul(U1, I);
    wjac<F> r;
    r.X = F::sub(F::sub(F::sqr(r2), J), F::dbl(V));
    r.Y = F::sub(F::mul(r2, F::sub(V, r.X)), F::dbl(F::mul(S1, J)));
    r.Z = F::mul(F::sub(F::sub(F::sqr(F::add(p.Z, q.Z)), Z1Z1), Z2Z2), H);
    return r;
}

// Jacobian -> affine (Montgomery form); identity -> (0, 0)
template <class F>
VMSM_HD_NOINLINE waff<F> wj_to_aff(const wjac<F> &p) {
    typedef typename F::T T;
    waff<F> r;
    if (F::is_zero(p.Z)) {
        r.x = F::zero();
        r.y = F::zero();
        return r;
    }
    T zi = F::inv(p.Z);
    T zi2 = F::sqr(zi);
    r.x = F::mul(p.X, zi2);
    r.y = F::mul(p.Y, F::mul(zi2, zi));
    return r;
}

// y^2 == x^3 + b  (or the identity encoding)
template <class F>
VMSM_HD bool wa_on_curve(const waff<F> &a) {
    if (wa_is_identity(a)) return true;
    typename F::T lhs = F::sqr(a.y);
    typename F::T rhs = F::add(F::mul(F::sqr(a.x), a.x), F::curve_b());
    return F::eq(lhs, rhs);
}

}  // namespace vmsm
