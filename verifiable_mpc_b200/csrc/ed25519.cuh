// Ed25519 group arithmetic (twisted Edwards, a = -1) on top of fe25519.cuh.
//
//   ge_ext    extended coordinates (X:Y:Z:T), x = X/Z, y = Y/Z, T = XY/Z           128 B
//   ge_niels  affine point cached as (y+x, y-x, 2d*x*y)                            96 B  (MSM / fold input form)
//   ge_aff    canonical affine (x, y), both < p                                     64 B  (wire / transcript form)
//
// Formulas: add-2008-hwcd-3 (unified and complete for a = -1, d non-square) and dbl-2008-hwcd.
// The reference computes the same group operation through MPyC's EllipticCurve('Ed25519','projective')
// (demos/demo_zkp_ac20.py:46) one Python operator at a time (pivot.py:143-144, compressed_pivot.py:64);
// results agree on the canonical affine form, which is what every parity test compares.
#pragma once
#include "fe25519.cuh"

namespace vmsm {

struct ge_ext {
    fe X, Y, Z, T;
};
struct ge_niels {
    fe ypx, ymx, t2d;
};
struct ge_aff {
    fe x, y;
};

// d and 2d, little-endian limbs
VMSM_HD fe fe_const_d() {
    fe r = {{0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du, 0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu}};
    return r;
}
VMSM_HD fe fe_const_2d() {
    fe r = {{0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au, 0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu}};
    return r;
}

// 1/d: turns the cached 2d*x*y of a niels point back into T = 2xy for the (2x : 2y : 2 : 2xy) representative
VMSM_HD fe fe_const_dinv() {
    fe r = {{0xcdc9f843u, 0x25e0f276u, 0x4279542eu, 0x0b5dd698u, 0xcdb9cf66u, 0x2b162114u, 0x14d5ce43u, 0x40907ed2u}};
    return r;
}

VMSM_HD ge_ext ge_identity() {
    ge_ext r;
    r.X = fe_zero();
    r.Y = fe_one();
    r.Z = fe_one();
    r.T = fe_zero();
    return r;
}

VMSM_HD ge_niels ge_niels_identity() {
    ge_niels r;
    r.ypx = fe_one();
    r.ymx = fe_one();
    r.t2d = fe_zero();
    return r;
}

// (neg ? -q : q) as an extended point, 1M: (X : Y : Z : T) = (2x : 2y : 2 : 2xy).  Used to start a bucket sum from its
// first base instead of adding it to the identity (saves 6 of the 7 multiplications of that addition).
VMSM_HD ge_ext ge_from_niels(const ge_niels &q, bool neg) {
    ge_ext r;
    fe X = fe_sub(q.ypx, q.ymx), T = fe_mul(q.t2d, fe_const_dinv());
    r.X = fe_select(neg, fe_neg(X), X);
    r.Y = fe_add(q.ypx, q.ymx);
    r.Z = fe_add(fe_one(), fe_one());
    r.T = fe_select(neg, fe_neg(T), T);
    return r;
}

// mixed addition r = p + (neg ? -q : q), 7M
VMSM_HD ge_ext ge_madd(const ge_ext &p, const ge_niels &q, bool neg) {
    fe qa = fe_select(neg, q.ypx, q.ymx);  // multiplies (Y1 - X1)
    fe qb = fe_select(neg, q.ymx, q.ypx);  // multiplies (Y1 + X1)
    fe A = fe_mul(fe_sub(p.Y, p.X), qa);
    fe B = fe_mul(fe_add(p.Y, p.X), qb);
    fe C = fe_mul(p.T, q.t2d);
    fe D = fe_dbl(p.Z);
    fe E = fe_sub(B, A);
    fe H = fe_add(B, A);
    fe DmC = fe_sub(D, C);
    fe DpC = fe_add(D, C);
    fe F = fe_select(neg, DpC, DmC);
    fe G = fe_select(neg, DmC, DpC);
    ge_ext r;
    r.X = fe_mul(E, F);
    r.Y = fe_mul(G, H);
    r.T = fe_mul(E, H);
    r.Z = fe_mul(F, G);
    return r;
}

// full addition, both extended, 9M
VMSM_HD ge_ext ge_add(const ge_ext &p, const ge_ext &q) {
    fe A = fe_mul(fe_sub(p.Y, p.X), fe_sub(q.Y, q.X));
    fe B = fe_mul(fe_add(p.Y, p.X), fe_add(q.Y, q.X));
    fe C = fe_mul(fe_mul(p.T, q.T), fe_const_2d());
    fe D = fe_dbl(fe_mul(p.Z, q.Z));
    fe E = fe_sub(B, A);
    fe F = fe_sub(D, C);
    fe G = fe_add(D, C);
    fe H = fe_add(B, A);
    ge_ext r;
    r.X = fe_mul(E, F);
    r.Y = fe_mul(G, H);
    r.T = fe_mul(E, H);
    r.Z = fe_mul(F, G);
    return r;
}

// doubling, 4M + 4S
VMSM_HD ge_ext ge_dbl(const ge_ext &p) {
    fe A = fe_sqr(p.X);
    fe B = fe_sqr(p.Y);
    fe C = fe_dbl(fe_sqr(p.Z));
    fe XY = fe_add(p.X, p.Y);
    fe E = fe_sub(fe_sub(fe_sqr(XY), A), B);
    fe G = fe_sub(B, A);          // D + B with D = -A
    fe F = fe_sub(G, C);
    fe H = fe_neg(fe_add(A, B));  // D - B
    ge_ext r;
    r.X = fe_mul(E, F);
    r.Y = fe_mul(G, H);
    r.T = fe_mul(E, H);
    r.Z = fe_mul(F, G);
    return r;
}

VMSM_HD ge_ext ge_neg(const ge_ext &p) {
    ge_ext r = p;
    r.X = fe_neg(p.X);
    r.T = fe_neg(p.T);
    return r;
}

VMSM_HD ge_niels ge_aff_to_niels(const ge_aff &a) {
    ge_niels r;
    r.ypx = fe_add(a.y, a.x);
    r.ymx = fe_sub(a.y, a.x);
    r.t2d = fe_mul(fe_mul(a.x, a.y), fe_const_2d());
    return r;
}

VMSM_HD ge_ext ge_aff_to_ext(const ge_aff &a) {
    ge_ext r;
    r.X = a.x;
    r.Y = a.y;
    r.Z = fe_one();
    r.T = fe_mul(a.x, a.y);
    return r;
}

// one inversion; canonical output
VMSM_HD ge_aff ge_ext_to_aff(const ge_ext &p) {
    fe zi = fe_inv(p.Z);
    ge_aff r;
    r.x = fe_canon(fe_mul(p.X, zi));
    r.y = fe_canon(fe_mul(p.Y, zi));
    return r;
}

// same, given zi = 1/Z computed elsewhere (batched inversion)
VMSM_HD ge_aff ge_ext_to_aff_zi(const ge_ext &p, const fe &zi) {
    ge_aff r;
    r.x = fe_canon(fe_mul(p.X, zi));
    r.y = fe_canon(fe_mul(p.Y, zi));
    return r;
}

// -x^2 + y^2 == 1 + d x^2 y^2
VMSM_HD bool ge_aff_on_curve(const ge_aff &a) {
    fe x2 = fe_sqr(a.x), y2 = fe_sqr(a.y);
    fe lhs = fe_sub(y2, x2);
    fe rhs = fe_add(fe_one(), fe_mul(fe_const_d(), fe_mul(x2, y2)));
    return fe_eq(lhs, rhs);
}

}  // namespace vmsm
