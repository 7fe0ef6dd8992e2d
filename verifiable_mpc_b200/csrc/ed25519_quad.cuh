// Quad-cooperative Ed25519 point arithmetic: FOUR adjacent lanes share one point operation.
//
// Why: a lone warp cannot issue more than one IMAD.WIDE every 4 cycles per SM sub-partition, whatever the number of
// active lanes, so the latency-bound tails of the MSM (upper levels of the bucket tree, the Horner chain over the
// windows: ~240 dependent doublings) ran at ~1.35 us per point addition when one thread did all 8-9 field
// multiplications of an addition back to back (profiles/r01: KFinal 0.66 ms, KReduce 0.45 ms of a 2.7 ms MSM).
// Here lane q of a quad holds coordinate q of every point (0:X 1:Y 2:Z 3:T, one field element per lane) and the four
// independent multiplications of each formula stage execute as ONE warp-level fe_mul; operands are exchanged with
// width-4 shuffles.  An addition costs 3 multiplication slots instead of 9, a doubling 2 instead of 8.
//
// Same formulas as ed25519.cuh (add-2008-hwcd-3, dbl-2008-hwcd); device only.  On the host (tests/hostemu) the quad
// kernels fall back to the scalar formulas executed by lane 0 of each quad (see kernels.cuh), which checks the
// indexing but not the shuffle choreography -- that is pinned by the GPU parity tests.
#pragma once
#include "ed25519.cuh"

#if defined(__CUDACC__)
namespace vmsm {

// broadcast a field element from lane `src` (0..3) of the caller's quad
VMSM_D fe quad_get(const fe &v, int src) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, v.v[i], src, 4);
    return r;
}

// r = (q == 0) ? a0 : (q == 1) ? a1 : (q == 2) ? a2 : a3
VMSM_D fe quad_pick(int q, const fe &a0, const fe &a1, const fe &a2, const fe &a3) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t lo = (q & 1) ? a1.v[i] : a0.v[i];
        uint32_t hi = (q & 1) ? a3.v[i] : a2.v[i];
        r.v[i] = (q & 2) ? hi : lo;
    }
    return r;
}

// c = this lane's coordinate of the point (lane 0: X, 1: Y, 2: Z, 3: T)
VMSM_D fe quad_identity(int q) { return (q == 1 || q == 2) ? fe_one() : fe_zero(); }

// P + Q, 3 multiplication slots
VMSM_D fe quad_add(int q, const fe &p, const fe &r) {
    fe X1 = quad_get(p, 0), Y1 = quad_get(p, 1), X2 = quad_get(r, 0), Y2 = quad_get(r, 1);
    fe d1 = fe_sub(Y1, X1), s1 = fe_add(Y1, X1), d2 = fe_sub(Y2, X2), s2 = fe_add(Y2, X2);
    fe a = quad_pick(q, d1, s1, p, p);  // lane 2: Z1, lane 3: T1
    fe b = quad_pick(q, d2, s2, r, r);  // lane 2: Z2, lane 3: T2
    fe m = fe_mul(a, b);                // A, B, Z1Z2, T1T2
    fe k = (q == 3) ? fe_const_2d() : fe_one();
    m = fe_mul(m, k);                   // lane 3: C = 2d T1T2 (other lanes: times one)
    fe A = quad_get(m, 0), B = quad_get(m, 1), D = fe_dbl(quad_get(m, 2)), C = quad_get(m, 3);
    fe E = fe_sub(B, A), F = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
    fe u = quad_pick(q, E, G, F, E);
    fe v = quad_pick(q, F, H, G, H);
    return fe_mul(u, v);  // X3 = EF, Y3 = GH, Z3 = FG, T3 = EH
}

// P + Q with Q precomputed (y+x, y-x, 2dxy), Z2 = 1: 2 multiplication slots.  `b` is this lane's operand of the first
// stage: lane 0: y2-x2, lane 1: y2+x2, lane 2: the constant 2 (D = 2 Z1), lane 3: 2d x2 y2 (see quad_niels_operand).
VMSM_D fe quad_madd(int q, const fe &p, const fe &b) {
    fe X1 = quad_get(p, 0), Y1 = quad_get(p, 1);
    fe a = quad_pick(q, fe_sub(Y1, X1), fe_add(Y1, X1), p, p);  // lane 2: Z1, lane 3: T1
    fe m = fe_mul(a, b);                                         // A, B, D, C
    fe A = quad_get(m, 0), B = quad_get(m, 1), D = quad_get(m, 2), C = quad_get(m, 3);
    fe E = fe_sub(B, A), F = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
    fe u = quad_pick(q, E, G, F, E);
    fe v = quad_pick(q, F, H, G, H);
    return fe_mul(u, v);
}

// this lane's first-stage operand for adding (neg ? -Q : Q): -Q swaps y+x with y-x and negates 2dxy
VMSM_D fe quad_niels_operand(int q, const ge_niels &n, bool neg) {
    fe two = fe_zero();
    two.v[0] = 2u;
    return neg ? quad_pick(q, n.ypx, n.ymx, two, fe_neg(n.t2d)) : quad_pick(q, n.ymx, n.ypx, two, n.t2d);
}

// 2P, 2 multiplication slots
VMSM_D fe quad_dbl(int q, const fe &p) {
    fe X = quad_get(p, 0), Y = quad_get(p, 1);
    fe a = quad_pick(q, X, Y, p, fe_add(X, Y));  // lane 2: Z
    fe m = fe_sqr(a);                            // A = X^2, B = Y^2, Z^2, (X+Y)^2
    fe A = quad_get(m, 0), B = quad_get(m, 1), C = fe_dbl(quad_get(m, 2)), S = quad_get(m, 3);
    fe E = fe_sub(fe_sub(S, A), B);
    fe G = fe_sub(B, A);
    fe F = fe_sub(G, C);
    fe H = fe_neg(fe_add(A, B));
    fe u = quad_pick(q, E, G, F, E);
    fe v = quad_pick(q, F, H, G, H);
    return fe_mul(u, v);
}

// coordinate q of the extended point at p (32 B per lane, 128 B per quad, coalesced)
VMSM_D fe quad_load(const ge_ext *p, int q) {
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(p) + 32 * q);
    uint4 lo = __ldg(src), hi = __ldg(src + 1);
    fe r = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
    return r;
}
VMSM_D void quad_store(ge_ext *p, int q, const fe &c) {
    uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(p) + 32 * q);
    dst[0] = make_uint4(c.v[0], c.v[1], c.v[2], c.v[3]);
    dst[1] = make_uint4(c.v[4], c.v[5], c.v[6], c.v[7]);
}

}  // namespace vmsm
#endif
