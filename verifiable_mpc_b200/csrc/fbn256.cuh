// Base field of BN256 (the 256-bit Barreto-Naehrig curve of verifiable_mpc/ac20/pairing.py:51-56) and its quadratic
// extension Fp2 = Fp[i]/(i^2+1), for the Pinocchio prover-key multi-exponentiations (trinocchio/pynocchio.py:228-273).
//
// p has its top bit set (2p > 2^256): no spare bits for lazy reduction, so elements are kept fully reduced in
// Montgomery form a*2^256 mod p on 8 x 32-bit limbs.  Multiplication = the shared 64-product schoolbook core
// (mp_mul8, fe25519.cuh) + a word-serial Montgomery reduction whose 8 rows are two carry-chained 4-product chains
// each (64 more wide MADs); the carries out of a row are deposited in a side array and added once at the end, since
// they only ever reach limbs >= 8 that no later row reads for its quotient digit.  M = 136 limb products.
#pragma once
#include "fe25519.cuh"

namespace vmsm {

// Code-size control.  On the device every Montgomery multiplication in Fp / Fp2 is a CALL to one real function that
// takes and returns its operands BY VALUE -- the device ABI keeps 32 / 64-byte aggregates in registers, so a call
// costs a dozen moves, no local memory -- and the point operations of bn256.cuh are inlined around those calls.
// Round 1 did the opposite (multiplication inlined, point operations as functions taking references): a mixed
// addition was 12 000 SASS instructions that missed the instruction cache on every pass (ncu: stall_no_instruction
// 1.2 warps per issue) and every point crossed a call boundary through local memory (350 B of LDL/STL per addition
// in G1, 2 KB in G2).  The field inversion stays a function of its own.
#if defined(__CUDACC__)
#define VMSM_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define VMSM_HD_NOINLINE static inline
#endif

struct fbn {
    uint32_t v[8];
};

#define FBN_LIMBS(a, b, c, d, e, f, g, h) {{a, b, c, d, e, f, g, h}}
VMSM_HD fbn fbn_p() { fbn r = FBN_LIMBS(0x5e089667u, 0x185cac6cu, 0x20b5b59eu, 0xee5b88d1u, 0x6184dc21u, 0xaa6fecb8u, 0x4aa387f9u, 0x8fb501e3u); return r; }
VMSM_HD fbn fbn_one() { fbn r = FBN_LIMBS(0xa1f76999u, 0xe7a35393u, 0xdf4a4a61u, 0x11a4772eu, 0x9e7b23deu, 0x55901347u, 0xb55c7806u, 0x704afe1cu); return r; }
VMSM_HD fbn fbn_r2() { fbn r = FBN_LIMBS(0x7e444f56u, 0x9c21c3ffu, 0xb2efb0c2u, 0x409ed151u, 0x80fb1651u, 0x0c6dc37bu, 0x2c2380b7u, 0x7c36e0e6u); return r; }
#define FBN_PINV 0x7f17daa9u  // -p^-1 mod 2^32

VMSM_HD fbn fbn_zero() {
    fbn r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}

// r = a - p if (carry_in || a >= p) else a
VMSM_HD fbn fbn_cond_sub_p(const fbn &a, uint32_t carry_in) {
    const fbn p = fbn_p();
    fbn d;
#if defined(__CUDA_ARCH__)
    uint32_t nb;  // 0 - borrow: 0 when a >= p, 0xffffffff when a < p
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]),
          "=r"(d.v[7]), "=r"(nb)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(p.v[0]), "r"(p.v[1]), "r"(p.v[2]), "r"(p.v[3]), "r"(p.v[4]), "r"(p.v[5]), "r"(p.v[6]), "r"(p.v[7]));
    bool use = carry_in != 0 || nb == 0;
#else
    int64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        bw += (int64_t)a.v[i] - (int64_t)p.v[i];
        d.v[i] = (uint32_t)bw;
        bw >>= 32;
    }
    bool use = carry_in != 0 || bw == 0;
#endif
    fbn r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = use ? d.v[i] : a.v[i];
    return r;
}

VMSM_HD fbn fbn_add(const fbn &a, const fbn &b) {
    fbn s;
#if defined(__CUDA_ARCH__)
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]), "=r"(s.v[5]), "=r"(s.v[6]),
          "=r"(s.v[7]), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    return fbn_cond_sub_p(s, c);
#else
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        s.v[i] = (uint32_t)c;
        c >>= 32;
    }
    return fbn_cond_sub_p(s, (uint32_t)c);
#endif
}

VMSM_HD fbn fbn_sub(const fbn &a, const fbn &b) {
    const fbn p = fbn_p();
    fbn d;
#if defined(__CUDA_ARCH__)
    uint32_t m;  // 0xffffffff on borrow
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]),
          "=r"(d.v[7]), "=r"(m)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    fbn r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(d.v[0]), "r"(d.v[1]), "r"(d.v[2]), "r"(d.v[3]), "r"(d.v[4]), "r"(d.v[5]), "r"(d.v[6]), "r"(d.v[7]),
          "r"(p.v[0] & m), "r"(p.v[1] & m), "r"(p.v[2] & m), "r"(p.v[3] & m), "r"(p.v[4] & m), "r"(p.v[5] & m),
          "r"(p.v[6] & m), "r"(p.v[7] & m));
    return r;
#else
    int64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        bw += (int64_t)a.v[i] - (int64_t)b.v[i];
        d.v[i] = (uint32_t)bw;
        bw >>= 32;
    }
    uint32_t mask = bw ? 0xffffffffu : 0u;
    uint64_t c = 0;
    fbn r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)d.v[i] + (p.v[i] & mask);
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    return r;
#endif
}

VMSM_HD fbn fbn_neg(const fbn &a) { return fbn_sub(fbn_zero(), a); }
VMSM_HD fbn fbn_dbl(const fbn &a) { return fbn_add(a, a); }

VMSM_HD bool fbn_is_zero(const fbn &a) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= a.v[i];
    return x == 0;
}
VMSM_HD bool fbn_eq(const fbn &a, const fbn &b) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= a.v[i] ^ b.v[i];
    return x == 0;
}
VMSM_HD fbn fbn_select(bool c, const fbn &a, const fbn &b) {
    fbn r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
}

// Montgomery reduction of a 512-bit value t < p * 2^256: returns t / 2^256 mod p
VMSM_HD fbn fbn_redc(uint32_t *t) {
    const fbn p = fbn_p();
    fbn r;
#if defined(__CUDA_ARCH__)
    uint32_t cy[18];
#pragma unroll
    for (int i = 0; i < 18; i++) cy[i] = 0;
    uint32_t tt[17];
#pragma unroll
    for (int i = 0; i < 16; i++) tt[i] = t[i];
    tt[16] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t m = tt[i] * FBN_PINV;
        // m * (p0,p2,p4,p6) at limbs i..i+7, carry -> cy[i+8];  m * (p1,p3,p5,p7) at limbs i+1..i+8, carry -> cy[i+9]
        fe_mad4(tt[i], tt[i + 1], tt[i + 2], tt[i + 3], tt[i + 4], tt[i + 5], tt[i + 6], tt[i + 7], cy[i + 8], p.v[0],
                p.v[2], p.v[4], p.v[6], m);
        fe_mad4(tt[i + 1], tt[i + 2], tt[i + 3], tt[i + 4], tt[i + 5], tt[i + 6], tt[i + 7], tt[i + 8], cy[i + 9], p.v[1],
                p.v[3], p.v[5], p.v[7], m);
    }
    uint32_t top;
    asm("add.cc.u32 %0, %9, %18;\n\t"
        "addc.cc.u32 %1, %10, %19;\n\t"
        "addc.cc.u32 %2, %11, %20;\n\t"
        "addc.cc.u32 %3, %12, %21;\n\t"
        "addc.cc.u32 %4, %13, %22;\n\t"
        "addc.cc.u32 %5, %14, %23;\n\t"
        "addc.cc.u32 %6, %15, %24;\n\t"
        "addc.cc.u32 %7, %16, %25;\n\t"
        "addc.u32 %8, %17, %26;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(top)
        : "r"(tt[8]), "r"(tt[9]), "r"(tt[10]), "r"(tt[11]), "r"(tt[12]), "r"(tt[13]), "r"(tt[14]), "r"(tt[15]),
          "r"(tt[16]), "r"(cy[8]), "r"(cy[9]), "r"(cy[10]), "r"(cy[11]), "r"(cy[12]), "r"(cy[13]), "r"(cy[14]),
          "r"(cy[15]), "r"(cy[16]));
    return fbn_cond_sub_p(r, top);
#else
    uint64_t top = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t m = t[i] * FBN_PINV;
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)t[i + j] + (uint64_t)m * p.v[j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        for (int k = i + 8; k < 16 && c; k++) {
            c += t[k];
            t[k] = (uint32_t)c;
            c >>= 32;
        }
        top += c;
    }
    for (int i = 0; i < 8; i++) r.v[i] = t[8 + i];
    return fbn_cond_sub_p(r, (uint32_t)top);
#endif
}

VMSM_HD fbn fbn_mul_inl(const fbn &a, const fbn &b) {
    uint32_t t[16];
    mp_mul8(a.v, b.v, t);
    return fbn_redc(t);
}
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ fbn fbn_mul_fn(fbn a, fbn b) { return fbn_mul_inl(a, b); }
VMSM_HD fbn fbn_mul(const fbn &a, const fbn &b) { return fbn_mul_fn(a, b); }
#else
VMSM_HD fbn fbn_mul(const fbn &a, const fbn &b) { return fbn_mul_inl(a, b); }
#endif
VMSM_HD fbn fbn_sqr(const fbn &a) { return fbn_mul(a, a); }

// plain integer (< p) <-> Montgomery form
VMSM_HD fbn fbn_to_mont(const fbn &plain) { return fbn_mul(plain, fbn_r2()); }
VMSM_HD fbn fbn_from_mont(const fbn &m) {
    uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = m.v[i], t[8 + i] = 0;
    return fbn_redc(t);
}
VMSM_HD bool fbn_plain_is_canonical(const fbn &a) {
    const fbn p = fbn_p();
    int64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        bw += (int64_t)a.v[i] - (int64_t)p.v[i];
        bw >>= 32;
    }
    return bw != 0;  // borrow <=> a < p
}

// a^(p-2) by square-and-multiply over the fixed exponent (Montgomery form in and out)
VMSM_HD_NOINLINE fbn fbn_inv(const fbn &a) {
    const uint32_t e[8] = {0x5e089665u, 0x185cac6cu, 0x20b5b59eu, 0xee5b88d1u, 0x6184dc21u, 0xaa6fecb8u, 0x4aa387f9u, 0x8fb501e3u};
    fbn r = fbn_one();
    for (int i = 255; i >= 0; i--) {
        r = fbn_sqr(r);
        if ((e[i >> 5] >> (i & 31)) & 1u) r = fbn_mul(r, a);
    }
    return r;
}

// ---------------------------------------------------------------------------------------------- Fp2 = Fp[i]/(i^2+1)
struct f2bn {
    fbn c0, c1;  // c0 + c1 * i
};

VMSM_HD f2bn f2bn_zero() { f2bn r = {fbn_zero(), fbn_zero()}; return r; }
VMSM_HD f2bn f2bn_one() { f2bn r = {fbn_one(), fbn_zero()}; return r; }
VMSM_HD f2bn f2bn_add(const f2bn &a, const f2bn &b) { f2bn r = {fbn_add(a.c0, b.c0), fbn_add(a.c1, b.c1)}; return r; }
VMSM_HD f2bn f2bn_sub(const f2bn &a, const f2bn &b) { f2bn r = {fbn_sub(a.c0, b.c0), fbn_sub(a.c1, b.c1)}; return r; }
VMSM_HD f2bn f2bn_neg(const f2bn &a) { f2bn r = {fbn_neg(a.c0), fbn_neg(a.c1)}; return r; }
VMSM_HD f2bn f2bn_dbl(const f2bn &a) { f2bn r = {fbn_dbl(a.c0), fbn_dbl(a.c1)}; return r; }
VMSM_HD bool f2bn_is_zero(const f2bn &a) { return fbn_is_zero(a.c0) && fbn_is_zero(a.c1); }
VMSM_HD bool f2bn_eq(const f2bn &a, const f2bn &b) { return fbn_eq(a.c0, b.c0) && fbn_eq(a.c1, b.c1); }
VMSM_HD f2bn f2bn_select(bool c, const f2bn &a, const f2bn &b) {
    f2bn r = {fbn_select(c, a.c0, b.c0), fbn_select(c, a.c1, b.c1)};
    return r;
}
// Karatsuba: 3 base-field multiplications
VMSM_HD f2bn f2bn_mul_inl(const f2bn &a, const f2bn &b) {
    fbn t0 = fbn_mul(a.c0, b.c0), t1 = fbn_mul(a.c1, b.c1);
    fbn t2 = fbn_mul(fbn_add(a.c0, a.c1), fbn_add(b.c0, b.c1));
    f2bn r = {fbn_sub(t0, t1), fbn_sub(fbn_sub(t2, t0), t1)};
    return r;
}
// (a0 + a1 i)^2 = (a0+a1)(a0-a1) + 2 a0 a1 i : 2 multiplications
VMSM_HD f2bn f2bn_sqr_inl(const f2bn &a) {
    fbn m = fbn_mul(a.c0, a.c1);
    f2bn r = {fbn_mul(fbn_add(a.c0, a.c1), fbn_sub(a.c0, a.c1)), fbn_dbl(m)};
    return r;
}
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ f2bn f2bn_mul_fn(f2bn a, f2bn b) { return f2bn_mul_inl(a, b); }
static __device__ __noinline__ f2bn f2bn_sqr_fn(f2bn a) { return f2bn_sqr_inl(a); }
VMSM_HD f2bn f2bn_mul(const f2bn &a, const f2bn &b) { return f2bn_mul_fn(a, b); }
VMSM_HD f2bn f2bn_sqr(const f2bn &a) { return f2bn_sqr_fn(a); }
#else
VMSM_HD f2bn f2bn_mul(const f2bn &a, const f2bn &b) { return f2bn_mul_inl(a, b); }
VMSM_HD f2bn f2bn_sqr(const f2bn &a) { return f2bn_sqr_inl(a); }
#endif
VMSM_HD f2bn f2bn_inv(const f2bn &a) {
    fbn d = fbn_inv(fbn_add(fbn_sqr(a.c0), fbn_sqr(a.c1)));
    f2bn r = {fbn_mul(a.c0, d), fbn_neg(fbn_mul(a.c1, d))};
    return r;
}

// ---------------------------------------------------------------------------------------------- field policies
// The curve code (bn256.cuh) is written once over these.
struct FpBN {
    typedef fbn T;
    enum { kWords = 8 };  // 32-bit words per element (wire and storage)
    static VMSM_HD T zero() { return fbn_zero(); }
    static VMSM_HD T one() { return fbn_one(); }
    static VMSM_HD T add(const T &a, const T &b) { return fbn_add(a, b); }
    static VMSM_HD T sub(const T &a, const T &b) { return fbn_sub(a, b); }
    static VMSM_HD T neg(const T &a) { return fbn_neg(a); }
    static VMSM_HD T dbl(const T &a) { return fbn_dbl(a); }
    static VMSM_HD T mul(const T &a, const T &b) { return fbn_mul(a, b); }
    static VMSM_HD T sqr(const T &a) { return fbn_sqr(a); }
    static VMSM_HD T inv(const T &a) { return fbn_inv(a); }
    static VMSM_HD bool is_zero(const T &a) { return fbn_is_zero(a); }
    static VMSM_HD bool eq(const T &a, const T &b) { return fbn_eq(a, b); }
    static VMSM_HD T select(bool c, const T &a, const T &b) { return fbn_select(c, a, b); }
    static VMSM_HD T to_mont(const T &a) { return fbn_to_mont(a); }
    static VMSM_HD T from_mont(const T &a) { return fbn_from_mont(a); }
    static VMSM_HD bool plain_ok(const T &a) { return fbn_plain_is_canonical(a); }
    // curve constant b = 3 (Montgomery form)
    static VMSM_HD T curve_b() { fbn r = FBN_LIMBS(0x29d50ffdu, 0x8630a1e2u, 0x5c7373e9u, 0x583653eau, 0x1867b356u, 0xabd06066u, 0x8ace581fu, 0x3176f68fu); return r; }
    static VMSM_HD T gen_x() { return fbn_one(); }
    static VMSM_HD T gen_y() { fbn r = FBN_LIMBS(0x7822599cu, 0x6172b1b1u, 0x82d6d678u, 0xb96e2344u, 0x86137087u, 0xa9bfb2e1u, 0x2a8e1fe6u, 0x3ed4078du); return r; }
};

struct Fp2BN {
    typedef f2bn T;
    enum { kWords = 16 };
    static VMSM_HD T zero() { return f2bn_zero(); }
    static VMSM_HD T one() { return f2bn_one(); }
    static VMSM_HD T add(const T &a, const T &b) { return f2bn_add(a, b); }
    static VMSM_HD T sub(const T &a, const T &b) { return f2bn_sub(a, b); }
    static VMSM_HD T neg(const T &a) { return f2bn_neg(a); }
    static VMSM_HD T dbl(const T &a) { return f2bn_dbl(a); }
    static VMSM_HD T mul(const T &a, const T &b) { return f2bn_mul(a, b); }
    static VMSM_HD T sqr(const T &a) { return f2bn_sqr(a); }
    static VMSM_HD T inv(const T &a) { return f2bn_inv(a); }
    static VMSM_HD bool is_zero(const T &a) { return f2bn_is_zero(a); }
    static VMSM_HD bool eq(const T &a, const T &b) { return f2bn_eq(a, b); }
    static VMSM_HD T select(bool c, const T &a, const T &b) { return f2bn_select(c, a, b); }
    static VMSM_HD T to_mont(const T &a) { T r = {fbn_to_mont(a.c0), fbn_to_mont(a.c1)}; return r; }
    static VMSM_HD T from_mont(const T &a) { T r = {fbn_from_mont(a.c0), fbn_from_mont(a.c1)}; return r; }
    static VMSM_HD bool plain_ok(const T &a) { return fbn_plain_is_canonical(a.c0) && fbn_plain_is_canonical(a.c1); }
    // twist constant b' = 3 / (i + 3)
    static VMSM_HD T curve_b() {
        fbn re = FBN_LIMBS(0xb4c5ee14u, 0xb94f760fu, 0x4c3b6eb4u, 0xdae9f8f2u, 0xe52f4fe4u, 0x77a675d2u, 0x9116c66bu, 0x736f31b0u);
        fbn im = FBN_LIMBS(0x386b8d71u, 0x75046774u, 0x46d36cf8u, 0x5bd0854au, 0xd41c8414u, 0x664327a1u, 0x932eeb2fu, 0x096c9abbu);
        T r = {re, im};
        return r;
    }
    static VMSM_HD T gen_x() {
        fbn re = FBN_LIMBS(0xa7cdc184u, 0x88f9f11du, 0xd69509d3u, 0x18293f95u, 0xa735d5a1u, 0xb5ce0c55u, 0x9bfd45a0u, 0x01513418u);
        fbn im = FBN_LIMBS(0x139e1404u, 0x402c4ab7u, 0x183d85a4u, 0xce1c368au, 0xcb8d3983u, 0xd67cf9a6u, 0xc2a9fbe8u, 0x3cf246bbu);
        T r = {re, im};
        return r;
    }
    static VMSM_HD T gen_y() {
        fbn re = FBN_LIMBS(0x63ea9e56u, 0xc2e07c14u, 0x2072ebd2u, 0xee444205u, 0x86036937u, 0x561a5194u, 0xcc0d2cceu, 0x05bd9394u);
        fbn im = FBN_LIMBS(0x1e9e87a2u, 0xbfac7d73u, 0x7962e441u, 0xa50bb800u, 0xe8270556u, 0xafe910a4u, 0x9d69159au, 0x5075c542u);
        T r = {re, im};
        return r;
    }
};

}  // namespace vmsm
