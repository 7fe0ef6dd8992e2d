// GF(2^255-19) on 8 x 32-bit limbs for sm_100a.
//
// Representation: a field element is ANY 256-bit value (not necessarily < p); 2^256 == 38 (mod p), so a
// carry/borrow out of limb 7 is folded back as +-38.  Every routine accepts and returns this "weak" form;
// fe_canon() produces the unique representative in [0, p) used on the wire and for equality.
//
// Device path: inline PTX mad.lo.cc / madc.hi.cc chains laid out so that ptxas fuses each lo/hi pair into
// one IMAD.WIDE.U32 with predicate carry (64 + 8 + 1 wide multiplies per field multiplication; SURVEY.md
// 8d counts 72 limb-products).  Host path (VMSM_HD without __CUDA_ARCH__): the same functions in portable
// 64-bit C so tests/ can exercise every kernel body on the CPU ("host emulation", tests only).
//
// Replaces (behaviourally) the base-field arithmetic of MPyC's GF(2^255-19) that the reference reaches
// through `g[i] ** x` / `a * b` (verifiable_mpc/ac20/pivot.py:143-144, compressed_pivot.py:64).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VMSM_HD __host__ __device__ __forceinline__
#define VMSM_D __device__ __forceinline__
#else
#define VMSM_HD inline
#define VMSM_D inline
#endif

namespace vmsm {

struct fe {
    uint32_t v[8];
};

VMSM_HD fe fe_zero() {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
VMSM_HD fe fe_one() {
    fe r = fe_zero();
    r.v[0] = 1;
    return r;
}

// --------------------------------------------------------------------------------------------- add / sub
VMSM_HD fe fe_add(const fe &a, const fe &b) {
    fe r;
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
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // fold the carry: += 38*c; a second carry can only happen when the wrapped value is < 38, so the last
    // fix-up touches limb 0 only.
    uint32_t f = c * 38u, c2;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7]), "=r"(c2)
        : "r"(f));
    r.v[0] += c2 * 38u;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    uint64_t f = c * 38u;
    for (int i = 0; i < 8; i++) {
        f += r.v[i];
        r.v[i] = (uint32_t)f;
        f >>= 32;
    }
    r.v[0] += (uint32_t)f * 38u;
#endif
    return r;
}

VMSM_HD fe fe_sub(const fe &a, const fe &b) {
    fe r;
#if defined(__CUDA_ARCH__)
    uint32_t bw;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"  // 0 - 0 - borrow = 0 or 0xFFFFFFFF
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(bw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // borrow: the stored value is a-b+2^256 == a-b+38, so subtract 38; a second borrow only when that value < 38.
    uint32_t f = bw & 38u, b2;
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.cc.u32 %7, %7, 0;\n\t"
        "subc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7]), "=r"(b2)
        : "r"(f));
    r.v[0] -= b2 & 38u;
#else
    int64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (int64_t)a.v[i] - (int64_t)b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;  // arithmetic shift: 0 or -1
    }
    int64_t f = c ? -38 : 0;
    for (int i = 0; i < 8; i++) {
        f += (int64_t)r.v[i];
        r.v[i] = (uint32_t)f;
        f >>= 32;
    }
    if (f) r.v[0] -= 38u;
#endif
    return r;
}

VMSM_HD fe fe_neg(const fe &a) { return fe_sub(fe_zero(), a); }
VMSM_HD fe fe_dbl(const fe &a) { return fe_add(a, a); }

// --------------------------------------------------------------------------------------------- reduce 512 -> 256
// t[0..15] -> r = t mod (2^256-38), weak form.
VMSM_HD fe fe_reduce512(const uint32_t *t) {
    fe r;
#if defined(__CUDA_ARCH__)
    uint32_t r8;
    const uint32_t k = 38u;
    // even columns: (r0,r1) = lo(0,1) + 38*hi0 ; (r2,r3) = lo(2,3) + 38*hi2 ; ...
    asm("mad.lo.cc.u32 %0, %17, %25, %9;\n\t"
        "madc.hi.cc.u32 %1, %17, %25, %10;\n\t"
        "madc.lo.cc.u32 %2, %19, %25, %11;\n\t"
        "madc.hi.cc.u32 %3, %19, %25, %12;\n\t"
        "madc.lo.cc.u32 %4, %21, %25, %13;\n\t"
        "madc.hi.cc.u32 %5, %21, %25, %14;\n\t"
        "madc.lo.cc.u32 %6, %23, %25, %15;\n\t"
        "madc.hi.cc.u32 %7, %23, %25, %16;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        // odd columns, shifted one limb
        "mad.lo.cc.u32 %1, %18, %25, %1;\n\t"
        "madc.hi.cc.u32 %2, %18, %25, %2;\n\t"
        "madc.lo.cc.u32 %3, %20, %25, %3;\n\t"
        "madc.hi.cc.u32 %4, %20, %25, %4;\n\t"
        "madc.lo.cc.u32 %5, %22, %25, %5;\n\t"
        "madc.hi.cc.u32 %6, %22, %25, %6;\n\t"
        "madc.lo.cc.u32 %7, %24, %25, %7;\n\t"
        "madc.hi.u32 %8, %24, %25, %8;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]),
          "=&r"(r.v[7]), "=&r"(r8)
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(t[8]),
          "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]), "r"(k));
    // r8 <= 39: fold r8 * 2^256 == r8 * 38
    uint32_t f = r8 * 38u, c2;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7]), "=r"(c2)
        : "r"(f));
    r.v[0] += c2 * 38u;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)t[i] + (uint64_t)t[i + 8] * 38u;
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    uint64_t f = c * 38u;
    for (int i = 0; i < 8; i++) {
        f += r.v[i];
        r.v[i] = (uint32_t)f;
        f >>= 32;
    }
    r.v[0] += (uint32_t)f * 38u;
#endif
    return r;
}

// --------------------------------------------------------------------------------------------- mul
#if defined(__CUDACC__)
// acc[0..7] += {a0,a1,a2,a3} * b laid out as four 64-bit columns; carry out -> acc8 (which holds at most a few
// earlier carries, so it cannot overflow).  Each lo/hi pair fuses into one IMAD.WIDE.U32 in SASS.
VMSM_D void fe_mad4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t &r4, uint32_t &r5, uint32_t &r6,
                    uint32_t &r7, uint32_t &r8, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(r8)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// Same without a carry-out limb (top of the 512-bit product: the carry is provably zero).
VMSM_D void fe_mad4_top(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t &r4, uint32_t &r5,
                        uint32_t &r6, uint32_t &r7, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
#endif

// 256 x 256 -> 512-bit schoolbook product, t[0..15].  Device: even/odd column accumulators so that every
// mad.lo.cc / madc.hi.cc pair fuses into one IMAD.WIDE.U32[.X]; shared with the BN256 Montgomery field (fbn256.cuh).
VMSM_HD void mp_mul8(const uint32_t *a, const uint32_t *b, uint32_t *t) {
#if defined(__CUDA_ARCH__)
    // e[k] is limb k of the sum of a_i*b_j with i+j even; o[k] is limb k+1 of the sum with i+j odd.
    uint32_t e[17], o[17];
#pragma unroll
    for (int i = 0; i < 17; i++) e[i] = o[i] = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if ((j & 1) == 0) {
            // even part: a0,a2,a4,a6 at limbs j..j+7 ; odd part: a1,a3,a5,a7 at limbs j+1..j+8 (o index j..j+7)
            fe_mad4(e[j], e[j + 1], e[j + 2], e[j + 3], e[j + 4], e[j + 5], e[j + 6], e[j + 7], e[j + 8], a[0],
                    a[2], a[4], a[6], b[j]);
            fe_mad4(o[j], o[j + 1], o[j + 2], o[j + 3], o[j + 4], o[j + 5], o[j + 6], o[j + 7], o[j + 8], a[1],
                    a[3], a[5], a[7], b[j]);
        } else {
            // even part: a1,a3,a5,a7 at limbs j+1..j+8 ; odd part: a0,a2,a4,a6 at limbs j..j+7 (o index j-1..j+6)
            if (j == 7)
                fe_mad4_top(e[8], e[9], e[10], e[11], e[12], e[13], e[14], e[15], a[1], a[3], a[5], a[7], b[j]);
            else
                fe_mad4(e[j + 1], e[j + 2], e[j + 3], e[j + 4], e[j + 5], e[j + 6], e[j + 7], e[j + 8], e[j + 9],
                        a[1], a[3], a[5], a[7], b[j]);
            fe_mad4(o[j - 1], o[j], o[j + 1], o[j + 2], o[j + 3], o[j + 4], o[j + 5], o[j + 6], o[j + 7], a[0],
                    a[2], a[4], a[6], b[j]);
        }
    }
    // t = e + (o << 32); o[15] (limb 16) is provably zero.
    t[0] = e[0];
    asm("add.cc.u32 %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32 %14, %29, %44;"
        : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]),
          "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]), "r"(e[9]),
          "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]), "r"(o[0]), "r"(o[1]), "r"(o[2]),
          "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]),
          "r"(o[12]), "r"(o[13]), "r"(o[14]));
#else
    for (int i = 0; i < 16; i++) t[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)t[i + j] + (uint64_t)a[i] * b[j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        t[i + 8] = (uint32_t)c;
    }
#endif
}

VMSM_HD fe fe_mul(const fe &a, const fe &b) {
    uint32_t t[16];
    mp_mul8(a.v, b.v, t);
    return fe_reduce512(t);
}

#if defined(__CUDACC__)
// shorter carry chains for the off-diagonal rows of a squaring (same layout as fe_mad4)
VMSM_D void fe_mad3(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t &r4, uint32_t &r5, uint32_t &r6,
                    uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32 %6, %6, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6)
        : "r"(a0), "r"(a1), "r"(a2), "r"(b));
}
VMSM_D void fe_mad2(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t &r4, uint32_t a0, uint32_t a1,
                    uint32_t b) {
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4)
        : "r"(a0), "r"(a1), "r"(b));
}
VMSM_D void fe_mad1(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t a0, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2)
        : "r"(a0), "r"(b));
}
#endif

// Squaring: 28 off-diagonal products (even/odd column accumulators as in fe_mul), doubled with funnel shifts, plus
// one 8-product carry chain for the diagonal: 36 + 8 + 1 wide MADs instead of 64 + 8 + 1.
VMSM_HD fe fe_sqr(const fe &a) {
#if defined(__CUDA_ARCH__)
    uint32_t e[17], o[17], t[16];
#pragma unroll
    for (int i = 0; i < 17; i++) e[i] = o[i] = 0;
    // multiplier a0: even a2,a4,a6 -> limbs 2..7 ; odd a1,a3,a5,a7 -> limbs 1..8 (o index 0..7)
    fe_mad3(e[2], e[3], e[4], e[5], e[6], e[7], e[8], a.v[2], a.v[4], a.v[6], a.v[0]);
    fe_mad4(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], a.v[1], a.v[3], a.v[5], a.v[7], a.v[0]);
    // a1: even a3,a5,a7 -> limbs 4..9 ; odd a2,a4,a6 -> limbs 3..8 (o 2..7)
    fe_mad3(e[4], e[5], e[6], e[7], e[8], e[9], e[10], a.v[3], a.v[5], a.v[7], a.v[1]);
    fe_mad3(o[2], o[3], o[4], o[5], o[6], o[7], o[8], a.v[2], a.v[4], a.v[6], a.v[1]);
    // a2: even a4,a6 -> limbs 6..9 ; odd a3,a5,a7 -> limbs 5..10 (o 4..9)
    fe_mad2(e[6], e[7], e[8], e[9], e[10], a.v[4], a.v[6], a.v[2]);
    fe_mad3(o[4], o[5], o[6], o[7], o[8], o[9], o[10], a.v[3], a.v[5], a.v[7], a.v[2]);
    // a3: even a5,a7 -> limbs 8..11 ; odd a4,a6 -> limbs 7..10 (o 6..9)
    fe_mad2(e[8], e[9], e[10], e[11], e[12], a.v[5], a.v[7], a.v[3]);
    fe_mad2(o[6], o[7], o[8], o[9], o[10], a.v[4], a.v[6], a.v[3]);
    // a4: even a6 -> limbs 10,11 ; odd a5,a7 -> limbs 9..12 (o 8..11)
    fe_mad1(e[10], e[11], e[12], a.v[6], a.v[4]);
    fe_mad2(o[8], o[9], o[10], o[11], o[12], a.v[5], a.v[7], a.v[4]);
    // a5: even a7 -> limbs 12,13 ; odd a6 -> limbs 11,12 (o 10,11)
    fe_mad1(e[12], e[13], e[14], a.v[7], a.v[5]);
    fe_mad1(o[10], o[11], o[12], a.v[6], a.v[5]);
    // a6: odd a7 -> limbs 13,14 (o 12,13)
    fe_mad1(o[12], o[13], o[14], a.v[7], a.v[6]);
    // u = e + (o << 32)
    uint32_t u[16];
    u[0] = e[0];
    asm("add.cc.u32 %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32 %14, %29, %44;"
        : "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]), "r"(e[9]),
          "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]), "r"(o[0]), "r"(o[1]), "r"(o[2]),
          "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]),
          "r"(o[12]), "r"(o[13]), "r"(o[14]));
    // t = 2u (u < 2^511, so nothing is shifted out)
    t[0] = u[0] << 1;
#pragma unroll
    for (int i = 1; i < 16; i++) t[i] = __funnelshift_l(u[i - 1], u[i], 1);
    // t += sum a_i^2 * 2^(64 i): one carry chain over all 16 limbs
    asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
        "madc.hi.u32 %15, %23, %23, %15;"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]),
          "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
    return fe_reduce512(t);
#else
    return fe_mul(a, a);
#endif
}


// a * small constant (< 2^32)
VMSM_HD fe fe_mul_small(const fe &a, uint32_t k) {
    uint32_t t[16];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] * k;
        t[i] = (uint32_t)c;
        c >>= 32;
    }
    t[8] = (uint32_t)c;
#pragma unroll
    for (int i = 9; i < 16; i++) t[i] = 0;
    return fe_reduce512(t);
}

// --------------------------------------------------------------------------------------------- canonical form
VMSM_HD fe fe_canon(const fe &a) {
    // 1) fold bit 255: v = (v mod 2^255) + 19*(v >> 255)  -> v < 2^255 + 19
    fe r = a;
    uint64_t c = (uint64_t)(r.v[7] >> 31) * 19u;
    r.v[7] &= 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += r.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    // 2) q = (v + 19) >> 255 is 1 iff v >= p ; v = (v + 19 q) mod 2^255
    uint64_t d = 19;
    uint32_t s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        d += r.v[i];
        s[i] = (uint32_t)d;
        d >>= 32;
    }
    uint32_t q = s[7] >> 31;
    c = (uint64_t)q * 19u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += r.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    r.v[7] &= 0x7fffffffu;
    return r;
}

VMSM_HD bool fe_is_zero(const fe &a) {
    fe c = fe_canon(a);
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= c.v[i];
    return x == 0;
}

VMSM_HD bool fe_eq(const fe &a, const fe &b) { return fe_is_zero(fe_sub(a, b)); }

// r = cond ? a : b   (branch-free select)
VMSM_HD fe fe_select(bool cond, const fe &a, const fe &b) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = cond ? a.v[i] : b.v[i];
    return r;
}

// --------------------------------------------------------------------------------------------- inversion
VMSM_HD fe fe_sqr_n(fe a, int n) {
    for (int i = 0; i < n; i++) a = fe_sqr(a);
    return a;
}

// z^(p-2), p-2 = 2^255 - 21 : the classic 254-squaring / 11-multiplication chain.
VMSM_HD fe fe_inv(const fe &z) {
    fe z2 = fe_sqr(z);                        // 2
    fe z9 = fe_mul(fe_sqr_n(z2, 2), z);       // 9
    fe z11 = fe_mul(z9, z2);                  // 11
    fe z2_5_0 = fe_mul(fe_sqr(z11), z9);      // 2^5 - 1
    fe z2_10_0 = fe_mul(fe_sqr_n(z2_5_0, 5), z2_5_0);
    fe z2_20_0 = fe_mul(fe_sqr_n(z2_10_0, 10), z2_10_0);
    fe z2_40_0 = fe_mul(fe_sqr_n(z2_20_0, 20), z2_20_0);
    fe z2_50_0 = fe_mul(fe_sqr_n(z2_40_0, 10), z2_10_0);
    fe z2_100_0 = fe_mul(fe_sqr_n(z2_50_0, 50), z2_50_0);
    fe z2_200_0 = fe_mul(fe_sqr_n(z2_100_0, 100), z2_100_0);
    fe z2_250_0 = fe_mul(fe_sqr_n(z2_200_0, 50), z2_50_0);
    return fe_mul(fe_sqr_n(z2_250_0, 5), z11);  // 2^255 - 21
}

// --------------------------------------------------------------------------------------------- bytes
// 32-byte little-endian <-> limbs.  fe_from_bytes does not reduce; callers that need "canonical input"
// check fe_is_canonical().
VMSM_HD bool fe_is_canonical(const fe &a) {
    // a < p  <=>  a + 19 < 2^255
    uint64_t d = 19;
    uint32_t top = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        d += a.v[i];
        top = (uint32_t)d;
        d >>= 32;
    }
    return d == 0 && (top >> 31) == 0;
}

}  // namespace vmsm
